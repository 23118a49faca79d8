"""
oracle/ -- CPU restatement of the tfrt ray-trace hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in float64 torch-CPU ops, the algorithm of the reference
`OpticalEngine.ray_trace` path (tfrt/geometry.py, tfrt/engine.py, tfrt/operation.py,
tfrt/materials.py, tfrt/boundaries.py).  It exists so that the CUDA path can be checked
against the reference semantics, and so that bench.py can time a CPU baseline.

Rules:
  * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
    legs may import this package.  The product package (tensorflowraytrace_b200/) never
    imports it and has no CPU fallback.
  * Every function cites the reference file:line it follows.

PARITY STATUS: **parity unpinned**.  The reference is pure Python on TensorFlow; TensorFlow
is not installed in the build container and not installable (no wheel, no network), so the
reference itself cannot be executed here, and the reference's own test-suite holds no golden
vectors for this path (its tests are stale TF1 code touching only 2D line/circle maths; see
SURVEY.md section 4).  The oracle is therefore pinned by:
  (i)   the known-answer *properties* of the reference's tests (tests/test_oracle_*.py),
  (ii)  an independent closed form (Moller-Trumbore vs. the reference's Cramer solve),
  (iii) finite-difference checks of its gradients in float64,
  (iv)  physical invariants (Snell's law, unit directions, reflection symmetry).
"""
