"""Collect the bench.py JSON lines brought back in gpurun_out/ into profiles/r2_bench_lines.md and
profiles/r2_scaling.md (tables + raw lines).

usage: collect_bench_lines.py <single-GPU prefix> [<scaling prefix>]
       e.g. collect_bench_lines.py r2_c18_bench_ r2_c19_bench_cfg5_
       reads gpurun_out/<prefix>cfg1..cfg5.log, <prefix>ref.log, <scaling prefix>{1,2,4,8}gpu.log
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
prefix = sys.argv[1] if len(sys.argv) > 1 else "r2_final_bench_"
scale_prefix = sys.argv[2] if len(sys.argv) > 2 else None


def load(f):
    try:
        return json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        return None


out = ["# Round 2 — bench.py lines (1 × B200, `gpurun`)\n",
       "Raw JSON printed by `python bench.py --config <cfg>` (default steps / warm-up; cfg5 is the default "
       "config). `value` = ray × boundary tests/s of the step with rays resident in HBM (config 3 / 5: "
       "regenerated on the device at every step); `e2e` = the same step from pinned host rays with gradients "
       "and loss copied back; `frac_executed` = 25 executed FLOP × tests ÷ `k_intersect3d` time ÷ the FP32 "
       "FMA peak measured in the same run, `FMA pipe` = 14 lane-slots per test ÷ the pipe's rate; "
       "`roofline.frac` uses the ALGORITHMIC 45-FLOP Möller–Trumbore count and therefore exceeds 1.\n",
       "| config | workload | ms/step | tests/s | fwd-only ms | k_intersect3d share | FMA pipe | "
       "frac_executed | roofline.frac | e2e ms/step | backward leg ms (frac of HBM) | re-validated / gross | "
       "CPU port tests/s (cores) |",
       "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for c in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
    d = load(f"gpurun_out/{prefix}{c}.log")
    if not d:
        continue
    r, cb = d.get("roofline", {}), d.get("cpu_baseline", {})
    kind = "fwd+bwd" if "fwd_bwd" in d["metric"] else "fwd"
    cfg = d["config"]
    rb = d.get("roofline_backward")
    rv = d.get("revalidation", {})
    if r.get("frac"):
        cols = (f"{r['share_of_step']:.3f} | {r['fma_pipe_utilisation']:.3f} | {r['frac_executed']:.3f} | "
                f"{r['frac']:.3f}")
    else:
        cols = (f"— | — | — | latency: {r.get('us_per_ray_trace_call', 0):.0f} µs per ray_trace call, "
                f"{r.get('us_per_trace_kernel_only', 0):.1f} µs kernel only")
    out.append(
        f"| {c} | {cfg.get('rays')} rays × {cfg.get('triangles')} × {cfg.get('iterations')} it, {kind} | "
        f"{d['ms_per_step']:.3f} | {d['value']:.4g} | {d['forward_only']['ms_per_step']:.3f} | {cols} | "
        f"{d['e2e'].get('ms_per_step', 0):.3f} | "
        + (f"{rb['ms_per_step']:.2f} ({rb['frac']:.2f})" if rb else "—")
        + f" | {rv.get('revalidated', 0)} / {rv.get('gross', 0)} | "
        + (f"{cb.get('value', 0):.3g} ({cb.get('cores')})" if cb else "—") + " |")
out.append("")
d5 = load(f"gpurun_out/{prefix}cfg5.log")
if d5 and d5.get("accel_grid"):
    a = d5["accel_grid"]
    out.append(f"Acceleration grid (`accel_grid`, optional mode, outside the tests/s metric), cfg5: "
               f"{a['ms_per_step']:.1f} ms per step ({a['speedup_over_brute_force_step']:.1f}× the brute-force "
               f"step), search kernel {a['search_kernel_ms_per_launch']:.2f} ms per launch, grid "
               f"{a['grid']['dims']} cells, {a['grid']['large']} large triangles.\n")
d1 = load(f"gpurun_out/{prefix}cfg1.log")
if d1 and d1.get("two_d_at_scale"):
    out.append("2D at scale (`two_d_at_scale`, the four sizes of `dev/performance.py:77-86`):\n")
    out.append("```json\n" + json.dumps(d1["two_d_at_scale"], indent=1) + "\n```\n")
ref = load(f"gpurun_out/{prefix}ref.log")
if ref:
    out.append(f"`bench.py --impl reference` (cfg5, CPU restatement, {ref['cpu_baseline']['cores']} host "
               f"cores): {ref['value']:.4g} tests/s, {ref['ms_per_step']:.0f} ms per step on the sample "
               f"\"{ref['cpu_baseline']['sample']}\".\n")
out.append("## Raw lines\n")
for c in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
    d = load(f"gpurun_out/{prefix}{c}.log")
    if d:
        out.append(f"### {c}\n```json\n{json.dumps(d)}\n```\n")
if ref:
    out.append(f"### --impl reference (cfg5)\n```json\n{json.dumps(ref)}\n```\n")
open("profiles/r2_bench_lines.md", "w").write("\n".join(out))
print("\n".join(out[:12]))

if scale_prefix:
    out = ["# Round 2 — multi-GPU scaling of the headline step (cfg5, strong scaling)\n",
           "`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N --steps 3 --warmup 3`: "
           "16 777 216 rays sharded over N B200s of one box — each rank generates only its own shard on "
           "its GPU (counter-based sampling, shard-invariant) — mesh (99 374 triangles) replicated, one NCCL "
           "all-reduce of the parameter gradients per step; CUDA-event timing bracketed by barriers, max "
           "over ranks. `check` = mean error and gradient checksums of the last timed step: equal between "
           "N = 1 and N > 1 up to the fp32 summation order of the all-reduce.\n",
           "| GPUs | ms/step | tests/s (whole job) | speed-up | efficiency | e2e ms/step | mean error | "
           "grad L1 |",
           "|---:|---:|---:|---:|---:|---:|---:|---:|"]
    base = None
    for n in (1, 2, 4, 8):
        d = load(f"gpurun_out/{scale_prefix}{n}gpu.log")
        if not d:
            continue
        base = base or d["value"] / d["n_gpus"]
        ck = d.get("check") or {}
        out.append(f"| {d['n_gpus']} | {d['ms_per_step']:.1f} | {d['value']:.4g} | "
                   f"{d['value'] / base:.2f} | {100 * d['value'] / base / d['n_gpus']:.1f}% | "
                   f"{d['e2e'].get('ms_per_step', 0):.1f} | {ck.get('mean_error', 0):.9g} | "
                   f"{ck.get('grad_l1', 0):.8g} |")
    out.append("\n## Raw lines\n")
    for n in (1, 2, 4, 8):
        d = load(f"gpurun_out/{scale_prefix}{n}gpu.log")
        if d:
            out.append(f"### {n} GPU(s)\n```json\n{json.dumps(d)}\n```\n")
    open("profiles/r2_scaling.md", "w").write("\n".join(out))
    print("\n".join(out[:10]))
