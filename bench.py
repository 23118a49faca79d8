#!/usr/bin/env python
"""
bench.py -- the headline benchmark of the tfrt ray-trace hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one fwd+bwd lens-optimisation step (SGD_Optimizer.process_gradient: system update,
OpticalEngine.ray_trace, imaging loss, back-propagation to the surface parameters, and for N > 1
the NCCL all-reduce of the parameter gradients) on BASELINE.json's headline configuration:
16 777 216 rays x 99 374 triangles x 3 trace iterations (config 5), synthetic data.

Metric: ray-triangle tests per second over the whole job, tests = sum_k live_rays_k x triangles
(SURVEY.md 8(d)); rays/s, forward-only numbers, the roofline of the dominant kernel, the CPU
baseline and the end-to-end (host buffers) number travel in the same JSON line.

--impl reference times the reference's own CPU algorithm (the float64 dense restatement in
oracle/, the only form of the reference that can run here: TensorFlow is not installable) on a
bounded ray sample of the same workload, on the box's host cores.
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

FLOP_PER_TEST = 45.0        # algorithmic Moller-Trumbore count
BWD_BYTES_PER_RAY_ITER = 100.0   # algorithmic bytes of the backward pass
METRIC = "ray_triangle_tests_per_s_fwd_bwd"
UNIT = "tests/s"

CONFIGS = {
    # name: (scene factory kwargs, has backward)
    "cfg1": dict(factory="config1", kwargs={}, backward=False),   # 2D, launch-latency bound
    "cfg2": dict(factory="config2", kwargs={}, backward=False),
    "cfg3": dict(factory="config3", kwargs={}, backward=True),
    "cfg4": dict(factory="config4", kwargs={}, backward=False),
    "cfg5": dict(factory="config5", kwargs={}, backward=True),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg5", choices=sorted(CONFIGS))
    ap.add_argument("--rays", type=int, default=0, help="override the ray count (debugging)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the named config's rays are sharded over the GPUs (BASELINE "
                         "config 5); weak: every GPU traces the full ray count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-accel", action="store_true", help="skip the acceleration-grid leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0,
                    help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


DEVICE_SOURCE_CONFIGS = ("cfg3", "cfg5")     # rays born on the GPU at every step (Philox)


def make_scene(args, ray_count=None, with_rays=True):
    """The config's scene.  with_rays=False (lens-optimisation configs): meshes only -- the rays
    are generated on the device (our arm) or, for a CPU sample, by `cpu_sample_rays`."""
    from tensorflowraytrace_b200 import scenes
    cfg = CONFIGS[args.config]
    kwargs = dict(cfg["kwargs"])
    if ray_count:
        kwargs["ray_count"] = int(ray_count)
    elif args.rays:
        kwargs["ray_count"] = int(args.rays)
    if args.config in DEVICE_SOURCE_CONFIGS and not with_rays:
        kwargs["with_rays"] = False
    scene = getattr(scenes, cfg["factory"])(**kwargs)
    scene.setdefault("ray_count", int(scene["sources"]["x_start"].shape[0]) if scene["sources"] else 0)
    return scene, cfg


def cpu_sample_rays(scene, n, epoch=2):
    """The first n rays of the device-born source of a lens-optimisation scene
    (scenes.lens_optimize_device_source), restated in NumPy from oracle/sampling.py (same Philox
    counters, same transforms): the CPU arm traces a sample of exactly the rays the GPU arm
    generates.  Returns a copy of the scene whose `sources` hold those n rays."""
    from oracle import sampling
    from tensorflowraytrace_b200 import scenes
    L = scenes.LENS_SOURCE
    seed = scene.get("seed", 1234)
    base, rank = sampling.square(seed, 0, epoch, 0, n, L["object_size"], L["object_size"])
    dirs, _ = sampling.sphere(seed, 1, epoch, 0, 0, n, scene["ray_count"], L["cone"], 1.0)
    starts = np.stack([np.full(n, -L["source_distance"]), base[:, 0].astype(np.float64),
                       base[:, 1].astype(np.float64)], axis=1)
    out = dict(scene)
    out["sources"] = scenes.make_rays(starts, dirs.astype(np.float64), L["wavelength"],
                                      extra={"rank": rank.astype(np.float64)})
    return out


def scene_ray_count(scene):
    return int(scene.get("ray_count") or scene["sources"]["x_start"].shape[0])


# ------------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="tfrt_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu_index), "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)),
                       power_w_max=float(max(power)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle): used for cpu_baseline and for --impl reference

def oracle_step(scene, n_rays, backward, threads):
    """One fwd(+bwd) step of the float64 dense restatement on the first n_rays rays of the scene.
    Returns (seconds, tests, live_counts)."""
    import torch
    import helpers
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    if not scene["sources"]:
        scene = cpu_sample_rays(scene, n_rays)
    if scene["dimension"] == 2:
        sub = dict(scene)
        sub["sources"] = {k: v[:n_rays] for k, v in scene["sources"].items()}
        eng, _ = helpers.oracle2d_from_scene(sub, compile_active_rays=False)
        eng.ray_trace(scene["max_iterations"])
        dt = time.perf_counter() - t0
        return dt, float(sum(eng.live_counts)) * scene_triangles(scene), eng.live_counts
    eng, params = helpers.oracle_from_scene(scene, ray_slice=slice(0, n_rays),
                                            requires_grad=backward, compile_active_rays=False)
    # cache-sized (triangles x ray-chunk) temporaries: 4x faster than the 256 MB default on the
    # build container, i.e. the most favourable setting for the CPU arm
    eng.optical_system.chunk_bytes = 4 << 20
    eng.ray_trace(scene["max_iterations"])
    if backward:
        fin = eng.finished_rays
        goal = scene["goal_scale"] * fin["rank"]
        err = (fin["y_end"] - goal[:, 0]) ** 2 + (fin["z_end"] - goal[:, 1]) ** 2
        torch.autograd.grad(err.sum(), params)
    dt = time.perf_counter() - t0
    return dt, float(eng.optical_system.tests_done), eng.live_counts


def size_cpu_sample(scene, backward, threads, target_s):
    """Pick a ray count whose oracle step takes about target_s seconds (probe, then scale)."""
    probe = 256
    dt, tests, _ = oracle_step(scene, probe, backward, threads)
    n_total = scene_ray_count(scene)
    n = int(probe * max(target_s / max(dt, 1e-3), 1.0))
    n = max(probe, min(n, 65536, n_total))
    return n


def reference_source_leg(scene, backward, threads, target_s=8.0, mem_bytes=6e9):
    """The reference's OWN Python (`tfrt`, pip-installed under baseline/_ref) executed on the stand-in
    tensorflow of oracle/tf_shim (torch float64 CPU ops): timed beside the oracle port.  The sample is
    bounded by memory: the reference materialises ~150 (triangles x rays) float64 tensors per
    iteration, all of them retained by the GradientTape.  Returns a dict or None (3D configs only;
    absent install -> None)."""
    if scene["dimension"] != 3:
        return None
    try:
        import torch
        from oracle import reference_runner as rr
        if rr.find_reference(installed_only=True) is None:
            return None
        torch.set_num_threads(threads)
        ref = rr.import_reference(installed_only=True)
        T = scene_triangles(scene)
        per_ray = T * 8.0 * 150 * (scene["max_iterations"] if backward else 1)
        n = int(max(8, min(mem_bytes / per_ray, 4096, scene_ray_count(scene))))
        if not scene["sources"]:
            scene = cpu_sample_rays(scene, n)
        dt, tests, live = rr.timed_reference_step(ref, scene, n, backward)        # warm-up / probe
        reps = int(max(1, min(5, target_s / max(dt, 1e-3))))
        total_s = total_tests = 0.0
        for _ in range(reps):
            dt, tests, live = rr.timed_reference_step(ref, scene, n, backward)
            total_s += dt
            total_tests += tests
        return {"value": total_tests / total_s, "unit": UNIT, "cores": threads,
                "kind": "reference source on a stand-in tensorflow",
                "sample": f"first {n} of {scene_ray_count(scene)} rays x {T} triangles x "
                          f"{scene['max_iterations']} iterations, {'fwd+bwd (GradientTape)' if backward else 'fwd'}, "
                          f"{reps} step(s) of {total_s / reps:.2f} s; unmodified tfrt.engine / geometry / "
                          f"operation / boundaries from baseline/_ref, every tf op served by torch float64 "
                          f"CPU (oracle/tf_shim), un-chunked dense (triangles x rays) tensors as the "
                          f"reference builds them"}
    except Exception as e:  # the leg is informative: never fail the bench line over it
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args):
    """--impl reference: the CPU restatement timed on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    scene, cfg = make_scene(args, with_rays=False)
    backward = cfg["backward"]
    T = scene_triangles(scene)
    n = size_cpu_sample(scene, backward, threads, target_s=max(2.0, 90.0 / max(args.steps + args.warmup, 1)))
    for _ in range(args.warmup):
        oracle_step(scene, n, backward, threads)
    total_s, total_tests, live = 0.0, 0.0, None
    for _ in range(args.steps):
        dt, tests, live = oracle_step(scene, n, backward, threads)
        total_s += dt
        total_tests += tests
    value = total_tests / total_s
    sample = (f"first {n} of {scene_ray_count(scene)} rays x {T} triangles x "
              f"{scene['max_iterations']} iterations, {'fwd+bwd' if backward else 'fwd'}, "
              f"float64 dense (triangles x ray-chunk) restatement, torch CPU")
    line = {
        "impl": "reference", "metric": METRIC if backward else "ray_triangle_tests_per_s_fwd",
        "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(args.steps, 1),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": bench_config(args, scene, 1),
        "rays_per_s": n / (total_s / max(args.steps, 1)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "TensorFlow is not installable in this image, so the reference cannot run as "
                "shipped; `value` is its algorithm restated op-for-op (oracle/, cache-chunked: the "
                "fastest CPU form we found); `reference_source` times the reference's own Python "
                "on a stand-in tensorflow when baseline/_ref holds the pip-installed package.",
    }
    src_leg = reference_source_leg(scene, backward, threads)
    if src_leg is not None:
        line["cpu_baseline"]["reference_source"] = src_leg
    print(json.dumps(line), flush=True)


def scene_triangles(scene):
    """Boundary count of the merged set: triangles (3D) or segments + arcs (2D)."""
    from tensorflowraytrace_b200 import scenes
    if scene["dimension"] == 2:
        n = 0
        for kind in ("optical", "stop", "target"):
            n += sum(np.asarray(e["segments"]).reshape(-1, 4).shape[0] for e in scene[f"{kind}_segments"])
            n += sum(np.asarray(e["arcs"]).reshape(-1, 5).shape[0] for e in scene[f"{kind}_arcs"])
        return n
    return scenes.triangle_count(scene)


def l2_note(scene, world):
    n = scene_ray_count(scene) // max(world, 1)
    work = n * (28 + 24 + 48 + 8 + 49 * scene["max_iterations"])
    if work < 2 * (126 << 20):
        return (f"per-step working set {work / 2**20:.0f} MiB fits the 126 MB L2: L2 flushed "
                "(256 MB write) between timed steps, each step timed on its own")
    return (f"per-step working set {work / 2**20:.0f} MiB per GPU (rays + state + tape) exceeds "
            "the 126 MB L2; no flush needed")


def bench_config(args, scene, world):
    n = scene_ray_count(scene)
    return {
        "workload": f"{args.config}: {scene['name']} -- {n} rays x {scene_triangles(scene)} "
                    f"{'segments+arcs' if scene['dimension'] == 2 else 'triangles'} x "
                    f"{scene['max_iterations']} trace iterations, "
                    f"{'fwd+bwd optimisation step' if CONFIGS[args.config]['backward'] else 'forward trace'}",
        "rays": int(n), "triangles": int(scene_triangles(scene)),
        "iterations": int(scene["max_iterations"]),
        "parallelism": f"rays sharded over {world} GPU(s), mesh replicated, "
                       f"NCCL all-reduce of parameter gradients" if world > 1 else "single GPU",
        "l2": l2_note(scene, world),
    }


# ------------------------------------------------------------------------------------------------
# our arm

def run_ours(args):
    import torch
    from tensorflowraytrace_b200 import kernels, optimizer, parallel, scenes

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    rank, world, local = parallel.init_from_env()
    dev = torch.device("cuda", torch.cuda.current_device())
    device_rays = args.config in DEVICE_SOURCE_CONFIGS
    scene, cfg = make_scene(args, with_rays=not device_rays)
    backward = cfg["backward"]
    n_total = scene_ray_count(scene)
    T = scene_triangles(scene)
    K = scene["max_iterations"]
    shard = (rank, world) if (args.scaling == "strong" and world > 1) else None
    build = scenes.build_engine_2d if scene["dimension"] == 2 else scenes.build_engine

    # ---- the engine of the device-resident legs
    if device_rays:
        # rays born on this rank's GPU at every system.update(), as the reference's optimiser
        # regenerates them every step (tfrt/optimizer.py:217): counter-based sampling, so the
        # ranks of a sharded job together trace exactly the rays one rank would
        source = scenes.lens_optimize_device_source(n_total, scene["seed"])
        engine, params, src = build(scene, source=source, ray_shard=shard,
                                    compile_active_rays=False, compile_finished_rays=True)
        host_fields = {k: v.detach() for k, v in src.items()}
    else:
        if shard is not None:
            lo, hi = parallel.shard_range(n_total, rank, world)
        else:
            lo, hi = 0, n_total
        scene_local = dict(scene)
        scene_local["sources"] = {k: v[lo:hi] for k, v in scene["sources"].items()}
        engine, params, src = build(scene_local, compile_active_rays=False,
                                    compile_finished_rays=True)
        host_fields = {k: torch.from_numpy(np.ascontiguousarray(v).astype(np.float32))
                       for k, v in scene_local["sources"].items()}
    n_local = int(next(iter(host_fields.values())).shape[0])

    # ---- this rank's rays as pinned HOST buffers, and a second engine fed from them: the e2e leg
    # copies them to the device every step through the public ManualSource API
    host = {k: t.to("cpu", torch.float32).contiguous().pin_memory() for k, t in host_fields.items()}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())
    if device_rays:
        from tensorflowraytrace_b200 import sources as tf_sources
        e2e_src = tf_sources.ManualSource(3).feed({k: t.to(dev) for k, t in host.items()})
        e2e_engine, e2e_params, _ = build(scene, source=e2e_src, compile_active_rays=False,
                                          compile_finished_rays=True)
    else:
        e2e_engine, e2e_params, e2e_src = engine, params, src
    errf = scenes.imaging_error(scene) if backward else None
    opt = optimizer.SGD_Optimizer(engine, params, errf, K) if backward else None
    e2e_opt = opt if e2e_engine is engine else (
        optimizer.SGD_Optimizer(e2e_engine, e2e_params, errf, K) if backward else None)
    acc = [None] * len(params)
    d2h_bytes = sum(p.numel() * 4 for p in params) + 4 if backward else 4
    tally = {"tests": 0.0, "steps": 0, "check": None, "diag": {"revalidated": 0, "gross": 0}}

    def account(eng, grads=None, err=None):
        """Per-step bookkeeping from numbers the step has already brought to the host."""
        tr = eng.last_trace
        live = tr.live_counts()
        tally["tests"] += float(sum(live)) * T
        tally["steps"] += 1
        tally["live"] = live
        if hasattr(tr, "diagnostics"):
            d = tr.diagnostics
            tally["diag"]["revalidated"] += d["revalidated"]
            tally["diag"]["gross"] += d["gross"]
        if grads is not None:
            tally["check"] = (grads, err)

    def step_resident():
        """Inputs already in HBM: update (rays regenerated on the device where the config says
        so) + trace (+ loss + backward + all-reduce)."""
        if backward:
            grads, err = opt.process_gradient(acc)
            account(engine, grads, err)
            return grads, err
        engine.optical_system.update()
        engine.ray_trace(K)
        engine.finished_rays              # materialise: the trace's result
        if scene["dimension"] == 3:
            account(engine)
        return None, None

    def step_e2e():
        """Public API from HOST buffers: H2D of this step's rays, step, D2H of loss + grads."""
        for k, t in host.items():
            e2e_src[k] = t.to(dev, non_blocking=True)
        if backward:
            grads, err = e2e_opt.process_gradient(acc)
            return [g.cpu() for g in grads] + [err.cpu()]
        e2e_engine.optical_system.update()
        e2e_engine.ray_trace(K)
        return [e2e_engine.finished_rays["x_end"][:1].cpu()]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # Working set of one step (rays + working state + tape).  Above the 126 MB L2 nothing of a
    # step survives to the next one; below it the L2 is flushed between timed steps by writing
    # a 256 MB buffer, and every step is timed on its own so the flush stays outside.
    work_bytes = n_local * (28 + 24 + 48 + 8 + 49 * K)
    flush_l2 = work_bytes < 2 * (126 << 20)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush_l2 else None

    def timed(fn, steps):
        if not flush_l2:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            return parallel.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
        total = 0.0
        for _ in range(steps):
            flush_buf.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            fn()
            e1.record()
            barrier()
            total += e0.elapsed_time(e1) * 1e-3
        return parallel.max_over_ranks(total)

    fp32_peak = kernels.fp32_peak(packed=0, iters=4096)
    fp32_peak_packed = kernels.fp32_peak(packed=1, iters=4096)

    # ---- warm-up, then the device-resident timed region
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    if scene["dimension"] == 2:
        live = engine.last_trace.live_counts()
        tally["live"] = live
    tally.update(tests=0.0, steps=0)
    tally["diag"] = {"revalidated": 0, "gross": 0}

    # clocks / throttle reasons are sampled on rank 0's GPU only: one nvidia-smi poller per rank
    # (NVML takes driver locks) can cost another rank a timed step on an 8-GPU box
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    kernels.profile_enable(True)
    kernels.profile_read()
    kernels.launch_count(reset=True)
    seconds = timed(step_resident, args.steps)
    launches = kernels.launch_count(reset=True)
    k_ms, k_n = kernels.profile_read()
    kernels.profile_enable(False)
    clocks = sampler.stop() if sampler is not None else {}
    live = tally["live"]
    if scene["dimension"] == 2:
        tests_local = float(sum(live)) * T * args.steps
    else:
        tests_local = tally["tests"]                 # summed over the timed steps
    tests_total = parallel.sum_over_ranks(tests_local) / args.steps     # per step, whole job
    rays_total = parallel.sum_over_ranks(n_local)
    diag = {k: int(parallel.sum_over_ranks(v)) for k, v in tally["diag"].items()}
    check = None
    if tally["check"] is not None:
        # loss and gradient checksums of the LAST timed step (already all-reduced: identical on
        # every rank).  The rays are a function of (seed, update number) only, so these agree
        # between N = 1 and N = 8 up to the fp32 summation order of the all-reduce.
        grads, err = tally["check"]
        g64 = [g.detach().double() for g in grads]
        check = {"mean_error": float(err),
                 "grad_sum": float(sum(g.sum() for g in g64)),
                 "grad_l1": float(sum(g.abs().sum() for g in g64)),
                 "update_number": int(max(args.warmup, 3) + args.steps)}

    # forward-only (update + trace), same inputs
    def step_forward():
        engine.optical_system.update()
        engine.ray_trace(K)
        engine.finished_rays
    step_forward()
    fwd_seconds = timed(step_forward, args.steps)

    value = tests_total * args.steps / seconds
    line = {
        "metric": METRIC if backward else "ray_triangle_tests_per_s_fwd",
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * seconds / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": bench_config(args, scene, world),
        "rays_per_s": rays_total * args.steps / seconds,
        "tests_per_step": tests_total, "live_rays_per_iteration_rank0": live,
        "forward_only": {"tests_per_s": tests_total * args.steps / fwd_seconds,
                         "rays_per_s": rays_total * args.steps / fwd_seconds,
                         "ms_per_step": 1e3 * fwd_seconds / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # fp64 re-validation of the fp32 search's winners over the timed steps, all ranks:
        # `revalidated` = proposals the exact rule rejected (searched again exactly; rounding at
        # triangle edges), `gross` = rejections beyond rounding (a search defect; must be 0)
        "revalidation": diag,
    }
    if check is not None:
        line["check"] = check
    line["config"]["rays_origin"] = (
        "regenerated on the device at every step by counter-based sampling kernels "
        "(Philox4x32-10 keyed by seed and global ray index; each rank generates only its shard)"
        if device_rays else "fixed NumPy scene, resident on the device")
    # ---- roofline of the dominant kernel (k_intersect3d), rank 0's launches
    if k_n > 0 and k_ms > 0:
        tests_per_launch = tests_local / k_n
        achieved = FLOP_PER_TEST * tests_per_launch / (k_ms / k_n * 1e-3) / 1e12
        launch_s = k_ms / k_n * 1e-3
        nominal = 148 * 128 * 2 * 1.965e9
        line["roofline"] = {
            "kernel": "k_intersect3d", "bound": "fp32_fma",
            # what the hardware does, first: the projected plane form EXECUTES 25 FLOP per test
            # (10 FFMA + 1 FMUL + 3 FADD + 1 reciprocal = 14 FMA-pipe lane-slots, issued as 7
            # packed f32x2 instructions, each testing one ray against two triangles)
            "executed_flop_per_test": 25.0,
            "frac_executed": 25.0 * tests_per_launch / launch_s / fp32_peak,
            "frac_executed_of_nominal": 25.0 * tests_per_launch / launch_s / nominal,
            "fma_pipe_slots_per_test": 14.0,
            "fma_pipe_utilisation": 14.0 * tests_per_launch / launch_s / (fp32_peak / 2.0),
            # the contract's figure: ALGORITHMIC count (45 FLOP per Moller-Trumbore test,
            # SURVEY.md 8(d)) over the measured launch time; above 1 because the kernel needs
            # fewer operations per test than the algorithmic count
            "achieved": achieved, "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
            "frac": achieved / (fp32_peak / 1e12),
            "frac_of_nominal": achieved / (nominal / 1e12),
            "flop_per_test": FLOP_PER_TEST,
            "traffic": intersect_traffic(n_local),
            "traffic_source": "ncu dram__bytes_read+write of one launch "
                              "(profiles/r2_intersect3d_traffic.json, else round 1's) scaled per "
                              "ray to this launch; DRAM is < 0.1 % of peak, the kernel is bound "
                              "by FP32 issue / register-operand bandwidth",
            "launches": int(k_n), "avg_launch_ms": k_ms / k_n,
            "share_of_step": k_ms * 1e-3 / seconds,
            "peak_source": "measured in this run by tfrt_fp32_peak (register-only FFMA loop on "
                           "all SMs); MEASURED_PEAKS.json carries no fp32 figure. Nominal "
                           "148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s",
            "peak_ffma2_tflops": fp32_peak_packed / 1e12,
        }
    elif scene["dimension"] == 2:
        # the 2D path is one launch per trace: launch-latency bound, no roofline fraction
        tr2 = engine.last_trace
        reps = 50
        tr2.forward()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            tr2.forward()
        e1.record()
        torch.cuda.synchronize()
        kernel_us = 1e3 * e0.elapsed_time(e1) / reps
        # OpticalEngine.ray_trace alone (system already updated, histories not read): the call
        # enqueues the launch and returns; wall clock including the drain of the GPU
        import time as _time
        for _ in range(20):
            engine.ray_trace(K)
        torch.cuda.synchronize()
        t0 = _time.perf_counter()
        for _ in range(200):
            engine.ray_trace(K)
        torch.cuda.synchronize()
        call_us = 1e6 * (_time.perf_counter() - t0) / 200
        line["roofline"] = {
            "kernel": "k_trace2d_forward", "bound": "latency", "achieved": None, "peak": None,
            "unit": "us/trace", "frac": None, "traffic": None,
            "us_per_trace": 1e6 * fwd_seconds / args.steps,
            "us_per_ray_trace_call": call_us,
            "us_per_trace_kernel_only": kernel_us,
            "note": "config 1 (1 000 rays x 3 segments x 3 iterations) is launch-latency bound "
                    "(SURVEY.md 8(d)): one kernel launch per trace, fp64 arithmetic; "
                    "us_per_trace is system.update + ray_trace + reading finished_rays through "
                    "the public API (L2 flushed before each), us_per_ray_trace_call the "
                    "OpticalEngine.ray_trace call alone on an updated system (200 back-to-back "
                    "calls, wall clock with the GPU drained; histories are built when first "
                    "read), us_per_trace_kernel_only the C-ABI launch alone (status memset + "
                    "k_trace2d_forward)"}
    if scene["dimension"] == 2 and rank == 0 and world == 1:
        line["two_d_at_scale"] = two_d_at_scale(torch, scenes, not args.no_cpu_baseline)
    if backward:
        # the backward leg (loss + K2 + K3 backward + all-reduce) timed on its own with CUDA
        # events around autograd, forward outside the timed span
        bwd_ms = []
        for _ in range(max(args.steps, 3)):
            engine.optical_system.update()
            engine.ray_trace(K)
            total = errf(engine).sum()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.autograd.grad(total, params, allow_unused=True)
            e1.record()
            torch.cuda.synchronize()
            bwd_ms.append(e0.elapsed_time(e1))
        bwd_s = parallel.max_over_ranks(float(np.median(bwd_ms)) * 1e-3)
        hbm = measured_peaks().get("hbm_gbs")
        gbs = BWD_BYTES_PER_RAY_ITER * float(sum(live)) / bwd_s / 1e9
        line["roofline_backward"] = {
            "kernel": "k_backward3d (+ faces backward, parameter projection)", "bound": "hbm",
            "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": (gbs / hbm) if hbm else None,
            "ms_per_step": 1e3 * bwd_s, "share_of_step": bwd_s * args.steps / seconds,
            "bytes_per_ray_iteration": BWD_BYTES_PER_RAY_ITER,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm else "absent"}

    # ---- end to end through the public API with host buffers
    if not args.no_e2e:
        for _ in range(max(args.warmup, 3)):     # own warm-up: fresh device buffers every step
            step_e2e()                           # (the caching allocator settles after two)
        e2e_seconds = timed(step_e2e, args.steps)
        if scene["dimension"] == 3:     # the fixed host rays' own live counts
            e2e_tests = parallel.sum_over_ranks(
                float(sum(e2e_engine.last_trace.live_counts())) * T)
        else:
            e2e_tests = tests_total
        line["e2e"] = {"value": e2e_tests * args.steps / e2e_seconds, "unit": UNIT,
                       "h2d_bytes_per_step": int(parallel.sum_over_ranks(h2d_bytes)),
                       "d2h_bytes_per_step": int(d2h_bytes * world),
                       "ms_per_step": 1e3 * e2e_seconds / args.steps,
                       "rays_per_s": rays_total * args.steps / e2e_seconds}

    # ---- optional acceleration mode, reported on its own: the same step
    # with the search walking a device-built uniform grid instead of testing every triangle.  Its
    # time is NEVER part of the tests/s metric above (that is the brute-force search).
    if scene["dimension"] == 3 and not args.no_accel:
        system = engine.optical_system
        system.accel = "grid"
        try:
            for _ in range(3):
                step_resident()
            kernels.profile_enable(True)
            kernels.profile_read()
            grid_seconds = timed(step_resident, args.steps)
            g_ms, g_n = kernels.profile_read()
            kernels.profile_enable(False)
            st = system.grid.stats()
            line["accel_grid"] = {
                "ms_per_step": 1e3 * grid_seconds / args.steps,
                "speedup_over_brute_force_step": seconds / grid_seconds,
                "rays_per_s": rays_total * args.steps / grid_seconds,
                "search_kernel_ms_per_launch": (g_ms / g_n) if g_n else None,
                "grid": st,
                "note": "same hit indices and gradients as the brute-force search "
                        "(tests/test_gpu_grid.py); optional mode, outside the tests/s metric"}
        finally:
            system.accel = None
            system.update()

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample, same workload
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = size_cpu_sample(scene, backward, threads, args.cpu_seconds)
        dt, tests, _ = oracle_step(scene, n, backward, threads)
        line["cpu_baseline"] = {
            "value": tests / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "rays_per_s": n / dt,
            "sample": f"first {n} of {n_total} rays x {T} triangles x {K} iterations, "
                      f"{'fwd+bwd' if backward else 'fwd'}, float64 dense restatement of the "
                      f"reference (oracle/), torch CPU, {dt:.1f} s"}
        src_leg = reference_source_leg(scene, backward, threads, target_s=4.0)
        if src_leg is not None:
            line["cpu_baseline"]["reference_source"] = src_leg
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def two_d_at_scale(torch, scenes, with_cpu):
    """the four sizes of the reference's own scaling script
    (dev/performance.py:77-86: 10 / 10 000 beam points x 6 wavelengths, 10 / 1 000 segments, 2
    iterations), through the public API (update + ray_trace + finished rays) and as the bare
    C-ABI launch, beside the CPU port of the reference's algorithm (oracle/engine2d.py)."""
    rows = []
    for beams, points in ((10, 11), (10000, 11), (10, 1001), (10000, 1001)):
        scene = scenes.performance_2d(beams, points)
        engine, _, _ = scenes.build_engine_2d(scene, compile_active_rays=False)

        def api():
            engine.optical_system.update()
            engine.ray_trace(2)
            return engine.finished_rays["x_end"]

        for _ in range(3):
            api()
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            api()
        torch.cuda.synchronize()
        api_s = (time.perf_counter() - t0) / reps
        tr = engine.last_trace
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tr.forward()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            tr.forward()
        e1.record()
        torch.cuda.synchronize()
        row = {"rays": beams * 6, "segments": points - 1 + 1, "iterations": 2,
               "us_per_trace_public_api": 1e6 * api_s,
               "us_per_trace_kernel_only": 1e3 * e0.elapsed_time(e1) / reps}
        if with_cpu and beams * (points - 1) <= 2_500_000:
            import helpers
            t0 = time.perf_counter()
            eng, _ = helpers.oracle2d_from_scene(scene, compile_active_rays=False)
            eng.ray_trace(2)
            row["cpu_port_us_per_trace"] = 1e6 * (time.perf_counter() - t0)
        rows.append(row)
    return rows


def intersect_traffic(n_rays):
    """DRAM bytes per k_intersect3d launch over n_rays rays, from the committed ncu capture."""
    try:
        for name in ("r2_intersect3d_traffic.json", "r1_intersect3d_traffic.json"):
            path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(path):
                t = json.load(open(path))
                return float(t["dram_bytes_per_ray_per_launch"]) * float(n_rays)
    except Exception:
        pass
    return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
