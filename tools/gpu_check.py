"""First-contact GPU check: fp32 peak, cfg2 parity vs the oracle on a ray subset, raw timings.
Run on the B200 box:  python tools/gpu_check.py [--big]"""

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tensorflowraytrace_b200 import kernels, materials, scenes  # noqa: E402
import helpers  # noqa: E402


def device_inputs(scene, dev):
    tri_list, norm_list, cat, mat_in, mat_out = [], [], [], [], []
    for s in scene["surfaces"]:
        zp = torch.tensor(s["zero_points"], dtype=torch.float32, device=dev)
        vec = torch.tensor(s["vectors"], dtype=torch.float32, device=dev)
        par = torch.tensor(s["parameters"], dtype=torch.float32, device=dev)
        verts = zp + par[:, None] * vec
        faces = torch.tensor(s["faces"], dtype=torch.int32, device=dev)
        tri, norm = kernels.faces_forward(verts, faces)
        tri_list.append(tri); norm_list.append(norm)
        F = faces.shape[0]
        cat.append(torch.zeros(F, dtype=torch.int32, device=dev))
        mat_in.append(torch.full((F,), s["mat_in"], dtype=torch.int32, device=dev))
        mat_out.append(torch.full((F,), s["mat_out"], dtype=torch.int32, device=dev))
    n_optical = sum(t.shape[1] for t in tri_list)
    for code, lst in ((1, scene["stops"]), (2, scene["targets"])):
        for m in lst:
            verts = torch.tensor(m["vertices"], dtype=torch.float32, device=dev)
            faces = torch.tensor(m["faces"], dtype=torch.int32, device=dev)
            tri, norm = kernels.faces_forward(verts, faces)
            tri_list.append(tri); norm_list.append(norm)
            cat.append(torch.full((faces.shape[0],), code, dtype=torch.int32, device=dev))
    src = torch.stack([torch.tensor(scene["sources"][k], dtype=torch.float32, device=dev)
                       for k in helpers.RAY_FIELDS + ("wavelength",)])
    mats = kernels.material_table([materials.BY_NAME[n].descriptor for n in scene["materials"]], dev)
    return dict(src=src, tri=torch.cat(tri_list, 1).contiguous(), norm=torch.cat(norm_list, 0).contiguous(),
                category=torch.cat(cat), n_optical=n_optical, mat_in=torch.cat(mat_in),
                mat_out=torch.cat(mat_out), materials=mats, n_materials=len(scene["materials"]))


def make_trace(scene, dev, eps=helpers.EPS):
    d = device_inputs(scene, dev)
    return kernels.Trace3D(d["src"], d["tri"], d["norm"], d["category"], d["n_optical"],
                           scene["max_iterations"], mat_in=d["mat_in"], mat_out=d["mat_out"],
                           materials=d["materials"], n_materials=d["n_materials"],
                           intersect_eps=eps["intersect_epsilion"], size_eps=eps["size_epsilion"],
                           ray_start_eps=eps["ray_start_epsilion"])


def timed(fn, warmup=2, reps=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), ts


def main():
    dev = torch.device("cuda:0")
    print("device:", torch.cuda.get_device_name(0))
    for mode in (0, 1):
        print(f"fp32 peak packed={mode}: {kernels.fp32_peak(mode, 8192) / 1e12:.2f} TFLOP/s")

    # ---- parity on a cfg2 subset
    scene = scenes.config2(ray_count=16384)
    tr = make_trace(scene, dev).forward()
    torch.cuda.synchronize()
    print("counters:\n", tr.counters_host)
    eng, _ = helpers.oracle_from_scene(scene)
    t0 = time.time(); eng.ray_trace(scene["max_iterations"]); print("oracle s:", time.time() - t0)
    print("oracle live counts:", eng.live_counts)
    fin, slot, it = tr.history("finished")
    ofin = eng.finished_rays
    print("finished: kernel", fin.shape[1], "oracle", ofin["x_start"].shape[0])
    if fin.shape[1] == ofin["x_start"].shape[0]:
        for i, k in enumerate(helpers.RAY_FIELDS):
            print(f"  {k}: rel err {helpers.rel_err(fin[i].cpu().numpy(), ofin[k].numpy()):.3e}")
    # per-iteration hit index parity
    for k in range(scene["max_iterations"]):
        n = tr.live_counts()[k]
        slots = tr.live[k, :n].long()
        ktri = tr.tape_tri[k][slots].cpu().numpy()
        ores = eng.last_intersect_result[k]
        otri = np.where(ores["valid"].numpy(), ores["gather_trig"].numpy(), -1)
        if len(otri) == len(ktri):
            print(f"  iter {k}: live {n} hit-index mismatches {(otri != ktri).sum()}")
        else:
            print(f"  iter {k}: live kernel {n} oracle {len(otri)}")

    # ---- timings
    for name, sc in (("cfg2", scenes.config2()), ("cfg3", scenes.config3())):
        tr = make_trace(sc, dev)
        best, ts = timed(lambda: tr.forward())
        live = tr.live_counts()
        tests = sum(live) * tr.T
        print(f"{name}: N={tr.N} T={tr.T} live={live} fwd {best*1e3:.3f} ms  "
              f"{tests/best/1e12:.3f} Ttests/s  {45*tests/best/1e12:.1f} TFLOP/s-equiv  all={ts}")
    if "--big" in sys.argv:
        sc = scenes.config5(ray_count=4 * 1048576)
        tr = make_trace(sc, dev)
        best, ts = timed(lambda: tr.forward(), warmup=1, reps=2)
        live = tr.live_counts(); tests = sum(live) * tr.T
        print(f"cfg5/4: N={tr.N} T={tr.T} live={live} fwd {best*1e3:.3f} ms  "
              f"{tests/best/1e12:.3f} Ttests/s  {45*tests/best/1e12:.1f} TFLOP/s-equiv")


if __name__ == "__main__":
    main()
