"""Collect the bench.py JSON lines brought back in gpurun_out/ into profiles/r1_bench_lines.md and
profiles/r1_scaling.md (tables + raw lines)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)


def load(f):
    try:
        return json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        return None


out = ["# Round 1 — bench.py lines (B200, `gpurun`)\n",
       "Raw JSON printed by `python bench.py --config <cfg>` (default steps 3, warm-up 3; cfg5 is the "
       "default config). `value` = ray×boundary tests/s of the step with rays resident in HBM; `e2e` = the "
       "same step from pinned host rays with gradients/loss copied back; `roofline.frac` = 45 FLOP × tests "
       "÷ kernel time ÷ the FP32 FMA peak measured in the same run — the ALGORITHMIC Möller–Trumbore count; the "
       "kernel's projected plane form executes 25 FLOP per test, so the fraction exceeds 1 while the FMA pipe "
       "(14 lane-slots per test, next column) is about 70 % busy.\n",
       "| config | workload | ms/step | tests/s | rays/s | fwd-only tests/s | roofline.frac | FMA-pipe "
       "slots busy | e2e tests/s | CPU baseline tests/s (cores) |",
       "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
for c in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
    d = load(f"gpurun_out/bench_{c}.log")
    if not d:
        continue
    r, cb = d.get("roofline", {}), d.get("cpu_baseline", {})
    if r.get("frac"):
        frac = f"{r['frac']:.3f}"
    else:
        frac = (f"latency: {r.get('us_per_trace', 0):.0f} µs/trace via the API, "
                f"{r.get('us_per_trace_kernel_only', 0):.1f} µs kernel only")
    util = f"{r['fma_pipe_utilisation']:.3f}" if r.get("fma_pipe_utilisation") else "—"
    kind = "fwd+bwd" if "fwd_bwd" in d["metric"] else "fwd"
    out.append(f"| {c} | {d['config']['rays']} rays × {d['config']['triangles']} × "
               f"{d['config']['iterations']} it, {kind} | {d['ms_per_step']:.3f} | {d['value']:.4g} | "
               f"{d['rays_per_s']:.4g} | {d['forward_only']['tests_per_s']:.4g} | {frac} | {util} | "
               f"{d['e2e']['value']:.4g} | {cb.get('value', 0):.3g} ({cb.get('cores')}) |")
out.append("")
ref = load("gpurun_out/bench_ref.log")
if ref:
    out.append(f"`bench.py --impl reference` (cfg5, CPU restatement, {ref['cpu_baseline']['cores']} host "
               f"cores): {ref['value']:.4g} tests/s, {ref['ms_per_step']:.0f} ms per step on the sample "
               f"\"{ref['cpu_baseline']['sample']}\".\n")
out.append("## Raw lines\n")
for c in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
    d = load(f"gpurun_out/bench_{c}.log")
    if d:
        out.append(f"### {c}\n```json\n{json.dumps(d)}\n```\n")
if ref:
    out.append(f"### --impl reference (cfg5)\n```json\n{json.dumps(ref)}\n```\n")
open("profiles/r1_bench_lines.md", "w").write("\n".join(out))

out = ["# Round 1 — multi-GPU scaling of the headline step (cfg5, strong scaling)\n",
       "`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N --steps 3 --warmup 3`: "
       "16 777 216 rays sharded over N B200s of one box, mesh (99 374 triangles) replicated, one NCCL "
       "all-reduce of the parameter gradients per step; CUDA-event timing bracketed by barriers, max over "
       "ranks.\n",
       "| GPUs | ms/step | tests/s (whole job) | speed-up | efficiency | e2e tests/s | roofline.frac (rank 0) |",
       "|---:|---:|---:|---:|---:|---:|---:|"]
base = None
files = ((1, "bench_cfg5"), (2, "bench_cfg5_2gpu"), (4, "bench_cfg5_4gpu"), (8, "bench_cfg5_8gpu"))
for n, f in files:
    d = load(f"gpurun_out/{f}.log")
    if not d:
        continue
    base = base or d["value"]
    out.append(f"| {d['n_gpus']} | {d['ms_per_step']:.1f} | {d['value']:.4g} | {d['value'] / base:.2f} | "
               f"{100 * d['value'] / base / d['n_gpus']:.1f}% | {d['e2e']['value']:.4g} | "
               f"{d['roofline']['frac']:.3f} |")
out.append("\n## Raw lines\n")
for n, f in files[1:]:
    d = load(f"gpurun_out/{f}.log")
    if d:
        out.append(f"### {n} GPUs\n```json\n{json.dumps(d)}\n```\n")
open("profiles/r1_scaling.md", "w").write("\n".join(out))
print(open("profiles/r1_bench_lines.md").read()[:1900])
