"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> <out.md> [title]
    python tools/summarize_ncu.py full <report.ncu-rep> <out.md> [title]
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


def launches(path, out, title):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    total = sum(t for _, t, _, _ in rows)
    agg = OrderedDict()
    for name, t, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += t
    with open(out, "w") as f:
        f.write(f"# {title}\n\n")
        f.write(f"Source: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
                f"({len(rows)} launches, {total/1e6:.3f} ms of kernel time; per-launch times are "
                f"cold-cache and serialised, compare SHARES not absolutes).\n\n")
        f.write("| kernel | launches | total ms | share | example grid | block |\n|---|---:|---:|---:|---|---|\n")
        for name, (n, t, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {t/1e6:.3f} | {100*t/total:.2f}% | {grid} | {block} |\n")
    print(open(out).read())


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def full(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` "
                f"(`{path.split('/')[-1]}`, read with `ncu -i ... --page raw --csv`).\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write(f"## `{short(d['Kernel Name'])}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d and d[k] not in ("", "n/a"):
                    f.write(f"| {k} | {d[k]} | {u.get(k, '')} |\n")
            f.write("\n")
    print(open(out).read())


if __name__ == "__main__":
    mode, path, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else path
    (launches if mode == "launches" else full)(path, out, title)
