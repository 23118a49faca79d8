#!/usr/bin/env python
"""Study: how far could a different ORDER of the search kernel's loop body cut its register-operand
reads (profiles/r2_k1_tuning.md sections 2 and 5)?

Takes the loop body of k_intersect3d<8, queue> out of `cuobjdump -sass`, builds its dependency graph
(RAW / WAR / WAW on registers, predicates and scoreboard barriers; memory instructions chained in
order), list-schedules it with "next instruction shares a source register in the same slot" as the
priority, and prints the operand-read model of the original and the reordered body.  Analysis only:
nothing is written.

usage: sass_sched.py file.sass first_line last_line     (1-based lines of the dump, inclusive)
"""
import re
import sys

sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), ".."))
from tensorflowraytrace_b200._reuse_patch import INSN, HIGH, PACKED  # noqa: E402


def load(path, first, last):
    out, pending = [], None
    for line in open(path).read().splitlines()[first - 1:last]:
        m = INSN.match(line)
        if m:
            pending = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)]
            continue
        m = HIGH.match(line)
        if m and pending is not None:
            out.append((pending[0], pending[1], pending[2], int(m.group(1), 16)))
            pending = None
    return out


def ctrl(hi):
    c = hi >> 41
    return dict(stall=c & 15, yield_=(c >> 4) & 1, wbar=(c >> 5) & 7, rbar=(c >> 8) & 7,
                wait=(c >> 11) & 63, reuse=(c >> 17) & 15)


def decode(text):
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*)", text)
    pred, op, rest = m.group(1), m.group(2), m.group(3)
    base = op.split(".")[0]
    args = [a.strip() for a in rest.split(",")] if rest else []
    reads, writes = [], []          # (kind, number)
    slots = []                      # per source slot: (reg, width) or None
    if pred:
        reads.append(("P", int(re.search(r"P(\d+)", pred).group(1))))
    wide = 2 if base in PACKED else 1
    if base == "LDS" and ".128" in op:
        wide = 4
    elif base == "LDS" and ".64" in op:
        wide = 2
    dst_done = False
    for idx, a in enumerate(args):
        if base in ("BRA", "BSSY", "BSYNC", "VOTE", "EXIT", "WARPSYNC"):
            break
        regs = re.findall(r"(?<![U\w])R(\d+)", a)
        preds = re.findall(r"(?<![U\w])P(\d+)", a)
        is_dst = (idx == 0 and not base.startswith(("ST", "RED", "ATOM"))) or \
                 (idx == 1 and base in ("FSETP", "ISETP", "DSETP") and False)
        if base in ("FSETP", "ISETP", "PLOP3") and idx in (0, 1):
            for p in preds:
                writes.append(("P", int(p)))
            continue
        if is_dst and not dst_done:
            dst_done = True
            for r in regs:
                for k in range(wide):
                    writes.append(("R", int(r) + k))
            continue
        for p in preds:
            reads.append(("P", int(p)))
        if regs:
            w = 2 if "F32x2" in a else 1
            if "[" in a:
                w = 1
            r = int(regs[0])
            for k in range(w):
                reads.append(("R", r + k))
            slots.append((r, w))
        else:
            slots.append(None)
    return base, op, reads, writes, slots


def graph(insns):
    n = len(insns)
    info = [decode(t) for _, t, _, _ in insns]
    cs = [ctrl(hi) for _, _, _, hi in insns]
    deps = [set() for _ in range(n)]
    last_w, last_r = {}, {}
    last_mem = None
    for i in range(n):
        base, op, reads, writes, _ = info[i]
        c = cs[i]
        reads, writes = list(reads), list(writes)
        for b in range(6):
            if c["wait"] >> b & 1:
                reads.append(("B", b))
        if c["wbar"] != 7:
            writes.append(("B", c["wbar"]))
        if c["rbar"] != 7:
            writes.append(("B", c["rbar"]))
        for r in reads:
            if r in last_w:
                deps[i].add(last_w[r])
        for w in writes:
            if w in last_w:
                deps[i].add(last_w[w])
            for j in last_r.get(w, ()):
                if j != i:
                    deps[i].add(j)
        if base in ("LDS", "LDG", "STS", "STG", "LDC", "ATOMS", "RED", "ATOM", "LDSM"):
            if last_mem is not None:
                deps[i].add(last_mem)
            last_mem = i
        for w in writes:
            last_w[w] = i
            last_r[w] = []
        for r in reads:
            last_r.setdefault(r, []).append(i)
    return info, cs, deps


def reads_model(order, info):
    prev, total, raw = {}, 0, 0
    for i in order:
        base, _, _, _, slots = info[i]
        cur = {}
        for s, v in enumerate(slots):
            if v is None:
                continue
            raw += v[1]
            if prev.get(s) != v or base not in PACKED:
                total += v[1]
            cur[s] = v
        prev = cur if base in PACKED else {}
    return total, raw


def schedule(info, deps, lookahead=True):
    n = len(info)
    users = [[] for _ in range(n)]
    indeg = [len(d) for d in deps]
    for i, d in enumerate(deps):
        for j in d:
            users[j].append(i)
    # critical-path length keeps the schedule from starving the tail
    depth = [0] * n
    for i in reversed(range(n)):
        depth[i] = 1 + max((depth[u] for u in users[i]), default=0)
    ready = [i for i in range(n) if indeg[i] == 0]
    order, prev = [], {}
    pos = {}
    while ready:
        def score(i):
            base, _, _, _, slots = info[i]
            share = 0
            if base in PACKED:
                for s, v in enumerate(slots):
                    if v is not None and prev.get(s) == v:
                        share += v[1]
            return (share, -i)      # original order breaks ties: keeps register pressure as ptxas had it
        best = max(ready, key=score)
        ready.remove(best)
        pos[best] = len(order)
        order.append(best)
        base, _, _, _, slots = info[best]
        prev = {s: v for s, v in enumerate(slots) if v is not None} if base in PACKED else {}
        for u in users[best]:
            indeg[u] -= 1
            if indeg[u] == 0:
                ready.append(u)
    return order


def main():
    path, first, last = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    insns = load(path, first, last)
    info, cs, deps = graph(insns)
    n = len(insns)
    t0, raw = reads_model(range(n), info)
    order = schedule(info, deps)
    t1, _ = reads_model(order, info)
    packed = sum(1 for b in info if b[0] in PACKED)
    print(f"{n} instructions, {packed} packed; operand reads without reuse {raw}; "
          f"ptxas order with every possible flag {t0}; reordered {t1}; FMA floor {packed * 4} reads-equivalent")
    if "-v" in sys.argv:
        for i in order:
            print(f"{i:4d}  {insns[i][1]}")


if __name__ == "__main__":
    main()
