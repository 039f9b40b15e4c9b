#!/usr/bin/env python
"""SASS-level view of an .ncu-rep captured with --import-source on: instructions sorted by address with executed counts;
prints the opcode mix of the instructions executed at least `frac` x the hottest instruction's count, and optionally the listing."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5; listing = len(sys.argv) > 3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
ins = []
for r in rows:
    if not r: continue
    if r[0] == 'Address' or (len(r) > 1 and r[1] == 'Source'):
        hdr = r; continue
    if hdr is None: continue
    try:
        ie = hdr.index('Instructions Executed'); src = hdr.index('Source'); ad = hdr.index('Address')
    except ValueError:
        continue
    if len(r) <= ie: continue
    try: n = int(r[ie])
    except Exception: continue
    tp = hdr.index('Thread Instructions Executed') if 'Thread Instructions Executed' in hdr else None
    ins.append((r[ad], r[src], n, int(r[tp]) if tp is not None and r[tp].isdigit() else 0))
if not ins:
    print(hdr); sys.exit(1)
mx = max(n for _, _, n, _ in ins)
tot = sum(n for _, _, n, _ in ins)
print('instructions', len(ins), 'total executed', tot, 'hottest', mx)
for lo, hi in ((0.5, 10), (0.1, 0.5), (0.01, 0.1), (0, 0.01)):
    sel = [(a, s, n, t) for a, s, n, t in ins if lo * mx <= n < hi * mx]
    ops = collections.Counter()
    for a, s, n, t in sel:
        tk = s.split()
        op = tk[1] if tk and tk[0].startswith('@') else (tk[0] if tk else '?')
        ops[op.split('.')[0]] += n
    st = sum(n for _, _, n, _ in sel)
    print(f"band [{lo}, {hi}) x hottest: {len(sel)} static instr, {st / tot * 100:.1f}% of executed: " + ', '.join(f"{k} {v / max(st,1) * 100:.1f}%" for k, v in ops.most_common(16)))
if listing:
    for a, s, n, t in ins:
        if n >= frac * mx: print(f"{a} {n:>12} {t / max(n,1):5.1f}  {s}")
