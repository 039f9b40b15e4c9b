#!/usr/bin/env python
"""Per-source-line instruction / stall breakdown of an .ncu-rep captured with --import-source on (kernels built -lineinfo)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; cur_file = None; cur_line = None; cur_src = ''
agg = collections.Counter(); src = {}; stall = collections.Counter(); tot = 0; ops = collections.Counter(); seen = set()
def num(x):
    try: return int(x)
    except Exception: return 0
for r in rows:
    if not r: continue
    if r[0] in ('File Path', 'File Name'): cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; ie = hdr.index('Instructions Executed'); ws = hdr.index('Warp Stall Sampling (All Samples)'); continue
    if hdr is None or len(r) <= ie: continue
    if r[0] != '': cur_line = r[0]; cur_src = r[1]
    addr = r[2]
    if not addr or addr in seen: continue
    seen.add(addr)
    n = num(r[ie]); s = num(r[ws])
    key = (cur_file, cur_line)
    agg[key] += n; stall[key] += s; src[key] = cur_src; tot += n
    t = r[3].split()
    op = t[0] if not t[0].startswith('@') else t[1]
    ops[op.split('.')[0]] += n
ts = max(1, sum(stall.values()))
print('total warp instructions', tot, 'stall samples', ts)
print('opcode mix: ' + ', '.join(f"{k} {v / tot * 100:.1f}%" for k, v in ops.most_common(18)))
for k, v in agg.most_common(top):
    print(f"{v / tot * 100:5.2f}% inst  {stall[k] / ts * 100:5.2f}% stall  {k[0]}:{k[1]}: {src[k].strip()[:110]}")
