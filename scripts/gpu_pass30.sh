#!/bin/bash
# pass 30: warp-granular slab marcher (B200RC_MARCHER=wslab): parity suite + timings
mkdir -p gpurun_out; O=gpurun_out/pass30_wslab.txt; : > $O
B200RC_MARCHER=wslab timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee -a $O
run() { echo "== $1" >> $O; shift; env "$@" timeout 300 python tests/perf_probe.py 2048 2>&1 | grep "cls=\|rror" | cut -c1-125 >> $O; }
run "CTA groups (default)" X=1
run "warp groups, depth 6" B200RC_MARCHER=wslab
run "warp groups, depth 4" B200RC_MARCHER=wslab B200RC_WSLAB_DEPTH=4
run "warp groups, depth 3" B200RC_MARCHER=wslab B200RC_WSLAB_DEPTH=3
run "warp groups, depth 8" B200RC_MARCHER=wslab B200RC_WSLAB_DEPTH=8
echo "== 1/8 band, warp groups depth 4 / CTA groups" >> $O
B200RC_MARCHER=wslab B200RC_WSLAB_DEPTH=4 timeout 300 python tests/perf_probe.py 2048 --band=8 2>&1 | grep "cls=" | cut -c1-125 >> $O
timeout 300 python tests/perf_probe.py 2048 --band=8 2>&1 | grep "cls=" | cut -c1-125 >> $O
cat $O
