#!/bin/bash
# pass 25: two lattice samples in flight per thread (SLAB_PAIR) -- parity + timing, 3 CTAs/SM (80 regs) vs 2 CTAs/SM (128 regs) vs unpaired
mkdir -p gpurun_out; O=gpurun_out/pass25_pair.txt; : > $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee -a $O
run() { echo "== $1" >> $O; shift; env "$@" timeout 300 python tests/perf_probe.py 2048 2>&1 | grep "cls=" | cut -c1-125 >> $O; }
run "paired, 3 CTAs/SM" X=1
run "unpaired (SLAB_PAIR=0), 3 CTAs/SM" B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_p0.so
run "paired, 2 CTAs/SM, 72 KB" B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_c2.so
run "paired, 2 CTAs/SM, 100 KB" B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_c2.so B200RC_SLAB_KB=100
run "paired, 2 CTAs/SM, 100 KB, depth 8" B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_c2.so B200RC_SLAB_KB=100 B200RC_SLAB_DEPTH=8
cat $O
