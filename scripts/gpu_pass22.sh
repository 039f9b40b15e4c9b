#!/bin/bash
# pass 22: 16x8-pixel tiles (128 threads, 6 CTAs/SM, 36 KB regions) against 16x16 (256 threads, 3 CTAs/SM, 72 KB), full image and 1/8 share
mkdir -p gpurun_out; O=gpurun_out/pass22_tile128.txt; : > $O
run() { echo "== $1" >> $O; b=$2; shift 2; env "$@" timeout 300 python tests/perf_probe.py 2048 --band=$b 2>&1 | grep "cls=" | cut -c1-125 >> $O; }
L=$PWD/voreen_b200/libb200rc_t128.so
run "256 threads (default), full" 1 X=1
run "128 threads, full" 1 B200RC_LIBRARY=$L
run "128 threads depth 4, full" 1 B200RC_LIBRARY=$L B200RC_SLAB_DEPTH=4
run "128 threads depth 8, full" 1 B200RC_LIBRARY=$L B200RC_SLAB_DEPTH=8
run "256 threads, 1/8" 8 X=1
run "128 threads, 1/8" 8 B200RC_LIBRARY=$L
cat $O
