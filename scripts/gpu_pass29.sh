#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/pass29_map.txt; : > $O
run() { echo "== $1" >> $O; shift; env "$@" timeout 300 python tests/perf_probe.py 2048 2>&1 | grep "cls=" | cut -c1-125 >> $O; }
run "8x4 block per warp (default)" X=1
run "two 8x2 blocks 8 rows apart" B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_m1.so
run "rows w and w+8, 16 wide" B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_m2.so
cat $O
