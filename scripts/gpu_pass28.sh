#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/pass28_batch.txt; : > $O
run() { echo "== $1" >> $O; shift; env "$@" timeout 300 python tests/perf_probe.py 2048 2>&1 | grep "cls=" | cut -c1-125 >> $O; }
run "batch 4 (default)" X=1
run "batch 2" B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_b2.so
run "batch 8" B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_b8.so
run "batch 8, depth 8" B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_b8.so B200RC_SLAB_DEPTH=8
run "batch 4, depth 5" B200RC_SLAB_DEPTH=5
run "batch 4, depth 7" B200RC_SLAB_DEPTH=7
cat $O
