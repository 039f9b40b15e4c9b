#!/bin/bash
# occupancy variants of the slab marcher: 2 CTAs/SM (123 regs, 88 KB) vs 3 (80 regs) vs 4 (64 regs)
mkdir -p gpurun_out; rm -f gpurun_out/perf_probe.txt
run() { echo "== $1" >> gpurun_out/perf_probe.txt; shift; env "$@" timeout 300 python tests/perf_probe.py 2048 >> gpurun_out/perf_probe.txt 2>&1; }
run "2 CTAs/SM KB=88" X=1
cp voreen_b200/libb200rc.so /tmp/keep.so
cp voreen_b200/libb200rc_c3.so voreen_b200/libb200rc.so
run "3 CTAs/SM KB=72" B200RC_SLAB_KB=72
run "3 CTAs/SM KB=56" B200RC_SLAB_KB=56
run "3 CTAs/SM KB=56 depth 6" B200RC_SLAB_KB=56 B200RC_SLAB_DEPTH=6
cp voreen_b200/libb200rc_c4.so voreen_b200/libb200rc.so
run "4 CTAs/SM KB=52" B200RC_SLAB_KB=52
run "4 CTAs/SM KB=40 depth 6" B200RC_SLAB_KB=40 B200RC_SLAB_DEPTH=6
cp /tmp/keep.so voreen_b200/libb200rc.so
grep -v "synth\|preint build" gpurun_out/perf_probe.txt | cut -c1-130
