#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/perf_probe.txt
run() { echo "== $1" >> gpurun_out/perf_probe.txt; shift; env "$@" timeout 300 python tests/perf_probe.py 2048 >> gpurun_out/perf_probe.txt 2>&1; }
run "3 CTAs/SM default (KB 72 / 55 with TF) depth 8" X=1
run "depth 6" B200RC_SLAB_DEPTH=6
run "depth 10" B200RC_SLAB_DEPTH=10
run "KB 64 depth 6" B200RC_SLAB_KB=64 B200RC_SLAB_DEPTH=6
run "KB 48 depth 6" B200RC_SLAB_KB=48 B200RC_SLAB_DEPTH=6
run "KB 48 depth 4" B200RC_SLAB_KB=48 B200RC_SLAB_DEPTH=4
grep -v "synth\|preint build" gpurun_out/perf_probe.txt | cut -c1-130
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
