#!/bin/bash
# pass 16: why does 1/8 of the image take 4.2 ms instead of 2.5?  band-limited single-GPU runs: tile order, slab depth
mkdir -p gpurun_out; O=gpurun_out/pass16_band.txt; : > $O
run() { echo "== $1" >> $O; shift; env "$@" timeout 300 python tests/perf_probe.py 2048 --band=8 2>&1 | grep "cls=pi" | cut -c1-120 >> $O; }
run "LPT order, depth 6" X=1
run "Z order, depth 6" B200RC_TILE_ORDER=0
run "LPT, depth 4" B200RC_SLAB_DEPTH=4
run "LPT, depth 10" B200RC_SLAB_DEPTH=10
run "LPT, depth 16" B200RC_SLAB_DEPTH=16
run "LPT, per-ray kernel" B200RC_MARCHER=ray
echo "== band 1/1 Z order" >> $O; B200RC_TILE_ORDER=0 timeout 300 python tests/perf_probe.py 2048 2>&1 | grep "cls=pi" | cut -c1-120 >> $O
cat $O
