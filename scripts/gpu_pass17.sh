#!/bin/bash
# pass 17: band height vs. efficiency of a 1/8 image share on one GPU (locality of interleaved 16-row bands)
mkdir -p gpurun_out; O=gpurun_out/pass17_bandrows.txt; : > $O
run() { echo "== $1" >> $O; shift; env "$@" timeout 300 python tests/perf_probe.py 2048 --band=8 2>&1 | grep "cls=pi" | cut -c1-120 >> $O; }
run "16 rows" X=1
run "32 rows" B200RC_BAND_ROWS=32
run "64 rows" B200RC_BAND_ROWS=64
run "128 rows" B200RC_BAND_ROWS=128
run "256 rows (one contiguous strip: rows 0-255 = top edge, little work)" B200RC_BAND_ROWS=256
cat $O
