#!/bin/bash
# pass 18: resident CTAs per SM vs. time, full image and a 1/8 share (throughput-bound or latency-chain-bound?)
mkdir -p gpurun_out; O=gpurun_out/pass18_ctas.txt; : > $O
run() { echo "== $1" >> $O; b=$2; shift 2; env "$@" timeout 300 python tests/perf_probe.py 2048 --band=$b 2>&1 | grep "cls=pi" | cut -c1-120 >> $O; }
for b in 8 1; do for c in 1 2 3; do run "band 1/$b, $c CTAs/SM" $b B200RC_SLAB_CTAS_PER_SM=$c; done; done
cat $O
