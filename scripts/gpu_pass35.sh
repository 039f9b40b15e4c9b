#!/bin/bash
# pass 35: ncu captures of the two other kernels DESIGN.md quotes: the streaming min/max build (HBM-bound) and the multi-volume marcher (C4)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:brick_partials_u16 -c 1 -o gpurun_out/prof_minmax python tests/perf_probe.py 2048 > gpurun_out/ncu_minmax.log 2>&1; tail -1 gpurun_out/ncu_minmax.log | cut -c1-150
timeout 600 ncu --set full --clock-control none -k regex:march_kernel -s 2 -c 1 -o gpurun_out/prof_multi python tests/c4_probe.py 1024 2048 4 > gpurun_out/ncu_multi.log 2>&1; tail -1 gpurun_out/ncu_multi.log | cut -c1-150
ls -la gpurun_out/prof_minmax.ncu-rep gpurun_out/prof_multi.ncu-rep
