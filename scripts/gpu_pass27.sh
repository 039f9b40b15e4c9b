#!/bin/bash
# pass 27: leader election by warp reduce, batched staging loads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tests/perf_probe.py 2048 2>&1 | grep "cls=" | cut -c1-125 > gpurun_out/pass27.txt
B200RC_DEBUG=1 B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_pf.so timeout 300 python tests/perf_probe.py 2048 2>&1 | grep "slab phases" | cut -c1-260 >> gpurun_out/pass27.txt
cat gpurun_out/pass27.txt
