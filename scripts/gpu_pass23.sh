#!/bin/bash
# pass 23: MIP max-culling (empty_space_skipping = 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "mip or full_size or compositing" 2>&1 | grep -v "^$" | tail -22 | cut -c1-300 > gpurun_out/pass23_pytest.txt; cat gpurun_out/pass23_pytest.txt
(timeout 300 python tests/perf_probe.py 2048 --comp=mip --ess=1; timeout 300 python tests/perf_probe.py 2048 --comp=mip --ess=2) 2>&1 | grep "cls=" | cut -c1-260 > gpurun_out/pass23_mip_cull.txt; cat gpurun_out/pass23_mip_cull.txt
