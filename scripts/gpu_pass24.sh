#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tests/perf_probe.py 2048 --comp=mip --ess=1; timeout 300 python tests/perf_probe.py 2048 --comp=mip --ess=2; timeout 300 python tests/perf_probe.py 2048 --comp=iso; timeout 300 python tests/perf_probe.py 2048) 2>&1 | grep "cls=" | cut -c1-250 > gpurun_out/pass24_mip_cull.txt; cat gpurun_out/pass24_mip_cull.txt
