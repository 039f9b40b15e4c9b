#!/bin/bash
# pass 14: lazy shading for MIP / ISO (gradient + Phong only for samples the compositing predicate keeps); full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pass14_pytest.txt; cat gpurun_out/pass14_pytest.txt
(timeout 300 python tests/perf_probe.py 2048 --comp=mip; timeout 300 python tests/perf_probe.py 2048 --comp=iso) > gpurun_out/pass14_mip_iso.txt 2>&1
grep -v "synth\|preint build" gpurun_out/pass14_mip_iso.txt | cut -c1-260
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/pass14_bench.json 2> gpurun_out/pass14_bench.err; cut -c1-400 gpurun_out/pass14_bench.json; tail -3 gpurun_out/pass14_bench.err
