#!/bin/bash
# pass 15: longest-first tile order in the slab marcher
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pass15_pytest.txt; cat gpurun_out/pass15_pytest.txt
timeout 300 python tests/perf_probe.py 2048 > gpurun_out/pass15_probe.txt 2>&1; grep -v "synth\|preint build" gpurun_out/pass15_probe.txt | cut -c1-200
timeout 600 python bench.py --steps 36 --warmup 5 --no-cpu-baseline > gpurun_out/pass15_bench.json 2> gpurun_out/pass15_bench.err; cut -c1-300 gpurun_out/pass15_bench.json; tail -3 gpurun_out/pass15_bench.err
