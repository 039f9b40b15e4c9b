#!/bin/bash
# pass 8: new tests (multi-volume, streaming min/max), min/max A/B timing, bench line with minmax_build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt | cut -c1-250
timeout 900 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_p8.json 2> gpurun_out/bench_c3_p8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c3_p8.json')); print(d['ms_per_step'], d['minmax_build'], d['roofline']['traffic'])"
tail -3 gpurun_out/bench_c3_p8.err
