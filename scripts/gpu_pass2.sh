#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
head -60 gpurun_out/pytest_gpu.txt | cut -c1-250
timeout 600 python tests/perf_probe.py 512 2048 > gpurun_out/perf_probe.txt 2>&1
cat gpurun_out/perf_probe.txt | cut -c1-330
