#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -6 gpurun_out/pytest_gpu.txt | cut -c1-250
bash scripts/gpu_multi.sh 2
