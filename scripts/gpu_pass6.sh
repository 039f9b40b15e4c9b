#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
grep -E "sharded|nucleon|passed|failed|rc=|Error|error" gpurun_out/pytest_gpu.txt | head -30 | cut -c1-250
