#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt | cut -c1-250
for d in 8 4 6 12; do
  echo "== SLAB_DEPTH=$d" >> gpurun_out/perf_probe.txt
  B200RC_SLAB_DEPTH=$d timeout 300 python tests/perf_probe.py 512 2048 >> gpurun_out/perf_probe.txt 2>&1
done
echo "== KB=104 depth 8" >> gpurun_out/perf_probe.txt
B200RC_SLAB_KB=104 timeout 300 python tests/perf_probe.py 2048 >> gpurun_out/perf_probe.txt 2>&1
cat gpurun_out/perf_probe.txt | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_slab -s 4 -c 1 -o gpurun_out/prof_slab python tests/perf_probe.py 2048 > gpurun_out/ncu_slab.log 2>&1
tail -3 gpurun_out/ncu_slab.log | cut -c1-200
