#!/bin/bash
# first GPU pass of the round: parity suite, bench, launch list, one full ncu capture of the marcher, variant probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 36 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 300 python bench.py --size 512 --image 1024 --classification tf --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 600 python tests/perf_probe.py 512 2048 > gpurun_out/perf_probe.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 40 -c 2 -o gpurun_out/prof_march python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err; cat gpurun_out/bench_c2.json
