#!/bin/bash
# pass 31 (final state of the round): full GPU parity suite, smoke, C3 bench line + reference arm, launch list, one full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt | cut -c1-250
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt | cut -c1-300
timeout 900 python bench.py --steps 36 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
cut -c1-300 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_slab -s 40 -c 1 -o gpurun_out/prof_slab_pi python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
timeout 300 python tests/c4_probe.py 1024 2048 4 2>&1 | grep "ess=1" > gpurun_out/pass31_c4.txt; cat gpurun_out/pass31_c4.txt
