#!/bin/bash
# multi-GPU pass (gpurun --gpus N): sort-last parity against the single-GPU image, then the bench line at N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/sortlast_gpu_check.py 512 1024 preint > gpurun_out/sortlast_check_$N.txt 2>&1; echo "check rc=$?" >> gpurun_out/sortlast_check_$N.txt
grep -E "sort-last|rc=|Error|error" gpurun_out/sortlast_check_$N.txt | tail -12 | cut -c1-220
timeout 600 $TR scripts/sortlast_gpu_check.py 512 1024 preint sort-first > gpurun_out/sortfirst_check_$N.txt 2>&1; echo "check rc=$?" >> gpurun_out/sortfirst_check_$N.txt
grep -E "sort-first|rc=|Error|error" gpurun_out/sortfirst_check_$N.txt | tail -8 | cut -c1-220
timeout 900 $TR bench.py --gpus $N --steps 12 --warmup 3 --parallelism sort-last > gpurun_out/bench_n${N}_sortlast.json 2> gpurun_out/bench_n${N}_sortlast.err; echo "bench sort-last rc=$?"
cut -c1-200 gpurun_out/bench_n${N}_sortlast.json
timeout 900 $TR bench.py --gpus $N --steps 12 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
cat gpurun_out/bench_n$N.json | cut -c1-900; tail -5 gpurun_out/bench_n$N.err | cut -c1-300
