#!/bin/bash
# 8-GPU pass: parity of the peer-memory sort-last path at 8 shards, C5 (4096^3 sharded 8 ways, 3840x2160), strong scaling of C3
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_8.txt
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511"; }
SL_TRANSPORT=p2p timeout 300 $(tr 8) scripts/sortlast_gpu_check.py 512 1024 preint sort-last > gpurun_out/tmp_check8.txt 2>&1; echo "check rc=$?"
grep -E "^sort-|Error|Traceback" gpurun_out/tmp_check8.txt | cut -c1-230 > gpurun_out/p2p_check_8.txt; cat gpurun_out/p2p_check_8.txt | tail -7
run() { name=$1; n=$2; shift 2; timeout 900 $(tr $n) bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/$name.json')); print('   ', d['config']['workload'][:40], d['config']['parallelism'], d['config'].get('transport'), 'ms/frame', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'nominal', round(d['gsamples_nominal_per_s'],1), d.get('sortlast_phases'))" || tail -5 gpurun_out/$name.err | cut -c1-300; }
run bench_c5_n8 8 --size 4096 --image 3840 --image-h 2160 --parallelism sort-last --steps 12 --warmup 3
run bench_n8 8 --steps 36 --warmup 5
run bench_n8_sortlast 8 --steps 12 --warmup 3 --parallelism sort-last
run bench_n4 4 --steps 36 --warmup 5
run bench_n2 2 --steps 36 --warmup 5
