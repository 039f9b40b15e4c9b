#!/bin/bash
# multi-GPU pass (gpurun --gpus N): peer-memory transport (symmetric memory) against NCCL send/recv, parity + frame times
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out/p2p_check_$N.txt; : > $O
for mode in sort-last sort-first; do for tr in p2p nccl; do
  SL_TRANSPORT=$tr timeout 300 $TR scripts/sortlast_gpu_check.py 512 1024 preint $mode > gpurun_out/tmp_check.txt 2>&1; rc=$?
  grep -E "^sort-|Error|error|Traceback" gpurun_out/tmp_check.txt | cut -c1-260 >> $O; echo "$mode $tr rc=$rc" >> $O
  [ $rc -ne 0 ] && tail -30 gpurun_out/tmp_check.txt | cut -c1-300 >> $O
done; done
cat $O | tail -40
timeout 600 $TR bench.py --gpus $N --size 1024 --image 960 --image-h 544 --parallelism sort-last --treat-as-sharded-only --steps 6 --warmup 3 > gpurun_out/bench_c5path_n$N.json 2> gpurun_out/bench_c5path_n$N.err; echo "bench c5-path rc=$?"; cut -c1-700 gpurun_out/bench_c5path_n$N.json; tail -4 gpurun_out/bench_c5path_n$N.err | cut -c1-300
for tr in p2p nccl; do for par in sort-last sort-first; do
  timeout 600 $TR bench.py --gpus $N --steps 12 --warmup 3 --parallelism $par --transport $tr > gpurun_out/bench_n${N}_${par}_$tr.json 2> gpurun_out/bench_n${N}_${par}_$tr.err; echo "bench $par $tr rc=$?"
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_n${N}_${par}_$tr.json')); print('   ', d['config']['parallelism'], d['config']['transport'], 'ms/frame', round(d['ms_per_step'],3), 'value', round(d['value'],1))" || tail -5 gpurun_out/bench_n${N}_${par}_$tr.err | cut -c1-300
done; done
