#!/bin/bash
# pass 20 (2 GPUs): compacted re-march list vs. tile re-march -- parity + phase times, sort-last bench
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard or sort" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out/pass20_remarch_$N.txt; : > $O
for rm in list slab; do
  echo "== re-march: $rm" >> $O
  B200RC_REMARCH=$rm SL_PHASES=1 timeout 300 $TR scripts/sortlast_gpu_check.py 1024 2048 preint sort-last > gpurun_out/tmp_check.txt 2>&1; echo "rc=$?" >> $O
  grep -E "^sort-|rank 0|Error|Traceback" gpurun_out/tmp_check.txt | cut -c1-330 >> $O
done
cat $O | cut -c1-330
for rm in list slab; do
B200RC_REMARCH=$rm timeout 600 $TR bench.py --gpus $N --steps 12 --warmup 3 --parallelism sort-last > gpurun_out/bench_n${N}_sortlast_$rm.json 2> gpurun_out/bench_n${N}_sortlast_$rm.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_n${N}_sortlast_$rm.json')); print('   $rm', d['config']['parallelism'], 'ms/frame', round(d['ms_per_step'],3), d.get('sortlast_phases'))"
done
