#!/bin/bash
# pass 32 (2 GPUs): slab depth of the sort-last re-march
N=2
mkdir -p gpurun_out; O=gpurun_out/pass32_remarch_depth.txt; : > $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for d in 6 10 16 24; do
  echo "== re-march depth $d" >> $O
  B200RC_SLAB_DEPTH_REMARCH=$d SL_PHASES=1 timeout 300 $TR scripts/sortlast_gpu_check.py 2048 2048 preint sort-last > gpurun_out/tmp_check.txt 2>&1; echo "rc=$?" >> $O
  grep -E "^sort-|rank 0|Error|Traceback" gpurun_out/tmp_check.txt | cut -c1-300 | head -8 >> $O
done
cat $O | cut -c1-300
