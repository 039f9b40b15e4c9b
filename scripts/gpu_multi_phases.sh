#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
SL_PHASES=1 SL_NOREF=1 timeout 600 $TR scripts/sortlast_gpu_check.py 2048 2048 preint > gpurun_out/sortlast_phases_$N.txt 2>&1; echo "rc=$?"
grep -E "view|rc=|Error|error" gpurun_out/sortlast_phases_$N.txt | head -60 | cut -c1-330
