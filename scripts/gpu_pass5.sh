#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/perf_probe.txt
echo "== 256 threads" >> gpurun_out/perf_probe.txt
timeout 300 python tests/perf_probe.py 2048 >> gpurun_out/perf_probe.txt 2>&1
cp voreen_b200/libb200rc.so /tmp/keep.so; cp voreen_b200/libb200rc_t128.so voreen_b200/libb200rc.so
for kb in 52 72; do
echo "== 128 threads KB=$kb" >> gpurun_out/perf_probe.txt
B200RC_SLAB_KB=$kb timeout 300 python tests/perf_probe.py 2048 >> gpurun_out/perf_probe.txt 2>&1
done
echo "== 128 threads KB=52 depth 6" >> gpurun_out/perf_probe.txt
B200RC_SLAB_DEPTH=6 timeout 300 python tests/perf_probe.py 2048 >> gpurun_out/perf_probe.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cp /tmp/keep.so voreen_b200/libb200rc.so
grep -v "synth\|preint" gpurun_out/perf_probe.txt | cut -c1-120
