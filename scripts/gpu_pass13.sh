#!/bin/bash
# pass 13: multi-volume ESS + specialised multi kernel (C4), C3 MIP / ISO timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multivolume" -s 2>&1 | grep -v "^$" | tail -15 > gpurun_out/pass13_multi_tests.txt
cat gpurun_out/pass13_multi_tests.txt
timeout 300 python tests/c4_probe.py 1024 2048 4 > gpurun_out/pass13_c4.txt 2>&1; cat gpurun_out/pass13_c4.txt
(timeout 300 python tests/perf_probe.py 2048 --comp=mip; timeout 300 python tests/perf_probe.py 2048 --comp=iso) > gpurun_out/pass13_mip_iso.txt 2>&1
grep -v "synth\|preint build" gpurun_out/pass13_mip_iso.txt | cut -c1-260
