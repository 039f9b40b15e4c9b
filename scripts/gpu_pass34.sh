#!/bin/bash
# pass 34: slab cases for MOP / MIDA / FHP / FHN
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in mop mida fhp fhn; do timeout 300 python tests/perf_probe.py 2048 --comp=$c 2>&1 | grep "cls=tf" | cut -c1-200; done | tee gpurun_out/pass34_modes.txt
for c in mop mida; do B200RC_MARCHER=ray timeout 300 python tests/perf_probe.py 2048 --comp=$c 2>&1 | grep "cls=tf" | cut -c1-200 | sed 's/^/per-ray kernel: /'; done | tee -a gpurun_out/pass34_modes.txt
