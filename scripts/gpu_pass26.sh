#!/bin/bash
# pass 26: where does a CTA of the slab marcher spend its cycles?  (-DSLAB_PROFILE build: clock64 of warp 0 at the phase boundaries)
mkdir -p gpurun_out
B200RC_DEBUG=1 B200RC_LIBRARY=$PWD/voreen_b200/libb200rc_pf.so timeout 300 python tests/perf_probe.py 2048 2>&1 | grep "cls=\|slab phases\|ray-rounds" | cut -c1-260 > gpurun_out/pass26_phases.txt
cat gpurun_out/pass26_phases.txt
