#!/bin/bash
mkdir -p gpurun_out
g++ -g -O1 -std=c++14 scripts/dbg/host_t2.cpp -Lvoreen_b200 -lb200rc_host -lb200rc -Wl,-rpath,$PWD/voreen_b200 -o /tmp/t2 && /tmp/t2
echo "--- python"
python - <<'PY'
import sys, faulthandler; faulthandler.enable()
sys.path.insert(0,'tests')
from voreen_b200 import host
p = host.SingleVolumeRaycaster(0)
print("created", p.h, flush=True)
p.set_property("compositing","dvr")
print("set ok", flush=True)
p.close()
import voreen_b200 as vb
c = vb.Context(0)
p = host.SingleVolumeRaycaster(0)
print("created 2", p.h, flush=True)
p.set_property("compositing","dvr")
print("set ok 2", flush=True)
PY
which gdb valgrind
