#!/bin/bash
# pass 21: .vvd / raw streaming loader, slab upload API; full suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pass21_pytest.txt; cat gpurun_out/pass21_pytest.txt
# streaming throughput from page cache: 2 GiB u16 raw -> HBM
python - <<'PY' 2>&1 | tee gpurun_out/pass21_stream.txt
import numpy as np, time, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import voreen_b200 as vb
from test_vvd import write_vvd
d = "/tmp/vvdtest"; os.makedirs(d, exist_ok=True)
n = 1024
vol = np.random.default_rng(1).integers(0, 65536, size=(n, n, n), dtype=np.uint16)
vol.tofile(f"{d}/big.raw"); write_vvd(f"{d}/big.vvd", "big.raw", "uint16", (n, n, n))
ctx = vb.Context(0)
for k in range(3):
    t0 = time.perf_counter(); ctx.load_vvd(0, f"{d}/big.vvd"); ctx.synchronize(); dt = time.perf_counter() - t0
    print(f"load_vvd {n}^3 u16 ({vol.nbytes / 2**30:.1f} GiB, page cache): {dt:.3f} s = {vol.nbytes / dt / 1e9:.2f} GB/s")
t0 = time.perf_counter(); ctx.upload_volume(0, vol); ctx.synchronize(); dt = time.perf_counter() - t0
print(f"upload_volume (pageable numpy, whole volume staged): {dt:.3f} s = {vol.nbytes / dt / 1e9:.2f} GB/s")
PY
