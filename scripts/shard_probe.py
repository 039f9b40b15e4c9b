#!/usr/bin/env python
"""Developer probe: first-pass march time of every shard of a sort-last split, on ONE GPU (no exchange)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import voreen_b200 as vb
from voreen_b200 import abi, host, sortlast

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
img = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
worlds = [int(a) for a in sys.argv[3:] if a.isdigit()] or [1, 2, 8]
SLABS = "slabs" in sys.argv


def slab_desc(dims, rank, world):
    b = [((dims[1] * i // world) + 7) // 8 * 8 for i in range(world + 1)]
    b[-1] = dims[1]
    from voreen_b200 import raycaster as rc
    return rc.make_volume_desc(dims, shard_lo=(0, b[rank], 0), shard_hi=(dims[0], b[rank + 1], dims[2]), shard_apron=8)

ctx = vb.Context(0)
dims = (size,) * 3
step = float(np.float32(1.0) / (np.float32(size) * np.float32(2.0)))
params = vb.default_params(step, classification=abi.CLASS_PREINT, out_format=abi.OUT_RGBA32F)
eep = vb.default_eep()
views = [0, 9, 14]
cams = {k: host.make_camera(bench.orbit_eye(k), W=img, H=img) for k in views}
for world in worlds:
    for rank in range(world):
        desc = slab_desc(dims, rank, world) if (SLABS and world > 1) else sortlast.shard_desc(dims, rank, world)
        ctx.synth_volume(0, desc, 1234)
        ctx.set_tf1d(0, host.tf1d_texture(bench.CT_KEYS, 1024), (0.0, 65535.0))
        ctx.build_preint(0, 1024, step, abi.PI_CPU)
        ctx.set_tf1d(0, host.tf1d_texture(bench.CT_KEYS, 4096), (0.0, 65535.0))
        params.partial_image = 1 if world > 1 else 0
        line = []
        for k in views:
            params.collect_stats = 0
            ctx.render_camera(params, eep, cams[k], img, img)
            ctx.render_camera(params, eep, cams[k], img, img)
            ms = ctx.stats().last_march_ms
            params.collect_stats = 1
            ctx.render_camera(params, eep, cams[k], img, img)
            st = ctx.stats()
            line.append(f"v{k}: {ms:6.2f} ms cls {st.samples_classified/1e6:6.1f}M shd {st.samples_shaded/1e6:6.1f}M skip {st.samples_skipped/1e6:7.1f}M rays {st.rays_cast}")
        print(f"world {world} rank {rank} lo {tuple(desc.shard_lo)} hi {tuple(desc.shard_hi)} | " + " | ".join(line), flush=True)
