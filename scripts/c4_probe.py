"""Developer probe (not a test): config C4 -- MultiVolumeRaycaster, 4 x 1024^3 u16 with per-volume TFs, DVR at 2048^2, one GPU."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import scenes  # noqa: E402
import voreen_b200 as vb  # noqa: E402
from voreen_b200 import abi, host  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
img = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
nvol = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = vb.Context(0)
desc = vb.make_volume_desc((n, n, n))
keys = [scenes.CT_KEYS, [(0.0, (0, 0, 0, 0)), (0.3, (0, 255, 0, 0)), (1.0, (0, 255, 0, 120))],
        [(0.0, (0, 0, 0, 0)), (0.4, (0, 0, 255, 0)), (1.0, (0, 0, 255, 90))], [(0.0, (0, 0, 0, 0)), (0.5, (255, 0, 255, 0)), (1.0, (255, 0, 255, 60))]]
for v in range(nvol):
    ctx.synth_volume(v, desc, 1234 + v)
    ctx.set_tf1d(v, host.tf1d_texture(keys[v], 4096), (0.0, 65535.0))
ctx.synchronize()
step_world = 2.0 / n / 2.0          # multivolumeraycaster.cpp:326-344: min_v max(cubeSizeWorld / dim) / samplingRate
step_tex = 1.0 / (n * 2.0)
for ess in (1, 0):
    p = vb.default_params(step_world, empty_space_skipping=ess, out_format=abi.OUT_RGBA16)
    p.num_volumes = nvol; p.multi_volume = 1
    p.mv_opacity_correction_factor = step_tex / step_world
    eep = vb.default_eep(world_coords=1)
    for k in (0, 9):
        cam = host.make_camera(scenes.orbit_eye(k), W=img, H=img)
        e, ed, x, xd = ctx.entry_exit(eep, cam, img, img)
        entry = np.stack([e] * (nvol + 1)); exitp = np.stack([x] * (nvol + 1))
        p.collect_stats = 0
        ctx.render(p, cam, img, img, entry, ed, exitp, xd)
        t0 = time.time()
        ctx.render(p, cam, img, img, entry, ed, exitp, xd)
        wall = time.time() - t0
        ms = ctx.stats().last_march_ms
        p.collect_stats = 1
        ctx.render(p, cam, img, img, entry, ed, exitp, xd)
        s = ctx.stats()
        print(f"[C4 {nvol}x{n}^3 @{img}^2 ess={ess}] view {k}: march {ms:.2f} ms (call wall {wall*1e3:.0f} ms) rays {s.rays_cast} classified {s.samples_classified/1e6:.1f}M "
              f"shaded {s.samples_shaded/1e6:.1f}M skipped {s.samples_skipped/1e6:.1f}M nominal {s.samples_nominal/1e6:.1f}M", flush=True)
