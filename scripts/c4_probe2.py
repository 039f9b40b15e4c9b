"""Developer probe (not a test): config C4 through b200rc_render_camera (co-registered volumes, shared world-space entry/exit):
the staging marcher (default) or the per-ray kernel (B200RC_MARCHER=ray)."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import scenes  # noqa: E402
import voreen_b200 as vb  # noqa: E402
from voreen_b200 import abi, host  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
img = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
nvol = int(sys.argv[3]) if len(sys.argv) > 3 else 4
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["tf", "preint"]
ctx = vb.Context(0)
desc = vb.make_volume_desc((n, n, n))
keys = [scenes.CT_KEYS, [(0.0, (0, 0, 0, 0)), (0.3, (0, 255, 0, 0)), (1.0, (0, 255, 0, 120))],
        [(0.0, (0, 0, 0, 0)), (0.4, (0, 0, 255, 0)), (1.0, (0, 0, 255, 90))], [(0.0, (0, 0, 0, 0)), (0.5, (255, 0, 255, 0)), (1.0, (255, 0, 255, 60))]]
step_world = 2.0 / n / 2.0
step_tex = 1.0 / (n * 2.0)
for v in range(nvol):
    ctx.synth_volume(v, desc, 1234 + v)
ctx.synchronize()
tag = f"marcher={os.environ.get('B200RC_MARCHER', 'slab')} depth={os.environ.get('B200RC_MSLAB_DEPTH', '6')} kb={os.environ.get('B200RC_MSLAB_KB', 'auto')}"
for cls in modes:
    for v in range(nvol):
        if cls == "preint":
            ctx.set_tf1d(v, host.tf1d_texture(keys[v], 1024), (0.0, 65535.0))
            ctx.build_preint(v, 1024, step_tex, abi.PI_CPU)
        ctx.set_tf1d(v, host.tf1d_texture(keys[v], 4096), (0.0, 65535.0))
    p = vb.default_params(step_world, empty_space_skipping=1, out_format=abi.OUT_RGBA16)
    p.num_volumes = nvol; p.multi_volume = 1
    p.mv_opacity_correction_factor = step_tex / step_world
    for v in range(nvol):
        p.vol[v].classification = abi.CLASS_PREINT if cls == "preint" else abi.CLASS_TF
    eep = vb.default_eep(world_coords=1)
    tot = []
    for k in ([int(a) for a in os.environ["C4_VIEWS"].split(",")] if os.environ.get("C4_VIEWS") else (0, 9, 18, 27)):
        cam = host.make_camera(scenes.orbit_eye(k), W=img, H=img)
        p.collect_stats = 0
        ctx.render_camera(p, eep, cam, img, img)
        ctx.render_camera(p, eep, cam, img, img)
        ms = ctx.stats().last_march_ms
        tot.append(ms)
        p.collect_stats = 1
        ctx.render_camera(p, eep, cam, img, img)
        s = ctx.stats()
        print(f"[C4 {nvol}x{n}^3 @{img}^2 {cls} {tag}] view {k}: march {ms:.2f} ms rays {s.rays_cast} classified {s.samples_classified/1e6:.1f}M "
              f"shaded {s.samples_shaded/1e6:.1f}M skipped {s.samples_skipped/1e6:.1f}M nominal {s.samples_nominal/1e6:.1f}M", flush=True)
    print(f"[C4 {nvol}x{n}^3 @{img}^2 {cls} {tag}] mean march {sum(tot)/len(tot):.2f} ms", flush=True)
