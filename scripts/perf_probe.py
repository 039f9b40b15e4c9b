"""Developer probe (not a test, not the bench): times the marcher on synthetic volumes for sampler / ESS variants."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import orc  # noqa: E402
import scenes  # noqa: E402
import voreen_b200 as vb  # noqa: E402
from voreen_b200 import abi  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith('--')] or [512]
    essmode = int(next((a.split('=')[1] for a in sys.argv if a.startswith('--ess=')), '1'))
    band = int(next((a.split('=')[1] for a in sys.argv if a.startswith('--band=')), '1'))
    comp = next((a.split('=')[1] for a in sys.argv if a.startswith('--comp=')), 'dvr')
    ctx = vb.Context(0)
    for n in sizes:
        img = 1024 if n <= 512 else 2048
        desc = vb.make_volume_desc((n, n, n))
        t0 = time.time()
        ctx.synth_volume(0, desc, 1234)
        ctx.synchronize()
        print(f"[{n}^3] synth+minmax {time.time() - t0:.2f}s", flush=True)
        tf = orc.make_tf(scenes.CT_KEYS, domain=(0, 65535))
        tex = orc.tf1d_texture(tf, 4096)
        ctx.set_tf1d(0, tex, (0.0, 65535.0))
        step = 1.0 / (n * 2.0)
        pi = None
        for cls in ("tf", "pi"):
            for sampler in ((abi.SAMPLER_EXACT, abi.SAMPLER_TEX) if '--tex' in sys.argv else (abi.SAMPLER_EXACT,)):
                for ess in ((1, 0) if '--noess' in sys.argv else (1,)):
                    if n >= 2048 and ess == 0 and sampler == abi.SAMPLER_EXACT:
                        continue
                    p = vb.default_params(step, sampler=sampler, empty_space_skipping=(essmode if ess else 0), out_format=abi.OUT_RGBA16, compositing=[abi.COMP[comp]] * 3)
                    if cls == "pi":
                        if pi is None:
                            tf1k = orc.tf1d_texture(tf, 1024)
                            ctx.set_tf1d(0, tf1k, (0.0, 65535.0))
                            t0 = time.time(); ctx.build_preint(0, 1024, step, abi.PI_CPU); ctx.synchronize()
                            print(f"   preint build {time.time() - t0:.3f}s")
                            ctx.set_tf1d(0, tex, (0.0, 65535.0))
                            pi = True
                        p.vol[0].classification = abi.CLASS_PREINT
                    p.image_band_modulus, p.image_band_remainder = band, 0
                    eep = vb.default_eep()
                    ms = []
                    for k in range(0, 36, 6):
                        cam = orc.make_camera(scenes.orbit_eye(k), W=img, H=img)
                        p.collect_stats = 0
                        ctx.render_camera(p, eep, cam, img, img)
                        ctx.render_camera(p, eep, cam, img, img)
                        s = ctx.stats()
                        ms.append((s.last_march_ms, s.last_eep_ms))
                    p.collect_stats = 1
                    ctx.render_camera(p, eep, cam, img, img)
                    s = ctx.stats()
                    m = np.median([a for a, _ in ms])
                    print(f"[{n}^3 @{img}^2] band 1/{band} comp={comp} cls={cls} sampler={'tex' if sampler else 'exact'} ess={ess}: march {m:.3f} ms (min {min(a for a,_ in ms):.3f} max {max(a for a,_ in ms):.3f}) "
                          f"eep {np.median([b for _, b in ms]):.3f} ms | rays {s.rays_cast} classified {s.samples_classified} shaded {s.samples_shaded} nominal {s.samples_nominal} "
                          f"cells16 {s.cells_touched} -> {s.samples_classified / m / 1e6:.2f} Gsamples/s", flush=True)


if __name__ == "__main__":
    main()
