"""Developer probe: how far is the texture-unit sampler (1.8 fixed-point weights) from the exact fp32 sampler?"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import orc  # noqa: E402
import scenes  # noqa: E402
import voreen_b200 as vb  # noqa: E402
from voreen_b200 import abi  # noqa: E402


def main():
    ctx = vb.Context(0)
    for n, img in ((64, 128), (128, 256), (256, 512), (512, 1024)):
        desc = vb.make_volume_desc((n, n, n))
        ctx.synth_volume(0, desc, 1234)
        tf = orc.make_tf(scenes.CT_KEYS, domain=(0, 65535))
        ctx.set_tf1d(0, orc.tf1d_texture(tf, 4096), (0.0, 65535.0))
        step = 1.0 / (n * 2.0)
        cam = orc.make_camera(scenes.orbit_eye(3), W=img, H=img)
        eep = vb.default_eep()
        for shade in ("none", "phong-diffuse-ambient", "phong"):
            for comp in ("dvr", "mip", "iso", "fhp"):
                res = {}
                for sampler in (abi.SAMPLER_EXACT, abi.SAMPLER_TEX):
                    p = vb.default_params(step, sampler=sampler, shading=abi.SHADE[shade], compositing=[abi.COMP[comp]] * 3, iso_value=0.2)
                    outs, depth = ctx.render_camera(p, eep, cam, img, img)
                    res[sampler] = outs[0]
                frac, mx = scenes.parity_report(res[0], res[1])
                d = np.abs(res[0] - res[1]).max(axis=2)
                print(f"{n}^3@{img}^2 shade={shade:22s} comp={comp}: within 1/255 {frac:.5f}  max {mx:.4f}  mean {d.mean():.2e}  p99.9 {np.quantile(d, 0.999):.4f}", flush=True)


if __name__ == "__main__":
    main()
