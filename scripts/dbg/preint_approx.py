import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import scenes, orc
import voreen_b200 as vb
from voreen_b200 import abi, host
ctx = vb.Context(0)
keys = scenes.CT_KEYS
res, step = 256, float(np.float32(1.0 / 96.0))
tf = orc.make_tf(keys, domain=(0.0, 65535.0))
want = orc.preint_cpu(tf, res, step, True)
ctx.set_tf1d(0, host.tf1d_texture(keys, res), (0.0, 65535.0))
ctx.build_preint(0, res, step, abi.PI_APPROX)
got = ctx.download_preint(0, res)
d = np.abs(got - want)
print("max", d.max(), "count > 1e-4:", int((d > 1e-4).sum()), "of", d.size)
idx = np.argwhere(d > 1e-4)
for (sb, sf, c) in idx[:12]:
    print(" sb", sb, "sf", sf, "ch", c, "got", got[sb, sf], "want", want[sb, sf])
tex = host.tf1d_texture(keys, res)
print("tf texels 60..90 alpha", tex[60:90, 3].tolist())
