import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import scenes, orc
import voreen_b200 as vb
from voreen_b200 import abi
ctx = vb.Context(0)
dims = (5, 7, 9); W, H = 17, 3
rng = np.random.default_rng(dims[0] * 100 + W)
vol = (rng.random(dims[::-1]) ** 2 * 65535).astype(np.uint16)
for shade in ("phong", "none"):
    s = scenes.Scene(vol, W=W, H=H, empty_space_skipping=0, shading=abi.SHADE[shade])
    o, od, ost = s.oracle_render()
    c, cd = s.cuda_render(ctx)
    e, ed, x, xd = s.oracle_eep()
    d = np.abs(np.asarray(o[0], np.float64) - np.asarray(c[0], np.float64)).max(axis=2)
    print(shade, "cast pixels", np.argwhere(np.any(e[..., :3] != x[..., :3], axis=2)).tolist())
    for (yy, xx) in np.argwhere(d > 1e-3):
        print("  diff at", yy, xx, "oracle", np.asarray(o[0])[yy, xx], "cuda", np.asarray(c[0])[yy, xx], "entry", e[yy, xx], "exit", x[yy, xx])
