"""Deterministic test scenes shared by the parity tests ("parity subset")."""
from __future__ import annotations

import math

import numpy as np

import orc
from voreen_b200 import abi, host
from voreen_b200 import raycaster as rc

# SURVEY 8d transfer function (intensity : rgba8)
CT_KEYS = [(0.00, (0, 0, 0, 0)), (0.25, (0, 0, 0, 0)), (0.30, (255, 140, 60, 8)), (0.55, (255, 230, 200, 40)),
           (0.80, (255, 255, 255, 200)), (1.00, (255, 255, 255, 255))]


def ct_like_volume(n, seed=1234, dtype=np.uint16, dims=None):
    """Small CPU stand-in of the synthetic CT-like volume: shell + soft ellipsoid + blobs + noise (numpy, seeded)."""
    dims = dims or (n, n, n)
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(*(np.arange(d, dtype=np.float32) for d in dims[::-1]), indexing="ij")
    px = (x + 0.5) / dims[0] - 0.5; py = (y + 0.5) / dims[1] - 0.5; pz = (z + 0.5) / dims[2] - 0.5
    r = np.sqrt(px * px + py * py + pz * pz)
    q = np.sqrt((px / 0.42) ** 2 + (py / 0.36) ** 2 + (pz / 0.40) ** 2)
    fall = np.clip((1.0 - q) * 0.36 * min(dims) / 4.0, 0, 1)
    f = 0.35 * fall * fall * (3 - 2 * fall)
    for _ in range(16):
        c = rng.uniform(-0.38, 0.38, 3); rad = rng.uniform(0.02, 0.05)
        f = np.where((px - c[0]) ** 2 + (py - c[1]) ** 2 + (pz - c[2]) ** 2 <= rad * rad, np.maximum(f, 0.6), f)
    f = np.where((r >= 0.30) & (r <= 0.32), np.maximum(f, 0.85), f)
    f = f + rng.normal(0, 0.01, f.shape).astype(np.float32)
    if dtype == np.uint16:
        return np.clip(f * 65535.0 + 0.5, 0, 65535).astype(np.uint16)
    if dtype == np.uint8:
        return np.clip(f * 255.0 + 0.5, 0, 255).astype(np.uint8)
    return np.clip(f, 0, 1).astype(np.float32)


def orbit_eye(k, n=36, radius=3.5):
    th = 2.0 * math.pi * k / n
    return (radius * math.sin(th), radius * 0.3, radius * math.cos(th))


class Scene:
    """One volume (or several) + TFs + params; renders through the oracle and through the C ABI."""

    def __init__(self, vol, keys=CT_KEYS, W=128, H=128, eye=(1.2, 1.05, 3.1), tf_width=None, tf_domain=None, **param_kw):
        self.vol = np.ascontiguousarray(vol)
        self.nch = vol.shape[3] if vol.ndim == 4 else 1
        dims = vol.shape[2::-1]
        self.dtype = {np.dtype(np.uint8): abi.U8, np.dtype(np.uint16): abi.U16, np.dtype(np.float32): abi.F32}[vol.dtype]
        self.desc = rc.make_volume_desc(dims, channels=self.nch, dtype=self.dtype)
        self.W, self.H = W, H
        self.cam = orc.make_camera(eye, W=W, H=H)
        rate = param_kw.pop("sampling_rate", 2.0)
        self.step = float(np.float32(1.0) / (np.float32(max(dims)) * np.float32(rate)))   # volumeraycaster.cpp:312-322
        self.params = rc.default_params(self.step, **param_kw)
        self.eep = rc.default_eep()
        scale = self.desc.rwm_scale
        self.tf_width = tf_width or {abi.U8: 256, abi.U16: 4096, abi.F32: 4096}[self.dtype]   # transfuncpropertybase.cpp:155-165
        keys_per_ch = keys if isinstance(keys[0], list) else [keys] * self.nch
        self.domain = tuple(tf_domain) if tf_domain is not None else (0.0, scale)
        self.keys_per_ch = keys_per_ch
        self.tfs = [orc.make_tf(k, domain=self.domain) for k in keys_per_ch]
        self.tf_tex = [orc.tf1d_texture(t, self.tf_width) for t in self.tfs]
        self.preint = None

    def use_preint(self, mode="cpu"):
        res = min(self.tf_width, 1024)      # preintegrationtablemap.cpp:66-67
        if mode == "cpu":
            self.preint = [orc.preint_cpu(t, res, self.step) for t in self.tfs]
        elif mode == "approx":
            self.preint = [orc.preint_cpu(t, res, self.step, True) for t in self.tfs]
        else:
            self.preint = [orc.preint_gpu(t, res, self.step) for t in self.tf_tex]
        for v in range(abi.MAX_VOLUMES):
            self.params.vol[v].classification = abi.CLASS_PREINT
        return self

    # ---- oracle ------------------------------------------------------------------------------------------------
    def oracle_eep(self):
        ov = orc.make_volume(self.vol, self.desc)
        boxes = orc.bbox_boxes(ov, self.eep)
        return orc.entry_exit(boxes, orc.texture_to_world(self.desc), self.eep, self.cam, self.W, self.H)

    def oracle_render(self, eep_images=None):
        ov = orc.make_volume(self.vol, self.desc)
        e, ed, x, xd = eep_images if eep_images is not None else self.oracle_eep()
        cls = []
        for c in range(self.nch):
            cls.append(orc.make_classifier(self.tf_tex[c], self.domain, preint=self.preint[c] if self.preint else None))
        return orc.raycast([ov], cls, self.params, self.cam, self.W, self.H, e, ed, x, xd)

    # ---- CUDA (through the C ABI) ----------------------------------------------------------------------------------
    def upload(self, ctx):
        ctx.upload_volume(0, self.vol, desc=self.desc)
        for c in range(self.nch):
            # the product's own TransFunc1DKeys::getTexture (host mirror), not the oracle's texels (pinned equal in test_oracle_pins)
            ctx.set_tf1d(c, host.tf1d_texture(self.keys_per_ch[c], self.tf_width), self.domain)
            if self.preint:
                ctx.set_preint(c, self.preint[c])

    def cuda_render(self, ctx, upload=True):
        if upload:
            self.upload(ctx)
        return ctx.render_camera(self.params, self.eep, self.cam, self.W, self.H)


def parity_report(a, b):
    """Fraction of pixels with max channel |diff| <= 1/255 and the max abs diff (NaNs must coincide)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    bad_a, bad_b = ~np.isfinite(a), ~np.isfinite(b)
    both = bad_a & bad_b            # NaN/inf in both images counts as equal, in only one as a failing pixel
    d = np.abs(np.where(bad_a, 0.0, a) - np.where(bad_b, 0.0, b))
    d = np.where(both, 0.0, np.where(bad_a ^ bad_b, np.inf, d))
    per_px = d.reshape(d.shape[0], d.shape[1], -1).max(axis=2)
    finite = d[np.isfinite(d)]
    return float((per_px <= 1.0 / 255.0).mean()), float(finite.max() if finite.size else 0.0)


# ---- the product's sort-last merge kernels on numpy images (upload -> b200rc_composite / b200rc_finalize -> download) ----------
def product_composite(ctx, mode, front, back):
    """front / back: (colour (n, 4) float32, depth (n,) float32).  Returns dst = front OVER back from b200rc_composite."""
    import torch
    dev = torch.device("cuda", 0)
    fc, fd, bc, bd = (torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev) for a in (front[0], front[1], back[0], back[1]))
    dc, dd = torch.empty_like(fc), torch.empty_like(fd)
    ctx.composite(mode, fc.data_ptr(), fd.data_ptr(), bc.data_ptr(), bd.data_ptr(), dc.data_ptr(), dd.data_ptr(), fd.shape[0])
    ctx.synchronize()
    return dc.cpu().numpy(), dd.cpu().numpy()


def product_finalize(ctx, color, out_format, snap=True):
    """(n, 4) float32 -> b200rc_finalize -> (n, 4) uint16 (RGBA16) or float32."""
    import torch
    from voreen_b200 import abi as _abi
    dev = torch.device("cuda", 0)
    c = torch.from_numpy(np.ascontiguousarray(color, np.float32)).to(dev)
    n = c.shape[0]
    out = torch.empty((n, 4), dtype=torch.int16 if out_format == _abi.OUT_RGBA16 else torch.float32, device=dev)
    ctx.finalize(c.data_ptr(), 0, n, out_format, out.data_ptr(), snap=snap)
    ctx.synchronize()
    o = out.cpu().numpy()
    return o.view(np.uint16) if out_format == _abi.OUT_RGBA16 else o


def numpy_finalize(color, out_format, snap=True):
    """mod_raysetup.frag:65-74 snap + GL_RGBA16 store (clamp, round half up) in numpy."""
    from voreen_b200 import abi as _abi
    c = np.array(color, np.float32, copy=True)
    if snap:
        c[:, 3] = np.where(c[:, 3] >= np.float32(0.95), np.float32(1.0), c[:, 3])
    if out_format != _abi.OUT_RGBA16:
        return c
    v = np.where(c > 0, np.minimum(c, np.float32(1.0)), np.float32(0.0)).astype(np.float32)        # NaN -> 0 like to_unorm16
    return np.floor(v * np.float32(65535.0) + np.float32(0.5)).astype(np.uint16)
