"""ctypes view of the C++ host-side mirror (include/b200rc_host.h, voreen_b200/host/): tgt::Camera, TransFunc1DKeys
texture generation and the SingleVolumeRaycaster processor with the reference's property IDs.  Thin binding only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200rc_host.so")

SYMBOLS = [
    "b200host_camera", "b200host_tf1d_keys_texture", "b200host_tf1d_width_for", "b200host_volume_desc", "b200host_svrc_create",
    "b200host_svrc_destroy", "b200host_svrc_last_error", "b200host_svrc_context", "b200host_svrc_set_property",
    "b200host_svrc_get_sampling_step", "b200host_svrc_set_volume", "b200host_svrc_set_synthetic_volume",
    "b200host_svrc_set_transfer_function", "b200host_svrc_set_camera", "b200host_svrc_connect_output",
    "b200host_svrc_set_entry_exit", "b200host_svrc_is_ready", "b200host_svrc_process", "b200host_vvd_parse", "b200host_vvd_count", "b200host_optimized_proxy_geometry",
    "b200host_tf1d_gaussian_texture", "b200host_tf2d_primitives_texture",
]


class TFCurve(C.Structure):
    _fields_ = [("intensity", C.c_float), ("width", C.c_float), ("base_value", C.c_float), ("color_l", C.c_uint8 * 4), ("color_r", C.c_uint8 * 4),
                ("base_color_l", C.c_uint8 * 4), ("base_color_r", C.c_uint8 * 4), ("active", C.c_int32)]


def make_curves(curves):
    """[dict(intensity, width, base_value, color_l, color_r, base_color_l, base_color_r, active)] -> TFCurve array (TransFuncMappingCurve)."""
    arr = (TFCurve * max(1, len(curves)))()
    for i, c in enumerate(curves):
        arr[i].intensity, arr[i].width, arr[i].base_value = c["intensity"], c["width"], c.get("base_value", 0.0)
        arr[i].color_l[:] = list(c["color_l"]); arr[i].color_r[:] = list(c.get("color_r", c["color_l"]))
        arr[i].base_color_l[:] = list(c.get("base_color_l", c["color_l"])); arr[i].base_color_r[:] = list(c.get("base_color_r", c.get("color_r", c["color_l"])))
        arr[i].active = 1 if c.get("active", True) else 0
    return arr


def tf1d_gaussian_texture(curves, width, threshold=(0.0, 1.0), gamma=1.0, alpha_mode=1):
    """TransFunc1DGaussian::getTexture: RGBA8 texels of a Gaussian-curve 1D transfer function."""
    arr = make_curves(curves)
    out = np.zeros((width, 4), np.uint8)
    thr = (C.c_float * 2)(*threshold)
    rc = load_library().b200host_tf1d_gaussian_texture(arr, len(curves), thr, gamma, alpha_mode, width, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise abi.B200RCError(f"b200host_tf1d_gaussian_texture: status {rc}")
    return out


class TF2DPrimitive(C.Structure):
    _fields_ = [("type", C.c_int32), ("fuzziness", C.c_float), ("position", (C.c_float * 2) * 4), ("color", (C.c_uint8 * 4) * 4)]


PRIMITIVE = {"triangle": 0, "quad": 1, "banana": 2}


def make_primitives(prims, cls=None):
    """[dict(type="triangle" | "quad" | "banana", points=[(x, y)] (3 or 4), colors=[rgba8] (one per point, or a single one), fuzziness=1.0)]
    -> TF2DPrimitive array (TransFuncTriangle / TransFuncQuad / TransFuncBanana)."""
    cls = cls or TF2DPrimitive
    arr = (cls * max(1, len(prims)))()
    for i, p in enumerate(prims):
        arr[i].type = PRIMITIVE[p["type"]]
        arr[i].fuzziness = p.get("fuzziness", 1.0)
        n = 3 if p["type"] == "triangle" else 4
        cols = p["colors"] if isinstance(p["colors"][0], (list, tuple)) else [p["colors"]] * n
        for k in range(n):
            arr[i].position[k][0], arr[i].position[k][1] = p["points"][k]
            for c in range(4):
                arr[i].color[k][c] = int(cols[k][c])
    return arr


def tf2d_primitives_texture(prims, width=256, height=256, threshold_x=(0.0, 1.0), threshold_y=(0.0, 1.0), gamma=(1.0, 1.0), alpha_mode=1):
    """TransFunc2DPrimitives::getTexture: (float32 (H, W, 4) texture the reference uploads, its uint8 quantisation for Context.set_tf2d)."""
    arr = make_primitives(prims)
    f = np.zeros((height, width, 4), np.float32)
    u = np.zeros((height, width, 4), np.uint8)
    tx, ty, g = (C.c_float * 2)(*threshold_x), (C.c_float * 2)(*threshold_y), (C.c_float * 2)(*gamma)
    rc = load_library().b200host_tf2d_primitives_texture(arr, len(prims), width, height, tx, ty, g, alpha_mode, f.ctypes.data_as(C.POINTER(C.c_float)),
                                                         u.ctypes.data_as(C.c_void_p))
    if rc:
        raise abi.B200RCError(f"b200host_tf2d_primitives_texture: status {rc}")
    return f, u


class TFKey(C.Structure):
    _fields_ = [("intensity", C.c_float), ("color_l", C.c_uint8 * 4), ("color_r", C.c_uint8 * 4), ("split", C.c_int32)]


class TF1D(C.Structure):
    _fields_ = [("keys", C.POINTER(TFKey)), ("n_keys", C.c_int32), ("domain", C.c_float * 2), ("threshold", C.c_float * 2),
                ("gamma", C.c_float), ("alpha_mode", C.c_int32)]


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    abi.load_library()      # libb200rc.so first (same directory, $ORIGIN rpath)
    if not os.path.exists(LIB_PATH):
        raise abi.B200RCError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
    lib = C.CDLL(LIB_PATH)
    pf, vp, i32, f32 = C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_float
    lib.b200host_camera.argtypes = [pf, pf, pf, f32, f32, f32, f32, i32, i32, C.POINTER(abi.Camera)]
    lib.b200host_tf1d_keys_texture.argtypes = [C.POINTER(TFKey), i32, pf, f32, i32, i32, vp]
    lib.b200host_tf1d_gaussian_texture.argtypes = [C.POINTER(TFCurve), i32, pf, f32, i32, i32, vp]
    lib.b200host_tf2d_primitives_texture.argtypes = [C.POINTER(TF2DPrimitive), i32, i32, i32, pf, pf, pf, i32, pf, vp]
    lib.b200host_tf1d_width_for.argtypes = [i32, C.POINTER(i32)]
    lib.b200host_volume_desc.argtypes = [C.POINTER(C.c_int32 * 3), i32, C.c_char_p, pf, pf, pf, C.POINTER(abi.VolumeDesc)]
    lib.b200host_svrc_create.argtypes = [i32, C.POINTER(vp)]
    lib.b200host_vvd_parse.argtypes = [C.c_char_p, i32, C.POINTER(abi.VolumeDesc), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.b200host_vvd_count.argtypes = [C.c_char_p]
    lib.b200host_optimized_proxy_geometry.argtypes = [vp, i32, C.POINTER(abi.VolumeDesc), C.POINTER(abi.EepParams), C.POINTER(TF1D), i32, pf, i32,
                                                      C.POINTER(C.c_int)]
    lib.b200host_svrc_destroy.argtypes = [vp]
    lib.b200host_svrc_destroy.restype = None
    lib.b200host_svrc_last_error.argtypes = [vp]
    lib.b200host_svrc_last_error.restype = C.c_char_p
    lib.b200host_svrc_context.argtypes = [vp]
    lib.b200host_svrc_context.restype = vp
    lib.b200host_svrc_set_property.argtypes = [vp, C.c_char_p, C.c_char_p]
    lib.b200host_svrc_get_sampling_step.argtypes = [vp, pf]
    lib.b200host_svrc_set_volume.argtypes = [vp, vp, C.POINTER(abi.VolumeDesc)]
    lib.b200host_svrc_set_synthetic_volume.argtypes = [vp, C.POINTER(abi.VolumeDesc), C.c_uint64]
    lib.b200host_svrc_set_transfer_function.argtypes = [vp, i32, C.POINTER(TFKey), i32, f32, f32, pf, f32]
    lib.b200host_svrc_set_camera.argtypes = [vp, pf, pf, pf, f32, f32, f32]
    lib.b200host_svrc_connect_output.argtypes = [vp, i32, i32]
    lib.b200host_svrc_set_entry_exit.argtypes = [vp, i32, i32, pf, pf, pf, pf]
    lib.b200host_svrc_is_ready.argtypes = [vp]
    lib.b200host_svrc_process.argtypes = [vp, i32, i32, C.POINTER(vp * 3), pf]
    _lib = lib
    return lib


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def make_camera(eye, focus=(0, 0, 0), up=(0, 1, 0), fovy=45.0, near=0.1, far=50.0, W=256, H=256, frustum_ratio=1.0):
    """tgt::Camera -> b200rc_camera (view, projection for the viewport, position, near/far)."""
    cam = abi.Camera()
    rc = load_library().b200host_camera(_f3(eye), _f3(focus), _f3(up), fovy, frustum_ratio, near, far, W, H, C.byref(cam))
    if rc:
        raise abi.B200RCError(f"b200host_camera: status {rc}")
    return cam


def optimized_proxy_geometry(ctx, slot, desc, eep, tfs, max_boxes=1 << 16):
    """OptimizedProxyGeometry "minboundingbox" / "visiblebricks" for the volume in `slot` of `ctx` (a raycaster.Context):
    tfs = [(keys, domain, threshold, gamma, alpha_mode)] per channel -> (n, 6) float32 cubes in texture space for eep.boxes."""
    arr = (TF1D * max(1, len(tfs)))()
    keep = []
    for i, (keys, domain, threshold, gamma, alpha_mode) in enumerate(tfs):
        k = make_keys(keys); keep.append(k)
        arr[i].keys = k; arr[i].n_keys = len(keys)
        arr[i].domain[:] = list(domain); arr[i].threshold[:] = list(threshold)
        arr[i].gamma = gamma; arr[i].alpha_mode = alpha_mode
    out = np.zeros((max_boxes, 6), np.float32)
    n = C.c_int(0)
    rc = load_library().b200host_optimized_proxy_geometry(ctx.h, slot, C.byref(desc), C.byref(eep), arr, len(tfs),
                                                          out.ctypes.data_as(C.POINTER(C.c_float)), max_boxes, C.byref(n))
    if rc:
        raise abi.B200RCError(f"b200host_optimized_proxy_geometry: status {rc}: {ctx.last_error() if hasattr(ctx, 'last_error') else ''}")
    return out[:min(n.value, max_boxes)].copy()


def parse_vvd(path, index=0):
    """VvdVolumeReader::read for one volume of a .vvd file -> (b200rc_volume_desc, path of the raw file)."""
    desc = abi.VolumeDesc()
    raw = C.create_string_buffer(4096)
    err = C.create_string_buffer(512)
    rc = load_library().b200host_vvd_parse(os.fsencode(path), index, C.byref(desc), raw, len(raw), err, len(err))
    if rc:
        raise abi.B200RCError(f"b200host_vvd_parse: {err.value.decode(errors='replace') or rc}")
    return desc, os.fsdecode(raw.value)


def vvd_count(path):
    return load_library().b200host_vvd_count(os.fsencode(path))


def make_keys(keys):
    """[(intensity, (r,g,b,a)[, (r,g,b,a) right])] sorted by intensity -> TFKey array."""
    arr = (TFKey * len(keys))()
    for i, k in enumerate(keys):
        arr[i].intensity = k[0]
        arr[i].color_l[:] = list(k[1])
        arr[i].color_r[:] = list(k[2] if len(k) > 2 else k[1])
        arr[i].split = 1 if len(k) > 2 else 0
    return arr


def tf1d_texture(keys, width, threshold=(0.0, 1.0), gamma=1.0, alpha_mode=1):
    """TransFunc1DKeys::getTexture: RGBA8 texels of a key-based 1D transfer function."""
    arr = make_keys(keys)
    out = np.zeros((width, 4), np.uint8)
    thr = (C.c_float * 2)(*threshold)
    rc = load_library().b200host_tf1d_keys_texture(arr, len(keys), thr, gamma, alpha_mode, width, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise abi.B200RCError(f"b200host_tf1d_keys_texture: status {rc}")
    return out


class SingleVolumeRaycaster:
    """Mirror of the reference processor: same property IDs / option strings, same ports."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.b200host_svrc_create(device, C.byref(h))
        if rc:
            raise abi.B200RCError(f"SingleVolumeRaycaster: initialize failed with status {rc} (3 = no sm_100 device; no CPU fallback)")
        self.h = h
        self.out_format = abi.OUT_RGBA32F
        self.connected = [True, False, False]
        self._keep = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200host_svrc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise abi.B200RCError(f"{what}: status {rc}: {self.lib.b200host_svrc_last_error(self.h).decode()}")

    def set_property(self, pid, value):
        if isinstance(value, bool):
            value = "true" if value else "false"
        elif isinstance(value, (tuple, list, np.ndarray)):
            value = " ".join(repr(float(v)) for v in value)
        self._check(self.lib.b200host_svrc_set_property(self.h, pid.encode(), str(value).encode()), f"setProperty {pid}")
        if pid == "outputFormat":
            self.out_format = abi.OUT_RGBA16 if value == "rgba16" else abi.OUT_RGBA32F

    def set_volume(self, data, desc):
        data = np.ascontiguousarray(data)
        self._keep["vol"] = data
        self._check(self.lib.b200host_svrc_set_volume(self.h, data.ctypes.data_as(C.c_void_p), C.byref(desc)), "setVolume")

    def set_synthetic_volume(self, desc, seed=1234):
        self._check(self.lib.b200host_svrc_set_synthetic_volume(self.h, C.byref(desc), seed), "setSyntheticVolume")

    def set_transfer_function(self, channel, keys, domain, threshold=(0.0, 1.0), gamma=1.0):
        arr = make_keys(keys)
        thr = (C.c_float * 2)(*threshold)
        self._check(self.lib.b200host_svrc_set_transfer_function(self.h, channel, arr, len(keys), domain[0], domain[1], thr, gamma), "setTransferFunction")

    def set_camera(self, eye, focus=(0, 0, 0), up=(0, 1, 0), fovy=45.0, near=0.1, far=50.0):
        self._check(self.lib.b200host_svrc_set_camera(self.h, _f3(eye), _f3(focus), _f3(up), fovy, near, far), "setCamera")

    def connect_output(self, port, connected=True):
        self.connected[port] = bool(connected)
        self._check(self.lib.b200host_svrc_connect_output(self.h, port, 1 if connected else 0), "connectOutput")

    def set_entry_exit(self, entry, entry_depth, exitp, exit_depth):
        arrs = [np.ascontiguousarray(a, np.float32) for a in (entry, entry_depth, exitp, exit_depth)]
        self._keep["eep"] = arrs
        H, W = arrs[0].shape[:2]
        pf = C.POINTER(C.c_float)
        self._check(self.lib.b200host_svrc_set_entry_exit(self.h, W, H, *[a.ctypes.data_as(pf) for a in arrs]), "setEntryExit")

    def is_ready(self):
        return bool(self.lib.b200host_svrc_is_ready(self.h))

    def process(self, W, H):
        dt = np.uint16 if self.out_format == abi.OUT_RGBA16 else np.float32
        outs = [np.empty((H, W, 4), dt) if c else None for c in self.connected]
        depth = np.empty((H, W), np.float32)
        ptrs = (C.c_void_p * 3)(*[o.ctypes.data if o is not None else None for o in outs])
        self._check(self.lib.b200host_svrc_process(self.h, W, H, C.byref(ptrs), depth.ctypes.data_as(C.POINTER(C.c_float))), "process")
        return outs, depth
