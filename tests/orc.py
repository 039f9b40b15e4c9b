"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from voreen_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORC_DIR, "liborc.so")
PROBE = os.path.join(ORC_DIR, "_ref", "tgt_probe")


class TFKey(C.Structure):
    _fields_ = [("intensity", C.c_float), ("color_l", C.c_uint8 * 4), ("color_r", C.c_uint8 * 4), ("split", C.c_int32)]


class TF1D(C.Structure):
    _fields_ = [("keys", C.POINTER(TFKey)), ("n_keys", C.c_int32), ("threshold", C.c_float * 2), ("gamma", C.c_float),
                ("alpha_mode", C.c_int32), ("domain", C.c_float * 2)]


class Volume(C.Structure):
    _fields_ = [("data", C.c_void_p), ("desc", abi.VolumeDesc)]


class Classifier(C.Structure):
    _fields_ = [("tf_rgba8", C.c_void_p), ("tf_width", C.c_int32), ("domain", C.c_float * 2),
                ("preint", C.POINTER(C.c_float)), ("preint_res", C.c_int32),
                ("tf2d_rgba8", C.c_void_p), ("tf2d_w", C.c_int32), ("tf2d_h", C.c_int32), ("domain_y", C.c_float * 2),
                ("seg_domains", C.POINTER(C.c_float))]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORC_DIR, "liborc.so"])
    if os.path.isdir("/root/reference/voreen/ext/tgt"):
        subprocess.check_call(["make", "-s", "-C", ORC_DIR, "ref"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        pf = C.POINTER(C.c_float)
        _lib.orc_mat4_mul.argtypes = [pf, pf, pf]
        _lib.orc_mat4_invert.argtypes = [pf, pf]
        _lib.orc_mat4_invert.restype = C.c_int
        _lib.orc_look_at.argtypes = [pf, pf, pf, pf]
        _lib.orc_perspective.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, pf]
        _lib.orc_camera_projection.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, pf]
        _lib.orc_texture_to_world.argtypes = [C.POINTER(abi.VolumeDesc), pf]
        _lib.orc_tf1d_texture.argtypes = [C.POINTER(TF1D), C.c_int, C.c_void_p]
        _lib.orc_preint_cpu.argtypes = [C.POINTER(TF1D), C.c_int, C.c_float, C.c_int, pf]
        _lib.orc_preint_gpu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, pf]
        _lib.orc_preint_classify.argtypes = [pf, C.c_int, C.c_float, C.c_float, pf]
        _lib.orc_region_minmax.argtypes = [C.POINTER(Volume), C.c_int, C.POINTER(C.c_int32 * 3), pf]
        _lib.orc_proxy_boxes.argtypes = [C.POINTER(Volume), C.POINTER(TF1D), C.c_int, C.POINTER(abi.EepParams), pf, C.c_int]
        _lib.orc_rescale_rgba16.argtypes = [C.c_void_p, pf, C.c_int, C.c_int, C.c_void_p, pf, C.c_int, C.c_int]
        _lib.orc_rescale_rgba16.restype = None
        _lib.orc_proxy_boxes.restype = C.c_int
        _lib.orc_entry_exit.argtypes = [pf, C.c_int, pf, C.POINTER(abi.EepParams), C.POINTER(abi.Camera), C.c_int, C.c_int, pf, pf, pf, pf]
        _lib.orc_raycast.argtypes = [C.POINTER(Volume), C.POINTER(Classifier), C.POINTER(abi.Params), C.POINTER(abi.Camera), C.c_int, C.c_int,
                                     pf, pf, pf, pf, C.POINTER(pf * 3), pf, C.POINTER(abi.Stats)]
        _lib.orc_sample3d.argtypes = [C.POINTER(Volume), C.c_int, C.c_int, C.c_float, pf, pf]
        _lib.orc_unorm16.argtypes = [C.c_float]; _lib.orc_unorm16.restype = C.c_float
        _lib.orc_depth24.argtypes = [C.c_float]; _lib.orc_depth24.restype = C.c_float
        _lib.orc_composite.argtypes = [C.c_int, pf, pf, pf, pf, pf, pf, C.c_size_t]
        _lib.orc_image_diff_count.argtypes = [pf, pf, C.c_int, C.c_int, C.c_float, C.c_int]
        _lib.orc_image_diff_count.restype = C.c_int64
        _lib.orc_set_threads.argtypes = [C.c_int]
        _lib.orc_image_compositor.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.orc_image_compositor.restype = None
        _lib.orc_image_background.argtypes = [C.POINTER(abi.BackgroundParams), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib.orc_image_background.restype = None
        _lib.orc_octree_raycast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32 * 3), C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                            C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, pf, pf, C.c_int, C.c_int, pf, pf]
        _lib.orc_octree_raycast.restype = None
        _lib.orc_synth_volume.argtypes = [C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), C.c_uint64, C.c_void_p]
        _lib.orc_tf2d_primitives_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, pf, pf, pf, C.c_int, pf]
        _lib.orc_tf2d_primitives_texture.restype = None
    return _lib


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


# ---- convenience wrappers ------------------------------------------------------------------------------------------
def make_tf(keys, domain=(0.0, 1.0), threshold=(0.0, 1.0), gamma=1.0, alpha_mode=1):
    """keys: [(intensity, (r,g,b,a))] or [(intensity, colorL, colorR)] sorted by intensity."""
    arr = (TFKey * len(keys))()
    for i, k in enumerate(keys):
        arr[i].intensity = k[0]
        arr[i].color_l[:] = list(k[1])
        if len(k) > 2:
            arr[i].color_r[:] = list(k[2]); arr[i].split = 1
        else:
            arr[i].color_r[:] = list(k[1]); arr[i].split = 0
    tf = TF1D()
    tf.keys = arr
    tf._keep = arr
    tf.n_keys = len(keys)
    tf.threshold[:] = list(threshold)
    tf.gamma = gamma
    tf.alpha_mode = alpha_mode
    tf.domain[:] = list(domain)
    return tf


def tf1d_texture(tf, width):
    out = np.zeros((width, 4), np.uint8)
    lib().orc_tf1d_texture(C.byref(tf), width, out.ctypes.data_as(C.c_void_p))
    return out


def preint_cpu(tf, res, step, use_integral=False):
    out = np.zeros((res, res, 4), np.float32)
    lib().orc_preint_cpu(C.byref(tf), res, step, 1 if use_integral else 0, fp(out))
    return out


def preint_gpu(tf_tex, res, step):
    tf_tex = np.ascontiguousarray(tf_tex, np.uint8)
    out = np.zeros((res, res, 4), np.float32)
    lib().orc_preint_gpu(tf_tex.ctypes.data_as(C.c_void_p), tf_tex.shape[0], res, step, fp(out))
    return out


def look_at(eye, focus, up):
    out = np.zeros(16, np.float32)
    e, f, u = (np.asarray(v, np.float32) for v in (eye, focus, up))
    lib().orc_look_at(fp(e), fp(f), fp(u), fp(out))
    return out


def camera_projection(fovy_deg, frustum_ratio, near, far, window_ratio):
    out = np.zeros(16, np.float32)
    lib().orc_camera_projection(fovy_deg, frustum_ratio, near, far, window_ratio, fp(out))
    return out


def make_camera(eye, focus=(0, 0, 0), up=(0, 1, 0), fovy=45.0, near=0.1, far=50.0, W=256, H=256, frustum_ratio=1.0):
    cam = abi.Camera()
    cam.view[:] = look_at(eye, focus, up).tolist()
    cam.proj[:] = camera_projection(fovy, frustum_ratio, near, far, float(np.float32(W) / np.float32(H))).tolist()
    cam.position[:] = [float(np.float32(v)) for v in eye]
    cam.near_dist = near
    cam.far_dist = far
    return cam


def texture_to_world(desc):
    out = np.zeros(16, np.float32)
    lib().orc_texture_to_world(C.byref(desc), fp(out))
    return out


def make_volume(data, desc):
    v = Volume()
    data = np.ascontiguousarray(data)
    v._keep = data
    v.data = data.ctypes.data
    v.desc = desc
    return v


def bbox_boxes(vol, eep, tfs=None):
    out = np.zeros((1 << 16, 6), np.float32)
    if tfs:
        arr = (TF1D * len(tfs))(*tfs)
        n = lib().orc_proxy_boxes(C.byref(vol), arr, len(tfs), C.byref(eep), fp(out), out.shape[0])
    else:
        n = lib().orc_proxy_boxes(C.byref(vol), None, 0, C.byref(eep), fp(out), out.shape[0])
    return out[:n].copy()


def entry_exit(boxes, model, eep, cam, W, H):
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    model = np.ascontiguousarray(model, np.float32)
    e = np.zeros((H, W, 4), np.float32); x = np.zeros((H, W, 4), np.float32)
    ed = np.zeros((H, W), np.float32); xd = np.zeros((H, W), np.float32)
    lib().orc_entry_exit(fp(boxes), boxes.shape[0], fp(model), C.byref(eep), C.byref(cam), W, H, fp(e), fp(ed), fp(x), fp(xd))
    return e, ed, x, xd


def make_classifier(tf_tex=None, domain=(0.0, 1.0), preint=None, tf2d=None, domain_y=(0.0, 1.0), seg_domains=None):
    c = Classifier()
    c._keep = []
    if tf_tex is not None:
        t = np.ascontiguousarray(tf_tex, np.uint8); c._keep.append(t)
        c.tf_rgba8 = t.ctypes.data; c.tf_width = t.shape[0]
    if preint is not None:
        t = np.ascontiguousarray(preint, np.float32); c._keep.append(t)
        c.preint = fp(t); c.preint_res = t.shape[0]
    if tf2d is not None:
        t = np.ascontiguousarray(tf2d, np.uint8); c._keep.append(t)
        c.tf2d_rgba8 = t.ctypes.data; c.tf2d_h, c.tf2d_w = t.shape[0], t.shape[1]
    if seg_domains is not None:
        t = np.ascontiguousarray(seg_domains, np.float32); c._keep.append(t)
        c.seg_domains = fp(t)
    c.domain[:] = list(domain)
    c.domain_y[:] = list(domain_y)
    return c


def raycast(vols, classifiers, params, cam, W, H, entry, entry_depth, exitp, exit_depth):
    va = (Volume * len(vols))(*vols)
    ca = (Classifier * len(classifiers))(*classifiers)
    outs = [np.zeros((H, W, 4), np.float32) if params.output_connected[o] else None for o in range(3)]
    depth = np.zeros((H, W), np.float32)
    pf = C.POINTER(C.c_float)
    ptrs = (pf * 3)(*[fp(o) if o is not None else None for o in outs])
    st = abi.Stats()
    entry = np.ascontiguousarray(entry, np.float32); exitp = np.ascontiguousarray(exitp, np.float32)
    lib().orc_raycast(va, ca, C.byref(params), C.byref(cam), W, H, fp(entry), fp(np.ascontiguousarray(entry_depth, np.float32)),
                      fp(exitp), fp(np.ascontiguousarray(exit_depth, np.float32)), C.byref(ptrs), fp(depth), C.byref(st))
    return outs, depth, st


def image_diff_count(a, b, tol=0.02, radius=1):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return int(lib().orc_image_diff_count(fp(a), fp(b), a.shape[1], a.shape[0], tol, radius))


def synth_volume(dims, seed=1234, lo=(0, 0, 0), rdims=None):
    """CPU twin of Context.synth_volume (identical bytes): voxels [lo, lo+rdims) of the global `dims` volume."""
    rdims = tuple(rdims) if rdims is not None else tuple(dims)
    out = np.empty((rdims[2], rdims[1], rdims[0]), np.uint16)
    a = (C.c_int32 * 3)(*dims); b = (C.c_int32 * 3)(*lo); c = (C.c_int32 * 3)(*rdims)
    lib().orc_synth_volume(C.byref(a), C.byref(b), C.byref(c), seed, out.ctypes.data_as(C.c_void_p))
    return out


def rescale_rgba16(color, depth, dw, dh):
    """VolumeRaycaster::rescaleRendering on an (H, W, 4) uint16 image + (H, W) float32 depth."""
    color = np.ascontiguousarray(color, np.uint16); depth = np.ascontiguousarray(depth, np.float32)
    sh, sw = color.shape[:2]
    out = np.empty((dh, dw, 4), np.uint16); od = np.empty((dh, dw), np.float32)
    lib().orc_rescale_rgba16(color.ctypes.data, fp(depth), sw, sh, out.ctypes.data, fp(od), dw, dh)
    return out, od


def image_compositor(mode, color0, depth0, color1, depth1, weighting_factor=0.5, weight_first=1.0, weight_second=1.0, add_depth=False):
    """voreen::Compositor (compositor.frag) on (H, W, 4) uint16 / float32 images + (H, W) depths (or None)."""
    color0 = np.ascontiguousarray(color0); color1 = np.ascontiguousarray(color1, color0.dtype)
    fmt = abi.OUT_RGBA16 if color0.dtype == np.uint16 else abi.OUT_RGBA32F
    H, W = color0.shape[:2]
    out = np.empty_like(color0); od = np.empty((H, W), np.float32)
    d0 = np.ascontiguousarray(depth0, np.float32) if depth0 is not None else None
    d1 = np.ascontiguousarray(depth1, np.float32) if depth1 is not None else None
    lib().orc_image_compositor(abi.COMPOSITOR[mode], weighting_factor, weight_first, weight_second, 1 if add_depth else 0, color0.ctypes.data,
                               d0.ctypes.data if d0 is not None else None, color1.ctypes.data, d1.ctypes.data if d1 is not None else None, fmt,
                               out.ctypes.data, od.ctypes.data, W * H)
    return out, od


def image_background(W, H, color=None, depth=None, mode="monochrome", blend="alpha-blending", color1=(1, 1, 1, 1), color2=(0.392, 0.392, 0.392, 1), angle=60, dtype=np.uint16):
    bp = abi.BackgroundParams()
    bp.mode, bp.blend_mode, bp.angle_deg = abi.BACKGROUND[mode], abi.BACKGROUND_BLEND[blend], int(angle)
    bp.color1[:] = [float(v) for v in color1]; bp.color2[:] = [float(v) for v in color2]
    if color is not None:
        color = np.ascontiguousarray(color); dtype = color.dtype
    fmt = abi.OUT_RGBA16 if dtype == np.uint16 else abi.OUT_RGBA32F
    out = np.empty((H, W, 4), dtype); od = np.empty((H, W), np.float32)
    d = np.ascontiguousarray(depth, np.float32) if depth is not None else None
    lib().orc_image_background(C.byref(bp), color.ctypes.data if color is not None else None, d.ctypes.data if d is not None else None, fmt,
                               out.ctypes.data, od.ctypes.data, W, H)
    return out, od


def octree_raycast(avg, flags, pool, levels, brick_dim, dims, octree_dim, tf_tex, domain, rwm, step, node_level, linear, entry, exitp, W, H, compositing=0):
    """SingleOctreeRaycasterCL DVR over the dense arrays of oracle/octree_ref.pool_arrays."""
    avg = np.ascontiguousarray(avg, np.uint16); flags = np.ascontiguousarray(flags, np.uint8); pool = np.ascontiguousarray(pool, np.uint16)
    tf_tex = np.ascontiguousarray(tf_tex, np.uint8)
    out = np.zeros((H, W, 4), np.float32); fh = np.zeros((H, W), np.float32)
    d = (C.c_int32 * 3)(*dims)
    entry = np.ascontiguousarray(entry, np.float32); exitp = np.ascontiguousarray(exitp, np.float32)
    lib().orc_octree_raycast(avg.ctypes.data, flags.ctypes.data, pool.ctypes.data, levels, brick_dim, C.byref(d), octree_dim, tf_tex.ctypes.data, tf_tex.shape[0],
                             domain[0], domain[1], rwm[0], rwm[1], step, node_level, 1 if linear else 0, compositing, fp(entry), fp(exitp), W, H, fp(out), fp(fh))
    return out, fh


def tf2d_primitives_texture(prims, width=256, height=256, threshold_x=(0.0, 1.0), threshold_y=(0.0, 1.0), gamma=(1.0, 1.0), alpha_mode=1):
    """TransFunc2DPrimitives texture (float32 (H, W, 4)); prims as voreen_b200.host.make_primitives takes them."""
    from voreen_b200 import host

    class Prim(C.Structure):        # orc_tf2d_primitive (same layout as the product's POD, declared independently in vrc_oracle.h)
        _fields_ = [("type", C.c_int32), ("fuzziness", C.c_float), ("position", (C.c_float * 2) * 4), ("color", (C.c_uint8 * 4) * 4)]
    arr = host.make_primitives(prims, cls=Prim)
    out = np.zeros((height, width, 4), np.float32)
    tx, ty, g = (C.c_float * 2)(*threshold_x), (C.c_float * 2)(*threshold_y), (C.c_float * 2)(*gamma)
    lib().orc_tf2d_primitives_texture(C.cast(arr, C.c_void_p), len(prims), width, height, tx, ty, g, alpha_mode, fp(out))
    return out
