"""Thin Python binding of the C ABI (include/b200rc.h) used by tests, bench.py and the sort-last driver.

Everything here is a direct call into libb200rc.so; no rendering arithmetic lives in Python.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_NP_DTYPE = {abi.U8: np.uint8, abi.U16: np.uint16, abi.F32: np.float32}
_DTYPE_OF = {np.dtype(np.uint8): abi.U8, np.dtype(np.uint16): abi.U16, np.dtype(np.float32): abi.F32}


def make_volume_desc(dims, channels=1, dtype=abi.U16, spacing=None, offset=None, phys_to_world=None, rwm_scale=None,
                     rwm_offset=0.0, shard_lo=(0, 0, 0), shard_hi=(0, 0, 0), shard_apron=0, data_on_device=False):
    """VolumeBase meta data as the reference passes it to the shader (V/src/core/utils/glsl.cpp:38-83).

    Defaults follow the bundled data: cube [-1,1]^3 * dims/max(dims) (nucleon.vvd:8-12) and the denormalising
    RealWorldMapping the VVD reader attaches (V/modules/core/io/vvdvolumereader.cpp:102-119)."""
    d = abi.VolumeDesc()
    dims = [int(v) for v in dims]
    d.dims[:] = dims
    d.channels = channels
    d.dtype = dtype
    m = float(max(dims))
    if spacing is None:
        spacing = [np.float32(2.0) / np.float32(m)] * 3
    if offset is None:
        offset = [np.float32(-(dims[a] - 1) * 0.5) * np.float32(spacing[a]) for a in range(3)]
    d.spacing[:] = [float(v) for v in spacing]
    d.offset[:] = [float(v) for v in offset]
    p2w = np.eye(4, dtype=np.float32) if phys_to_world is None else np.asarray(phys_to_world, np.float32)
    d.phys_to_world[:] = p2w.reshape(-1).tolist()
    if rwm_scale is None:
        rwm_scale = {abi.U8: 255.0, abi.U16: 65535.0, abi.F32: 1.0}[dtype]
    d.rwm_scale = rwm_scale
    d.rwm_offset = rwm_offset
    d.shard_lo[:] = list(shard_lo)
    d.shard_hi[:] = list(shard_hi)
    d.shard_apron = shard_apron
    d.data_on_device = 1 if data_on_device else 0
    return d


def default_params(step, **kw):
    """SingleVolumeRaycaster defaults (V/src/core/processors/volumeraycaster.cpp:42-157, volumerenderer.cpp:41-53)."""
    p = abi.Params()
    p.num_volumes = 1
    p.multi_volume = 0
    for v in range(abi.MAX_VOLUMES):
        p.vol[v].classification = abi.CLASS_TF
        p.vol[v].shading = abi.SHADE["phong"]
        p.vol[v].filter = abi.FILTER_LINEAR
        p.vol[v].clamp = abi.CLAMP_TO_EDGE
        p.vol[v].border_intensity = 0.0
    p.sampling_step = step
    p.mv_opacity_correction_factor = 0.0
    p.gradient = abi.GRAD["central-differences"]
    p.compositing[:] = [abi.COMP["dvr"]] * 3
    p.output_connected[:] = [1, 0, 0]
    p.iso_value = 0.5
    p.mida_gamma[:] = [0.0, 0.0, 0.0]
    p.enable_channel_shift = 0
    p.light.position[:] = [2.3, 1.5, 1.5, 1.0]
    p.light.ambient[:] = [0.4, 0.4, 0.4]
    p.light.diffuse[:] = [0.8, 0.8, 0.8]
    p.light.specular[:] = [0.6, 0.6, 0.6]
    p.light.attenuation[:] = [1.0, 0.0, 0.0]
    p.light.shininess = 60.0
    p.light.apply_attenuation = 0
    p.sampler = abi.SAMPLER_EXACT
    p.empty_space_skipping = 1
    p.out_format = abi.OUT_RGBA32F
    p.quantize_depth24 = 1
    p.collect_stats = 0
    for k, v in kw.items():
        if k in ("classification", "shading", "filter", "clamp", "border_intensity"):
            for i in range(abi.MAX_VOLUMES):
                setattr(p.vol[i], k, v)
        elif k in ("compositing", "output_connected", "mida_gamma"):
            getattr(p, k)[:] = list(v)
        else:
            setattr(p, k, v)
    return p


def default_eep(**kw):
    """MeshEntryExitPoints defaults (meshentryexitpoints.cpp:47-94) over a bounding-box proxy geometry."""
    e = abi.EepParams()
    e.proxy_mode = abi.PROXY_BBOX
    e.proxy_cell_size = 16
    e.proxy_threshold = 1.0
    e.inside_technique = abi.INSIDE_GPU
    e.jitter = 0
    e.jitter_step_length = 0.005
    e.world_coords = 0
    e.enable_clipping = 0
    e.n_boxes = 0
    for k, v in kw.items():
        if k == "boxes":
            arr = np.ascontiguousarray(v, np.float32).reshape(-1, 6)
            e._boxes_keepalive = arr
            e.boxes = arr.ctypes.data_as(C.POINTER(C.c_float))
            e.n_boxes = arr.shape[0]
        elif k in ("clip_llf", "clip_urb"):
            getattr(e, k)[:] = list(v)
        else:
            setattr(e, k, v)
    return e


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class Context:
    """One GPU, one stream: owner of the resident volume(s), transfer functions and frame buffers."""

    def __init__(self, device=0):
        self.lib = abi.load_library()
        h = C.c_void_p()
        rc = self.lib.b200rc_create(device, C.byref(h))
        if rc != abi.OK:
            raise abi.B200RCError(f"b200rc_create failed with status {rc} "
                                  "(3 = no sm_100 device; this path has no CPU fallback)")
        self.h = h
        self.volume_desc = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200rc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != abi.OK:
            raise abi.B200RCError(f"{what}: status {rc}: {self.lib.b200rc_last_error(self.h).decode()}")

    # -- plumbing ----------------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        self._check(self.lib.b200rc_set_stream(self.h, C.c_void_p(cuda_stream_handle)), "set_stream")

    def synchronize(self):
        self._check(self.lib.b200rc_synchronize(self.h), "synchronize")

    def stats(self):
        s = abi.Stats()
        self._check(self.lib.b200rc_get_stats(self.h, C.byref(s)), "get_stats")
        return s

    # -- volume --------------------------------------------------------------------------------------------------
    def upload_volume(self, slot, data, desc=None, **desc_kw):
        data = np.ascontiguousarray(data)
        if desc is None:
            ch = data.shape[3] if data.ndim == 4 else 1
            desc = make_volume_desc(data.shape[2::-1], channels=ch, dtype=_DTYPE_OF[data.dtype], **desc_kw)
        self._check(self.lib.b200rc_upload_volume(self.h, slot, data.ctypes.data_as(C.c_void_p), C.byref(desc)), "upload_volume")
        self.volume_desc[slot] = desc
        return desc

    def upload_volume_device(self, slot, device_ptr, desc):
        desc.data_on_device = 1
        self._check(self.lib.b200rc_upload_volume(self.h, slot, C.c_void_p(device_ptr), C.byref(desc)), "upload_volume")
        self.volume_desc[slot] = desc
        return desc

    def upload_volume_slabs(self, slot, data, desc, slab_planes=8):
        """Streaming upload (b200rc_upload_volume_begin / _slab / _end): `data` is the resident range, handed over in z slabs."""
        data = np.ascontiguousarray(data)
        self._check(self.lib.b200rc_upload_volume_begin(self.h, slot, C.byref(desc)), "upload_volume_begin")
        for z0 in range(0, data.shape[0], slab_planes):
            part = np.ascontiguousarray(data[z0:z0 + slab_planes])
            self._check(self.lib.b200rc_upload_volume_slab(self.h, slot, z0, part.shape[0], part.ctypes.data_as(C.c_void_p)), "upload_volume_slab")
        self._check(self.lib.b200rc_upload_volume_end(self.h, slot), "upload_volume_end")
        self.volume_desc[slot] = desc
        return desc

    def load_raw(self, slot, path, desc, header_skip=0, slab_planes=0):
        """Raw file -> HBM through two pinned slab buffers (b200rc_load_raw_file); reads only the resident range of desc."""
        self._check(self.lib.b200rc_load_raw_file(self.h, slot, os.fsencode(path), C.byref(desc), header_skip, slab_planes), "load_raw_file")
        self.volume_desc[slot] = desc
        return desc

    def load_vvd(self, slot, path, index=0, slab_planes=0, **shard):
        """A Voreen .vvd volume (host.parse_vvd) streamed into the brick pool; shard_lo / shard_hi / shard_apron select a shard."""
        from . import host
        desc, raw = host.parse_vvd(path, index)
        for k, val in shard.items():
            if k in ("shard_lo", "shard_hi"):
                getattr(desc, k)[:] = list(val)
            else:
                setattr(desc, k, val)
        return self.load_raw(slot, raw, desc, 0, slab_planes)

    def synth_volume(self, slot, desc, seed=1234):
        self._check(self.lib.b200rc_synth_volume(self.h, slot, C.byref(desc), seed), "synth_volume")
        self.volume_desc[slot] = desc
        return desc

    def download_volume(self, slot):
        d = self.volume_desc[slot]
        lo = [max(0, d.shard_lo[a] - d.shard_apron) for a in range(3)]
        hi = [min(d.dims[a], d.shard_hi[a] + d.shard_apron) for a in range(3)]
        dims = list(d.dims) if d.shard_hi[0] <= d.shard_lo[0] else [hi[a] - lo[a] for a in range(3)]
        shape = (dims[2], dims[1], dims[0]) + ((d.channels,) if d.channels > 1 else ())
        out = np.empty(shape, _NP_DTYPE[d.dtype])
        self._check(self.lib.b200rc_download_volume(self.h, slot, out.ctypes.data_as(C.c_void_p), out.nbytes), "download_volume")
        return out

    def region_minmax(self, slot, cell_size):
        cells = (C.c_int32 * 3)()
        self._check(self.lib.b200rc_region_minmax(self.h, slot, cell_size, C.byref(cells), None, 0), "region_minmax")
        nch = self.volume_desc[slot].channels
        out = np.empty((cells[2], cells[1], cells[0], nch, 2), np.float32)
        self._check(self.lib.b200rc_region_minmax(self.h, slot, cell_size, C.byref(cells), _fp(out), out.size), "region_minmax")
        return out

    def download_minmax(self, slot):
        """The 8^3 macro-cell min/max grid (normalised values) and the device time of its build in ms."""
        cells = (C.c_int32 * 3)(); ms = C.c_double(0.0)
        self._check(self.lib.b200rc_download_minmax(self.h, slot, C.byref(cells), None, 0, C.byref(ms)), "download_minmax")
        out = np.empty((cells[2], cells[1], cells[0], self.volume_desc[slot].channels, 2), np.float32)
        self._check(self.lib.b200rc_download_minmax(self.h, slot, C.byref(cells), _fp(out), out.size, None), "download_minmax")
        return out, float(ms.value)

    # -- transfer functions ----------------------------------------------------------------------------------------
    def set_tf1d(self, chan, rgba8, domain):
        rgba8 = np.ascontiguousarray(rgba8, np.uint8).reshape(-1, 4)
        self._check(self.lib.b200rc_set_tf1d(self.h, chan, rgba8.ctypes.data_as(C.c_void_p), rgba8.shape[0], domain[0], domain[1]), "set_tf1d")

    def set_preint(self, chan, table):
        table = np.ascontiguousarray(table, np.float32)
        res = table.shape[0]
        self._check(self.lib.b200rc_set_preint(self.h, chan, _fp(table), res), "set_preint")

    def build_preint(self, chan, res, step, mode):
        self._check(self.lib.b200rc_build_preint(self.h, chan, res, step, mode), "build_preint")

    def download_preint(self, chan, res):
        out = np.empty((res, res, 4), np.float32)
        self._check(self.lib.b200rc_download_preint(self.h, chan, _fp(out), res), "download_preint")
        return out

    def set_tf2d(self, chan, rgba8, domain_x, domain_y):
        rgba8 = np.ascontiguousarray(rgba8, np.uint8)
        h, w = rgba8.shape[:2]
        self._check(self.lib.b200rc_set_tf2d(self.h, chan, rgba8.ctypes.data_as(C.c_void_p), w, h, domain_x[0], domain_x[1], domain_y[0], domain_y[1]), "set_tf2d")

    def set_segmentation_tf(self, rgba8, domains):
        """SegmentationRaycaster: (3 * 256, width, 4) uint8 texture (three rows per segment) + (256, 2) float32 per-segment domains."""
        rgba8 = np.ascontiguousarray(rgba8, np.uint8); domains = np.ascontiguousarray(domains, np.float32)
        h, w = rgba8.shape[:2]
        self._check(self.lib.b200rc_set_segmentation_tf(self.h, rgba8.ctypes.data_as(C.c_void_p), w, h, domains.ctypes.data_as(C.POINTER(C.c_float))), "set_segmentation_tf")

    # -- entry / exit ---------------------------------------------------------------------------------------------
    def entry_exit(self, eep, cam, W, H):
        e = np.empty((H, W, 4), np.float32); x = np.empty((H, W, 4), np.float32)
        ed = np.empty((H, W), np.float32); xd = np.empty((H, W), np.float32)
        self._check(self.lib.b200rc_entry_exit(self.h, C.byref(eep), C.byref(cam), W, H, _fp(e), _fp(ed), _fp(x), _fp(xd)), "entry_exit")
        return e, ed, x, xd

    # -- render -----------------------------------------------------------------------------------------------------
    @staticmethod
    def _alloc_outputs(params, W, H):
        outs = []
        for o in range(3):
            if params.output_connected[o]:
                outs.append(np.empty((H, W, 4), np.uint16 if params.out_format == abi.OUT_RGBA16 else np.float32))
            else:
                outs.append(None)
        return outs, np.empty((H, W), np.float32)

    def render(self, params, cam, W, H, entry, entry_depth, exitp, exit_depth):
        outs, depth = self._alloc_outputs(params, W, H)
        ptrs = (C.c_void_p * 3)(*[o.ctypes.data if o is not None else None for o in outs])
        entry = np.ascontiguousarray(entry, np.float32); exitp = np.ascontiguousarray(exitp, np.float32)
        entry_depth = np.ascontiguousarray(entry_depth, np.float32); exit_depth = np.ascontiguousarray(exit_depth, np.float32)
        self._check(self.lib.b200rc_render(self.h, C.byref(params), C.byref(cam), W, H, _fp(entry), _fp(entry_depth), _fp(exitp), _fp(exit_depth),
                                           C.byref(ptrs), _fp(depth)), "render")
        return outs, depth

    def render_camera(self, params, eep, cam, W, H, outs=None, depth=None):
        if outs is None:
            outs, depth = self._alloc_outputs(params, W, H)
        ptrs = (C.c_void_p * 3)(*[o.ctypes.data if o is not None else None for o in outs])
        self._check(self.lib.b200rc_render_camera(self.h, C.byref(params), C.byref(eep), C.byref(cam), W, H, C.byref(ptrs), _fp(depth)), "render_camera")
        return outs, depth

    def render_camera_ptrs(self, params, eep, cam, W, H, out_ptrs, depth_ptr, device=False):
        """Raw-pointer variant (pinned host tensors or device tensors owned by the caller, e.g. torch)."""
        ptrs = (C.c_void_p * 3)(*[C.c_void_p(p) if p else None for p in out_ptrs])
        if device:
            rc = self.lib.b200rc_render_camera_device(self.h, C.byref(params), C.byref(eep), C.byref(cam), W, H, C.byref(ptrs), C.c_void_p(depth_ptr) if depth_ptr else None)
        else:
            rc = self.lib.b200rc_render_camera(self.h, C.byref(params), C.byref(eep), C.byref(cam), W, H, C.byref(ptrs),
                                               C.cast(C.c_void_p(depth_ptr), C.POINTER(C.c_float)) if depth_ptr else None)
        self._check(rc, "render_camera")

    # -- sort-last ---------------------------------------------------------------------------------------------------
    def composite(self, mode, front, front_depth, back, back_depth, dst, dst_depth, n):
        self._check(self.lib.b200rc_composite(self.h, mode, front, front_depth, back, back_depth, dst, dst_depth, n), "composite")

    def rescale(self, color, depth, dw, dh):
        """VolumeRaycaster::rescaleRendering: (H, W, 4) uint16 / float32 colour + (H, W) depth -> (dh, dw) with GL_LINEAR filtering."""
        color = np.ascontiguousarray(color)
        fmt = abi.OUT_RGBA16 if color.dtype == np.uint16 else abi.OUT_RGBA32F
        sh, sw = color.shape[:2]
        out = np.empty((dh, dw, 4), color.dtype)
        od = np.empty((dh, dw), np.float32) if depth is not None else None
        d = np.ascontiguousarray(depth, np.float32) if depth is not None else None
        self._check(self.lib.b200rc_rescale(self.h, color.ctypes.data_as(C.c_void_p), _fp(d) if d is not None else None, sw, sh, fmt,
                                            out.ctypes.data_as(C.c_void_p), _fp(od) if od is not None else None, dw, dh, 0), "rescale")
        return out, od

    # -- octree / multi-resolution brick pool ---------------------------------------------------------------------------
    def octree_build(self, slot, brick_dim, homogeneity_threshold=-1.0, half_sample="mean"):
        fn = {"mean": 0, "max": 1, "min": 2}[half_sample]
        self._check(self.lib.b200rc_octree_build(self.h, slot, brick_dim, homogeneity_threshold, fn), "octree_build")
        info = abi.OctreeInfo()
        self._check(self.lib.b200rc_octree_get_info(self.h, slot, C.byref(info)), "octree_get_info")
        return info

    def octree_level(self, slot, level, brick_dim, bricks=True):
        """(avg, mn, mx, flags) as (G, G, G) arrays of level `level` (0 = root) and, optionally, the bricks (G^3, B, B, B)."""
        G = 1 << level
        avg = np.empty((G, G, G), np.uint16); mn = np.empty((G, G, G), np.uint16); mx = np.empty((G, G, G), np.uint16); fl = np.empty((G, G, G), np.uint8)
        br = np.empty((G ** 3, brick_dim, brick_dim, brick_dim), np.uint16) if bricks else None
        self._check(self.lib.b200rc_octree_download(self.h, slot, level, avg.ctypes.data, mn.ctypes.data, mx.ctypes.data, fl.ctypes.data,
                                                    br.ctypes.data if br is not None else None), "octree_download")
        return avg, mn, mx, fl, br

    def octree_render(self, slot, step, node_level, eep, cam, W, H, linear=True, compositing="dvr"):
        rp = abi.OctreeRenderParams(); rp.sampling_step = step; rp.node_level = node_level; rp.filter = abi.FILTER_LINEAR if linear else abi.FILTER_NEAREST
        rp.compositing = abi.COMP[compositing] if isinstance(compositing, str) else int(compositing)
        out = np.empty((H, W, 4), np.float32); fh = np.empty((H, W), np.float32)
        self._check(self.lib.b200rc_octree_render(self.h, slot, C.byref(rp), C.byref(eep), C.byref(cam), W, H, _fp(out), _fp(fh)), "octree_render")
        return out, fh

    # -- image-space post chain (host numpy images) -------------------------------------------------------------------
    def image_compositor(self, mode, color0, depth0, color1, depth1, weighting_factor=0.5, weight_first=1.0, weight_second=1.0, add_depth=False):
        """voreen::Compositor on two (H, W, 4) uint16 / float32 images (+ (H, W) float32 depths or None)."""
        color0 = np.ascontiguousarray(color0); color1 = np.ascontiguousarray(color1, color0.dtype)
        fmt = abi.OUT_RGBA16 if color0.dtype == np.uint16 else abi.OUT_RGBA32F
        H, W = color0.shape[:2]
        out = np.empty_like(color0); od = np.empty((H, W), np.float32)
        d0 = np.ascontiguousarray(depth0, np.float32) if depth0 is not None else None
        d1 = np.ascontiguousarray(depth1, np.float32) if depth1 is not None else None
        self._check(self.lib.b200rc_image_compositor(self.h, abi.COMPOSITOR[mode] if isinstance(mode, str) else mode, weighting_factor, weight_first, weight_second,
                                                     1 if add_depth else 0, color0.ctypes.data, d0.ctypes.data if d0 is not None else None, color1.ctypes.data,
                                                     d1.ctypes.data if d1 is not None else None, fmt, out.ctypes.data, od.ctypes.data, W, H, 0), "image_compositor")
        return out, od

    def image_background(self, W, H, color=None, depth=None, mode="monochrome", blend="alpha-blending", color1=(1, 1, 1, 1), color2=(0.392, 0.392, 0.392, 1), angle=60,
                         dtype=np.uint16):
        """voreen::Background under an (H, W, 4) image (or alone when color is None)."""
        bp = abi.BackgroundParams()
        bp.mode, bp.blend_mode, bp.angle_deg = abi.BACKGROUND[mode], abi.BACKGROUND_BLEND[blend], int(angle)
        bp.color1[:] = [float(v) for v in color1]; bp.color2[:] = [float(v) for v in color2]
        if color is not None:
            color = np.ascontiguousarray(color); dtype = color.dtype
        fmt = abi.OUT_RGBA16 if dtype == np.uint16 else abi.OUT_RGBA32F
        out = np.empty((H, W, 4), dtype); od = np.empty((H, W), np.float32)
        d = np.ascontiguousarray(depth, np.float32) if depth is not None else None
        self._check(self.lib.b200rc_image_background(self.h, C.byref(bp), color.ctypes.data if color is not None else None, d.ctypes.data if d is not None else None, fmt,
                                                     out.ctypes.data, od.ctypes.data, W, H, 0), "image_background")
        return out, od

    def composite_keyed(self, fc, fk, fd, bc, bk, bd, dc, dk, dd, n):
        self._check(self.lib.b200rc_composite_keyed(self.h, fc, fk, fd, bc, bk, bd, dc, dk, dd, n), "composite_keyed")

    def finalize(self, rgba, depth, n, out_format, out, snap=True):
        """snap: early ray termination was active for this output (alpha >= 0.95 -> 1.0); False for mip / mop / mida."""
        self._check(self.lib.b200rc_finalize(self.h, rgba, depth, n, out_format, 1 if snap else 0, out), "finalize")

    def ert_plan(self, alphas, world, front_mask, color, depth, n, prefix, flags, count=False):
        c = C.c_uint64(0)
        self._check(self.lib.b200rc_ert_plan(self.h, alphas, world, front_mask, color, depth, n, prefix, flags, C.byref(c) if count else None), "ert_plan")
        return int(c.value)

    def ert_plan_planes(self, planes, front_mask, color, depth, n, prefix, flags, count=False):
        """b200rc_ert_plan over one alpha-plane pointer per rank (e.g. peer-mapped key planes): no all-gather."""
        c = C.c_uint64(0)
        arr = (C.c_void_p * len(planes))(*[C.c_void_p(int(p)) if p else None for p in planes])
        self._check(self.lib.b200rc_ert_plan_planes(self.h, arr, len(planes), front_mask, color, depth, n, prefix, flags, C.byref(c) if count else None), "ert_plan_planes")
        return int(c.value)

    def remarch_device(self, params, cam, W, H, prefix, flags, out_ptrs, depth_ptr):
        ptrs = (C.c_void_p * 3)(*[C.c_void_p(p) if p else None for p in out_ptrs])
        self._check(self.lib.b200rc_remarch_device(self.h, C.byref(params), C.byref(cam), W, H, prefix, flags, C.byref(ptrs), C.c_void_p(depth_ptr) if depth_ptr else None), "remarch_device")
