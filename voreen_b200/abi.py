"""ctypes view of the C ABI in include/b200rc.h (the drop-in boundary).

Loads voreen_b200/libb200rc.so -- hand-written sm_100a CUDA.  There is no CPU fallback: if the shared library is
missing the import of the product path fails loudly (B200RCError), it never silently degrades.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_VOLUMES, MAX_CHANNELS, MAX_OUTPUTS = 4, 4, 3

# enums (include/b200rc.h)
OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_NOT_READY, ERR_OOM = range(6)
U8, U16, F32 = 1, 2, 3
GRAD = {"none": 0, "forward-differences": 1, "central-differences": 2, "sobel": 3, "filtered": 4}
CLASS_TF, CLASS_PREINT, CLASS_TF2D = 0, 1, 2
PI_CPU, PI_APPROX, PI_GPU = 0, 1, 2
SHADE = {"none": 0, "phong-diffuse": 1, "phong-specular": 2, "phong-diffuse-ambient": 3, "phong-diffuse-specular": 4,
         "phong": 5, "toon": 6, "cook-torrance": 7, "oren-nayar": 8, "ward": 9}
COMP = {"dvr": 0, "mip": 1, "mop": 2, "mida": 3, "iso": 4, "fhp": 5, "fhn": 6}
FILTER_NEAREST, FILTER_LINEAR = 0, 1
CLAMP_TO_EDGE, CLAMP_TO_BORDER = 0, 1
SAMPLER_EXACT, SAMPLER_TEX = 0, 1
OUT_RGBA32F, OUT_RGBA16 = 0, 1
INSIDE_NONE, INSIDE_GPU = 0, 1
ABI_VERSION = 4          # include/b200rc.h B200RC_ABI_VERSION
RAYCASTER_SINGLE, RAYCASTER_SIMPLE, RAYCASTER_RGB, RAYCASTER_SEGMENTATION = 0, 1, 2, 3
PROXY_BBOX, PROXY_MINBBOX, PROXY_VISIBLE_BRICKS, PROXY_OUTER_FACES, PROXY_VOLUME_OCTREE = 0, 1, 2, 3, 4
COMPOSITE_ERT = 0x100


class VolumeDesc(C.Structure):
    _fields_ = [("dims", C.c_int32 * 3), ("channels", C.c_int32), ("dtype", C.c_int32),
                ("spacing", C.c_float * 3), ("offset", C.c_float * 3), ("phys_to_world", C.c_float * 16),
                ("rwm_scale", C.c_float), ("rwm_offset", C.c_float),
                ("shard_lo", C.c_int32 * 3), ("shard_hi", C.c_int32 * 3), ("shard_apron", C.c_int32),
                ("data_on_device", C.c_int32)]


class Light(C.Structure):
    _fields_ = [("position", C.c_float * 4), ("ambient", C.c_float * 3), ("diffuse", C.c_float * 3),
                ("specular", C.c_float * 3), ("attenuation", C.c_float * 3), ("shininess", C.c_float),
                ("apply_attenuation", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("position", C.c_float * 3),
                ("near_dist", C.c_float), ("far_dist", C.c_float)]


class VolumeParams(C.Structure):
    _fields_ = [("classification", C.c_int32), ("shading", C.c_int32), ("filter", C.c_int32), ("clamp", C.c_int32),
                ("border_intensity", C.c_float)]


class Params(C.Structure):
    _fields_ = [("num_volumes", C.c_int32), ("multi_volume", C.c_int32), ("vol", VolumeParams * MAX_VOLUMES),
                ("sampling_step", C.c_float), ("mv_opacity_correction_factor", C.c_float), ("gradient", C.c_int32),
                ("compositing", C.c_int32 * MAX_OUTPUTS), ("output_connected", C.c_int32 * MAX_OUTPUTS),
                ("iso_value", C.c_float), ("mida_gamma", C.c_float * MAX_OUTPUTS),
                ("enable_channel_shift", C.c_int32), ("channel_shift", (C.c_float * 3) * MAX_CHANNELS),
                ("light", Light),
                ("sampler", C.c_int32), ("empty_space_skipping", C.c_int32), ("out_format", C.c_int32),
                ("quantize_depth24", C.c_int32), ("collect_stats", C.c_int32), ("partial_image", C.c_int32),
                ("image_band_modulus", C.c_int32), ("image_band_remainder", C.c_int32), ("raycaster", C.c_int32), ("rgb_color_modulation", C.c_int32),
                ("apply_segmentation", C.c_int32), ("segmentation_bit_depth", C.c_int32)]


class EepParams(C.Structure):
    _fields_ = [("proxy_mode", C.c_int32), ("proxy_cell_size", C.c_int32), ("proxy_threshold", C.c_float),
                ("inside_technique", C.c_int32), ("jitter", C.c_int32), ("jitter_step_length", C.c_float),
                ("world_coords", C.c_int32), ("enable_clipping", C.c_int32),
                ("clip_llf", C.c_float * 3), ("clip_urb", C.c_float * 3),
                ("boxes", C.POINTER(C.c_float)), ("n_boxes", C.c_int32)]


class OctreeInfo(C.Structure):
    _fields_ = [("octree_dim", C.c_int32), ("brick_dim", C.c_int32), ("num_levels", C.c_int32), ("actual_depth", C.c_int32),
                ("num_nodes", C.c_uint64), ("num_bricks", C.c_uint64), ("build_ms", C.c_double)]


class OctreeRenderParams(C.Structure):
    _fields_ = [("sampling_step", C.c_float), ("node_level", C.c_int32), ("filter", C.c_int32), ("compositing", C.c_int32)]


class BackgroundParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("blend_mode", C.c_int32), ("color1", C.c_float * 4), ("color2", C.c_float * 4), ("angle_deg", C.c_int32)]


COMPOSITOR = {"depth-test": 0, "alpha-compositing": 1, "alpha-blending": 2, "alpha-blending-b-over-a": 3, "weighted-average": 4, "take-first": 5,
              "take-second": 6, "take-second-if-ready": 7, "first-has-priority": 8, "second-has-priority": 9, "maximum-alpha": 10, "difference": 11,
              "difference-alpha-mul": 12, "add": 13}       # compositor.cpp:47-60 option keys
BACKGROUND = {"monochrome": 0, "gradient": 1}
BACKGROUND_BLEND = {"alpha-blending": 0, "add": 1}


class Stats(C.Structure):
    _fields_ = [("last_render_ms", C.c_double), ("last_march_ms", C.c_double), ("last_eep_ms", C.c_double),
                ("rays_cast", C.c_uint64), ("samples_classified", C.c_uint64), ("samples_shaded", C.c_uint64),
                ("samples_nominal", C.c_uint64), ("cells_touched", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("samples_skipped", C.c_uint64)]


class B200RCError(RuntimeError):
    pass


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200RC_LIBRARY") or os.path.join(_HERE, "libb200rc.so")     # override: A/B builds of the same ABI (developer knob)

# every symbol include/b200rc.h declares
SYMBOLS = [
    "b200rc_abi_version", "b200rc_device_count", "b200rc_create", "b200rc_destroy", "b200rc_last_error",
    "b200rc_set_stream", "b200rc_synchronize", "b200rc_get_stats", "b200rc_upload_volume", "b200rc_release_volume",
    "b200rc_synth_volume", "b200rc_download_volume", "b200rc_region_minmax", "b200rc_download_minmax", "b200rc_set_tf1d", "b200rc_set_preint",
    "b200rc_build_preint", "b200rc_download_preint", "b200rc_set_tf2d", "b200rc_entry_exit", "b200rc_render",
    "b200rc_render_camera", "b200rc_render_camera_device", "b200rc_composite", "b200rc_finalize", "b200rc_ert_plan", "b200rc_ert_plan_planes",
    "b200rc_remarch_device", "b200rc_upload_volume_begin", "b200rc_upload_volume_slab", "b200rc_upload_volume_end", "b200rc_load_raw_file", "b200rc_rescale", "b200rc_composite_keyed",
    "b200rc_host_alloc", "b200rc_host_free", "b200rc_image_compositor", "b200rc_image_background", "b200rc_set_segmentation_tf",
    "b200rc_octree_build", "b200rc_octree_get_info", "b200rc_octree_download", "b200rc_octree_render",
]

_lib = None


def load_library():
    """dlopen the CUDA library; raises B200RCError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200RCError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(hand-written sm_100a CUDA; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, f32, u64, sz = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_size_t
    pf = C.POINTER(C.c_float)
    lib.b200rc_abi_version.restype = i32
    if lib.b200rc_abi_version() != ABI_VERSION:
        raise B200RCError(f"{LIB_PATH} has ABI version {lib.b200rc_abi_version()}, this binding expects {ABI_VERSION}: rebuild (g.build())")
    lib.b200rc_device_count.restype = i32
    lib.b200rc_create.argtypes = [i32, C.POINTER(vp)]
    lib.b200rc_destroy.argtypes = [vp]
    lib.b200rc_last_error.argtypes = [vp]
    lib.b200rc_last_error.restype = C.c_char_p
    lib.b200rc_set_stream.argtypes = [vp, vp]
    lib.b200rc_synchronize.argtypes = [vp]
    lib.b200rc_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.b200rc_upload_volume.argtypes = [vp, i32, vp, C.POINTER(VolumeDesc)]
    lib.b200rc_release_volume.argtypes = [vp, i32]
    lib.b200rc_upload_volume_begin.argtypes = [vp, i32, C.POINTER(VolumeDesc)]
    lib.b200rc_upload_volume_slab.argtypes = [vp, i32, i32, i32, vp]
    lib.b200rc_upload_volume_end.argtypes = [vp, i32]
    lib.b200rc_load_raw_file.argtypes = [vp, i32, C.c_char_p, C.POINTER(VolumeDesc), u64, i32]
    lib.b200rc_synth_volume.argtypes = [vp, i32, C.POINTER(VolumeDesc), u64]
    lib.b200rc_download_volume.argtypes = [vp, i32, vp, sz]
    lib.b200rc_region_minmax.argtypes = [vp, i32, i32, C.POINTER(C.c_int32 * 3), pf, sz]
    lib.b200rc_download_minmax.argtypes = [vp, i32, C.POINTER(C.c_int32 * 3), pf, sz, C.POINTER(C.c_double)]
    lib.b200rc_set_tf1d.argtypes = [vp, i32, vp, i32, f32, f32]
    lib.b200rc_set_preint.argtypes = [vp, i32, pf, i32]
    lib.b200rc_build_preint.argtypes = [vp, i32, i32, f32, i32]
    lib.b200rc_download_preint.argtypes = [vp, i32, pf, i32]
    lib.b200rc_set_tf2d.argtypes = [vp, i32, vp, i32, i32, f32, f32, f32, f32]
    lib.b200rc_set_segmentation_tf.argtypes = [vp, vp, i32, i32, pf]
    lib.b200rc_entry_exit.argtypes = [vp, C.POINTER(EepParams), C.POINTER(Camera), i32, i32, pf, pf, pf, pf]
    lib.b200rc_render.argtypes = [vp, C.POINTER(Params), C.POINTER(Camera), i32, i32, pf, pf, pf, pf, C.POINTER(vp * 3), pf]
    lib.b200rc_render_camera.argtypes = [vp, C.POINTER(Params), C.POINTER(EepParams), C.POINTER(Camera), i32, i32, C.POINTER(vp * 3), pf]
    lib.b200rc_render_camera_device.argtypes = [vp, C.POINTER(Params), C.POINTER(EepParams), C.POINTER(Camera), i32, i32, C.POINTER(vp * 3), vp]
    lib.b200rc_composite.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, sz]
    lib.b200rc_finalize.argtypes = [vp, vp, vp, sz, i32, i32, vp]
    lib.b200rc_composite_keyed.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz]
    lib.b200rc_rescale.argtypes = [vp, vp, pf, i32, i32, i32, vp, pf, i32, i32, i32]
    lib.b200rc_ert_plan.argtypes = [vp, vp, i32, C.c_uint32, vp, vp, sz, vp, vp, C.POINTER(C.c_uint64)]
    lib.b200rc_ert_plan_planes.argtypes = [vp, C.POINTER(C.c_void_p), i32, C.c_uint32, vp, vp, sz, vp, vp, C.POINTER(C.c_uint64)]
    lib.b200rc_remarch_device.argtypes = [vp, C.POINTER(Params), C.POINTER(Camera), i32, i32, vp, vp, C.POINTER(vp * 3), vp]
    lib.b200rc_image_compositor.argtypes = [vp, i32, f32, f32, f32, i32, vp, vp, vp, vp, i32, vp, vp, i32, i32, i32]
    lib.b200rc_image_background.argtypes = [vp, C.POINTER(BackgroundParams), vp, vp, i32, vp, vp, i32, i32, i32]
    lib.b200rc_octree_build.argtypes = [vp, i32, i32, f32, i32]
    lib.b200rc_octree_get_info.argtypes = [vp, i32, C.POINTER(OctreeInfo)]
    lib.b200rc_octree_download.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    lib.b200rc_octree_render.argtypes = [vp, i32, C.POINTER(OctreeRenderParams), C.POINTER(EepParams), C.POINTER(Camera), i32, i32, pf, pf]
    lib.b200rc_host_alloc.argtypes = [sz, C.POINTER(vp)]
    lib.b200rc_host_free.argtypes = [vp]
    for s in SYMBOLS:
        fn = getattr(lib, s)
        if s not in ("b200rc_last_error",):
            fn.restype = i32
    _lib = lib
    return lib
