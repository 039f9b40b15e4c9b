/*
 * vrc_oracle.h -- CPU restatement of Voreen's direct-volume-ray-casting path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker / the CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" at image level.  The reference's rendered goldens live in the un-vendored
 * `voreen-testdata` submodule and the reference itself (GLSL behind an OpenGL driver, Qt, DevIL) cannot be built or
 * run in this container.  What IS pinned:
 *   - camera / matrix arithmetic against the reference's own tgt headers compiled in place (oracle/_ref/tgt_probe,
 *     see oracle/Makefile and tests/test_oracle_pins.py);
 *   - every numeric constant and formula (0.95, 200, +-0.02, Sobel weights, /16, Phong defaults, step formula,
 *     pre-integration recurrence, unorm16 formats), each cited below at the reference file:line it follows;
 *   - the reference's bundled scene (nucleon.vvd + rendering-raycasting.vws) parsed from /root/reference by the
 *     committed generator tests/golden/make_golden.py.
 * Texture filtering follows the OpenGL 4.x specification with float weights, which is what the reference's CI driver
 * (Mesa llvmpipe, GALLIVM_PERF=no_aos_sampling; /root/reference/.gitlab-ci.yml:27-30) implements.
 *
 * V/ = /root/reference/voreen/
 */
#ifndef VRC_ORACLE_H
#define VRC_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/b200rc.h"   /* POD parameter blocks only (no code is shared with the product) */

#ifdef __cplusplus
extern "C" {
#endif

/* TransFuncMappingKey (V/src/core/datastructures/transfunc/1d/1dkeys/utils/transfuncmappingkey.cpp:30-35) */
typedef struct orc_tf_key {
    float   intensity;
    uint8_t color_l[4];
    uint8_t color_r[4];     /* == color_l unless split */
    int32_t split;
} orc_tf_key;

/* TransFuncBase::AlphaMode (V/include/voreen/core/datastructures/transfunc/transfuncbase.h:165-169) */
enum { ORC_TF_ZERO_ALPHA = 0, ORC_TF_USE_ALPHA = 1, ORC_TF_ONE_ALPHA = 2 };

typedef struct orc_tf1d {
    const orc_tf_key* keys;     /* sorted by intensity */
    int32_t n_keys;
    float   threshold[2];
    float   gamma;
    int32_t alpha_mode;
    float   domain[2];
} orc_tf1d;

/* A volume as the GL path sees it. */
typedef struct orc_volume {
    const void* data;           /* x fastest, channels interleaved */
    b200rc_volume_desc desc;
} orc_volume;

/* per-channel / per-volume classification inputs handed to the marcher */
typedef struct orc_classifier {
    const uint8_t* tf_rgba8;    /* 1D TF texture, NEAREST (transfunc1d.h:55) */
    int32_t tf_width;
    float   domain[2];
    const float* preint;        /* res*res rgba32f, LINEAR, CLAMP_TO_EDGE (preintegrationtable.cpp:294-312) */
    int32_t preint_res;
    const uint8_t* tf2d_rgba8;  /* 2D TF, LINEAR */
    int32_t tf2d_w, tf2d_h;
    float   domain_y[2];
    const float* seg_domains;   /* SegmentationRaycaster (classifier 1): 256 x {lower, upper} per-segment domains; tf2d_* = the segmentation TF texture */
} orc_classifier;

int  orc_abi_version(void);
void orc_set_threads(int n);    /* OpenMP threads; 0 = default */

/* --- tgt math restated (V/ext/tgt/matrix.h:1196-1248,1327-1376; V/ext/tgt/camera.cpp:188-217) --- */
void orc_mat4_mul(const float a[16], const float b[16], float out[16]);
int  orc_mat4_invert(const float m[16], float out[16]);
void orc_look_at(const float eye[3], const float focus[3], const float up[3], float out[16]);
void orc_perspective(float fovy_rad, float aspect, float pnear, float pfar, float out[16]);
/* Camera::getProjectionMatrix(windowRatio) for PERSPECTIVE mode */
void orc_camera_projection(float fovy_deg, float frustum_ratio, float pnear, float pfar, float window_ratio, float out[16]);
/* VolumeBase coordinate systems (V/src/core/datastructures/volume/volumebase.cpp:158-231) */
void orc_texture_to_world(const b200rc_volume_desc* d, float out[16]);
void orc_world_to_physical(const b200rc_volume_desc* d, float out[16]);

/* --- transfer functions --- */
/* TransFunc1D::createTextureData + TransFuncBase::applyAlpha (transfunc1d.cpp:85-116, transfuncbase.cpp:177-210) */
void orc_tf1d_texture(const orc_tf1d* tf, int width, uint8_t* out_rgba8);
/* PreIntegrationTable::computeTable (preintegrationtable.cpp:67-204) */
void orc_preint_cpu(const orc_tf1d* tf, int res, float sampling_step, int use_integral, float* out_rgba32f);
/* preintegration.frag:35-87 rendered into GL_RGBA16 (preintegrationtable.cpp:48) */
void orc_preint_gpu(const uint8_t* tf_rgba8, int tf_width, int res, float sampling_step, float* out_rgba32f);
/* PreIntegrationTable::classify (preintegrationtable.cpp:256-279) */
void orc_preint_classify(const float* table, int res, float fs, float fe, float out[4]);

/* --- proxy geometry --- */
/* computeRegionStructure (optimizedproxygeometry.cpp:894-987): out = cells * channels * {min,max}, real-world units */
void orc_region_minmax(const orc_volume* vol, int cell_size, int32_t cells[3], float* out_minmax);
/* Proxy boxes in texture space for the chosen mode.  Returns number of boxes written (<= max_boxes), each 6 floats
 * llf.xyz urb.xyz.  BBOX: processCube (:420-448).  MINBBOX: computeMinCube (:1014-1116).  VISIBLE_BRICKS: every
 * non-empty cell as its own cube (the union rasterises identically to the merged boxes of :1139-1372). */
int orc_proxy_boxes(const orc_volume* vol, const orc_tf1d* tfs, int n_tfs, const b200rc_eep_params* eep,
                    float* out_boxes, int max_boxes);

/* --- entry / exit points (meshentryexitpoints.cpp:188-273, eep_*.frag) ---
 * boxes: proxy cubes in texture space; model = mesh transformation (textureToWorld).  Outputs float4 + depth. */
void orc_entry_exit(const float* boxes, int n_boxes, const float model[16], const b200rc_eep_params* eep,
                    const b200rc_camera* cam, int width, int height,
                    float* entry_rgba, float* entry_depth, float* exit_rgba, float* exit_depth);

/* --- the ray marcher (rc_singlevolume.frag:116-340 / rc_multivolume.frag:97-350) ---
 * entry/exit: for multi-volume, (1 + num_volumes) consecutive images [global, vol1..]; single: one image.
 * out_color[i]: float4 images (NULL to skip); out_depth: float.  stats may be NULL. */
void orc_raycast(const orc_volume* vols, const orc_classifier* cls, const b200rc_params* params,
                 const b200rc_camera* cam, int width, int height,
                 const float* entry_rgba, const float* entry_depth, const float* exit_rgba, const float* exit_depth,
                 float* out_color[B200RC_MAX_OUTPUTS], float* out_depth, b200rc_stats* stats);

/* single filtered fetch, exposed for unit tests (mod_sampler3d.frag:90-101 + GL 4.x 8.14 texture filtering) */
void orc_sample3d(const orc_volume* vol, int filter, int clamp, float border, const float pos[3], float out[4]);

/* GL_RGBA16 store + fetch round trip, GL_DEPTH_COMPONENT24 likewise */
float orc_unorm16(float v);
float orc_depth24(float v);

/* --- sort-last reference: dst = front over back for premultiplied colours --- */
void orc_composite(int mode, const float* front_rgba, const float* front_depth, const float* back_rgba,
                   const float* back_depth, float* dst_rgba, float* dst_depth, size_t n);

/* --- the reference's own image acceptance test (V/src/core/utils/regressiontest/filecomparators.cpp:482-632):
 * returns the number of pixels whose RGBA distance/2 exceeds `tolerance` after a radius-`radius` neighbour search. */
int64_t orc_image_diff_count(const float* a, const float* b, int width, int height, float tolerance, int radius);

/* --- synthetic "CT-like" benchmark volume: CPU twin of b200rc_synth_volume, identical bytes.
 * Writes voxels [lo, lo+rdims) of a global volume of `dims` voxels, x fastest. --- */
void orc_synth_volume(const int32_t dims[3], const int32_t lo[3], const int32_t rdims[3], uint64_t seed, uint16_t* out);

/* VolumeRaycaster::rescaleRendering (volumeraycaster.cpp:348-377) with copyimage.frag:37-55: GL_LINEAR / CLAMP_TO_EDGE copy of an
 * RGBA16-unorm colour image (as uint16) + DEPTH24 depth to another size. */
void orc_rescale_rgba16(const uint16_t* src, const float* src_depth, int sw, int sh, uint16_t* dst, float* dst_depth, int dw, int dh);

/* image-space post chain (compositor.frag:41-140; background.frag:35-57 + background.cpp:262-300) */
void orc_image_compositor(int mode, float weighting_factor, float weight_first, float weight_second, int add_depth, const void* color0, const float* depth0,
                          const void* color1, const float* depth1, int format, void* dst_color, float* dst_depth, size_t n);
void orc_image_background(const b200rc_background_params* bg, const void* src_color, const float* src_depth, int format, void* dst_color, float* dst_depth, int W, int H);

/* TransFuncMappingCurve + TransFunc1DGaussian::getTexture (transfunc1dgaussian.cpp:122-178, transfuncmappingcurve.cpp:257-288) */
typedef struct orc_tf_curve {
    float intensity, width, base_value;
    uint8_t color_l[4], color_r[4], base_color_l[4], base_color_r[4];
    int32_t active;
} orc_tf_curve;
void orc_tf1d_gaussian_texture(const orc_tf_curve* curves, int n, const float threshold[2], float gamma, int alpha_mode, int width, uint8_t* out_rgba8);

/* TransFunc2DPrimitives: triangle lists of the primitives (transfuncprimitive.cpp:98-189,240-352,405-490), the help texture a GL_MAX
 * rasterisation of them gives (transfunc2dprimitives.cpp:165-250,313-333: ortho [0,1]^2, texel centres, inclusive edges, barycentric
 * colours in double precision -- OpenGL itself leaves sub-pixel snapping to the implementation, so there is no reference bit pattern
 * to pin this against: PARITY UNPINNED for the rasteriser), then TransFunc2D::createTextureData (transfunc2d.cpp:75-124, TF_FLOAT) and
 * TransFuncBase::applyAlpha (transfuncbase.cpp:177-210). */
typedef struct orc_tf2d_primitive {
    int32_t type;               /* 0 triangle, 1 quad, 2 banana */
    float fuzziness;
    float position[4][2];
    uint8_t color[4][4];
} orc_tf2d_primitive;
void orc_tf2d_primitives_texture(const orc_tf2d_primitive* prims, int n, int width, int height, const float thr_x[2], const float thr_y[2],
                                 const float gamma[2], int alpha_mode, float* out_rgba32f);

/* SingleOctreeRaycasterCL DVR traversal over dense per-level node arrays (oracle/octree_ref.py builds them) */
void orc_octree_raycast(const uint16_t* avg, const uint8_t* flags, const uint16_t* bricks, int levels, int brick_dim, const int32_t dims[3], int octree_dim,
                        const uint8_t* tf_rgba8, int tf_width, float dom_lo, float dom_hi, float rwm_scale, float rwm_offset,
                        float step, int node_level, int filter_linear, int compositing /* B200RC_COMP_DVR / _MIP / _MOP */, const float* entry_rgba, const float* exit_rgba, int W, int H, float* out_rgba, float* out_firsthit);

#ifdef __cplusplus
}
#endif
#endif
