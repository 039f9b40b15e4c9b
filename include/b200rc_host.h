/*
 * b200rc_host.h -- flat C view of the C++ host-side mirror (voreen_b200/host/) of the reference's processor
 * interface for the ray-casting path: tgt::Camera, TransFunc1DKeys, VolumeRaycaster / SingleVolumeRaycaster
 * properties (identical IDs and option strings), MeshEntryExitPoints over an OptimizedProxyGeometry cube.
 *
 * Inside Voreen none of this is needed -- the drop-in processor (custommodules/b200raycast/) takes camera, transfer
 * function and properties from Voreen itself and calls include/b200rc.h directly.  Outside Voreen (tests, bench.py,
 * standalone tools) this library plays Voreen's role, so that parity tests read like the reference's own regression
 * workspaces: set properties by ID, attach a volume, process(), compare images.
 *
 * All functions return 0 on success (B200RC_OK) unless stated.  V/ = /root/reference/voreen/
 */
#ifndef B200RC_HOST_H
#define B200RC_HOST_H

#include "b200rc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* TransFuncMappingKey (V/src/core/datastructures/transfunc/1d/1dkeys/utils/transfuncmappingkey.cpp:30-35) */
typedef struct b200host_tf_key {
    float   intensity;          /* [0,1] position in the TF domain */
    uint8_t color_l[4];         /* RGBA left of the key */
    uint8_t color_r[4];         /* RGBA right of the key (== color_l unless split) */
    int32_t split;
} b200host_tf_key;

/* ---- tgt::Camera (V/ext/tgt/camera.cpp:133-217, V/ext/tgt/matrix.h:1196-1248) ------------------------------ */
/* view = createLookAt(position, focus, up); proj = Camera::getProjectionMatrix(width/height) (PERSPECTIVE, fovy clamped
 * to [6,175] degrees).  Fills every field of `out`. */
int b200host_camera(const float position[3], const float focus[3], const float up[3], float fovy_deg, float frustum_ratio,
                    float near_dist, float far_dist, int width, int height, b200rc_camera* out);

/* ---- TransFunc1DKeys::getTexture (V/src/core/datastructures/transfunc/1d/transfunc1d.cpp:85-116,
 *      1dkeys/transfunc1dkeys.cpp:119-154, transfuncbase.cpp:177-210): keys -> RGBA8 texels.  alpha_mode: 0 zero, 1 use, 2 one. */
int b200host_tf1d_keys_texture(const b200host_tf_key* keys, int n_keys, const float threshold[2], float gamma, int alpha_mode,
                               int width, uint8_t* out_rgba8);
/* ---- TransFunc1DGaussian::getTexture ("1D Gaussian Color Map", transFuncGaussian1..4): the same texture generation with
 *      TransFunc1DGaussian::getMappingForValueUByte (V/src/core/datastructures/transfunc/1d/1dgaussian/transfunc1dgaussian.cpp:122-178)
 *      over TransFuncMappingCurve::getColorAt / getOpacityAt (.../1dgaussian/utils/transfuncmappingcurve.cpp:257-288): every active
 *      curve contributes exp(-(x - intensity)^2 / (2 width)) between its base and peak colours; colours are mixed weighted by
 *      opacity and normalised by min(sum of opacities, 1), the alpha is the maximum over the curves; every conversion truncates to
 *      uint8 as the reference does. */
typedef struct b200host_tf_curve {
    float   intensity, width, base_value;   /* peak position [0,1], variance-like width, opacity far from the peak */
    uint8_t color_l[4], color_r[4];         /* peak colour (+ peak alpha) left / right of the peak */
    uint8_t base_color_l[4], base_color_r[4];   /* colour far from the peak (alpha ignored) */
    int32_t active;
} b200host_tf_curve;
int b200host_tf1d_gaussian_texture(const b200host_tf_curve* curves, int n_curves, const float threshold[2], float gamma, int alpha_mode,
                                   int width, uint8_t* out_rgba8);
/* ---- TransFunc2DPrimitives (V/src/core/datastructures/transfunc/2d/2dprimitives/transfunc2dprimitives.cpp:150-250,313-333,
 *      utils/transfuncprimitive.cpp:98-189 triangle, :240-352 quad, :405-490 banana; TransFunc2D::createTextureData
 *      transfunc2d.cpp:75-124; TransFuncBase::applyAlpha transfuncbase.cpp:177-210).  The reference paints the primitives with
 *      OpenGL into a float RGBA help texture (orthographic [0,1]^2, per-vertex colours, glBlendEquation(GL_MAX), clear colour 0)
 *      and resamples that into the transfer-function texture (thresholds, gamma, alpha mode).  Here the same triangle lists are
 *      rasterised in software: a texel receives a fragment when its centre ((x + 0.5) / W, (y + 0.5) / H) lies inside or on the
 *      edge of a non-degenerate triangle, colours are interpolated barycentrically (double precision), fragments combine with max.
 *      OpenGL leaves the sub-pixel snapping and the tie-breaking rule to the implementation, so this is the GL result up to what
 *      implementations differ in among themselves.  type: 0 triangle (3 points), 1 quad (4 points: the reference's corner order),
 *      2 banana (4 points: left end, upper middle, lower middle, right end; 20 steps).
 *      out_rgba32f (W x H x 4, may be NULL) is the texture the reference uploads (TF_FLOAT); out_rgba8 (may be NULL) its
 *      round-to-nearest uint8 quantisation for b200rc_set_tf2d. */
typedef struct b200host_tf2d_primitive {
    int32_t type;
    float   fuzziness;              /* TransFuncPrimitive::fuzziness_ (1 = no fade-out border) */
    float   position[4][2];         /* control points, [0,1]^2: x = intensity, y = gradient magnitude */
    uint8_t color[4][4];
} b200host_tf2d_primitive;
int b200host_tf2d_primitives_texture(const b200host_tf2d_primitive* prims, int n_prims, int width, int height,
                                     const float threshold_x[2], const float threshold_y[2], const float gamma[2], int alpha_mode,
                                     float* out_rgba32f, uint8_t* out_rgba8);
/* TransFuncPropertyBase texture width for a volume (transfuncpropertybase.cpp:155-165, transfunc1dkeysproperty.cpp:94-97) */
int b200host_tf1d_width_for(int dtype, int* out_width);

/* ---- VolumeBase meta data with the reference's defaults (denormalising RealWorldMapping for integer formats,
 *      V/include/voreen/core/datastructures/meta/realworldmappingmetadata.h:72-88); spacing/offset/phys_to_world may be NULL
 *      (unit cube [-1,1]^3 * dims/max(dims) like the bundled data, identity). */
int b200host_volume_desc(const int32_t dims[3], int channels, const char* format /* "uint8" "uint16" "float" */,
                         const float* spacing, const float* offset, const float* phys_to_world, b200rc_volume_desc* out);

/* ---- OptimizedProxyGeometry, the transfer-function dependent modes (V/modules/base/processors/proxygeometry/
 *      optimizedproxygeometry.cpp:878-880,1014-1116 "minboundingbox", :1119-1290 "visiblebricks"): the region structure comes
 *      from the device (b200rc_region_minmax over cells of eep->proxy_cell_size voxels, [llf-1, urb+1] aprons, real-world units);
 *      every region is classified with the 256^2 pre-integration table of its channel's transfer function at sampling step 1.0
 *      (PreIntegrationTable::computeTable, preintegrationtable.cpp:67-133, alpha channel) and is empty iff
 *      classify(min, max).a <= 1e-4 * eep->proxy_threshold; the non-empty regions give one bounding cube (MINBBOX) or one cube each
 *      (VISIBLE_BRICKS; the reference merges neighbouring bricks into larger cubes, which covers the same union), clipped to
 *      the clip region when enabled.  Cubes are written as (llf.xyz, urb.xyz) in texture space, ready for
 *      b200rc_eep_params.boxes.  B200RC_PROXY_BBOX (or n_tfs == 0) yields the (clipped) bounding box.  *n_boxes receives the
 *      number of cubes found (which may exceed max_boxes; only max_boxes are written). */
typedef struct b200host_tf1d {
    const b200host_tf_key* keys; int32_t n_keys;
    float domain[2];            /* real-world units */
    float threshold[2];         /* [0,1] */
    float gamma;
    int32_t alpha_mode;         /* 0 zero, 1 use, 2 one (TransFuncBase::AlphaMode) */
} b200host_tf1d;
int b200host_optimized_proxy_geometry(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc, const b200rc_eep_params* eep,
                                      const b200host_tf1d* tfs, int n_tfs, float* out_boxes, int max_boxes, int* n_boxes);

/* ---- .vvd volume descriptions (V/modules/core/io/vvdvolumereader.cpp:58-128, vvdformat.cpp:59-120,290-330): the XML
 *      wrapper every bundled workspace loads.  Parses volume number `index` of the file: <RawData format x y z> with the raw
 *      file either as the old `filename` attribute (relative to the .vvd) or as <Paths><paths><item value>, and the MetaData
 *      items Offset, Spacing (Vec3MetaData), Transformation (Mat4MetaData, rows value.row0..3) and RealWorldMapping (scale,
 *      offset); a volume without RealWorldMapping gets the denormalising one, as VvdVolumeReader::read does (:103-118).
 *      DerivedData (hash, min/max, histogram, preview) is ignored: the library rebuilds what it needs on the device.
 *      Formats: uint8 / uint16 / float and VectorN(<those>) for N = 2..4 channels; anything else (signed, 32/64-bit integer,
 *      RawDataNoExternalFile, OctreeData) returns B200RC_ERR_INVALID with the reason in `error` (if given).
 *      raw_path receives the absolute/relative path of the raw file (resolved against the directory of the .vvd). */
int b200host_vvd_parse(const char* vvd_path, int index, b200rc_volume_desc* out_desc, char* raw_path, size_t raw_path_len,
                       char* error, size_t error_len);
/* Number of <Volume> entries in the file, or a negative status. */
int b200host_vvd_count(const char* vvd_path);

/* ---- SingleVolumeRaycaster mirror (V/modules/base/processors/render/singlevolumeraycaster.cpp:46-170,246-425) ---- */
typedef struct b200host_svrc b200host_svrc;

int  b200host_svrc_create(int device, b200host_svrc** out);
void b200host_svrc_destroy(b200host_svrc* p);
const char* b200host_svrc_last_error(const b200host_svrc* p);
b200rc_ctx* b200host_svrc_context(b200host_svrc* p);

/* "-c Processor.Property=value" (V/apps/voreentool/voreentool.cpp:338-350): property IDs and option strings of
 * VolumeRaycaster / VolumeRenderer / SingleVolumeRaycaster / MeshEntryExitPoints / VolumePort, e.g.
 *   samplingRate=2  compositing=dvr compositing2=mip gradient=central-differences classification=pre-integrated-cpu
 *   shading=phong isoValue=0.5 gammaValue=0.1 lightPosition="2.3 1.5 1.5 1" materialShininess=60
 *   textureFilterMode=linear textureClampMode=clamp-to-edge textureBorderIntensity=0 enableChannelShift=true
 *   channelShift1="1 0 0" applyLightAttenuation=false lightAttenuation="1 0 0" interactionQuality=1
 *   cameraInsideVolumeTechnique=gpu jitterEntryPoints=false jitterStepLength=0.005 useFloatRenderTargets=false
 * and the implementation knobs emptySpaceSkipping=true sampler=exact|texture outputFormat=rgba16|rgba32f.
 * Unknown IDs / options return B200RC_ERR_INVALID (the reference logs "property not found"). */
int b200host_svrc_set_property(b200host_svrc* p, const char* id, const char* value);
int b200host_svrc_get_sampling_step(b200host_svrc* p, float* out);     /* VolumeRaycaster::getSamplingStepSize */

/* volumehandle.volumehandle: the VolumeRAM pointer + meta data; uploads to the brick pool. */
int b200host_svrc_set_volume(b200host_svrc* p, const void* data, const b200rc_volume_desc* desc);
/* synthetic benchmark volume generated in HBM (no host copy) */
int b200host_svrc_set_synthetic_volume(b200host_svrc* p, const b200rc_volume_desc* desc, uint64_t seed);
/* transferFunction, transferFunction1..3 (channel 0..3): keys + domain (real-world units) + threshold + gamma */
int b200host_svrc_set_transfer_function(b200host_svrc* p, int channel, const b200host_tf_key* keys, int n_keys,
                                        float domain_lo, float domain_hi, const float threshold[2], float gamma);
/* camera property */
int b200host_svrc_set_camera(b200host_svrc* p, const float position[3], const float focus[3], const float up[3],
                             float fovy_deg, float near_dist, float far_dist);
/* image.output / image.output1 / image.output2 connected? (default: only image.output) */
int b200host_svrc_connect_output(b200host_svrc* p, int port, int connected);
/* image.entrypoints / image.exitpoints handed over as host float buffers (what the reference's RenderPorts hold);
 * pass NULLs to detach: process() then runs the MeshEntryExitPoints mirror on the device. */
int b200host_svrc_set_entry_exit(b200host_svrc* p, int width, int height, const float* entry_rgba, const float* entry_depth,
                                 const float* exit_rgba, const float* exit_depth);
/* Processor::isReady (singlevolumeraycaster.cpp:192-205): 1 ready, 0 not */
int b200host_svrc_is_ready(b200host_svrc* p);
/* beforeProcess + process: out_color[i] / out_depth are host buffers (pixel type per outputFormat) or NULL. */
int b200host_svrc_process(b200host_svrc* p, int width, int height, void* out_color[B200RC_MAX_OUTPUTS], float* out_depth);

#ifdef __cplusplus
}
#endif
#endif
