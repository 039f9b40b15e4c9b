/*
 * b200rc.h -- C ABI of the B200-native direct-volume-ray-casting path.
 *
 * This is the drop-in boundary: everything a Voreen
 * `Processor` needs to replace the GL draw call of SingleVolumeRaycaster /
 * MultiVolumeRaycaster / MeshEntryExitPoints / OptimizedProxyGeometry with
 * hand-written sm_100a CUDA.  Plain pointers and sizes only; no C++ types, no
 * torch types.  All entry points return an int status (0 = B200RC_OK); none
 * throws.  `b200rc_last_error()` returns a human readable message for the last
 * failing call on a context.
 *
 * Conventions (all taken from the reference):
 *   - matrices are float[16], row-major, column-vector convention (M * v),
 *     i.e. the memory order of tgt::mat4::elem (V/ext/tgt/matrix.h).
 *   - texture space = [0,1]^3 with voxel i centred at (i+0.5)/dim
 *     (V/src/core/datastructures/volume/volumebase.cpp:221-231).
 *   - images are row-major, row 0 = bottom row (GL window coordinates,
 *     gl_FragCoord.y = row + 0.5), pixel (x,y) at index y*width + x.
 *   - colours are RGBA, premultiplied exactly as the GLSL compositing leaves them
 *     (V/modules/core/glsl/modules/mod_compositing.frag:47-59).
 *
 * V/ = /root/reference/voreen/
 */
#ifndef B200RC_H
#define B200RC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RC_ABI_VERSION 4     /* 3: b200rc_params grew (raycaster, rgb_color_modulation); streaming upload, rescale, raw loader
                                    4: b200rc_finalize takes ert_snap; pinned host buffers; post chain; octree; segmentation */
#define B200RC_MAX_VOLUMES 4   /* MultiVolumeRaycaster: volume1..4 (multivolumeraycaster.cpp:41-65) */
#define B200RC_MAX_CHANNELS 4  /* SingleVolumeRaycaster: NUM_CHANNELS 1..4 (singlevolumeraycaster.cpp:440-445) */
#define B200RC_MAX_OUTPUTS 3   /* image.output, image.output1, image.output2 (volumeraycaster.cpp:44-49) */

typedef struct b200rc_ctx b200rc_ctx;

/* ---- status codes ------------------------------------------------------------------------- */
enum {
    B200RC_OK = 0,
    B200RC_ERR_INVALID = 1,      /* bad argument / unsupported combination (reference logs LERROR) */
    B200RC_ERR_CUDA = 2,         /* CUDA runtime error, see b200rc_last_error */
    B200RC_ERR_NO_DEVICE = 3,    /* no sm_100 device: there is NO CPU fallback */
    B200RC_ERR_NOT_READY = 4,    /* missing volume / TF (Processor::isReady() == false) */
    B200RC_ERR_OOM = 5
};

/* ---- enums mirroring the reference's option strings ------------------------------------------ */
/* volume element types the GL path accepts (V/src/core/datastructures/volume/volumegl.cpp:122-139) */
enum { B200RC_U8 = 1, B200RC_U16 = 2, B200RC_F32 = 3 };

/* "gradient" property (V/src/core/processors/volumeraycaster.cpp:126-131,197-207) */
enum { B200RC_GRAD_NONE = 0, B200RC_GRAD_FORWARD = 1, B200RC_GRAD_CENTRAL = 2, B200RC_GRAD_SOBEL = 3, B200RC_GRAD_FILTERED = 4 };

/* "classification" property (V/src/core/utils/classificationmodes.cpp:84-90).  The three pre-integrated modes differ
 * only in how the table is built; the marcher sees B200RC_CLASS_PREINT. */
enum { B200RC_CLASS_TF = 0, B200RC_CLASS_PREINT = 1, B200RC_CLASS_TF2D = 2 };

/* pre-integration table builders (classificationmodes.cpp:56-73, preintegrationtable.cpp:67-204, preintegration.frag) */
enum { B200RC_PI_CPU = 0, B200RC_PI_APPROX = 1, B200RC_PI_GPU = 2 };

/* "shading" property (V/src/core/utils/glsl.cpp:397-437) */
enum {
    B200RC_SHADE_NONE = 0, B200RC_SHADE_PHONG_D = 1, B200RC_SHADE_PHONG_S = 2, B200RC_SHADE_PHONG_DA = 3,
    B200RC_SHADE_PHONG_DS = 4, B200RC_SHADE_PHONG = 5, B200RC_SHADE_TOON = 6, B200RC_SHADE_COOK_TORRANCE = 7,
    B200RC_SHADE_OREN_NAYAR = 8, B200RC_SHADE_WARD = 9
};

/* "compositing[,2,3]" properties (volumeraycaster.cpp:250-272,379-388) */
enum { B200RC_COMP_DVR = 0, B200RC_COMP_MIP = 1, B200RC_COMP_MOP = 2, B200RC_COMP_MIDA = 3, B200RC_COMP_ISO = 4, B200RC_COMP_FHP = 5, B200RC_COMP_FHN = 6 };

/* VolumePort texture access properties (V/src/core/ports/volumeport.cpp:262-287) */
enum { B200RC_FILTER_NEAREST = 0, B200RC_FILTER_LINEAR = 1 };
enum { B200RC_CLAMP_TO_EDGE = 0, B200RC_CLAMP_TO_BORDER = 1 };

/* how the CUDA marcher fetches voxels (north star: "hardware-trilinear 3D texture fetches or ... staging") */
enum {
    B200RC_SAMPLER_EXACT = 0,   /* fp32 trilinear from the bricked pool, GL-spec float weights (bit-faithful to the oracle) */
    B200RC_SAMPLER_TEX = 1      /* cudaArray + texture units (1.8 fixed-point weights) */
};

/* output pixel formats */
enum {
    B200RC_OUT_RGBA32F = 0,     /* unclamped float4, for analysis */
    B200RC_OUT_RGBA16 = 1       /* GL_RGBA16 unorm: clamp to [0,1], round(v*65535) (RenderPort default, renderport.h:116) */
};

/* MeshEntryExitPoints "cameraInsideVolumeTechnique" (meshentryexitpoints.cpp:71) */
enum { B200RC_INSIDE_NONE = 0, B200RC_INSIDE_GPU = 1 };

/* OptimizedProxyGeometry "modeString" (optimizedproxygeometry.cpp:107-114).  The library itself builds the
 * bounding-box cube (processCube :420-448); for the TF-dependent modes the host mirror classifies the min/max
 * regions (b200rc_region_minmax) with the TF keys and hands the resulting cubes over as explicit boxes. */
enum { B200RC_RAYCASTER_SINGLE = 0, B200RC_RAYCASTER_SIMPLE = 1, B200RC_RAYCASTER_RGB = 2, B200RC_RAYCASTER_SEGMENTATION = 3 };
/* OUTER_FACES ("outerfaces", optimizedproxygeometry.cpp:1392-1604) emits only the cell faces between a non-empty and an empty cell:
 * the boundary of the union of the non-empty cells.  Rendered by MeshEntryExitPoints (nearest front face / farthest back face) that
 * boundary gives the same entry/exit images as the cubes of VISIBLE_BRICKS, whose interior faces are never nearest or farthest;
 * the cell classification (min/max over [llf-1, urb+1], 256^2 pre-integration table, threshold) is the same code (:1436-1455).
 * VOLUME_OCTREE ("volumeoctree", optimizedproxygeometry.cpp:1620-1737): one cube per non-transparent LEAF of the volume's octree
 * (b200rc_octree_build must have run on the slot), classified by the leaf's min / max values, stretched by one voxel (:1689-1694). */
enum { B200RC_PROXY_BBOX = 0, B200RC_PROXY_MINBBOX = 1, B200RC_PROXY_VISIBLE_BRICKS = 2, B200RC_PROXY_OUTER_FACES = 3, B200RC_PROXY_VOLUME_OCTREE = 4 };

/* voreen::Compositor "blendMode" options, in the order of V/modules/base/processors/image/compositor.cpp:47-60 */
enum {
    B200RC_COMPOSITOR_DEPTH_TEST = 0, B200RC_COMPOSITOR_ALPHA_COMPOSITING = 1, B200RC_COMPOSITOR_ALPHA_BLENDING = 2,
    B200RC_COMPOSITOR_ALPHA_BLENDING_B_OVER_A = 3, B200RC_COMPOSITOR_WEIGHTED_AVERAGE = 4, B200RC_COMPOSITOR_TAKE_FIRST = 5,
    B200RC_COMPOSITOR_TAKE_SECOND = 6, B200RC_COMPOSITOR_TAKE_SECOND_IF_READY = 7, B200RC_COMPOSITOR_FIRST_HAS_PRIORITY = 8,
    B200RC_COMPOSITOR_SECOND_HAS_PRIORITY = 9, B200RC_COMPOSITOR_MAXIMUM_ALPHA = 10, B200RC_COMPOSITOR_DIFFERENCE = 11,
    B200RC_COMPOSITOR_DIFFERENCE_ALPHA_MUL = 12, B200RC_COMPOSITOR_ADD = 13
};
/* voreen::Background "backgroundModeAsString" / "blendMode" (V/modules/base/processors/image/background.cpp:107-131) */
enum { B200RC_BACKGROUND_MONOCHROME = 0, B200RC_BACKGROUND_GRADIENT = 1 };
enum { B200RC_BACKGROUND_BLEND_ALPHA = 0, B200RC_BACKGROUND_BLEND_ADD = 1 };

/* ---- PODs ------------------------------------------------------------------------------------ */

/* What VolumeBase exposes to GLSL::setUniform (V/src/core/utils/glsl.cpp:38-83). */
typedef struct b200rc_volume_desc {
    int32_t dims[3];            /* voxels */
    int32_t channels;           /* 1..4 */
    int32_t dtype;              /* B200RC_U8 / U16 / F32 */
    float   spacing[3];         /* VolumeBase::getSpacing */
    float   offset[3];          /* VolumeBase::getOffset (position of voxel 0 centre in physical space) */
    float   phys_to_world[16];  /* VolumeBase::getPhysicalToWorldMatrix */
    float   rwm_scale;          /* RealWorldMapping slope  (u16 default 65535, realworldmappingmetadata.h:72-88) */
    float   rwm_offset;         /* RealWorldMapping intercept */
    /* Sort-last sharding: this context holds voxels [shard_lo, shard_hi) of a global volume of `dims` voxels
     * (plus an apron of `shard_apron` voxels on every interior face, already included in the uploaded data).
     * All zero = the whole volume is resident. */
    int32_t shard_lo[3];
    int32_t shard_hi[3];
    int32_t shard_apron;
    int32_t data_on_device;     /* 1: `data` passed to b200rc_upload_volume is a device pointer */
} b200rc_volume_desc;

/* VolumeRenderer light + material properties (V/src/core/processors/volumerenderer.cpp:41-71). */
typedef struct b200rc_light {
    float position[4];          /* world space, lightPosition */
    float ambient[3];
    float diffuse[3];
    float specular[3];
    float attenuation[3];       /* constant, linear, quadratic */
    float shininess;            /* materialShininess */
    int32_t apply_attenuation;  /* PHONG_APPLY_ATTENUATION */
} b200rc_light;

/* tgt::Camera as the raycaster and MeshEntryExitPoints consume it. */
typedef struct b200rc_camera {
    float view[16];             /* Camera::getViewMatrix */
    float proj[16];             /* Camera::getProjectionMatrix(viewport) */
    float position[3];          /* Camera::getPositionWithOffsets (world) */
    float near_dist, far_dist;  /* for const_to_z_* (volumeraycaster.cpp:280-287) */
} b200rc_camera;

/* Per-volume settings (SingleVolumeRaycaster uses entry 0 only). */
typedef struct b200rc_volume_params {
    int32_t classification;     /* B200RC_CLASS_* (MultiVolumeRaycaster: classification, classification2..4) */
    int32_t shading;            /* B200RC_SHADE_* (MultiVolumeRaycaster: shading1..4) */
    int32_t filter;             /* B200RC_FILTER_* */
    int32_t clamp;              /* B200RC_CLAMP_* */
    float   border_intensity;   /* textureBorderIntensity */
} b200rc_volume_params;

/* Everything VolumeRaycaster::generateHeader + process() turn into #defines and uniforms. */
typedef struct b200rc_params {
    int32_t num_volumes;        /* 1 = SingleVolumeRaycaster semantics (texture-space rays);
                                   >=1 with multi_volume=1 = MultiVolumeRaycaster semantics (world-space rays) */
    int32_t multi_volume;
    b200rc_volume_params vol[B200RC_MAX_VOLUMES];
    float   sampling_step;      /* samplingStepSize_ : texture space (single) / world space (multi) */
    float   mv_opacity_correction_factor;   /* mvOpacityCorrectionFactor_ (multivolumeraycaster.cpp:352-357) */
    int32_t gradient;           /* B200RC_GRAD_* */
    int32_t compositing[B200RC_MAX_OUTPUTS];
    int32_t output_connected[B200RC_MAX_OUTPUTS];
    float   iso_value;
    float   mida_gamma[B200RC_MAX_OUTPUTS]; /* gammaValue1_..3_ as the shader receives them */
    int32_t enable_channel_shift;
    float   channel_shift[B200RC_MAX_CHANNELS][3];  /* already divided by dims (singlevolumeraycaster.cpp:270-279) */
    b200rc_light light;
    /* implementation knobs (no reference counterpart; never change pixels beyond the stated tolerance) */
    int32_t sampler;            /* B200RC_SAMPLER_* */
    int32_t empty_space_skipping;   /* 0: off.  1: macro-cells the TF maps to alpha 0 are skipped inside [entry,exit]; pixels AND sample
                                       counters identical.  2: additionally, with MIP on a single connected output, cells whose maximum cannot
                                       exceed the ray's running maximum are skipped; pixels identical, samples_shaded counts only the samples
                                       actually classified */
    int32_t out_format;         /* B200RC_OUT_* */
    int32_t quantize_depth24;   /* 1: depth rounded to GL_DEPTH_COMPONENT24 */
    int32_t collect_stats;      /* 1: count classified samples / cast rays (slower) */
    int32_t partial_image;      /* 1: sort-last partial image of a sharded volume: early ray termination on the local alpha
                                   without the snap to 1.0 (b200rc_finalize snaps after the merge); use B200RC_OUT_RGBA32F */
    /* Sort-first (image-space) partition for volumes replicated on several GPUs: only pixel rows whose 16-row band index
     * (row / 16) satisfies band % image_band_modulus == image_band_remainder are rendered; all other pixels of the outputs
     * are left untouched.  modulus 0 or 1 = the whole image. */
    int32_t image_band_modulus;
    int32_t image_band_remainder;
    /* Which raycaster of the family (B200RC_RAYCASTER_*).  SIMPLE = voreen::SimpleRaycaster (rc_simple.frag:50-102): one output,
     * 1D TF on voxel.a, no gradients / shading, DVR with opacity correction, first-hit depth at the first sample that leaves
     * result.a > 0, early termination at alpha >= 1.0, at most 255*255 loop iterations.  gradient / shading / classification
     * / compositing fields are ignored. */
    int32_t raycaster;
    /* B200RC_RAYCASTER_RGB = voreen::RGBRaycaster (rc_rgb.frag:65-140) for 3- or 4-channel volumes: the sample's colour is the
     * un-mapped texture value, its alpha comes from the 1D TF looked up at the hue of that colour; rgb_color_modulation =
     * the "applyColorModulation" property (colour *= TF colour).  Otherwise like SIMPLE.  No empty-space skipping (the alpha
     * of a cell is not a function of a scalar range). */
    int32_t rgb_color_modulation;
    /* B200RC_RAYCASTER_SEGMENTATION = voreen::SegmentationRaycaster (V/modules/base/glsl/rc_segmentation.frag:104-160): one-channel
     * intensity volume in slot 0; the loop of SingleVolumeRaycaster with ka = kd = ks = the classified colour (:137) and early ray
     * termination on image.output's alpha alone (:145).  apply_segmentation ("applySegmentation"): the segmentation volume in slot 1
     * (uint8 / uint16, NEAREST, clamp to edge) selects per sample one of 256 transfer functions (b200rc_set_segmentation_tf): segment =
     * min(value * (2^bits - 1), 255); the intensity is mapped through that segment's domain and looked up with LINEAR filtering in row
     * (3 * segment + 1.5) / height of the segmentation TF texture (:67-98).  segmentation_bit_depth = 8, 12 or 16 (bitDepth_). */
    int32_t apply_segmentation;
    int32_t segmentation_bit_depth;
} b200rc_params;

/* MeshEntryExitPoints + proxy geometry settings for b200rc_render_camera / b200rc_entry_exit. */
typedef struct b200rc_eep_params {
    int32_t proxy_mode;         /* B200RC_PROXY_* */
    int32_t proxy_cell_size;    /* macro-cell edge in voxels (OptimizedProxyGeometry stepSize) */
    float   proxy_threshold;    /* "threshold" property (alpha <= 1e-4*threshold is empty) */
    int32_t inside_technique;   /* B200RC_INSIDE_* */
    int32_t jitter;             /* jitterEntryPoints */
    float   jitter_step_length; /* jitterStepLength */
    int32_t world_coords;       /* outputCoordinateSystem == "world" (RGBA16F storage) */
    int32_t enable_clipping;    /* OptimizedProxyGeometry clip region, voxel coords */
    float   clip_llf[3], clip_urb[3];
    /* explicit proxy cubes in texture space (llf.xyz, urb.xyz each); n_boxes == 0: the (clipped) bounding box of the
     * volume in slot 0 is used.  Host pointer. */
    const float* boxes;
    int32_t n_boxes;
} b200rc_eep_params;

/* voreen::Background properties (background.cpp:96-131): color1 / color2 / angle / backgroundModeAsString / blendMode.  The
 * "radial", "texture" and "cloud" modes draw a texture loaded or generated on the host; hand such a background over as the first
 * input of b200rc_image_compositor (mode ALPHA_BLENDING_B_OVER_A is the same blend as Background's alpha mode on straight alpha). */
typedef struct b200rc_background_params {
    int32_t mode;               /* B200RC_BACKGROUND_* */
    int32_t blend_mode;         /* B200RC_BACKGROUND_BLEND_* */
    float   color1[4];          /* "color1", default (1, 1, 1, 1) */
    float   color2[4];          /* "color2", default (0.392, 0.392, 0.392, 1) */
    int32_t angle_deg;          /* "angle", 0..359, default 60 */
} b200rc_background_params;

/* VolumeOctree facts (V/include/voreen/core/datastructures/octree/volumeoctreebase.h): getOctreeDim, getBrickDim, getNumLevels,
 * getActualTreeDepth, getNumNodes, getNumBricks. */
typedef struct b200rc_octree_info {
    int32_t octree_dim;         /* next power of two >= the largest volume dimension (cubic octree) */
    int32_t brick_dim;
    int32_t num_levels;         /* log2(octree_dim / brick_dim) + 1 */
    int32_t actual_depth;       /* levels that hold at least one node after the homogeneity optimisation */
    uint64_t num_nodes;         /* nodes of the (optimised) tree, dummy nodes outside the volume included, as VolumeOctreeNode::getNodeCount */
    uint64_t num_bricks;
    double   build_ms;          /* device time of the construction kernels */
} b200rc_octree_info;

/* SingleOctreeRaycasterCL render parameters (singleoctreeraycastercl.cpp:731-749): samplingStepSize = 1 / (avgDim * samplingRate),
 * node_level = level of detail counted from the root (num_levels - 1 = full resolution, "Full" quality; < 0 = full resolution). */
typedef struct b200rc_octree_render_params {
    float   sampling_step;
    int32_t node_level;
    int32_t filter;             /* B200RC_FILTER_* inside a brick (TEXTURE_FILTER_LINEAR define) */
    int32_t compositing;        /* B200RC_COMP_DVR, _MIP or _MOP ("compositingMode", singleoctreeraycastercl.cpp:345-347; octree_dvr/_mip/_mop.cl) */
} b200rc_octree_render_params;

typedef struct b200rc_stats {
    double   last_render_ms;        /* device time of the last render call's kernels (CUDA events) */
    double   last_march_ms;         /* ... of the ray-marching kernel alone */
    double   last_eep_ms;
    uint64_t rays_cast;             /* pixels with entry != exit */
    uint64_t samples_classified;    /* samples that reached classification */
    uint64_t samples_shaded;        /* samples with alpha > 0 */
    uint64_t samples_nominal;       /* sum over rays of floor(tEnd/step)+1 */
    uint64_t cells_touched;         /* 16^3 macro-cells marched (for V_touched) */
    uint64_t kernel_launches;       /* launches since context creation */
    uint64_t samples_skipped;       /* lattice samples jumped over by empty-space skipping (classified + skipped =
                                       the samples the reference shader executes for the same frame) */
} b200rc_stats;

/* ---- lifetime -------------------------------------------------------------------------------- */
/* Replaces Processor::initialize()/deinitialize() GPU resource handling
 * (pattern: V/modules/opencl/processors/openclprocessor.cpp:44-58). */
int b200rc_abi_version(void);
int b200rc_device_count(void);
int b200rc_create(int device, b200rc_ctx** out_ctx);
int b200rc_destroy(b200rc_ctx* ctx);
const char* b200rc_last_error(const b200rc_ctx* ctx);
/* Run all work of this context on `cuda_stream` (a cudaStream_t); NULL = the context's own stream. */
int b200rc_set_stream(b200rc_ctx* ctx, void* cuda_stream);
int b200rc_synchronize(b200rc_ctx* ctx);
int b200rc_get_stats(b200rc_ctx* ctx, b200rc_stats* out);
/* Page-locked host memory (cudaHostAlloc) for the buffers that cross PCIe every frame: the entry/exit images read back from the
 * reference's RenderPorts (cpuraycaster.cpp:166-169 downloads them into heap memory) and the colour / depth images on their way
 * to RenderPort::getColorTexture()->uploadTexture().  Pageable buffers work too, at roughly half the copy rate. */
int b200rc_host_alloc(size_t bytes, void** out);
int b200rc_host_free(void* p);

/* ---- volume: replaces VolumeGL::generateTexture + GLSL::bindVolumeTexture
 *      (V/src/core/datastructures/volume/volumegl.cpp:106-139,336-381; V/src/core/utils/glsl.cpp:85-112).
 *      Builds the bricked pool (and the cudaArray for B200RC_SAMPLER_TEX) plus the min/max macro-cell grid. */
int b200rc_upload_volume(b200rc_ctx* ctx, int slot, const void* data, const b200rc_volume_desc* desc);
/* Streaming upload, for volumes that should not (or cannot) sit in host memory as a whole: begin allocates the pool, every
 * slab call hands over planes [z0, z0 + nz) of the RESIDENT range (x fastest, channels interleaved; host memory, or device
 * memory when desc->data_on_device) -- z0 a multiple of 8, nz a multiple of 8 except for the last slab -- and end builds the
 * min/max grid.  The brick pool stores the 8-plane brick layers contiguously, so a slab lands in one contiguous range of it. */
int b200rc_upload_volume_begin(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc);
int b200rc_upload_volume_slab(b200rc_ctx* ctx, int slot, int z0, int nz, const void* data);
int b200rc_upload_volume_end(b200rc_ctx* ctx, int slot);
/* Raw file -> HBM: replaces RawVolumeReader / VolumeDiskRaw::loadVolume (V/modules/core/io/rawvolumereader.cpp:460-560,
 * V/src/core/datastructures/volume/volumedisk.cpp) for the formats of b200rc_volume_desc.  The file holds the WHOLE volume
 * (desc->dims, x fastest, little endian) after `header_skip` bytes; only the resident range of desc is read (a sort-last
 * rank streams just its shard + apron out of a volume no single GPU could hold).  Slabs of `slab_planes` planes (0 = about
 * 64 MiB) go through two pinned buffers: the read of slab k+1 overlaps the host->device copy and bricking of slab k. */
int b200rc_load_raw_file(b200rc_ctx* ctx, int slot, const char* path, const b200rc_volume_desc* desc, uint64_t header_skip, int slab_planes);
int b200rc_release_volume(b200rc_ctx* ctx, int slot);
/* Deterministic synthetic "CT-like" volume generated directly in HBM; u16, 1 channel. */
int b200rc_synth_volume(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc, uint64_t seed);
/* Copies the resident (linear, x-fastest) volume back, for tests. */
int b200rc_download_volume(b200rc_ctx* ctx, int slot, void* host_dst, size_t bytes);
/* Min/max macro-cell grid: replaces StructureProxyGeometryBackgroundThread::computeRegionStructure
 * (V/modules/base/processors/proxygeometry/optimizedproxygeometry.cpp:894-987).  out_minmax holds, per cell
 * (x fastest) and channel, {min,max} in real-world units.  cells[] receives the grid size. */
int b200rc_region_minmax(b200rc_ctx* ctx, int slot, int cell_size, int32_t cells[3], float* out_minmax, size_t out_floats);
/* The empty-space-skipping grid itself, for tests: per 8^3 macro-cell (x fastest) and channel {min,max} of the NORMALISED
 * texture values over the cell's tap region [8c-1, 8c+9] (clamped to the resident range).  cells[] receives the grid size;
 * out_minmax may be NULL to query it.  build_ms (may be NULL) returns the device time of the last build. */
int b200rc_download_minmax(b200rc_ctx* ctx, int slot, int32_t cells[3], float* out_minmax, size_t out_floats, double* build_ms);

/* ---- transfer functions: replaces TransFuncBase::getTexture()->bind + TransFunc1D::setUniform
 *      (V/src/core/datastructures/transfunc/1d/transfunc1d.cpp:184-194) and ClassificationModes::bindTexture
 *      (V/src/core/utils/classificationmodes.cpp:56-73).  `chan` = channel (single volume) or volume slot (multi). */
int b200rc_set_tf1d(b200rc_ctx* ctx, int chan, const uint8_t* rgba8, int width, float domain_lo, float domain_hi);
/* Upload a host-built pre-integration table (PreIntegrationTable::getTable, rgba32f, res*res, row = sb). */
int b200rc_set_preint(b200rc_ctx* ctx, int chan, const float* rgba32f, int res);
/* Build the table on the device from the TF set with b200rc_set_tf1d: mode B200RC_PI_GPU follows
 * preintegration.frag:35-87 incl. RGBA16 quantisation; B200RC_PI_CPU / _APPROX follow
 * PreIntegrationTable::computeTable (the TF texture must then have exactly `res` texels). */
int b200rc_build_preint(b200rc_ctx* ctx, int chan, int res, float sampling_step, int mode);
int b200rc_download_preint(b200rc_ctx* ctx, int chan, float* rgba32f, int res);
/* SegmentationRaycaster: the segmentation transfer-function texture (segmentationraycaster.cpp:175-178,381-415): rgba8, `width` x
 * `height` (= 3 rows per segment, 256 segments: 768), LINEAR + CLAMP, every segment's 1D TF resampled into the three rows
 * 3 i .. 3 i + 2; domains = 256 x {lower, upper} real-world domain per segment (segmentationTransferFuncDomains_). */
int b200rc_set_segmentation_tf(b200rc_ctx* ctx, const uint8_t* rgba8, int width, int height, const float* domains);
/* 2D TF (intensity x gradient magnitude), rgba8, LINEAR filtered (transfunc2dprimitives). */
int b200rc_set_tf2d(b200rc_ctx* ctx, int chan, const uint8_t* rgba8, int width, int height,
                    float domain_lo_x, float domain_hi_x, float domain_lo_y, float domain_hi_y);

/* ---- entry/exit points: replaces MeshEntryExitPoints::process
 *      (V/modules/base/processors/entryexitpoints/meshentryexitpoints.cpp:188-273) with the proxy geometry of
 *      OptimizedProxyGeometry.  Buffers are float4 per pixel (rgb = position, a = coverage) + float depth; any
 *      output pointer may be NULL.  Host pointers. */
int b200rc_entry_exit(b200rc_ctx* ctx, const b200rc_eep_params* eep, const b200rc_camera* cam, int width, int height,
                      float* entry_rgba, float* entry_depth, float* exit_rgba, float* exit_depth);

/* ---- render: replaces SingleVolumeRaycaster::process / MultiVolumeRaycaster::process
 *      (V/modules/base/processors/render/singlevolumeraycaster.cpp:246-425, multivolumeraycaster.cpp:225-417).
 *
 * b200rc_render consumes entry/exit buffers produced elsewhere (download of the reference's RenderPorts,
 * pattern V/modules/base/processors/render/cpuraycaster.cpp:166-169): float4 RGBA + float depth, host memory.
 * For multi-volume rendering entry/exit arrays hold [global, vol1..volN] consecutively.
 * out_color[i] (i < 3) receives output i when non-NULL and output_connected[i]; out_depth the depth of output 0.
 * Output memory: host pointers (pinned for best throughput); pixel type per params->out_format. */
int b200rc_render(b200rc_ctx* ctx, const b200rc_params* params, const b200rc_camera* cam, int width, int height,
                  const float* entry_rgba, const float* entry_depth, const float* exit_rgba, const float* exit_depth,
                  void* out_color[B200RC_MAX_OUTPUTS], float* out_depth);

/* Entry/exit generated on the device, then marched; nothing but parameters crosses PCIe on the way in.
 * With params->multi_volume: all volumes must have identical geometry (dims, spacing, offset, transformation, shard range) -- the
 * co-registered channels of config C4 -- and share one world-space entry/exit pair over their common proxy cube; volumes placed
 * differently take per-volume entry/exit buffers through b200rc_render (MultiVolumeProxyGeometry, multivolumeraycaster.cpp:225-417).
 * Sharded volumes (sort-last): a lattice sample -- with every volume's contribution at it, in the shader's order -- belongs to the
 * rank that owns its tap base index, so the shards partition the single-GPU sample set for multi-volume rendering as well. */
int b200rc_render_camera(b200rc_ctx* ctx, const b200rc_params* params, const b200rc_eep_params* eep,
                         const b200rc_camera* cam, int width, int height,
                         void* out_color[B200RC_MAX_OUTPUTS], float* out_depth);

/* Same as b200rc_render_camera but the outputs stay in HBM: out_color_dev / out_depth_dev are DEVICE pointers
 * (used by the sort-last compositor and by device-resident benchmarks). */
int b200rc_render_camera_device(b200rc_ctx* ctx, const b200rc_params* params, const b200rc_eep_params* eep,
                                const b200rc_camera* cam, int width, int height,
                                void* out_color_dev[B200RC_MAX_OUTPUTS], float* out_depth_dev);

/* ---- sort-last compositing (new work; semantics = premultiplied front-to-back over, depth = min;
 *      SURVEY section 5 / 8e).  All pointers are DEVICE pointers; float4 colour + float depth, n pixels.
 *      dst = front OVER back.  `front` may alias dst.  With mode = compositing mode of the output:
 *      DVR -> over; MIP/MOP -> max-alpha/intensity select (w channel carries the key); FHP/FHN/ISO -> nearer depth.
 *      mode | B200RC_COMPOSITE_ERT: early ray termination is active for this output (volumeraycaster.cpp:225-233): a
 *      front image that already reached alpha >= 0.95 hides everything behind it, as the terminated ray would. */
#define B200RC_COMPOSITE_ERT 0x100
int b200rc_composite(b200rc_ctx* ctx, int mode, const float* front_rgba, const float* front_depth,
                     const float* back_rgba, const float* back_depth, float* dst_rgba, float* dst_depth, size_t n);
/* Exact early ray termination across shards (optional second pass between rendering and merging).
 * b200rc_ert_plan: alphas_dev holds the alpha planes of ALL ranks' first-pass partial images (world x n floats, rank major),
 * front_mask the ranks whose brick lies in front of this context's brick (camera side of every differing split plane).
 * Erases this rank's pixels behind a terminated ray, and flags (2) those whose ray crosses the global 0.95 threshold inside
 * this shard; prefix_dev receives the alpha in front of the shard.  remarch_count (host, may be NULL) returns the number of
 * flagged pixels (synchronises).  b200rc_remarch_device re-marches exactly the flagged pixels with the alpha accumulator
 * started at the prefix -- mod_raysetup.frag:65-74 then fires on the GLOBAL alpha -- reusing the entry/exit images of the
 * preceding b200rc_render_camera_device call, and overwrites those pixels of the partial image. */
int b200rc_ert_plan(b200rc_ctx* ctx, const float* alphas_dev, int world, uint32_t front_mask, float* color_dev, float* depth_dev, size_t n,
                    float* prefix_dev, uint8_t* flags_dev, uint64_t* remarch_count);
/* The same plan without the all-gather: alpha_planes_dev is a HOST array of `world` device pointers, entry r = rank r's alpha plane
 * (n floats; peer-mapped memory of another GPU is fine, the kernel reads it over NVLink; entries of ranks outside front_mask may be NULL).
 * A DVR partial image rendered by b200rc_render_camera_device with an unconnected out_color_dev[1] writes its alpha plane there. */
int b200rc_ert_plan_planes(b200rc_ctx* ctx, const float* const* alpha_planes_dev, int world, uint32_t front_mask, float* color_dev, float* depth_dev,
                           size_t n, float* prefix_dev, uint8_t* flags_dev, uint64_t* remarch_count);
int b200rc_remarch_device(b200rc_ctx* ctx, const b200rc_params* params, const b200rc_camera* cam, int width, int height,
                          const float* prefix_dev, const uint8_t* flags_dev, void* out_color_dev[B200RC_MAX_OUTPUTS], float* out_depth_dev);

/* VolumeRaycaster::rescaleRendering (V/src/core/processors/volumeraycaster.cpp:348-377, copyimage.frag:37-55): the internal
 * render targets (rendered at outport size / interactionCoarseness) are copied to the outports with GL_LINEAR filtering and
 * CLAMP_TO_EDGE: destination pixel (x, y) fetches the source at ((x + 0.5) / dst_w, (y + 0.5) / dst_h), colour and depth alike;
 * RGBA16 sources/destinations are unorm16 (B200RC_OUT_RGBA16) or float (B200RC_OUT_RGBA32F), depth is re-quantised to DEPTH24.
 * Host buffers unless on_device != 0; src_depth / dst_depth may both be NULL. */
int b200rc_rescale(b200rc_ctx* ctx, const void* src_color, const float* src_depth, int src_w, int src_h, int format,
                   void* dst_color, float* dst_depth, int dst_w, int dst_h, int on_device);

/* ---- octree / multi-resolution brick pool: replaces VolumeOctree construction
 *      (V/src/core/datastructures/octree/volumeoctree.cpp:637-961,1186-1536) and the DVR traversal of SingleOctreeRaycasterCL
 *      (V/modules/opencl/cl/singleoctreeraycaster.cl:190-300, octree_config.cl, octree_dvr.cl) for one-channel uint8 / uint16 volumes.
 * build: from the resident volume of `slot`; brick_dim a power of two; homogeneity_threshold relative to the value range
 *        (< 0: complete tree, no optimisation; >= 0: nodes with max - min <= threshold * 65535 keep no brick and no subtree);
 *        half_sample: 0 mean, 1 max, 2 min (VolumeOctree::HalfSampleAggregateFunction).
 * download: node values of `level` (0 = root; (2^level)^3 nodes, x fastest) and, optionally, their bricks (brick_dim^3 uint16 each,
 *        node order); flags bit 0 = node exists, 1 = inside the volume, 2 = has a brick, 3 = has children.  Any pointer may be NULL.
 * render: DVR through the octree at `node_level` with the 1D TF of classifier slot 0 (LINEAR-filtered, as read_imagef does) --
 *        no shading, early termination at 0.95, out_rgba32f W x H x 4 floats, out_firsthit (may be NULL) the first-hit ray
 *        parameter / ray length.  Entry/exit points are generated on the device (eep, texture coordinates). */
int b200rc_octree_build(b200rc_ctx* ctx, int slot, int brick_dim, float homogeneity_threshold, int half_sample);
int b200rc_octree_get_info(b200rc_ctx* ctx, int slot, b200rc_octree_info* out);
int b200rc_octree_download(b200rc_ctx* ctx, int slot, int level, uint16_t* avg, uint16_t* min_values, uint16_t* max_values, uint8_t* flags, uint16_t* bricks);
int b200rc_octree_render(b200rc_ctx* ctx, int slot, const b200rc_octree_render_params* params, const b200rc_eep_params* eep, const b200rc_camera* cam,
                         int width, int height, float* out_rgba32f, float* out_firsthit);

/* ---- image-space post chain: the two processors that follow the raycaster in the standard workspaces, so that a
 *      frame can stay on the device up to the canvas.  Per-pixel maps over equal-sized images; `format` = B200RC_OUT_RGBA16 (what
 *      RenderPorts hold) or B200RC_OUT_RGBA32F for all colour images; depth images are float and may be NULL (= 1.0 / not written).
 *      Host pointers unless on_device != 0.
 * voreen::Compositor (V/modules/base/glsl/image/compositor.frag:41-140): image 0 = "image.inport0", image 1 = "image.inport1";
 * weighting_factor = "weightingFactor" (WEIGHTED_AVERAGE), weight_first / weight_second / add_depth = "weightFirst", "weightSecond",
 * "addDepth" (ADD). */
int b200rc_image_compositor(b200rc_ctx* ctx, int mode, float weighting_factor, float weight_first, float weight_second, int add_depth,
                            const void* color0, const float* depth0, const void* color1, const float* depth1, int format,
                            void* dst_color, float* dst_depth, int width, int height, int on_device);
/* voreen::Background (background.cpp:161-230, background.frag:35-57): the generated background under `src` (src_color may be NULL:
 * the background alone, as when "image.input" is unconnected).  Output depth = min(1, src depth). */
int b200rc_image_background(b200rc_ctx* ctx, const b200rc_background_params* bg, const void* src_color, const float* src_depth, int format,
                            void* dst_color, float* dst_depth, int width, int height, int on_device);

/* Max-key merge for sharded MIP / MOP: dst = the image with the larger key (front on ties); colour, key and depth planes move together.
 * A partial MIP / MOP image is rendered with partial_image = 1 and out_color[1] (output 1 unconnected) pointing at a float-per-pixel
 * key plane: the marcher stores the ray's maximum intensity (MIP) / maximum alpha (MOP) there.  All pointers are device pointers. */
int b200rc_composite_keyed(b200rc_ctx* ctx, const float* front_rgba, const float* front_key, const float* front_depth,
                           const float* back_rgba, const float* back_key, const float* back_depth,
                           float* dst_rgba, float* dst_key, float* dst_depth, size_t n);

/* ERT finalisation after merging (ert_snap != 0: alpha >= 0.95 -> 1.0, mod_raysetup.frag:65-74 -- only when early ray
 * termination is active for the output, i.e. not for mip / mop / mida, volumeraycaster.cpp:225-233) and conversion to out_format. */
int b200rc_finalize(b200rc_ctx* ctx, const float* rgba_dev, const float* depth_dev, size_t n, int out_format, int ert_snap,
                    void* out_color_dev);

#ifdef __cplusplus
}
#endif
#endif /* B200RC_H */
