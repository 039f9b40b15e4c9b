// common.cuh -- device-side parameter blocks and small helpers shared by the b200rc kernels.
//
// Arithmetic contract of the EXACT sampler (DESIGN.md "Arithmetic contract"): every operation that feeds a
// discrete decision of the reference shader (ray cast or not, sample position, trilinear weights, TF texel index,
// alpha > 0, MIP compare) is written with explicitly rounded intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn,
// fmaf only where the GL-spec lerp is a fused a + w*(b-a)) so that nvcc cannot contract or reassociate it.
// Everything that only feeds continuous colour math (pow, Phong, compositing) is left to the compiler.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200rc.h"

#define B200RC_ESS_CELL 8          // level-0 macro-cell edge in voxels (min/max + occupancy grid)
#define B200RC_ESS_COARSE 8        // level-1 cell = 8x8x8 level-0 cells (64 voxels)
#define B200RC_BRICK 8             // bricked pool: 8x8x8 voxels per brick, bricks x-fastest
#define B200RC_STAT_CELL 16        // V_touched accounting granularity (16^3 macro-cells)

struct VolDev {
    const void* bricks;             // bricked pool (dtype x channels), see brick_offset()
    cudaTextureObject_t tex;        // cudaArray view, unnormalised coords, filter/clamp as requested
    int dims[3];                    // GLOBAL volume dimensions (texture space refers to these)
    int res_lo[3], res_dims[3];     // resident voxel range [res_lo, res_lo+res_dims) (== whole volume unless sharded)
    int own_lo[3], own_hi[3];       // sample ownership range of this shard (tap base index), whole volume otherwise
    int sharded;
    int clamp_lo[3], clamp_hi[3];   // texel index clamp: [0,dims-1] intersected with the resident range
    int nbx, nby;                   // bricks per row / column of the resident range
    float dimf[3], dims_rcp[3], spacing[3], spacing_rcp[3], offset[3];
    int nch, dtype, stored_nch;     // stored_nch: 3-channel volumes are padded to 4
    float rwm_scale, rwm_offset;
    int filter, clamp;
    float border;
    float w2t[16];                  // worldToTexture (multi-volume rays are world space)
    float cam_phys[3], light_phys[3];
    // empty-space skipping: occupancy bits for the current TF(s); level 0 = 8^3 voxels, level 1 = 64^3 voxels
    const uint32_t* occ0; const uint32_t* occ1;
    int occ0_dims[3], occ1_dims[3]; // cells per axis
    int occ0_wpr, occ1_wpr;         // 32-bit words per x-row
    int ess;                        // occupancy valid for this render
    // MIP culling (b200rc_params.empty_space_skipping == 2): level-0 {min,max} per cell and level-1 max, normalised texture values
    const float2* minmax; const float* max1; int mip_cull;
};

struct ClsDev {
    const uchar4* tf; int tf_width;
    float dom_lo, dom_hi;
    const float* corr_lut;          // 256 floats: 1 - pow(1 - a/255, step*200)  (DVR/MIDA opacity correction)
    const float4* preint; int pres;
    cudaTextureObject_t preint_tex;
    const uchar4* tf2d; int tf2d_w, tf2d_h;
    float domy_lo, domy_hi;
};

struct DevParams {
    int W, H;
    const float4* entry; const float* entry_depth;   // (1 + nvol) images for multi-volume
    const float4* exitp; const float* exit_depth;
    void* out[3];
    float* out_depth;
    float* out_key;                 // sort-last partial MIP / MOP image: the merge key (max intensity / max alpha) per pixel, or null
    int out_format, quantize_depth24;
    int nvol, multi;
    int eep_shared;                 // multi-volume: every volume uses entry/exit plane 0 (identical proxy geometry, in-kernel entry/exit)
    VolDev vol[B200RC_MAX_VOLUMES];
    ClsDev cls[B200RC_MAX_CHANNELS];
    int classification[B200RC_MAX_VOLUMES], shading[B200RC_MAX_VOLUMES];
    float step, mv_factor;
    int gradient;
    int comp[3], connected[3];
    float iso, gamma[3];
    int channel_shift; float shift[4][3];
    b200rc_light light;
    float near_d, far_d;
    int ert;                        // USE_EARLY_TERMINATION
    int ert_snap;                   // 1: alpha := 1 on termination (reference); 0: sort-last partial image
    int preint_comp;                // compositing header generated for pre-integration (no opacity correction)
    int lattice_exact;              // k*step exactly representable for all samples -> closed-form skipping
    int collect_stats;
    unsigned long long* stats;      // [0] rays [1] classified [2] shaded [3] nominal [4] skipped
    unsigned char* touched;         // per 16^3 macro-cell flag (collect_stats only)
    int touched_dims[3];
    // sort-last exact early ray termination (second pass): only pixels with flags[pix] == 2 are marched, their alpha
    // accumulator starts at the alpha of everything in front of this shard (prefix), so the 0.95 test is the global one
    const float* prefix; const unsigned char* flags;
    // ... compacted by ert_plan_kernel: the per-ray kernel marches pixel_list[0 .. *pixel_count) instead of image tiles
    const unsigned int* pixel_list; const unsigned int* pixel_count;
    int seg, seg_apply; float seg_scale; const float2* seg_dom;      // SegmentationRaycaster (rc_segmentation.frag); the segmentation TF sits in cls[1].tf2d, the segmentation volume in vol[1]
    int simple;                     // 1: SimpleRaycaster (rc_simple.frag), 2: RGBRaycaster (rc_rgb.frag); per-ray generic kernel
    int rgb_modulate;               // RGBRaycaster applyColorModulation
    int band_mod, band_rem;         // sort-first: render only the pixel-row bands with band % band_mod == band_rem (band_mod <= 1: all)
    int band_rows;                  // rows per band (a multiple of 16; b200rc_params.image_band_rows, default 16)
    unsigned int* tile_counter;     // persistent tile scheduler
    int tiles_x, tiles_y;
};

// ---- small vector helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// un-contracted: cross(v, v) must be exactly zero (fixClipBorderGradient at the image border / next to uncovered pixels feeds two
// identical vectors; an fma residue would turn the resulting NaN normal into an arbitrary unit vector)
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return f3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)), __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// GLSL normalize: a zero vector gives NaN (0 * rsq(0)); max()/min() below drop NaNs like FMNMX / maxps do.
__device__ __forceinline__ float3 normalize3(float3 a) { float r = rsqrtf(dot3(a, a)); return f3(a.x * r, a.y * r, a.z * r); }
// The same with rsqrt.approx.ftz: identical MUFU.RSQ result for normal operands, without the denormal pre-/post-scaling of rsqrtf
// (a squared length below 1.2e-38 means a vector shorter than 1e-19: the differences of normalised voxel values that feed the
// gradients are either exactly zero -- NaN normal either way -- or above 1e-12)
__device__ __forceinline__ float rsqrt_ftz(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float3 normalize3_ftz(float3 a) { float r = rsqrt_ftz(dot3(a, a)); return f3(a.x * r, a.y * r, a.z * r); }
__device__ __forceinline__ float length3(float3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float3 ld3(const float* p) { return f3(p[0], p[1], p[2]); }
__device__ __forceinline__ float3 mix3(float3 a, float3 b, float t) { return a * (1.f - t) + b * t; }

// exact (non-contractable) helpers
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xlerp(float a, float b, float w) { return __fmaf_rn(w, __fsub_rn(b, a), a); }
// row-major mat4 * (p,1) with the summation order ((a+b)+c)+d
__device__ __forceinline__ float4 xmat_mul(const float* m, float4 v) {
    float4 r;
    r.x = xadd(xadd(xadd(xmul(m[0], v.x), xmul(m[1], v.y)), xmul(m[2], v.z)), xmul(m[3], v.w));
    r.y = xadd(xadd(xadd(xmul(m[4], v.x), xmul(m[5], v.y)), xmul(m[6], v.z)), xmul(m[7], v.w));
    r.z = xadd(xadd(xadd(xmul(m[8], v.x), xmul(m[9], v.y)), xmul(m[10], v.z)), xmul(m[11], v.w));
    r.w = xadd(xadd(xadd(xmul(m[12], v.x), xmul(m[13], v.y)), xmul(m[14], v.z)), xmul(m[15], v.w));
    return r;
}

// GL_RGBA16 / GL_DEPTH_COMPONENT24 stores
__device__ __forceinline__ unsigned short to_unorm16(float v) {
    if (!(v > 0.f)) return 0;
    if (v > 1.f) v = 1.f;
    return (unsigned short)floorf(v * 65535.f + 0.5f);
}
__device__ __forceinline__ float unorm16_roundtrip(float v) { return __fdiv_rn((float)to_unorm16(v), 65535.f); }
__device__ __forceinline__ float depth24_roundtrip(float v) {
    if (!(v > 0.f)) return 0.f;
    if (v > 1.f) v = 1.f;
    double q = floor((double)v * 16777215.0 + 0.5);
    return (float)(q / 16777215.0);
}

// bricked pool addressing: element index of voxel (x,y,z) given in RESIDENT coordinates
__device__ __forceinline__ long long brick_off_x(int x) { return (long long)(x >> 3) * 512 + (x & 7); }
__device__ __forceinline__ long long brick_off_y(int y, int nbx) { return (long long)(y >> 3) * nbx * 512 + (y & 7) * 8; }
__device__ __forceinline__ long long brick_off_z(int z, int nbx, int nby) { return (long long)(z >> 3) * nbx * nby * 512 + (z & 7) * 64; }

// host-side launchers (one per translation unit)
struct b200rc_ctx;
