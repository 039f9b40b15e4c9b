// ctx.hpp -- host-side context of the b200rc library (one context = one GPU, one stream).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "common.cuh"

#define RC_CHECK(call)                                                                                           \
    do {                                                                                                         \
        cudaError_t e__ = (call);                                                                                \
        if (e__ != cudaSuccess) {                                                                                \
            ctx->set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                                 \
            return B200RC_ERR_CUDA;                                                                              \
        }                                                                                                        \
    } while (0)

struct VolumeSlot {
    bool valid = false;
    b200rc_volume_desc desc{};
    int res_lo[3] = {0, 0, 0}, res_dims[3] = {0, 0, 0};     // resident voxel range
    int own_lo[3] = {0, 0, 0}, own_hi[3] = {0, 0, 0};
    bool sharded = false;
    int stored_nch = 1;
    size_t elem_bytes = 0;
    int nbx = 0, nby = 0, nbz = 0;
    void* bricks = nullptr; size_t brick_bytes = 0;
    cudaArray_t array = nullptr;
    cudaTextureObject_t tex = 0; int tex_filter = -1, tex_clamp = -1; float tex_border = 0.f;
    // min/max grid, level-0 cells (B200RC_ESS_CELL voxels): float2 per cell and channel, normalised texture values
    float2* minmax = nullptr; int cells[3] = {0, 0, 0}; double minmax_build_ms = 0.0;
    float* max1 = nullptr;      // level-1 (64^3 voxels) maximum of the level-0 cell maxima, channel 0 (MIP culling)
    uint32_t* occ0 = nullptr; uint32_t* occ1 = nullptr;
    int occ0_wpr = 0, occ1_wpr = 0, occ1_dims[3] = {0, 0, 0};
    unsigned long long occ_key = 0;     // (tf generation, classification) the occupancy was built for
    unsigned char* touched = nullptr; int touched_dims[3] = {0, 0, 0};
};

struct ClassifierSlot {
    uchar4* tf = nullptr; int tf_width = 0; float dom[2] = {0.f, 1.f};
    std::vector<uint8_t> tf_host;
    float* corr_lut = nullptr; float lut_step = -1.f;
    float4* preint = nullptr; int pres = 0;
    cudaArray_t preint_array = nullptr; cudaTextureObject_t preint_tex = 0;
    uchar4* tf2d = nullptr; int tf2d_w = 0, tf2d_h = 0; float domy[2] = {0.f, 1.f};
    // "any alpha > 0" prefix sums along the intensity axis for the three lookup kinds (device, int[R+1])
    int* nz_tf = nullptr; int* nz_pi = nullptr; int* nz_2d = nullptr;
    unsigned long long generation = 0;  // bumped whenever anything the occupancy grids depend on changed
    // change detection (a drop-in processor hands its TFs over every frame; nothing is re-uploaded or rebuilt unless it changed)
    unsigned long long tf_gen = 0;      // bumped when the 1D TF texels / domain changed
    unsigned long long pi_built_gen = 0; int pi_built_mode = -1; float pi_built_step = 0.f;     // table built from tf_gen with (mode, step)
    unsigned long long pi_hash = 0; bool pi_uploaded = false;                                   // table uploaded from the host: its content hash
    std::vector<uint8_t> tf2d_host;
    int preint_array_res = 0; bool preint_tex_dirty = true; int nz_pi_res = 0;     // texture view of the table (B200RC_SAMPLER_TEX only), rebuilt lazily
};

struct OctreeSlot;

struct b200rc_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string error;
    int num_sms = 148;
    VolumeSlot vol[B200RC_MAX_VOLUMES];
    ClassifierSlot cls[B200RC_MAX_CHANNELS];
    // per-frame device buffers
    float4* d_entry = nullptr; float4* d_exit = nullptr; float* d_entry_depth = nullptr; float* d_exit_depth = nullptr;
    size_t eep_capacity = 0;        // pixels x images
    int eep_w = 0, eep_h = 0;       // size of the entry/exit images currently held (b200rc_remarch_device reuses them)
    void* d_out[3] = {nullptr, nullptr, nullptr}; float* d_out_depth = nullptr; size_t out_capacity = 0;
    unsigned long long* d_stats = nullptr;
    unsigned int* d_tile_counter = nullptr;     // [0] scheduler counter, [1] number of tiles with work (slab marcher)
    unsigned int* d_remarch_list = nullptr; size_t remarch_capacity = 0;   // pixels flagged by the last b200rc_ert_plan (count: d_tile_counter[2])
    const void* plan_flags = nullptr; size_t plan_n = 0;                   // ... and the flags buffer / pixel count that call was made for
    float* d_tile_cost = nullptr; unsigned int* d_tile_order = nullptr; size_t tile_capacity = 0;
    float* d_boxes = nullptr; int boxes_capacity = 0;
    OctreeSlot* octree[B200RC_MAX_VOLUMES] = {nullptr, nullptr, nullptr, nullptr};    // multi-resolution brick pools (octree.cu)
    float2* d_seg_domains = nullptr;    // SegmentationRaycaster: 256 per-segment TF domains
    b200rc_stats stats{};
    unsigned long long launches = 0;
    bool use_slab = true;               // B200RC_MARCHER=ray (environment) selects the per-ray kernels for A/B measurements
    void set_error(const std::string& s) { error = s; }
};

// ---- launchers implemented in the individual translation units (all asynchronous on ctx->stream) ---------------
// volume.cu
int rc_volume_upload(b200rc_ctx* ctx, int slot, const void* data, const b200rc_volume_desc* desc);
int rc_volume_synth(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc, uint64_t seed);
int rc_volume_release(b200rc_ctx* ctx, int slot);
int rc_volume_stream_begin(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc);
int rc_volume_stream_slab(b200rc_ctx* ctx, int slot, int z0, int nz, const void* data);
int rc_volume_stream_end(b200rc_ctx* ctx, int slot);
int rc_volume_load_raw(b200rc_ctx* ctx, int slot, const char* path, const b200rc_volume_desc* desc, uint64_t header_skip, int slab_planes);
int rc_volume_download(b200rc_ctx* ctx, int slot, void* host_dst, size_t bytes);
int rc_volume_ensure_texture(b200rc_ctx* ctx, int slot, int filter, int clamp, float border);
int rc_volume_region_minmax(b200rc_ctx* ctx, int slot, int cell_size, int32_t cells[3], float* out, size_t out_floats);
int rc_volume_build_occupancy(b200rc_ctx* ctx, int slot, const int* cls_modes, int n_channels, bool multi_slot, int clamp, float border);
// preint.cu
int rc_tf_set_1d(b200rc_ctx* ctx, int chan, const uint8_t* rgba8, int width, float lo, float hi);
int rc_tf_set_preint(b200rc_ctx* ctx, int chan, const float* rgba32f, int res);
int rc_tf_build_preint(b200rc_ctx* ctx, int chan, int res, float step, int mode);
int rc_tf_set_2d(b200rc_ctx* ctx, int chan, const uint8_t* rgba8, int w, int h, float lox, float hix, float loy, float hiy);
int rc_tf_update_lut(b200rc_ctx* ctx, int chan, float step);
int rc_tf_ensure_preint_texture(b200rc_ctx* ctx, int chan);
// eep.cu
int rc_eep_launch(b200rc_ctx* ctx, const b200rc_eep_params* eep, const b200rc_camera* cam, const float* boxes, int n_boxes,
                  const float model[16], int W, int H, float4* entry, float* entry_depth, float4* exitp, float* exit_depth);
int rc_proxy_boxes(b200rc_ctx* ctx, int slot, const b200rc_eep_params* eep, float* out_boxes_host, int max_boxes, int* n_out);
// march_generic.cu / march_fast.cu
int rc_march_generic(b200rc_ctx* ctx, const DevParams& p, int sampler, int nch);
bool rc_march_fast(b200rc_ctx* ctx, const DevParams& p, int sampler, int nch, int dtype, int* rc_out);
bool rc_march_fast_multi(b200rc_ctx* ctx, const DevParams& p, int sampler, int* rc_out);
bool rc_march_mslab(b200rc_ctx* ctx, const DevParams& p, int sampler, int* rc_out);
bool rc_march_slab(b200rc_ctx* ctx, const DevParams& p, int sampler, int nch, int dtype, int* rc_out);
int rc_march_launch(b200rc_ctx* ctx, const DevParams& p, int sampler, int nch, int dtype);
// composite.cu
int rc_composite_keyed_launch(b200rc_ctx* ctx, const float* fc, const float* fk, const float* fd, const float* bc, const float* bk, const float* bd,
                              float* dc, float* dk, float* dd, size_t n);
int rc_rescale_launch(b200rc_ctx* ctx, const void* src, const float* sdepth, int sw, int sh, int format, void* dst, float* ddepth, int dw, int dh);
int rc_composite_launch(b200rc_ctx* ctx, int mode, const float* fc, const float* fd, const float* bc, const float* bd, float* dc, float* dd, size_t n);
int rc_finalize_launch(b200rc_ctx* ctx, const float* rgba, const float* depth, size_t n, int out_format, int snap, void* out);
int rc_ert_plan_launch(b200rc_ctx* ctx, const float* alphas, const float* const* planes, int world, unsigned front_mask, float* color, float* depth, size_t n,
                       float* prefix, unsigned char* flags, unsigned long long* counter);

// post.cu
int rc_compositor_launch(b200rc_ctx* ctx, int mode, float wf, float w0, float w1, int add_depth, const void* c0, const float* d0, const void* c1, const float* d1,
                         int format, void* dc, float* dd, size_t n);
int rc_background_launch(b200rc_ctx* ctx, const b200rc_background_params* bp, const void* src, const float* sdepth, int format, void* dc, float* dd, int W, int H);

// octree.cu
int rc_octree_build(b200rc_ctx* ctx, int slot, int brick_dim, float homogeneity_threshold, int halfsample_fn, const VolDev& v, const VolumeSlot& vs);
int rc_octree_release(b200rc_ctx* ctx, int slot);
void rc_octree_destroy_all(b200rc_ctx* ctx);
int rc_octree_info(b200rc_ctx* ctx, int slot, b200rc_octree_info* out);
int rc_octree_download(b200rc_ctx* ctx, int slot, int level, uint16_t* avg, uint16_t* mn, uint16_t* mx, uint8_t* flags, uint16_t* bricks);
int rc_octree_render(b200rc_ctx* ctx, int slot, const b200rc_octree_render_params* rp, const float4* entry, const float4* exitp, int W, int H, float4* out, float* out_firsthit);

// host matrix helpers (api.cu), tgt semantics (V/ext/tgt/matrix.h)
void rc_mat4_mul(const float a[16], const float b[16], float out[16]);
bool rc_mat4_invert(const float m[16], float out[16]);
void rc_texture_to_world(const b200rc_volume_desc* d, float out[16]);
