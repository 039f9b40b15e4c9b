// api.cu -- the C ABI declared in include/b200rc.h: context lifetime, parameter translation, frame orchestration.
// No entry point throws; every failure is an error code + b200rc_last_error().  There is no CPU fallback: without an
// sm_100 device b200rc_create fails with B200RC_ERR_NO_DEVICE.
#include "ctx.hpp"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <cfloat>
#include <new>

// ---- host matrix helpers, tgt semantics (V/ext/tgt/matrix.h:1327-1376 and operator*) -------------------------------
void rc_mat4_mul(const float a[16], const float b[16], float out[16]) {
    float r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            r[i * 4 + j] = s;
        }
    memcpy(out, r, sizeof r);
}
bool rc_mat4_invert(const float m[16], float out[16]) {
    float tmp[16], res[16];
    memcpy(tmp, m, sizeof tmp);
    for (int i = 0; i < 16; ++i) res[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 4; ++i) {
        int swap = i;
        for (int j = i + 1; j < 4; ++j)
            if (fabsf(tmp[j * 4 + i]) > fabsf(tmp[i * 4 + i])) swap = j;
        if (swap != i)
            for (int k = 0; k < 4; ++k) {
                float t = tmp[i * 4 + k]; tmp[i * 4 + k] = tmp[swap * 4 + k]; tmp[swap * 4 + k] = t;
                t = res[i * 4 + k]; res[i * 4 + k] = res[swap * 4 + k]; res[swap * 4 + k] = t;
            }
        if (tmp[i * 4 + i] == 0.f) return false;
        float t = tmp[i * 4 + i];
        for (int k = 0; k < 4; ++k) { tmp[i * 4 + k] /= t; res[i * 4 + k] /= t; }
        for (int j = 0; j < 4; ++j)
            if (j != i) {
                t = tmp[j * 4 + i];
                for (int k = 0; k < 4; ++k) { tmp[j * 4 + k] -= tmp[i * 4 + k] * t; res[j * 4 + k] -= res[i * 4 + k] * t; }
            }
    }
    memcpy(out, res, sizeof res);
    return true;
}
// VolumeBase::getTextureToWorldMatrix (V/src/core/datastructures/volume/volumebase.cpp:190-231)
void rc_texture_to_world(const b200rc_volume_desc* d, float out[16]) {
    const float T1[16] = {1, 0, 0, -0.5f, 0, 1, 0, -0.5f, 0, 0, 1, -0.5f, 0, 0, 0, 1};
    const float S1[16] = {(float)d->dims[0], 0, 0, 0, 0, (float)d->dims[1], 0, 0, 0, 0, (float)d->dims[2], 0, 0, 0, 0, 1};
    const float T2[16] = {1, 0, 0, d->offset[0], 0, 1, 0, d->offset[1], 0, 0, 1, d->offset[2], 0, 0, 0, 1};
    const float S2[16] = {d->spacing[0], 0, 0, 0, 0, d->spacing[1], 0, 0, 0, 0, d->spacing[2], 0, 0, 0, 0, 1};
    float a[16], b[16], v2w[16];
    rc_mat4_mul(T1, S1, a);
    rc_mat4_mul(T2, S2, b);
    rc_mat4_mul(d->phys_to_world, b, v2w);
    rc_mat4_mul(v2w, a, out);
}
static void mat4_mul_v4(const float m[16], const float v[4], float r[4]) {
    for (int i = 0; i < 4; ++i) r[i] = ((m[i * 4] * v[0] + m[i * 4 + 1] * v[1]) + m[i * 4 + 2] * v[2]) + m[i * 4 + 3] * v[3];
}

// ---- lifetime -------------------------------------------------------------------------------------------------------
extern "C" int b200rc_abi_version(void) { return B200RC_ABI_VERSION; }

extern "C" int b200rc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int b200rc_create(int device, b200rc_ctx** out_ctx) {
    if (!out_ctx) return B200RC_ERR_INVALID;
    *out_ctx = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) { cudaGetLastError(); return B200RC_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return B200RC_ERR_NO_DEVICE; }
    if (prop.major != 10) return B200RC_ERR_NO_DEVICE;      // kernels are built for sm_100a only
    b200rc_ctx* ctx = new (std::nothrow) b200rc_ctx();
    if (!ctx) return B200RC_ERR_OOM;
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    { const char* m = getenv("B200RC_MARCHER"); ctx->use_slab = !(m && strcmp(m, "ray") == 0); }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError(); delete ctx; return B200RC_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    for (int i = 0; i < 6; ++i) cudaEventCreate(&ctx->ev[i]);
    cudaMalloc(&ctx->d_stats, 24 * sizeof(unsigned long long));     // [0..7] counters, [8..17] SLAB_PROFILE phase clocks
    cudaMalloc(&ctx->d_tile_counter, 4 * sizeof(unsigned int));
    *out_ctx = ctx;
    return B200RC_OK;
}

extern "C" int b200rc_destroy(b200rc_ctx* ctx) {
    if (!ctx) return B200RC_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    rc_octree_destroy_all(ctx);
    for (int s = 0; s < B200RC_MAX_VOLUMES; ++s) rc_volume_release(ctx, s);
    for (int c = 0; c < B200RC_MAX_CHANNELS; ++c) {
        ClassifierSlot& k = ctx->cls[c];
        if (k.preint_tex) cudaDestroyTextureObject(k.preint_tex);
        if (k.preint_array) cudaFreeArray(k.preint_array);
        cudaFree(k.tf); cudaFree(k.corr_lut); cudaFree(k.preint); cudaFree(k.tf2d); cudaFree(k.nz_tf); cudaFree(k.nz_pi); cudaFree(k.nz_2d);
    }
    cudaFree(ctx->d_entry); cudaFree(ctx->d_exit); cudaFree(ctx->d_entry_depth); cudaFree(ctx->d_exit_depth);
    for (int o = 0; o < 3; ++o) cudaFree(ctx->d_out[o]);
    cudaFree(ctx->d_seg_domains);
    cudaFree(ctx->d_out_depth); cudaFree(ctx->d_stats); cudaFree(ctx->d_tile_counter); cudaFree(ctx->d_remarch_list); cudaFree(ctx->d_tile_cost); cudaFree(ctx->d_tile_order); cudaFree(ctx->d_boxes);
    for (int i = 0; i < 6; ++i) cudaEventDestroy(ctx->ev[i]);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return B200RC_OK;
}

extern "C" const char* b200rc_last_error(const b200rc_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }

extern "C" int b200rc_set_stream(b200rc_ctx* ctx, void* cuda_stream) {
    if (!ctx) return B200RC_ERR_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return B200RC_OK;
}

extern "C" int b200rc_synchronize(b200rc_ctx* ctx) {
    if (!ctx) return B200RC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200RC_OK;
}

// page-locked host memory for the buffers that cross PCIe every frame (entry/exit images in, colour + depth out)
extern "C" int b200rc_host_alloc(size_t bytes, void** out) {
    if (!out) return B200RC_ERR_INVALID;
    *out = nullptr;
    if (cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); *out = nullptr; return B200RC_ERR_OOM; }
    return B200RC_OK;
}
extern "C" int b200rc_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return B200RC_ERR_CUDA; }
    return B200RC_OK;
}

extern "C" int b200rc_get_stats(b200rc_ctx* ctx, b200rc_stats* out) {
    if (!ctx || !out) return B200RC_ERR_INVALID;
    ctx->stats.kernel_launches = ctx->launches;
    *out = ctx->stats;
    return B200RC_OK;
}

#define CTX_GUARD()                                            \
    if (!ctx) return B200RC_ERR_INVALID;                       \
    if (cudaSetDevice(ctx->device) != cudaSuccess) { cudaGetLastError(); ctx->set_error("cudaSetDevice failed"); return B200RC_ERR_CUDA; }

// ---- volume / TF pass-throughs ---------------------------------------------------------------------------------------
extern "C" int b200rc_upload_volume(b200rc_ctx* ctx, int slot, const void* data, const b200rc_volume_desc* desc) { CTX_GUARD(); return rc_volume_upload(ctx, slot, data, desc); }
extern "C" int b200rc_upload_volume_begin(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc) { CTX_GUARD(); return rc_volume_stream_begin(ctx, slot, desc); }
extern "C" int b200rc_upload_volume_slab(b200rc_ctx* ctx, int slot, int z0, int nz, const void* data) { CTX_GUARD(); return rc_volume_stream_slab(ctx, slot, z0, nz, data); }
extern "C" int b200rc_upload_volume_end(b200rc_ctx* ctx, int slot) { CTX_GUARD(); return rc_volume_stream_end(ctx, slot); }
extern "C" int b200rc_load_raw_file(b200rc_ctx* ctx, int slot, const char* path, const b200rc_volume_desc* desc, uint64_t header_skip, int slab_planes) {
    CTX_GUARD(); if (slot < 0 || slot >= B200RC_MAX_VOLUMES) return B200RC_ERR_INVALID; return rc_volume_load_raw(ctx, slot, path, desc, header_skip, slab_planes);
}
extern "C" int b200rc_release_volume(b200rc_ctx* ctx, int slot) { CTX_GUARD(); if (slot < 0 || slot >= B200RC_MAX_VOLUMES) return B200RC_ERR_INVALID; return rc_volume_release(ctx, slot); }
extern "C" int b200rc_synth_volume(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc, uint64_t seed) { CTX_GUARD(); return rc_volume_synth(ctx, slot, desc, seed); }
extern "C" int b200rc_download_volume(b200rc_ctx* ctx, int slot, void* host_dst, size_t bytes) {
    CTX_GUARD(); if (slot < 0 || slot >= B200RC_MAX_VOLUMES || !host_dst) return B200RC_ERR_INVALID; return rc_volume_download(ctx, slot, host_dst, bytes);
}
extern "C" int b200rc_region_minmax(b200rc_ctx* ctx, int slot, int cell_size, int32_t cells[3], float* out_minmax, size_t out_floats) {
    CTX_GUARD(); if (slot < 0 || slot >= B200RC_MAX_VOLUMES || !cells) return B200RC_ERR_INVALID; return rc_volume_region_minmax(ctx, slot, cell_size, cells, out_minmax, out_floats);
}
extern "C" int b200rc_download_minmax(b200rc_ctx* ctx, int slot, int32_t cells[3], float* out, size_t out_floats, double* build_ms) {
    CTX_GUARD();
    if (slot < 0 || slot >= B200RC_MAX_VOLUMES || !cells || !ctx->vol[slot].valid) { ctx->set_error("download_minmax: no such volume"); return B200RC_ERR_INVALID; }
    const VolumeSlot& v = ctx->vol[slot];
    for (int a = 0; a < 3; ++a) cells[a] = v.cells[a];
    if (build_ms) *build_ms = v.minmax_build_ms;
    if (!out) return B200RC_OK;
    const size_t need = (size_t)v.cells[0] * v.cells[1] * v.cells[2] * v.desc.channels * 2;
    if (out_floats < need) { ctx->set_error("download_minmax: output buffer too small"); return B200RC_ERR_INVALID; }
    RC_CHECK(cudaMemcpyAsync(out, v.minmax, need * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200RC_OK;
}
extern "C" int b200rc_set_tf1d(b200rc_ctx* ctx, int chan, const uint8_t* rgba8, int width, float lo, float hi) { CTX_GUARD(); return rc_tf_set_1d(ctx, chan, rgba8, width, lo, hi); }
extern "C" int b200rc_set_preint(b200rc_ctx* ctx, int chan, const float* rgba32f, int res) { CTX_GUARD(); return rc_tf_set_preint(ctx, chan, rgba32f, res); }
extern "C" int b200rc_build_preint(b200rc_ctx* ctx, int chan, int res, float step, int mode) { CTX_GUARD(); return rc_tf_build_preint(ctx, chan, res, step, mode); }
extern "C" int b200rc_download_preint(b200rc_ctx* ctx, int chan, float* rgba32f, int res) {
    CTX_GUARD();
    if (chan < 0 || chan >= B200RC_MAX_CHANNELS || !rgba32f || ctx->cls[chan].pres != res || !ctx->cls[chan].preint) { ctx->set_error("no such pre-integration table"); return B200RC_ERR_INVALID; }
    RC_CHECK(cudaMemcpyAsync(rgba32f, ctx->cls[chan].preint, (size_t)res * res * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200RC_OK;
}
extern "C" int b200rc_set_segmentation_tf(b200rc_ctx* ctx, const uint8_t* rgba8, int width, int height, const float* domains) {
    CTX_GUARD();
    if (!rgba8 || !domains || width < 2 || height < 3) { ctx->set_error("set_segmentation_tf: bad arguments"); return B200RC_ERR_INVALID; }
    int rc = rc_tf_set_2d(ctx, 1, rgba8, width, height, 0.f, 1.f, 0.f, 1.f);        // LINEAR-filtered rgba8 image: the 2D-TF storage of classifier slot 1
    if (rc) return rc;
    if (!ctx->d_seg_domains) RC_CHECK(cudaMalloc(&ctx->d_seg_domains, 256 * sizeof(float2)));
    RC_CHECK(cudaMemcpyAsync(ctx->d_seg_domains, domains, 256 * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200RC_OK;
}
extern "C" int b200rc_set_tf2d(b200rc_ctx* ctx, int chan, const uint8_t* rgba8, int w, int h, float lox, float hix, float loy, float hiy) {
    CTX_GUARD(); return rc_tf_set_2d(ctx, chan, rgba8, w, h, lox, hix, loy, hiy);
}

// ---- frame buffers ------------------------------------------------------------------------------------------------------
static int ensure_frame_buffers(b200rc_ctx* ctx, size_t pixels, int images) {
    const size_t need = pixels * (size_t)images;
    if (need > ctx->eep_capacity) {
        cudaFree(ctx->d_entry); cudaFree(ctx->d_exit); cudaFree(ctx->d_entry_depth); cudaFree(ctx->d_exit_depth);
        ctx->d_entry = ctx->d_exit = nullptr; ctx->d_entry_depth = ctx->d_exit_depth = nullptr; ctx->eep_capacity = 0;
        RC_CHECK(cudaMalloc(&ctx->d_entry, need * sizeof(float4)));
        RC_CHECK(cudaMalloc(&ctx->d_exit, need * sizeof(float4)));
        RC_CHECK(cudaMalloc(&ctx->d_entry_depth, need * sizeof(float)));
        RC_CHECK(cudaMalloc(&ctx->d_exit_depth, need * sizeof(float)));
        ctx->eep_capacity = need;
    }
    if (pixels > ctx->out_capacity) {
        for (int o = 0; o < 3; ++o) { cudaFree(ctx->d_out[o]); ctx->d_out[o] = nullptr; }
        cudaFree(ctx->d_out_depth); ctx->d_out_depth = nullptr; ctx->out_capacity = 0;
        for (int o = 0; o < 3; ++o) RC_CHECK(cudaMalloc(&ctx->d_out[o], pixels * sizeof(float4)));
        RC_CHECK(cudaMalloc(&ctx->d_out_depth, pixels * sizeof(float)));
        ctx->out_capacity = pixels;
    }
    return B200RC_OK;
}

// ---- parameter translation: VolumeRaycaster::generateHeader + process() -> DevParams ------------------------------------
static bool lattice_is_exact(float step, int max_samples) {
    // t_k = fl(t_{k-1} + step) equals k*step for every k iff k*m fits the 24-bit significand, m = odd part of step
    if (!(step > 0.f) || !std::isfinite(step)) return false;
    int e; float fr = frexpf(step, &e);
    uint32_t m = (uint32_t)ldexpf(fr, 24);
    while (m && !(m & 1u)) m >>= 1;
    return (unsigned long long)m * (unsigned long long)max_samples < (1ull << 24);
}

static int build_params(b200rc_ctx* ctx, const b200rc_params* P, const b200rc_camera* cam, int W, int H, DevParams& p, int* out_nch, int* out_dtype) {
    memset(&p, 0, sizeof p);
    if (!P || !cam || W < 1 || H < 1) { ctx->set_error("render: bad arguments"); return B200RC_ERR_INVALID; }
    const int nv = P->multi_volume ? P->num_volumes : 1;
    if (nv < 1 || nv > B200RC_MAX_VOLUMES) { ctx->set_error("render: bad volume count"); return B200RC_ERR_INVALID; }
    if (!(P->sampling_step > 0.f)) { ctx->set_error("render: sampling step must be positive"); return B200RC_ERR_INVALID; }
    p.W = W; p.H = H;
    p.nvol = nv; p.multi = P->multi_volume ? 1 : 0;
    p.step = P->sampling_step; p.mv_factor = P->mv_opacity_correction_factor;
    p.gradient = P->gradient;
    p.simple = P->multi_volume ? 0 : (P->raycaster == B200RC_RAYCASTER_SIMPLE ? 1 : (P->raycaster == B200RC_RAYCASTER_RGB ? 2 : 0));
    p.rgb_modulate = P->rgb_color_modulation ? 1 : 0;
    if (p.simple == 2 && ctx->vol[0].valid && ctx->vol[0].desc.channels < 3) { ctx->set_error("render: RGBRaycaster needs a 3- or 4-channel volume"); return B200RC_ERR_INVALID; }
    if (P->raycaster < B200RC_RAYCASTER_SINGLE || P->raycaster > B200RC_RAYCASTER_SEGMENTATION) { ctx->set_error("render: unknown raycaster kind"); return B200RC_ERR_INVALID; }
    if (!P->multi_volume && P->raycaster == B200RC_RAYCASTER_SEGMENTATION) {
        if (ctx->vol[0].valid && ctx->vol[0].desc.channels != 1) { ctx->set_error("render: SegmentationRaycaster needs a one-channel intensity volume"); return B200RC_ERR_INVALID; }
        p.seg = 1; p.seg_apply = P->apply_segmentation ? 1 : 0;
        if (p.seg_apply) {
            const VolumeSlot& sv = ctx->vol[1];
            if (!sv.valid || sv.desc.channels != 1 || sv.desc.dtype == B200RC_F32 || sv.sharded) { ctx->set_error("render: applySegmentation needs a one-channel uint8 / uint16 segmentation volume in slot 1"); return B200RC_ERR_NOT_READY; }
            if (!ctx->cls[1].tf2d || !ctx->d_seg_domains) { ctx->set_error("render: no segmentation transfer functions (b200rc_set_segmentation_tf)"); return B200RC_ERR_NOT_READY; }
            const int bits = P->segmentation_bit_depth;
            p.seg_scale = bits == 12 ? 4095.0f : (bits == 16 ? 65535.0f : 255.0f);      // rc_segmentation.frag:72-76
            p.seg_dom = ctx->d_seg_domains;
        }
    }
    p.iso = P->iso_value;
    p.light = P->light;
    p.near_d = cam->near_dist; p.far_d = cam->far_dist;
    p.out_format = P->out_format; p.quantize_depth24 = P->quantize_depth24;
    p.collect_stats = P->collect_stats;
    p.band_mod = P->image_band_modulus > 1 ? P->image_band_modulus : 1;
    p.band_rem = P->image_band_modulus > 1 ? P->image_band_remainder : 0;
    p.band_rows = 16;
    { static const int env_rows = getenv("B200RC_BAND_ROWS") ? atoi(getenv("B200RC_BAND_ROWS")) : 0;      // experiments
      if (env_rows >= 16 && env_rows % 16 == 0) p.band_rows = env_rows; }
    if (p.band_rem < 0 || p.band_rem >= p.band_mod) { ctx->set_error("image_band_remainder out of range"); return B200RC_ERR_INVALID; }
    p.ert_snap = P->partial_image ? 0 : 1;      // sort-last partial images keep the un-snapped alpha; b200rc_finalize snaps after the merge
    bool ert = true;
    for (int o = 0; o < 3; ++o) {
        p.comp[o] = P->compositing[o]; p.connected[o] = P->output_connected[o] ? 1 : 0; p.gamma[o] = P->mida_gamma[o];
        if (p.comp[o] < B200RC_COMP_DVR || p.comp[o] > B200RC_COMP_FHN) { ctx->set_error("render: unknown compositing mode"); return B200RC_ERR_INVALID; }
        // USE_EARLY_TERMINATION (volumeraycaster.cpp:225-233, usePortgroup_)
        const bool can = ((p.comp[o] != B200RC_COMP_MIP) && (p.comp[o] != B200RC_COMP_MOP) && (p.comp[o] != B200RC_COMP_MIDA)) || !p.connected[o];
        ert = ert && can;
    }
    p.ert = ert ? 1 : 0;
    if (!p.connected[0] && !p.connected[1] && !p.connected[2]) { ctx->set_error("render: no output connected (isReady() == false)"); return B200RC_ERR_NOT_READY; }
    p.preint_comp = (P->vol[0].classification == B200RC_CLASS_PREINT) ? 1 : 0;     // volumeraycaster.cpp:216-221
    p.channel_shift = (!P->multi_volume && P->enable_channel_shift) ? 1 : 0;
    memcpy(p.shift, P->channel_shift, sizeof p.shift);
    float w2p[16];
    const int nv_fill = p.seg_apply ? 2 : nv;      // SegmentationRaycaster: slot 1 = the segmentation volume (NEAREST, clamp to edge)
    for (int v = 0; v < nv_fill; ++v) {
        VolumeSlot& vs = ctx->vol[v];
        if (!vs.valid) { ctx->set_error("render: volume slot is empty"); return B200RC_ERR_NOT_READY; }
        b200rc_volume_params vp = P->vol[v];
        if (p.seg_apply && v == 1) { vp.filter = B200RC_FILTER_NEAREST; vp.clamp = B200RC_CLAMP_TO_EDGE; vp.border_intensity = 0.f; vp.classification = B200RC_CLASS_TF; vp.shading = B200RC_SHADE_NONE; }
        VolDev& d = p.vol[v];
        d.bricks = vs.bricks;
        for (int a = 0; a < 3; ++a) {
            d.dims[a] = vs.desc.dims[a]; d.res_lo[a] = vs.res_lo[a]; d.res_dims[a] = vs.res_dims[a];
            // a shard that touches the volume border also owns the samples of the outer half voxel (tap base index -1, or
            // dims when the exit point sits on the high face): they are composited with clamp-to-edge by the unsharded marcher
            d.own_lo[a] = vs.own_lo[a] <= 0 ? -2 : vs.own_lo[a];
            d.own_hi[a] = vs.own_hi[a] >= vs.desc.dims[a] ? vs.desc.dims[a] + 2 : vs.own_hi[a];
            d.clamp_lo[a] = vs.res_lo[a] > 0 ? vs.res_lo[a] : 0;
            const int hi = vs.res_lo[a] + vs.res_dims[a] - 1;
            d.clamp_hi[a] = hi < vs.desc.dims[a] - 1 ? hi : vs.desc.dims[a] - 1;
            d.dimf[a] = (float)vs.desc.dims[a]; d.dims_rcp[a] = 1.f / (float)vs.desc.dims[a];       // glsl.cpp:43-44
            d.spacing[a] = vs.desc.spacing[a]; d.spacing_rcp[a] = 1.f / vs.desc.spacing[a];
            d.offset[a] = vs.desc.offset[a];
            d.occ0_dims[a] = vs.cells[a]; d.occ1_dims[a] = vs.occ1_dims[a];
        }
        d.sharded = vs.sharded ? 1 : 0;
        d.nbx = vs.nbx; d.nby = vs.nby;
        d.nch = vs.desc.channels; d.dtype = vs.desc.dtype; d.stored_nch = vs.stored_nch;
        d.rwm_scale = vs.desc.rwm_scale; d.rwm_offset = vs.desc.rwm_offset;
        d.filter = vp.filter; d.clamp = vp.clamp; d.border = vp.border_intensity;
        d.occ0 = vs.occ0; d.occ1 = vs.occ1; d.occ0_wpr = vs.occ0_wpr; d.occ1_wpr = vs.occ1_wpr;
        d.minmax = vs.minmax; d.max1 = vs.max1; d.mip_cull = 0;
        p.classification[v] = vp.classification; p.shading[v] = vp.shading;
        if (vp.classification < B200RC_CLASS_TF || vp.classification > B200RC_CLASS_TF2D || vp.shading < 0 || vp.shading > B200RC_SHADE_WARD) { ctx->set_error("render: unknown classification/shading mode"); return B200RC_ERR_INVALID; }
        if (P->sampler == B200RC_SAMPLER_TEX) {
            int rc = rc_volume_ensure_texture(ctx, v, vp.filter, vp.clamp, vp.border_intensity);
            if (rc) return rc;
            d.tex = vs.tex;
        }
        // camera / light position in volume physical coordinates (glsl.cpp:62-72)
        if (!rc_mat4_invert(vs.desc.phys_to_world, w2p)) for (int i = 0; i < 16; ++i) w2p[i] = (i % 5 == 0) ? 1.f : 0.f;
        const float cp[4] = {cam->position[0], cam->position[1], cam->position[2], 1.f};
        float r[4];
        mat4_mul_v4(w2p, cp, r); d.cam_phys[0] = r[0]; d.cam_phys[1] = r[1]; d.cam_phys[2] = r[2];
        mat4_mul_v4(w2p, P->light.position, r); d.light_phys[0] = r[0]; d.light_phys[1] = r[1]; d.light_phys[2] = r[2];
        float t2w[16];
        rc_texture_to_world(&vs.desc, t2w);
        if (!rc_mat4_invert(t2w, d.w2t)) { ctx->set_error("render: singular textureToWorld matrix"); return B200RC_ERR_INVALID; }
    }
    // classifiers: per channel (single volume) or per volume (multi); SimpleRaycaster: one 1D TF on voxel.a
    const int ncls = P->multi_volume ? nv : (p.simple ? 1 : ctx->vol[0].desc.channels);
    for (int c = 0; c < ncls; ++c) {
        ClassifierSlot& k = ctx->cls[c];
        const int mode = p.simple ? B200RC_CLASS_TF : (P->multi_volume ? P->vol[c].classification : P->vol[0].classification);
        ClsDev& d = p.cls[c];
        d.dom_lo = k.dom[0]; d.dom_hi = k.dom[1]; d.domy_lo = k.domy[0]; d.domy_hi = k.domy[1];
        if (p.seg_apply) continue;      // classified through the segmentation TF texture
        if (mode == B200RC_CLASS_TF) {
            if (!k.tf) { ctx->set_error("render: no 1D transfer function for channel/volume"); return B200RC_ERR_NOT_READY; }
            int rc = rc_tf_update_lut(ctx, c, P->sampling_step);
            if (rc) return rc;
            d.tf = k.tf; d.tf_width = k.tf_width; d.corr_lut = k.corr_lut;
        } else if (mode == B200RC_CLASS_PREINT) {
            if (!k.preint) { ctx->set_error("render: no pre-integration table for channel/volume"); return B200RC_ERR_NOT_READY; }
            if (P->sampler == B200RC_SAMPLER_TEX) { int rc = rc_tf_ensure_preint_texture(ctx, c); if (rc) return rc; }
            d.preint = k.preint; d.pres = k.pres; d.preint_tex = k.preint_tex;
        } else {
            if (!k.tf2d) { ctx->set_error("render: no 2D transfer function for channel/volume"); return B200RC_ERR_NOT_READY; }
            d.tf2d = k.tf2d; d.tf2d_w = k.tf2d_w; d.tf2d_h = k.tf2d_h;
        }
    }
    if (p.seg_apply) {      // the segmentation TF texture: LINEAR-filtered rgba8 image in the 2D-TF storage of classifier slot 1
        p.cls[1].tf2d = ctx->cls[1].tf2d; p.cls[1].tf2d_w = ctx->cls[1].tf2d_w; p.cls[1].tf2d_h = ctx->cls[1].tf2d_h;
    }
    // empty-space skipping (no channel shift): occupancy for the current TFs
    if (p.simple) { p.classification[0] = B200RC_CLASS_TF; p.channel_shift = 0; p.preint_comp = 0; }
    if (p.seg) p.channel_shift = 0;
    if (P->empty_space_skipping && !P->multi_volume && !p.channel_shift && !p.seg_apply && !(p.simple && (p.simple == 2 || ctx->vol[0].desc.channels != 1))) {
        int modes[4];
        for (int c = 0; c < ncls; ++c) modes[c] = p.simple ? B200RC_CLASS_TF : P->vol[0].classification;
        int rc = rc_volume_build_occupancy(ctx, 0, modes, ncls, false, P->vol[0].clamp, P->vol[0].border_intensity);
        if (rc) return rc;
        p.vol[0].ess = 1;
        // MIP culling (mode 2): cells whose maximum cannot beat the running maximum are skipped like transparent ones.  One
        // channel, un-bordered sampling, increasing real-world mapping, only image.output connected and compositing MIP.
        if (P->empty_space_skipping >= 2 && ncls == 1 && p.comp[0] == B200RC_COMP_MIP && p.connected[0] && !p.connected[1] && !p.connected[2] &&
            P->vol[0].clamp == B200RC_CLAMP_TO_EDGE && ctx->vol[0].desc.rwm_scale > 0.f && ctx->vol[0].max1)
            p.vol[0].mip_cull = 1;
    }
    // multi-volume: one occupancy grid per volume for its own classifier (rc_multivolume.frag reads voxel.a of every
    // volume, i.e. the last channel; multi-channel volumes keep marching every sample)
    if (P->empty_space_skipping && P->multi_volume) {
        for (int v = 0; v < nv; ++v) {
            if (ctx->vol[v].desc.channels != 1) continue;
            int mode = P->vol[v].classification;
            int rc = rc_volume_build_occupancy(ctx, v, &mode, 1, true, P->vol[v].clamp, P->vol[v].border_intensity);
            if (rc) return rc;
            p.vol[v].ess = 1;
        }
    }
    p.lattice_exact = (!P->multi_volume && lattice_is_exact(P->sampling_step, 65536)) ? 1 : 0;
    p.stats = ctx->d_stats;
    p.tile_counter = ctx->d_tile_counter;
    p.tiles_x = (((W + 7) / 8) + 7) / 8 * 8;
    p.tiles_y = (((H + 3) / 4) + 7) / 8 * 8;
    if (P->collect_stats && !P->multi_volume) {
        VolumeSlot& vs = ctx->vol[0];
        const size_t n = (size_t)vs.touched_dims[0] * vs.touched_dims[1] * vs.touched_dims[2];
        if (!vs.touched) RC_CHECK(cudaMalloc(&vs.touched, n));
        RC_CHECK(cudaMemsetAsync(vs.touched, 0, n, ctx->stream));
        p.touched = vs.touched;
        for (int a = 0; a < 3; ++a) p.touched_dims[a] = vs.touched_dims[a];
    }
    *out_nch = P->multi_volume ? 1 : ctx->vol[0].desc.channels;
    *out_dtype = ctx->vol[0].desc.dtype;
    return B200RC_OK;
}

__global__ void count_touched_kernel(const unsigned char* __restrict__ f, size_t n, unsigned long long* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned v = (i < n && f[i]) ? 1u : 0u;
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, (unsigned long long)v);
}

int rc_march_launch(b200rc_ctx* ctx, const DevParams& p, int sampler, int nch, int dtype) {
    int rc = B200RC_OK;
    // sort-last second pass (p.flags): ~2 % of the pixels along a band.  Measured at 2048^3 x2 (profiles/r1_pass9_*): slab
    // marcher 2.4 ms, per-ray kernel 3.4 ms -> the slab marcher stays; B200RC_REMARCH=ray selects the per-ray kernel (A/B)
    static const bool remarch_ray = getenv("B200RC_REMARCH") && !strcmp(getenv("B200RC_REMARCH"), "ray");
    const bool slab_ok = ctx->use_slab && !(p.flags && (remarch_ray || p.pixel_list));
    if (p.simple || p.seg) return rc_march_generic(ctx, p, sampler, nch);      // Simple / RGB / SegmentationRaycaster: the run-time-switched per-ray kernel
    if (slab_ok && rc_march_slab(ctx, p, sampler, nch, dtype, &rc)) return rc;
    if (rc_march_fast(ctx, p, sampler, nch, dtype, &rc)) return rc;
    if (slab_ok && rc_march_mslab(ctx, p, sampler, &rc)) return rc;      // co-registered multi-volume: one staged region per volume
    if (rc_march_fast_multi(ctx, p, sampler, &rc)) return rc;
    return rc_march_generic(ctx, p, sampler, nch);
}

static int run_march(b200rc_ctx* ctx, const b200rc_params* P, DevParams& p, int nch, int dtype) {
    if (P->collect_stats) RC_CHECK(cudaMemsetAsync(ctx->d_stats, 0, 24 * sizeof(unsigned long long), ctx->stream));
    RC_CHECK(cudaEventRecord(ctx->ev[2], ctx->stream));
    int rc = rc_march_launch(ctx, p, P->sampler, nch, dtype);
    if (rc) return rc;
    RC_CHECK(cudaEventRecord(ctx->ev[3], ctx->stream));
    if (P->collect_stats && p.touched) {
        const size_t n = (size_t)p.touched_dims[0] * p.touched_dims[1] * p.touched_dims[2];
        count_touched_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p.touched, n, ctx->d_stats + 5);
        ctx->launches++;
    }
    return B200RC_OK;
}

static int finish_frame(b200rc_ctx* ctx, const b200rc_params* P, bool had_eep) {
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    if (had_eep) { cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->stats.last_eep_ms = ms; } else ctx->stats.last_eep_ms = 0.0;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->stats.last_march_ms = ms;
    cudaEventElapsedTime(&ms, had_eep ? ctx->ev[0] : ctx->ev[2], ctx->ev[3]); ctx->stats.last_render_ms = ms;
    if (P->collect_stats) {
        unsigned long long h[8];
        RC_CHECK(cudaMemcpy(h, ctx->d_stats, sizeof h, cudaMemcpyDeviceToHost));
        ctx->stats.rays_cast = h[0]; ctx->stats.samples_classified = h[1]; ctx->stats.samples_shaded = h[2];
        ctx->stats.samples_nominal = h[3]; ctx->stats.samples_skipped = h[4]; ctx->stats.cells_touched = h[5];
        if (getenv("B200RC_DEBUG")) {
            fprintf(stderr, "[b200rc] slab ray-rounds with work %llu of %llu\n", h[6], h[7]);
            unsigned long long ph[10];
            RC_CHECK(cudaMemcpy(ph, ctx->d_stats + 8, sizeof ph, cudaMemcpyDeviceToHost));
            if (ph[7])      // only a -DSLAB_PROFILE build fills these: clock cycles of lane 0 of every warp, summed
                fprintf(stderr, "[b200rc] slab phases (cycles of lane 0 of every warp, summed): skip+zmin %llu wait1 %llu bbox %llu wait2 %llu leader+wait3 %llu stage %llu march %llu warp-rounds %llu | march-loop iterations %llu active lanes %.2f\n",
                        ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7], ph[8], ph[8] ? (double)ph[9] / (double)ph[8] : 0.0);
        }
    }
    return B200RC_OK;
}

static size_t out_pixel_bytes(int fmt) { return fmt == B200RC_OUT_RGBA16 ? 8 : 16; }

// ---- entry/exit -----------------------------------------------------------------------------------------------------------
static int default_boxes(b200rc_ctx* ctx, const b200rc_eep_params* eep, const VolumeSlot& vs, float box[6]) {
    (void)ctx;
    // OptimizedProxyGeometry::processCube (optimizedproxygeometry.cpp:420-448)
    for (int a = 0; a < 3; ++a) {
        float lo = 0.f - 0.5f, hi = (float)(vs.desc.dims[a] - 1) + 0.5f;
        if (eep->enable_clipping) { lo = fmaxf(lo, eep->clip_llf[a] - 0.5f); hi = fminf(hi, eep->clip_urb[a] + 0.5f); }
        box[a] = (1.0f / (float)vs.desc.dims[a]) * (lo + 0.5f);        // getVoxelToTextureMatrix = scale(1/dim) * translate(0.5)
        box[3 + a] = (1.0f / (float)vs.desc.dims[a]) * (hi + 0.5f);
    }
    return B200RC_OK;
}

static int run_eep(b200rc_ctx* ctx, const b200rc_eep_params* eep, const b200rc_camera* cam, int W, int H) {
    const VolumeSlot& vs = ctx->vol[0];
    if (!vs.valid) { ctx->set_error("entry/exit: volume slot 0 is empty"); return B200RC_ERR_NOT_READY; }
    float model[16];
    rc_texture_to_world(&vs.desc, model);
    float box[6];
    const float* boxes = eep->boxes; int nb = eep->n_boxes;
    if (!boxes || nb <= 0) { default_boxes(ctx, eep, vs, box); boxes = box; nb = 1; }
    RC_CHECK(cudaEventRecord(ctx->ev[0], ctx->stream));
    int rc = rc_eep_launch(ctx, eep, cam, boxes, nb, model, W, H, ctx->d_entry, ctx->d_entry_depth, ctx->d_exit, ctx->d_exit_depth);
    if (rc) return rc;
    RC_CHECK(cudaEventRecord(ctx->ev[1], ctx->stream));
    return B200RC_OK;
}

extern "C" int b200rc_entry_exit(b200rc_ctx* ctx, const b200rc_eep_params* eep, const b200rc_camera* cam, int W, int H,
                                 float* entry_rgba, float* entry_depth, float* exit_rgba, float* exit_depth) {
    CTX_GUARD();
    if (!eep || !cam || W < 1 || H < 1) { ctx->set_error("entry_exit: bad arguments"); return B200RC_ERR_INVALID; }
    const size_t px = (size_t)W * H;
    int rc = ensure_frame_buffers(ctx, px, 1);
    if (rc) return rc;
    rc = run_eep(ctx, eep, cam, W, H);
    if (rc) return rc;
    if (entry_rgba) RC_CHECK(cudaMemcpyAsync(entry_rgba, ctx->d_entry, px * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    if (exit_rgba) RC_CHECK(cudaMemcpyAsync(exit_rgba, ctx->d_exit, px * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    if (entry_depth) RC_CHECK(cudaMemcpyAsync(entry_depth, ctx->d_entry_depth, px * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (exit_depth) RC_CHECK(cudaMemcpyAsync(exit_depth, ctx->d_exit_depth, px * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->stats.last_eep_ms = ms;
    return B200RC_OK;
}

// ---- render ---------------------------------------------------------------------------------------------------------------
static int copy_outputs(b200rc_ctx* ctx, const b200rc_params* P, size_t px, void* out_color[3], float* out_depth, cudaMemcpyKind kind, const DevParams& p) {
    for (int o = 0; o < 3; ++o)
        if (out_color && out_color[o] && p.out[o] && p.out[o] != out_color[o])
            RC_CHECK(cudaMemcpyAsync(out_color[o], p.out[o], px * out_pixel_bytes(P->out_format), kind, ctx->stream));
    if (out_depth && p.out_depth && p.out_depth != out_depth) RC_CHECK(cudaMemcpyAsync(out_depth, p.out_depth, px * sizeof(float), kind, ctx->stream));
    return B200RC_OK;
}

extern "C" int b200rc_render(b200rc_ctx* ctx, const b200rc_params* P, const b200rc_camera* cam, int W, int H,
                             const float* entry_rgba, const float* entry_depth, const float* exit_rgba, const float* exit_depth,
                             void* out_color[B200RC_MAX_OUTPUTS], float* out_depth) {
    CTX_GUARD();
    if (!entry_rgba || !exit_rgba || !entry_depth || !exit_depth) { ctx->set_error("render: entry/exit buffers missing (isReady() == false)"); return B200RC_ERR_NOT_READY; }
    DevParams p; int nch = 1, dtype = 0;
    int rc = build_params(ctx, P, cam, W, H, p, &nch, &dtype);
    if (rc) return rc;
    if (P->multi_volume && ctx->vol[0].sharded) { ctx->set_error("render: sharded multi-volume rendering goes through b200rc_render_camera_device (identical geometry, in-kernel entry/exit)"); return B200RC_ERR_INVALID; }
    const size_t px = (size_t)W * H;
    const int images = P->multi_volume ? 1 + P->num_volumes : 1;
    rc = ensure_frame_buffers(ctx, px, images);
    if (rc) return rc;
    RC_CHECK(cudaMemcpyAsync(ctx->d_entry, entry_rgba, px * images * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    RC_CHECK(cudaMemcpyAsync(ctx->d_exit, exit_rgba, px * images * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    RC_CHECK(cudaMemcpyAsync(ctx->d_entry_depth, entry_depth, px * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    RC_CHECK(cudaMemcpyAsync(ctx->d_exit_depth, exit_depth, px * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    p.entry = ctx->d_entry; p.exitp = ctx->d_exit; p.entry_depth = ctx->d_entry_depth; p.exit_depth = ctx->d_exit_depth;
    for (int o = 0; o < 3; ++o) p.out[o] = (p.connected[o] && out_color && out_color[o]) ? ctx->d_out[o] : nullptr;
    p.out_depth = out_depth ? ctx->d_out_depth : nullptr;
    rc = run_march(ctx, P, p, nch, dtype);
    if (rc) return rc;
    rc = copy_outputs(ctx, P, px, out_color, out_depth, cudaMemcpyDeviceToHost, p);
    if (rc) return rc;
    return finish_frame(ctx, P, false);
}

static int render_camera_impl(b200rc_ctx* ctx, const b200rc_params* P, const b200rc_eep_params* eep, const b200rc_camera* cam, int W, int H,
                              void* out_color[3], float* out_depth, bool device_out, const float* prefix = nullptr, const unsigned char* flags = nullptr) {
    if (!eep && !flags) { ctx->set_error("render_camera: eep params missing"); return B200RC_ERR_INVALID; }
    // MultiVolumeRaycaster with in-kernel entry/exit: one proxy cube for all volumes, i.e. volumes of identical geometry (config C4:
    // four co-registered channels); world-space entry/exit points (multivolumeraycaster.cpp:326-344), every volume reads plane 0.
    // Volumes placed differently take explicit per-volume entry/exit buffers (b200rc_render), as MultiVolumeProxyGeometry provides.
    b200rc_eep_params eep_world;
    if (P && P->multi_volume) {
        for (int v = 1; v < P->num_volumes && v < B200RC_MAX_VOLUMES; ++v) {
            const b200rc_volume_desc &a = ctx->vol[0].desc, &b = ctx->vol[v].desc;
            if (!ctx->vol[v].valid || memcmp(a.dims, b.dims, sizeof a.dims) || memcmp(a.spacing, b.spacing, sizeof a.spacing) || memcmp(a.offset, b.offset, sizeof a.offset) ||
                memcmp(a.phys_to_world, b.phys_to_world, sizeof a.phys_to_world) || memcmp(a.shard_lo, b.shard_lo, sizeof a.shard_lo) || memcmp(a.shard_hi, b.shard_hi, sizeof a.shard_hi)) {
                ctx->set_error("render_camera: multi-volume rendering with in-kernel entry/exit needs volumes of identical geometry (and identical shards); use b200rc_render with per-volume entry/exit buffers otherwise");
                return B200RC_ERR_INVALID;
            }
        }
        if (eep) { eep_world = *eep; eep_world.world_coords = 1; eep = &eep_world; }
    }
    DevParams p; int nch = 1, dtype = 0;
    int rc = build_params(ctx, P, cam, W, H, p, &nch, &dtype);
    if (rc) return rc;
    if (P->multi_volume) p.eep_shared = 1;
    const size_t px = (size_t)W * H;
    rc = ensure_frame_buffers(ctx, px, 1);
    if (rc) return rc;
    if (flags) {        // second pass: the entry/exit images of the first pass are still resident
        if (ctx->eep_w != W || ctx->eep_h != H) { ctx->set_error("remarch: no entry/exit images of this size (call b200rc_render_camera_device first)"); return B200RC_ERR_NOT_READY; }
        p.prefix = prefix; p.flags = flags;
        // B200RC_REMARCH=list (A/B measurements): the flagged pixels were compacted by the b200rc_ert_plan call that wrote this
        // flags buffer; march that list with the per-ray kernel (one warp per 32 flagged rays).  Measured at 2048^3 x2 shards,
        // 77 k flagged pixels (profiles/r1_pass20_remarch_2.txt): list 6.1 ms, slab marcher over the tiles that hold a flagged
        // pixel 2.9 ms -- a warp of the per-ray kernel is a ~40-load dependency chain per sample, the staged tile is not, and
        // the longest-first tile order already drops every tile without a flagged pixel.  The slab marcher stays the default.
        static const bool remarch_list = getenv("B200RC_REMARCH") && !strcmp(getenv("B200RC_REMARCH"), "list");
        if (remarch_list && flags == ctx->plan_flags && px == ctx->plan_n && ctx->d_remarch_list) {
            p.pixel_list = ctx->d_remarch_list; p.pixel_count = ctx->d_tile_counter + 2;
        }
    } else {
        rc = run_eep(ctx, eep, cam, W, H);
        if (rc) return rc;
        ctx->eep_w = W; ctx->eep_h = H;
    }
    p.entry = ctx->d_entry; p.exitp = ctx->d_exit; p.entry_depth = ctx->d_entry_depth; p.exit_depth = ctx->d_exit_depth;
    for (int o = 0; o < 3; ++o) {
        if (!(p.connected[o] && out_color && out_color[o])) p.out[o] = nullptr;
        else p.out[o] = device_out ? out_color[o] : ctx->d_out[o];
    }
    p.out_depth = out_depth ? (device_out ? out_depth : ctx->d_out_depth) : nullptr;
    // sort-last partial MIP / MOP image: out_color[1] (unconnected) receives the merge key plane, one float per pixel
    if (P->partial_image && (p.connected[1] || p.connected[2])) { ctx->set_error("render_camera: a sort-last partial image renders image.output only (outputs 1 and 2 must be unconnected)"); return B200RC_ERR_INVALID; }
    if (P->partial_image && device_out && out_color && out_color[1] && (p.comp[0] == B200RC_COMP_MIP || p.comp[0] == B200RC_COMP_MOP)) {
        // multi-channel MIP composites per channel and adds the results (rc_singlevolume.frag:272-309): one key cannot merge that
        if (nch != 1) { ctx->set_error("render_camera: the mip / mop merge key needs a single-channel volume"); return B200RC_ERR_INVALID; }
        p.out_key = (float*)out_color[1];
    }
    // sort-last partial DVR image: the same plane receives the image's alpha, compact (4 B / pixel), so that the ranks behind this one can
    // read it straight out of peer memory for the exact-ERT plan (b200rc_ert_plan_planes) -- no strided copy, no all-gather
    if (P->partial_image && device_out && out_color && out_color[1] && p.comp[0] == B200RC_COMP_DVR && !flags && !p.simple) p.out_key = (float*)out_color[1];
    rc = run_march(ctx, P, p, nch, dtype);
    if (rc) return rc;
    if (!device_out) {
        rc = copy_outputs(ctx, P, px, out_color, out_depth, cudaMemcpyDeviceToHost, p);
        if (rc) return rc;
    }
    return finish_frame(ctx, P, !flags);
}

extern "C" int b200rc_render_camera(b200rc_ctx* ctx, const b200rc_params* P, const b200rc_eep_params* eep, const b200rc_camera* cam, int W, int H,
                                    void* out_color[B200RC_MAX_OUTPUTS], float* out_depth) {
    CTX_GUARD();
    return render_camera_impl(ctx, P, eep, cam, W, H, out_color, out_depth, false);
}

extern "C" int b200rc_render_camera_device(b200rc_ctx* ctx, const b200rc_params* P, const b200rc_eep_params* eep, const b200rc_camera* cam, int W, int H,
                                           void* out_color_dev[B200RC_MAX_OUTPUTS], float* out_depth_dev) {
    CTX_GUARD();
    return render_camera_impl(ctx, P, eep, cam, W, H, out_color_dev, out_depth_dev, true);
}

// ---- sort-last ---------------------------------------------------------------------------------------------------------------
extern "C" int b200rc_composite(b200rc_ctx* ctx, int mode, const float* front_rgba, const float* front_depth, const float* back_rgba, const float* back_depth,
                                float* dst_rgba, float* dst_depth, size_t n) {
    CTX_GUARD();
    if (!front_rgba || !back_rgba || !dst_rgba) { ctx->set_error("composite: null image"); return B200RC_ERR_INVALID; }
    return rc_composite_launch(ctx, mode, front_rgba, front_depth, back_rgba, back_depth, dst_rgba, dst_depth, n);
}

extern "C" int b200rc_rescale(b200rc_ctx* ctx, const void* src_color, const float* src_depth, int sw, int sh, int format,
                              void* dst_color, float* dst_depth, int dw, int dh, int on_device) {
    CTX_GUARD();
    if (!src_color || !dst_color || sw < 1 || sh < 1 || dw < 1 || dh < 1 || (format != B200RC_OUT_RGBA16 && format != B200RC_OUT_RGBA32F) ||
        ((src_depth == nullptr) != (dst_depth == nullptr))) { ctx->set_error("rescale: bad arguments"); return B200RC_ERR_INVALID; }
    if (on_device) return rc_rescale_launch(ctx, src_color, src_depth, sw, sh, format, dst_color, dst_depth, dw, dh);
    const size_t pb = out_pixel_bytes(format), sp = (size_t)sw * sh, dp = (size_t)dw * dh;
    void* d_src = nullptr; void* d_dst = nullptr; float* d_sd = nullptr; float* d_dd = nullptr;
    int rc = B200RC_OK;
    if (cudaMalloc(&d_src, sp * pb) != cudaSuccess || cudaMalloc(&d_dst, dp * pb) != cudaSuccess ||
        (src_depth && (cudaMalloc(&d_sd, sp * sizeof(float)) != cudaSuccess || cudaMalloc(&d_dd, dp * sizeof(float)) != cudaSuccess))) {
        cudaGetLastError(); ctx->set_error("rescale: out of device memory"); rc = B200RC_ERR_OOM;
    }
    if (!rc && cudaMemcpyAsync(d_src, src_color, sp * pb, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = B200RC_ERR_CUDA;
    if (!rc && src_depth && cudaMemcpyAsync(d_sd, src_depth, sp * sizeof(float), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = B200RC_ERR_CUDA;
    if (!rc) rc = rc_rescale_launch(ctx, d_src, d_sd, sw, sh, format, d_dst, d_dd, dw, dh);
    if (!rc && cudaMemcpyAsync(dst_color, d_dst, dp * pb, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = B200RC_ERR_CUDA;
    if (!rc && dst_depth && cudaMemcpyAsync(dst_depth, d_dd, dp * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = B200RC_ERR_CUDA;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) rc = B200RC_ERR_CUDA;
    if (rc == B200RC_ERR_CUDA) ctx->set_error("rescale: copy failed");
    cudaFree(d_src); cudaFree(d_dst); cudaFree(d_sd); cudaFree(d_dd);
    return rc;
}

extern "C" int b200rc_composite_keyed(b200rc_ctx* ctx, const float* front_rgba, const float* front_key, const float* front_depth,
                                      const float* back_rgba, const float* back_key, const float* back_depth,
                                      float* dst_rgba, float* dst_key, float* dst_depth, size_t n) {
    CTX_GUARD();
    if (!front_rgba || !back_rgba || !dst_rgba || !front_key || !back_key || !dst_key) { ctx->set_error("composite_keyed: null image"); return B200RC_ERR_INVALID; }
    return rc_composite_keyed_launch(ctx, front_rgba, front_key, front_depth, back_rgba, back_key, back_depth, dst_rgba, dst_key, dst_depth, n);
}

extern "C" int b200rc_finalize(b200rc_ctx* ctx, const float* rgba_dev, const float* depth_dev, size_t n, int out_format, int ert_snap, void* out_color_dev) {
    CTX_GUARD();
    if (!rgba_dev || !out_color_dev) { ctx->set_error("finalize: null image"); return B200RC_ERR_INVALID; }
    return rc_finalize_launch(ctx, rgba_dev, depth_dev, n, out_format, ert_snap, out_color_dev);
}

extern "C" int b200rc_ert_plan(b200rc_ctx* ctx, const float* alphas_dev, int world, uint32_t front_mask, float* color_dev, float* depth_dev, size_t n,
                               float* prefix_dev, uint8_t* flags_dev, uint64_t* remarch_count) {
    CTX_GUARD();
    if (!alphas_dev || !color_dev || !depth_dev || !prefix_dev || !flags_dev || world < 1 || world > 32) { ctx->set_error("ert_plan: bad arguments"); return B200RC_ERR_INVALID; }
    unsigned long long* counter = nullptr;
    if (remarch_count) { counter = ctx->d_stats + 7; RC_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream)); }
    const int rc = rc_ert_plan_launch(ctx, alphas_dev, nullptr, world, front_mask, color_dev, depth_dev, n, prefix_dev, flags_dev, counter);
    if (rc) return rc;
    if (remarch_count) {
        unsigned long long h = 0;
        RC_CHECK(cudaMemcpyAsync(&h, counter, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        RC_CHECK(cudaStreamSynchronize(ctx->stream));
        *remarch_count = h;
    }
    return B200RC_OK;
}

extern "C" int b200rc_ert_plan_planes(b200rc_ctx* ctx, const float* const* alpha_planes_dev, int world, uint32_t front_mask, float* color_dev, float* depth_dev,
                                      size_t n, float* prefix_dev, uint8_t* flags_dev, uint64_t* remarch_count) {
    CTX_GUARD();
    if (!alpha_planes_dev || !color_dev || !depth_dev || !prefix_dev || !flags_dev || world < 1 || world > 32) { ctx->set_error("ert_plan_planes: bad arguments"); return B200RC_ERR_INVALID; }
    for (int r = 0; r < world; ++r)
        if (((front_mask >> r) & 1u) && !alpha_planes_dev[r]) { ctx->set_error("ert_plan_planes: no alpha plane for a rank in front_mask"); return B200RC_ERR_INVALID; }
    unsigned long long* counter = nullptr;
    if (remarch_count) { counter = ctx->d_stats + 7; RC_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream)); }
    const int rc = rc_ert_plan_launch(ctx, nullptr, alpha_planes_dev, world, front_mask, color_dev, depth_dev, n, prefix_dev, flags_dev, counter);
    if (rc) return rc;
    if (remarch_count) {
        unsigned long long h = 0;
        RC_CHECK(cudaMemcpyAsync(&h, counter, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        RC_CHECK(cudaStreamSynchronize(ctx->stream));
        *remarch_count = h;
    }
    return B200RC_OK;
}

extern "C" int b200rc_remarch_device(b200rc_ctx* ctx, const b200rc_params* P, const b200rc_camera* cam, int W, int H, const float* prefix_dev,
                                     const uint8_t* flags_dev, void* out_color_dev[B200RC_MAX_OUTPUTS], float* out_depth_dev) {
    CTX_GUARD();
    if (!prefix_dev || !flags_dev) { ctx->set_error("remarch: prefix / flags missing"); return B200RC_ERR_INVALID; }
    if (!P || !P->partial_image) { ctx->set_error("remarch: needs partial_image rendering"); return B200RC_ERR_INVALID; }
    return render_camera_impl(ctx, P, nullptr, cam, W, H, out_color_dev, out_depth_dev, true, prefix_dev, flags_dev);
}

// ---- image-space post chain -----------------------------------------------------------------------------------------------------
// host <-> device staging for the per-pixel image maps (stream-ordered allocations; nothing persists)
struct PostBuf {
    b200rc_ctx* ctx; void* d = nullptr; size_t bytes = 0; bool ok = true;
    PostBuf(b200rc_ctx* c, const void* host, size_t n, bool upload) : ctx(c), bytes(n) {
        if (!n) return;
        if (cudaMallocAsync(&d, n, ctx->stream) != cudaSuccess) { cudaGetLastError(); d = nullptr; ok = false; return; }
        if (upload && host && cudaMemcpyAsync(d, host, n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { cudaGetLastError(); ok = false; }
    }
    ~PostBuf() { if (d) cudaFreeAsync(d, ctx->stream); }
};

extern "C" int b200rc_image_compositor(b200rc_ctx* ctx, int mode, float wf, float w0, float w1, int add_depth, const void* color0, const float* depth0,
                                       const void* color1, const float* depth1, int format, void* dst_color, float* dst_depth, int W, int H, int on_device) {
    CTX_GUARD();
    if (!color0 || !color1 || !dst_color || W < 1 || H < 1 || mode < B200RC_COMPOSITOR_DEPTH_TEST || mode > B200RC_COMPOSITOR_ADD ||
        (format != B200RC_OUT_RGBA16 && format != B200RC_OUT_RGBA32F)) { ctx->set_error("image_compositor: bad arguments"); return B200RC_ERR_INVALID; }
    const size_t n = (size_t)W * H, cb = n * out_pixel_bytes(format), db = n * sizeof(float);
    if (on_device) return rc_compositor_launch(ctx, mode, wf, w0, w1, add_depth, color0, depth0, color1, depth1, format, dst_color, dst_depth, n);
    PostBuf c0(ctx, color0, cb, true), c1(ctx, color1, cb, true), d0(ctx, depth0, depth0 ? db : 0, true), d1(ctx, depth1, depth1 ? db : 0, true);
    PostBuf oc(ctx, nullptr, cb, false), od(ctx, nullptr, dst_depth ? db : 0, false);
    if (!c0.ok || !c1.ok || !d0.ok || !d1.ok || !oc.ok || !od.ok) { ctx->set_error("image_compositor: out of device memory"); return B200RC_ERR_OOM; }
    int rc = rc_compositor_launch(ctx, mode, wf, w0, w1, add_depth, c0.d, (const float*)d0.d, c1.d, (const float*)d1.d, format, oc.d, (float*)od.d, n);
    if (rc) return rc;
    RC_CHECK(cudaMemcpyAsync(dst_color, oc.d, cb, cudaMemcpyDeviceToHost, ctx->stream));
    if (dst_depth) RC_CHECK(cudaMemcpyAsync(dst_depth, od.d, db, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200RC_OK;
}

extern "C" int b200rc_image_background(b200rc_ctx* ctx, const b200rc_background_params* bg, const void* src_color, const float* src_depth, int format,
                                       void* dst_color, float* dst_depth, int W, int H, int on_device) {
    CTX_GUARD();
    if (!bg || !dst_color || W < 1 || H < 1 || (format != B200RC_OUT_RGBA16 && format != B200RC_OUT_RGBA32F) || bg->mode < B200RC_BACKGROUND_MONOCHROME ||
        bg->mode > B200RC_BACKGROUND_GRADIENT || bg->blend_mode < B200RC_BACKGROUND_BLEND_ALPHA || bg->blend_mode > B200RC_BACKGROUND_BLEND_ADD || bg->angle_deg < 0 || bg->angle_deg > 359) {
        ctx->set_error("image_background: bad arguments"); return B200RC_ERR_INVALID;
    }
    const size_t n = (size_t)W * H, cb = n * out_pixel_bytes(format), db = n * sizeof(float);
    if (on_device) return rc_background_launch(ctx, bg, src_color, src_depth, format, dst_color, dst_depth, W, H);
    PostBuf sc(ctx, src_color, src_color ? cb : 0, true), sd(ctx, src_depth, (src_color && src_depth) ? db : 0, true), oc(ctx, nullptr, cb, false), od(ctx, nullptr, dst_depth ? db : 0, false);
    if (!sc.ok || !sd.ok || !oc.ok || !od.ok) { ctx->set_error("image_background: out of device memory"); return B200RC_ERR_OOM; }
    int rc = rc_background_launch(ctx, bg, sc.d, (const float*)sd.d, format, oc.d, (float*)od.d, W, H);
    if (rc) return rc;
    RC_CHECK(cudaMemcpyAsync(dst_color, oc.d, cb, cudaMemcpyDeviceToHost, ctx->stream));
    if (dst_depth) RC_CHECK(cudaMemcpyAsync(dst_depth, od.d, db, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200RC_OK;
}

// ---- octree / multi-resolution brick pool -------------------------------------------------------------------------------------------
extern "C" int b200rc_octree_build(b200rc_ctx* ctx, int slot, int brick_dim, float homogeneity_threshold, int half_sample) {
    CTX_GUARD();
    if (slot < 0 || slot >= B200RC_MAX_VOLUMES || !ctx->vol[slot].valid) { ctx->set_error("octree_build: volume slot is empty"); return B200RC_ERR_NOT_READY; }
    const VolumeSlot& vs = ctx->vol[slot];
    VolDev v; memset(&v, 0, sizeof v);
    v.bricks = vs.bricks; v.nbx = vs.nbx; v.nby = vs.nby;
    return rc_octree_build(ctx, slot, brick_dim, homogeneity_threshold, half_sample, v, vs);
}
extern "C" int b200rc_octree_get_info(b200rc_ctx* ctx, int slot, b200rc_octree_info* out) {
    CTX_GUARD();
    if (!out) return B200RC_ERR_INVALID;
    return rc_octree_info(ctx, slot, out);
}
extern "C" int b200rc_octree_download(b200rc_ctx* ctx, int slot, int level, uint16_t* avg, uint16_t* mn, uint16_t* mx, uint8_t* flags, uint16_t* bricks) {
    CTX_GUARD();
    return rc_octree_download(ctx, slot, level, avg, mn, mx, flags, bricks);
}
extern "C" int b200rc_octree_render(b200rc_ctx* ctx, int slot, const b200rc_octree_render_params* rp, const b200rc_eep_params* eep, const b200rc_camera* cam,
                                    int W, int H, float* out_rgba32f, float* out_firsthit) {
    CTX_GUARD();
    if (!rp || !eep || !cam || !out_rgba32f || W < 1 || H < 1 || slot != 0) { ctx->set_error("octree_render: bad arguments (the octree of slot 0 is rendered)"); return B200RC_ERR_INVALID; }
    const size_t px = (size_t)W * H;
    int rc = ensure_frame_buffers(ctx, px, 1);
    if (rc) return rc;
    b200rc_eep_params e = *eep; e.world_coords = 0;
    rc = run_eep(ctx, &e, cam, W, H);
    if (rc) return rc;
    RC_CHECK(cudaEventRecord(ctx->ev[2], ctx->stream));
    rc = rc_octree_render(ctx, slot, rp, ctx->d_entry, ctx->d_exit, W, H, (float4*)ctx->d_out[0], out_firsthit ? ctx->d_out_depth : nullptr);
    if (rc) return rc;
    RC_CHECK(cudaEventRecord(ctx->ev[3], ctx->stream));
    RC_CHECK(cudaMemcpyAsync(out_rgba32f, ctx->d_out[0], px * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_firsthit) RC_CHECK(cudaMemcpyAsync(out_firsthit, ctx->d_out_depth, px * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->stats.last_march_ms = ms;
    return B200RC_OK;
}
