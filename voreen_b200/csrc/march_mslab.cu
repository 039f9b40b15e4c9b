// march_mslab.cu -- instantiations + launcher of the multi-volume staging marcher (march_mslab.cuh).
#ifdef MSLAB_THREADS               // A/B builds: 128 = 16x8-pixel tiles, twice the CTAs per SM
#define SLAB_THREADS MSLAB_THREADS
#endif
#include "ctx.hpp"
#include "march_mslab.cuh"
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>

#ifndef MSLAB_SHARDED_DEFAULT
#define MSLAB_SHARDED_DEFAULT true      // B200RC_MSLAB_SHARDED=0 (environment) or -DMSLAB_SHARDED_DEFAULT=false keeps the sort-last passes on the per-ray kernel
#endif
template <int CLS> using MSlabCfg = MarchCfg<B200RC_SAMPLER_EXACT, 1, B200RC_U16, CLS, B200RC_GRAD_CENTRAL, B200RC_SHADE_PHONG, -1, true, true>;

// t_k = fl(t_{k-1} + step) equals k * step for every k <= max_samples iff k * m fits the 24-bit significand, m = odd part of step
static bool mslab_lattice_exact(float step, int max_samples) {
    if (!(step > 0.f) || !std::isfinite(step)) return false;
    int e; const float fr = frexpf(step, &e);
    uint32_t m = (uint32_t)ldexpf(fr, 24);
    while (m && !(m & 1u)) m >>= 1;
    return (unsigned long long)m * (unsigned long long)max_samples < (1ull << 24);
}

template <class C>
static int launch_mslab(b200rc_ctx* ctx, const DevParams& p0) {
    DevParams p = p0;
    p.tiles_x = (((p.W + SLAB_TILE - 1) / SLAB_TILE) + 7) / 8 * 8;
    p.tiles_y = (((p.H + SLAB_TILE_H - 1) / SLAB_TILE_H) + 7) / 8 * 8;
    static int env_depth = -1, env_cap = 0;
    if (env_depth < 0) {
        const char* d = getenv("B200RC_MSLAB_DEPTH"); env_depth = d ? atoi(d) : 0;
        if (env_depth < 2 || env_depth > 32) env_depth = 0;
        const char* c = getenv("B200RC_MSLAB_KB"); env_cap = c ? atoi(c) : 0;        // per-volume region, KiB (0: what the SM's share allows)
    }
    auto kern = march_mslab_kernel<C>;
    cudaFuncAttributes fa;
    RC_CHECK(cudaFuncGetAttributes(&fa, kern));
    // MSLAB_MIN_CTAS CTAs share the SM's 227 KB; the nv regions of a CTA split what is left after the static scratch
    const int share_kb = (int)((227 * 1024 / MSLAB_MIN_CTAS - 1024 - (long long)fa.sharedSizeBytes) / 1024);
    int cap_kb = share_kb / p.nvol;
    if (env_cap >= 8 && env_cap < cap_kb) cap_kb = env_cap;
    if (cap_kb < 8) return -1;
    SlabArgs sa; sa.cap_floats = cap_kb * 256; sa.order = nullptr; sa.exact_lattice = mslab_lattice_exact(p.step, 65537) ? 1 : 0;
    // slab depth: chosen per tile in the kernel -- the deepest slab (4..12 planes) whose region fits the capacity, given the lateral drift
    // of the tile's rays per plane and `ext`, the lateral extent a tile's region has at depth 0: 16 pixels' footprint in voxels (+ 20 %
    // perspective spread) + the gradient apron and slack (calibrated on config C4: profiles/r2_c4_mslab_depth_sweep.txt);
    // B200RC_MSLAB_DEPTH fixes the depth for A/B runs
    {
        const float dmax = fmaxf(p.vol[0].dimf[0], fmaxf(p.vol[0].dimf[1], p.vol[0].dimf[2]));
        const float vpp = 1.2f * dmax / (float)(p.W > p.H ? p.W : p.H);
        sa.ext = (int)(vpp * SLAB_TILE) + 12;
        sa.depth = env_depth; sa.dmin = 4; sa.dmax = 12; sa.prefetch = 0;
    }
    const size_t smem = (size_t)p.nvol * sa.cap_floats * sizeof(float);
    RC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    RC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SLAB_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long tiles = (long long)p.tiles_x * p.tiles_y;
    long long grid = (long long)ctx->num_sms * per_sm;      // persistent: one wave, a multiple of the SM count
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    if (ctx->tile_capacity < (size_t)tiles) {
        cudaFree(ctx->d_tile_cost); cudaFree(ctx->d_tile_order); ctx->d_tile_cost = nullptr; ctx->d_tile_order = nullptr; ctx->tile_capacity = 0;
        RC_CHECK(cudaMalloc(&ctx->d_tile_cost, sizeof(float) * tiles));
        RC_CHECK(cudaMalloc(&ctx->d_tile_order, sizeof(unsigned int) * tiles));
        ctx->tile_capacity = (size_t)tiles;
    }
    slab_tile_cost_kernel<<<(unsigned)tiles, SLAB_THREADS, 0, ctx->stream>>>(p, ctx->d_tile_cost);
    slab_tile_order_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_tile_cost, (unsigned)tiles, ctx->d_tile_order, p.tile_counter);
    ctx->launches += 2;
    sa.order = ctx->d_tile_order;
    kern<<<(unsigned)grid, SLAB_THREADS, smem, ctx->stream>>>(p, sa);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}

// Returns true when the configuration is covered (and launched); the caller falls back to the per-ray multi-volume kernels otherwise.
bool rc_march_mslab(b200rc_ctx* ctx, const DevParams& p, int sampler, int* rc_out) {
    if (!p.multi || !p.eep_shared || sampler != B200RC_SAMPLER_EXACT || p.nvol < 1 || p.nvol > B200RC_MAX_VOLUMES) return false;
    if (!p.connected[0] || p.connected[1] || p.connected[2]) return false;
    if (p.gradient != B200RC_GRAD_CENTRAL || p.comp[0] == B200RC_COMP_FHP) return false;
    if (p.pixel_list) return false;                                       // compacted second pass: per-ray kernel
    if (p.out_key && p.comp[0] != B200RC_COMP_DVR) return false;          // only the DVR alpha plane is exported here
    const VolDev& A = p.vol[0];
    {   // sort-last passes (shards, second pass, alpha-plane export) on this kernel: B200RC_MSLAB_SHARDED=0 keeps them on the per-ray kernel
        static const bool sharded_ok = !(getenv("B200RC_MSLAB_SHARDED") && atoi(getenv("B200RC_MSLAB_SHARDED")) == 0) && MSLAB_SHARDED_DEFAULT;
        if ((A.sharded || p.flags || p.out_key) && !sharded_ok) return false;
    }
    const int cls = p.classification[0];
    if (cls != B200RC_CLASS_TF && cls != B200RC_CLASS_PREINT) return false;
    bool ess_any = false;
    for (int v = 0; v < p.nvol; ++v) {
        const VolDev& B = p.vol[v];
        if (B.dtype != B200RC_U16 || B.nch != 1 || B.filter != B200RC_FILTER_LINEAR || B.clamp != B200RC_CLAMP_TO_EDGE) return false;
        if (p.classification[v] != cls || p.shading[v] != B200RC_SHADE_PHONG) return false;
        if (cls == B200RC_CLASS_TF && !p.cls[v].tf) return false;
        if (cls == B200RC_CLASS_PREINT && !p.cls[v].preint) return false;
        // co-registered: one texture coordinate for all volumes at every lattice sample
        if (memcmp(B.dims, A.dims, sizeof A.dims) || memcmp(B.res_lo, A.res_lo, sizeof A.res_lo) || memcmp(B.res_dims, A.res_dims, sizeof A.res_dims) ||
            memcmp(B.clamp_lo, A.clamp_lo, sizeof A.clamp_lo) || memcmp(B.clamp_hi, A.clamp_hi, sizeof A.clamp_hi) || memcmp(B.w2t, A.w2t, sizeof A.w2t) ||
            memcmp(B.spacing, A.spacing, sizeof A.spacing) || memcmp(B.offset, A.offset, sizeof A.offset) || B.nbx != A.nbx || B.nby != A.nby || B.sharded != A.sharded ||
            memcmp(B.own_lo, A.own_lo, sizeof A.own_lo) || memcmp(B.own_hi, A.own_hi, sizeof A.own_hi)) return false;
        ess_any = ess_any || B.ess;
    }
    if (cls == B200RC_CLASS_PREINT && ess_any) {
        // lastIntensity after skipped samples is re-fetched from the staged footprint: needs <= 1 voxel per lattice step on every axis
        for (int a = 0; a < 3; ++a) {
            const float* r = A.w2t + 4 * a;
            if (p.step * sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * A.dimf[a] > 1.0f) return false;
        }
    }
    int rc = -1;
    if (cls == B200RC_CLASS_TF) rc = launch_mslab<MSlabCfg<B200RC_CLASS_TF>>(ctx, p);
    else rc = launch_mslab<MSlabCfg<B200RC_CLASS_PREINT>>(ctx, p);
    if (rc < 0) return false;
    *rc_out = rc;
    return true;
}
