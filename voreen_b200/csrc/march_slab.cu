// march_slab.cu -- instantiations + launcher of the shared-memory brick-staging marcher (march_slab.cuh).
#include "ctx.hpp"
#include "march_slab.cuh"
#include <cstdlib>

template <int DT, int CLS, int COMP, int SHADE> using SlabCfg = MarchCfg<B200RC_SAMPLER_EXACT, 1, DT, CLS, B200RC_GRAD_CENTRAL, SHADE, COMP, true, false>;

template <class C, bool STATS>
static int launch_slab(b200rc_ctx* ctx, const DevParams& p0) {
    DevParams p = p0;
    p.tiles_x = (((p.W + SLAB_TILE - 1) / SLAB_TILE) + 7) / 8 * 8;
    p.tiles_y = (((p.H + SLAB_TILE_H - 1) / SLAB_TILE_H) + 7) / 8 * 8;
    // shared memory: SLAB_MIN_CTAS CTAs per SM share 227 KB; the 1D TF (+ LUT) sits behind the staged region
    size_t tf_bytes = 0;
    if (C::CLS == B200RC_CLASS_TF && p.cls[0].tf) tf_bytes = (size_t)p.cls[0].tf_width * 4 + 1024;
    if (tf_bytes > 40 * 1024) return -1;            // TF too wide for the staged layout: caller falls back
    static int env_depth = -1, env_cap = -1, env_depth_remarch = 6;
    if (env_depth < 0) {
        // slab depth 0 (default): chosen per tile from the rays' lateral drift and the region capacity -- 19.3 -> 18.4 ms over the 36-view
        // orbit of C3 against the best fixed depth of round 1, 6 planes (profiles/r2_slab_adaptive_depth.txt); B200RC_SLAB_DEPTH=n fixes it
        const char* d = getenv("B200RC_SLAB_DEPTH"); env_depth = d ? atoi(d) : 0;
        const char* c = getenv("B200RC_SLAB_KB"); env_cap = c ? atoi(c) : (SLAB_THREADS == 256 ? 72 : 52);
        if (env_depth != 0 && (env_depth < 2 || env_depth > 32)) env_depth = 6;
        // sort-last second pass: a handful of flagged rays per tile -> narrow regions, deeper slabs fit and mean fewer rounds
        const char* dr = getenv("B200RC_SLAB_DEPTH_REMARCH"); env_depth_remarch = dr ? atoi(dr) : 6;
        if (env_depth_remarch < 2 || env_depth_remarch > 32) env_depth_remarch = 6;
        if (env_cap < 32 || env_cap > 200) env_cap = 88;
    }
    // SLAB_MIN_CTAS CTAs share the SM's 227 KB: the staged region takes what the TF (+ LUT) leaves of one CTA's share
    int cap_kb = env_cap;
    // per CTA: 1 KB reserved by the runtime + this instantiation's static shared memory (scheduling scratch, MIP cell maxima)
    auto kern = march_slab_kernel<C, STATS>;
    cudaFuncAttributes fa;
    RC_CHECK(cudaFuncGetAttributes(&fa, kern));
    const int share_kb = (int)((227 * 1024 / SLAB_MIN_CTAS - 1024 - (long long)fa.sharedSizeBytes) / 1024);
    if (cap_kb + (int)((tf_bytes + 1023) / 1024) > share_kb) cap_kb = share_kb - (int)((tf_bytes + 1023) / 1024);
    SlabArgs sa; sa.depth = p0.flags ? env_depth_remarch : env_depth; sa.cap_floats = cap_kb * 256; sa.exact_lattice = 0;
    { static const int env_pf = getenv("B200RC_SLAB_PREFETCH") ? atoi(getenv("B200RC_SLAB_PREFETCH")) : 0; sa.prefetch = env_pf; }
    {   // lateral extent of a tile's region at depth 0: 16 pixels' footprint in voxels (+ 20 % perspective spread) + apron and slack
        static const int env_ext = getenv("B200RC_SLAB_EXT") ? atoi(getenv("B200RC_SLAB_EXT")) : 16;
        const float dmax = fmaxf(p.vol[0].dimf[0], fmaxf(p.vol[0].dimf[1], p.vol[0].dimf[2]));
        sa.ext = (int)(1.2f * dmax / (float)(p.W > p.H ? p.W : p.H) * SLAB_TILE) + env_ext;
        static const int env_dmin = getenv("B200RC_SLAB_DMIN") ? atoi(getenv("B200RC_SLAB_DMIN")) : 4, env_dmax = getenv("B200RC_SLAB_DMAX") ? atoi(getenv("B200RC_SLAB_DMAX")) : 12;
        sa.dmin = env_dmin; sa.dmax = env_dmax;
    }
    // pre-integration + empty-space skipping restores lastIntensity from the staged footprint: needs <= 1 voxel per step
    if (C::CLS == B200RC_CLASS_PREINT && p.vol[0].ess) {
        const float m = fmaxf(p.vol[0].dimf[0], fmaxf(p.vol[0].dimf[1], p.vol[0].dimf[2]));
        if (p.step * m > 1.0f) return -1;
    }
    const size_t smem = (size_t)sa.cap_floats * sizeof(float) + tf_bytes;
    RC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    RC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SLAB_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    { static const int env_per_sm = getenv("B200RC_SLAB_CTAS_PER_SM") ? atoi(getenv("B200RC_SLAB_CTAS_PER_SM")) : 0;     // experiments
      if (env_per_sm >= 1 && env_per_sm < per_sm) per_sm = env_per_sm; }
    const long long tiles = (long long)p.tiles_x * p.tiles_y;
    long long grid = (long long)ctx->num_sms * per_sm;      // persistent: one wave, a multiple of the SM count
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    // tile order: most expensive first, tiles without work dropped (also resets the scheduler counter)
    if (ctx->tile_capacity < (size_t)tiles) {
        cudaFree(ctx->d_tile_cost); cudaFree(ctx->d_tile_order); ctx->d_tile_cost = nullptr; ctx->d_tile_order = nullptr; ctx->tile_capacity = 0;
        RC_CHECK(cudaMalloc(&ctx->d_tile_cost, sizeof(float) * tiles));
        RC_CHECK(cudaMalloc(&ctx->d_tile_order, sizeof(unsigned int) * tiles));
        ctx->tile_capacity = (size_t)tiles;
    }
    slab_tile_cost_kernel<<<(unsigned)tiles, SLAB_THREADS, 0, ctx->stream>>>(p, ctx->d_tile_cost);
    static const bool plain_order = getenv("B200RC_TILE_ORDER") && atoi(getenv("B200RC_TILE_ORDER")) == 0;     // A/B measurements
    if (plain_order) slab_tile_order_plain_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_tile_cost, (unsigned)tiles, ctx->d_tile_order, p.tile_counter);
    else slab_tile_order_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_tile_cost, (unsigned)tiles, ctx->d_tile_order, p.tile_counter);
    ctx->launches += 2;
    sa.order = ctx->d_tile_order;
    kern<<<(unsigned)grid, SLAB_THREADS, smem, ctx->stream>>>(p, sa);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}

#define SLAB_CASE(DT, CLS, COMP, SHADE)                                                                       \
    if ((DT == 0 || dtype == DT) && cls == CLS && comp == COMP && (SHADE < 0 || p.shading[0] == SHADE)) {    \
        const int rc = p.collect_stats ? launch_slab<SlabCfg<DT, CLS, COMP, SHADE>, true>(ctx, p)           \
                                       : launch_slab<SlabCfg<DT, CLS, COMP, SHADE>, false>(ctx, p);         \
        if (rc >= 0) { *rc_out = rc; return true; }                                                          \
    }

// Returns true when the configuration is covered (and launched); the caller falls back to the per-ray kernels otherwise.
bool rc_march_slab(b200rc_ctx* ctx, const DevParams& p, int sampler, int nch, int dtype, int* rc_out) {
    if (sampler != B200RC_SAMPLER_EXACT || p.multi || nch != 1 || p.channel_shift) return false;
    if (!p.connected[0] || p.connected[1] || p.connected[2]) return false;
    if (p.gradient != B200RC_GRAD_CENTRAL) return false;
    if (p.vol[0].filter != B200RC_FILTER_LINEAR || p.vol[0].clamp != B200RC_CLAMP_TO_EDGE) return false;
    const int cls = p.classification[0], comp = p.comp[0];
    SLAB_CASE(B200RC_U16, B200RC_CLASS_PREINT, B200RC_COMP_DVR, B200RC_SHADE_PHONG)
    SLAB_CASE(B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_DVR, B200RC_SHADE_PHONG)
    SLAB_CASE(B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_MIP, B200RC_SHADE_PHONG)
    SLAB_CASE(B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_ISO, B200RC_SHADE_PHONG)
    // no pre-integrated MIP instantiation: SingleVolumeRaycaster switches pre-integration off when an output composites with MIP
    // (singlevolumeraycaster.cpp:523-527); callers of the C ABI that ask for it anyway get the run-time-switched per-ray kernel
    SLAB_CASE(B200RC_U16, B200RC_CLASS_PREINT, B200RC_COMP_ISO, B200RC_SHADE_PHONG)
    SLAB_CASE(B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_MOP, B200RC_SHADE_PHONG)
    SLAB_CASE(B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_MIDA, B200RC_SHADE_PHONG)
    SLAB_CASE(B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_FHP, B200RC_SHADE_PHONG)
    SLAB_CASE(B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_FHN, B200RC_SHADE_PHONG)
    SLAB_CASE(0, B200RC_CLASS_TF, B200RC_COMP_DVR, -1)
    SLAB_CASE(0, B200RC_CLASS_PREINT, B200RC_COMP_DVR, -1)
    return false;
}
