// march_slab.cuh -- the shared-memory brick-staging marcher for the hot configuration (one u8/u16/f32 channel, linear
// filtering, clamp-to-edge, one connected output, central-difference gradients).
//
// Why it exists (profiles/r1_pass1_*): the per-ray marcher of march.cuh is instruction-issue bound -- 76 % issue-slot
// utilisation, ~1300 warp instructions per shaded sample, half of them 64-bit brick addressing, index clamps and
// u16->f32 conversions for the 8 + 32 taps, and its gathers cost ~20 sectors per request in L1.  HBM traffic is
// irrelevant at that point (2.8 % of peak).  Here one CTA owns a 16x16 pixel tile and marches all 256 rays through the
// volume in lock-step SLABS along the tile's major axis: per slab the CTA stages the sub-volume its rays traverse
// (+ the -1..+2 voxel apron of the gradient footprint) into shared memory ONCE, already clamped to the volume and
// already normalised to fp32 exactly as the sampler does (raw * 1/65535), and every tap of every sample in the slab
// is then a 32-bit shared-memory load at a constant offset from one base address.  The arithmetic per sample -- sample
// lattice, trilinear weights, classification, shading, compositing, early ray termination, empty-space skipping -- is
// the same source as march.cuh, so images are identical to the per-ray kernel.
//
// A sample is taken from the staged region only if its whole 4x4x4 footprint lies inside it (exact integer test); rays
// whose next sample is outside simply wait for a later slab.  The region always contains the footprint of the ray that
// is furthest behind, so every round makes progress.
#pragma once
#include "march.cuh"

#ifndef SLAB_THREADS
#define SLAB_THREADS 256        // 256: 16x16-pixel tiles, 128: 16x8-pixel tiles
#endif
#define SLAB_TILE 16            // pixels per tile edge (x)
#define SLAB_TILE_H (SLAB_THREADS / 16)     // tile rows: each warp owns an 8x4 sub-tile, two warps side by side
// 3 CTAs (24 warps) per SM: ptxas fits the kernel into 80 registers with 16-40 bytes of spills; measured against 2 CTAs/SM
// (123 registers, 88 KB regions) at 2048^3: TF 18.7 -> 16.2 ms, pre-integration 23.2 -> 20.7 ms (profiles/r1_pass11_occupancy.txt)
#ifndef SLAB_MIN_CTAS
#define SLAB_MIN_CTAS (768 / SLAB_THREADS)
#endif
// per-sample counters (classified / shaded / skipped): only the STATS instantiation of the kernel keeps them -- three registers and
// two increments per sample are worth 1.1 % of the frame (profiles/r2_slab_nostats_ab.txt); b200rc_params.collect_stats selects it
#define SLAB_STAT(x) do { if (STATS) { x; } } while (0)
#ifndef SLAB_STAGE_BATCH
#define SLAB_STAGE_BATCH 4      // 16-byte loads in flight per thread while staging
#endif
#define SLAB_OCC_WORDS 160      // (max lateral extent / 8 + 2)^2 occupancy words per region
#define SLAB_CMAX 512           // staged level-0 cell maxima per region (MIP culling)

struct SlabShared {
    int zmin;                   // min over rays of the direction-normalised major index of their next sample
    int leader;                 // thread id of a ray attaining zmin
    int lo[3], hi[3];           // bounding box (base indices) of the participating ray segments
    int rlo[3], rn[3];          // staged region: origin (voxel index) and extent
    int pitch;                  // floats per staged row (odd: conflict-free staging stores)
    unsigned int tile;
    int axis_vote[SLAB_THREADS / 32];
    // occupancy bits of the level-0 ESS cells overlapping the region: word (jz * occ_ny + jy), bit jx, where
    // j = (l + occ_off) >> 3 for the region-local voxel index l
    int occ_off[3], occ_ny;
    int cm_nx, cm_ok;           // MIP culling: cells per x-row of the staged cell maxima; 0 when the region has more cells than the table
    float dom_rcp;              // refined reciprocal of the TF domain span (rw_to_tex_rcp), 0 = generic division
    float step_rcp;             // ~1 / sampling step (first guess of the lattice index behind an empty cell; fixed up exactly)
    unsigned int occ[SLAB_OCC_WORDS];
};

struct SlabArgs { int depth; int cap_floats; const unsigned int* order; int exact_lattice; int ext; int dmin, dmax; int prefetch; };   // march_mslab: exact_lattice = (t_k = fl(t_{k-1} + step) == k * step for every k); ext = expected lateral extent of a tile's region at depth 0 (depth <= 0: per-tile depth)   // order: tile ids, most expensive first; p.tile_counter[1] of them

__device__ __forceinline__ int pick3(int a, int b, int c, int axis) { return axis == 0 ? a : (axis == 1 ? b : c); }
__device__ __forceinline__ float pick3f(float a, float b, float c, int axis) { return axis == 0 ? a : (axis == 1 ? b : c); }

template <int DT>
__device__ __forceinline__ float ld_norm_raw(const void* bricks, long long e, int rt_dtype) {
    const int dt = DT ? DT : rt_dtype;
    if (dt == B200RC_U16) return xmul((float)__ldg((const unsigned short*)bricks + e), 1.0f / 65535.0f);
    if (dt == B200RC_U8) return xmul((float)__ldg((const unsigned char*)bricks + e), 1.0f / 255.0f);
    return __ldg((const float*)bricks + e);
}

// exact u16 -> float on the integer + fma pipes: 0x4B000000 | v is the float 2^23 + v
__device__ __forceinline__ float u16lo_norm(unsigned w) { return xmul(xsub(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)), 8388608.0f), 1.0f / 65535.0f); }
__device__ __forceinline__ float u16hi_norm(unsigned w) { return xmul(xsub(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)), 8388608.0f), 1.0f / 65535.0f); }

// Cooperative staging: region voxel (lx,ly,lz) <- normalised voxel clamp(rlo + l, clamp_lo, clamp_hi).
// u16: one item = one 8-voxel brick row segment (a single 16-byte load); a warp takes 32 consecutive (ly,lz) rows of the
// same x chunk, so 8 lanes share one 128-byte line of the brick pool and -- the row pitch being odd -- the eight
// 32-bit stores of the warp hit 32 different banks.
// L2 prefetch of the region the NEXT round is expected to stage: this round's region moved by (ax, ay, az) voxels (one slab along the
// major axis).  One prefetch per 128-byte brick line (8 rows of a brick share one); issued right after staging, so the DRAM reads
// overlap the march phase.  A hint only: a wrong prediction costs bandwidth, never correctness.
__device__ __forceinline__ void slab_prefetch_next(const VolDev& v, const SlabShared& sh, int ax, int ay, int az) {
    const int nx = sh.rn[0], ny = sh.rn[1], nz = sh.rn[2];
    const int xr0 = sh.rlo[0] + ax - v.res_lo[0];
    const int c_first = xr0 >> 3, nchunks = ((xr0 + nx - 1) >> 3) - c_first + 1;
    const int y0 = sh.rlo[1] + ay, z0 = sh.rlo[2] + az;
    const int yb0 = (min(max(y0, v.clamp_lo[1]), v.clamp_hi[1]) - v.res_lo[1]) >> 3, yb1 = (min(max(y0 + ny - 1, v.clamp_lo[1]), v.clamp_hi[1]) - v.res_lo[1]) >> 3;
    const int nyb = yb1 - yb0 + 1;
    const int items = nchunks * nyb * nz;                            // (x chunk, y brick, z plane): one 128-byte line each
    for (int it = threadIdx.x; it < items; it += SLAB_THREADS) {
        const int c = it % nchunks, r = it / nchunks, yb = r % nyb, lz = r / nyb;
        const int z = min(max(z0 + lz, v.clamp_lo[2]), v.clamp_hi[2]) - v.res_lo[2];
        const int cxc = min(max(c_first + c, 0), v.nbx - 1);
        const long long e = (long long)cxc * 512 + brick_off_y((yb0 + yb) << 3, v.nbx) + brick_off_z(z, v.nbx, v.nby);
        asm volatile("prefetch.global.L2 [%0];" ::"l"((const unsigned short*)v.bricks + e));
    }
}

template <class C>
__device__ __forceinline__ void slab_stage(const VolDev& v, const SlabShared& sh, float* __restrict__ S) {
    const int nx = sh.rn[0], ny = sh.rn[1], nz = sh.rn[2], pitch = sh.pitch;
    const int rows = ny * nz;
    const int dt = C::DT ? C::DT : v.dtype;
    if (dt == B200RC_U16) {
        const int xr0 = sh.rlo[0] - v.res_lo[0];                    // resident x of region column 0
        const int c_first = xr0 >> 3;                               // arithmetic shift: floor for negative x
        const int nchunks = ((xr0 + nx - 1) >> 3) - c_first + 1;
        const float rcp_ny = 1.0f / (float)ny;
        // items = (x chunk, row) pairs, chunk-major; every thread takes SLAB_STAGE_BATCH items per trip and issues all their
        // 16-byte loads before it converts and stores the first one (the loads of a trip are in flight together)
        const int items = nchunks * rows;
        const float rcp_rows = 1.0f / (float)rows;
        for (int base = threadIdx.x; base < items; base += SLAB_THREADS * SLAB_STAGE_BATCH) {
            uint4 q[SLAB_STAGE_BATCH]; int rr[SLAB_STAGE_BATCH], cxs[SLAB_STAGE_BATCH];
#pragma unroll
            for (int k = 0; k < SLAB_STAGE_BATCH; ++k) {
                const int it = base + k * SLAB_THREADS;
                rr[k] = -1; cxs[k] = 0; q[k] = make_uint4(0u, 0u, 0u, 0u);
                if (it < items) {
                    const int c = __float2int_rd(((float)it + 0.5f) * rcp_rows), r = it - c * rows;         // it / rows (exact: a half-integer over rows is never near an integer)
                    const int lz = __float2int_rd(((float)r + 0.5f) * rcp_ny), ly = r - lz * ny;            // r / ny (exact: (r + .5) / ny is never near an integer)
                    const int y = min(max(sh.rlo[1] + ly, v.clamp_lo[1]), v.clamp_hi[1]) - v.res_lo[1];
                    const int z = min(max(sh.rlo[2] + lz, v.clamp_lo[2]), v.clamp_hi[2]) - v.res_lo[2];
                    const int cx = c_first + c;
                    const int cxc = min(max(cx, 0), v.nbx - 1);
                    const long long e = (long long)cxc * 512 + brick_off_y(y, v.nbx) + brick_off_z(z, v.nbx, v.nby);
                    q[k] = __ldg((const uint4*)((const unsigned short*)v.bricks + e));
                    rr[k] = r; cxs[k] = cx;
                }
            }
#pragma unroll
            for (int k = 0; k < SLAB_STAGE_BATCH; ++k) {
                if (rr[k] < 0) continue;
                const int cx = cxs[k];
                float f[8] = {u16lo_norm(q[k].x), u16hi_norm(q[k].x), u16lo_norm(q[k].y), u16hi_norm(q[k].y), u16lo_norm(q[k].z), u16hi_norm(q[k].z), u16lo_norm(q[k].w), u16hi_norm(q[k].w)};
                // a chunk entirely left of the resident range replicates its first voxel (CLAMP_TO_EDGE / shard clamp); the pool
                // itself replicates the last voxel of a row into the brick padding on the right
                if (cx < 0) { for (int i = 1; i < 8; ++i) f[i] = f[0]; }
                else if (cx > v.nbx - 1) { for (int i = 0; i < 7; ++i) f[i] = f[7]; }
                float* row = S + rr[k] * pitch;
                const int lx0 = cx * 8 - xr0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int lx = lx0 + i;
                    if ((unsigned)lx < (unsigned)nx) row[lx] = f[i];
                }
            }
        }
    } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int r = warp; r < rows; r += SLAB_THREADS / 32) {
            const int lz = r / ny, ly = r - lz * ny;
            const int y = min(max(sh.rlo[1] + ly, v.clamp_lo[1]), v.clamp_hi[1]) - v.res_lo[1];
            const int z = min(max(sh.rlo[2] + lz, v.clamp_lo[2]), v.clamp_hi[2]) - v.res_lo[2];
            const long long oyz = brick_off_y(y, v.nbx) + brick_off_z(z, v.nbx, v.nby);
            float* row = S + r * pitch;
            for (int lx = lane; lx < nx; lx += 32) {
                const int x = min(max(sh.rlo[0] + lx, v.clamp_lo[0]), v.clamp_hi[0]) - v.res_lo[0];
                row[lx] = ld_norm_raw<C::DT>(v.bricks, brick_off_x(x) + oyz, v.dtype);
            }
        }
    }
}

// occupancy bits of the region's level-0 cells -> shared memory (one word per (jy, jz) row of cells)
template <bool CULL>
__device__ __forceinline__ void slab_stage_occ(const VolDev& v, SlabShared& sh, float* __restrict__ cmax, unsigned int* __restrict__ occ_dst) {
    int off[3], n[3], c0[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int rl = sh.rlo[a] - v.res_lo[a];                     // the grids cover the resident range: resident coordinates
        off[a] = rl & 7;                                            // table cell j covers resident voxels rl - off + 8 j .. + 7
        n[a] = (sh.rn[a] - 1 + off[a]) / 8 + 1;
        c0[a] = (rl - off[a]) >> 3;                                 // (possibly negative) cell index of table cell 0
    }
    if (threadIdx.x == 0) { sh.occ_off[0] = off[0]; sh.occ_off[1] = off[1]; sh.occ_off[2] = off[2]; sh.occ_ny = n[1]; }
    const int words = n[1] * n[2];
    for (int w = threadIdx.x; w < words && w < SLAB_OCC_WORDS; w += SLAB_THREADS) {
        const int jz = w / n[1], jy = w - jz * n[1];
        const int cy = min(max(c0[1] + jy, 0), v.occ0_dims[1] - 1), cz = min(max(c0[2] + jz, 0), v.occ0_dims[2] - 1);
        const uint32_t* rowp = v.occ0 + ((size_t)cz * v.occ0_dims[1] + cy) * v.occ0_wpr;
        unsigned bits = 0;
        for (int jx = 0; jx < n[0] && jx < 32; ++jx) {
            const int cx = min(max(c0[0] + jx, 0), v.occ0_dims[0] - 1);
            bits |= ((__ldg(rowp + (cx >> 5)) >> (cx & 31)) & 1u) << jx;
        }
        occ_dst[w] = bits;
    }
    if (CULL && v.mip_cull) {       // level-0 cell maxima of the same cells (normalised texture values)
        const int tot = n[0] * n[1] * n[2];
        if (threadIdx.x == 0) { sh.cm_nx = n[0]; sh.cm_ok = (tot <= SLAB_CMAX && words <= SLAB_OCC_WORDS && n[0] <= 32) ? 1 : 0; }
        if (tot <= SLAB_CMAX)
            for (int i = threadIdx.x; i < tot; i += SLAB_THREADS) {
                const int jx = i % n[0], jy = (i / n[0]) % n[1], jz = i / (n[0] * n[1]);
                const int cx = min(max(c0[0] + jx, 0), v.occ0_dims[0] - 1), cy = min(max(c0[1] + jy, 0), v.occ0_dims[1] - 1), cz = min(max(c0[2] + jz, 0), v.occ0_dims[2] - 1);
                cmax[i] = __ldg(&v.minmax[(((size_t)cz * v.occ0_dims[1] + cy) * v.occ0_dims[0] + cx) * v.nch].y);
            }
    }
}

// jump over the empty cell the sample at t lies in: returns false when the cell is not empty
template <bool CULL, bool STATS>
__device__ __forceinline__ bool slab_skip_empty(const DevParams& p, const VolDev& V, float3 first, float3 dir, float tEnd, float step,
                                                float& t, float& t_last, int& loop, bool& finished, RayStats& st, float maxI) {
    const float3 sp = f3(xadd(first.x, xmul(t, dir.x)), xadd(first.y, xmul(t, dir.y)), xadd(first.z, xmul(t, dir.z)));
    float t_exit = 0.f;
    if (!((CULL && V.mip_cull) ? ess_empty_cell_t<true>(V, sp, dir, t, maxI, t_exit) : ess_empty_cell(V, sp, dir, t, t_exit))) return false;
    if (t_exit > tEnd) { SLAB_STAT(st.skipped += (unsigned)floorf((tEnd - t) / step) + 1u); finished = true; return true; }
    int n = 1; float tn;
    if (p.lattice_exact) {
        int k1 = max(loop + 1, __float2int_ru(t_exit / step));
        while ((float)k1 * step < t_exit) ++k1;
        while (k1 - 1 > loop && (float)(k1 - 1) * step >= t_exit) --k1;
        n = k1 - loop; tn = (float)k1 * step;
        t_last = (float)(k1 - 1) * step;
    } else {
        t_last = t;
        tn = xadd(t, step);
        while (tn < t_exit && tn <= tEnd) { t_last = tn; tn = xadd(tn, step); ++n; }
    }
    SLAB_STAT(st.skipped += n); loop += n; t = tn;
    finished = (t > tEnd) || loop >= 256 * 256;
    return true;
}

// The same jump for a sample the STAGED occupancy bits already classified as lying in an empty level-0 cell (march loop): no grid
// lookups, reciprocals instead of divisions.  (su*, f*) = the sample's continuous voxel coordinate u = p * dim - 0.5 and its floor.
// The exit parameter is rounded DOWN (1e-6 relative + directed rounding), so it never lies behind the exit the divisions would
// give: the jump may stop one lattice sample early -- that sample is still inside the empty cell, is tested against the staged bit
// again and skipped by the next call -- but never late.  Pixels and the classified / skipped counters are unchanged.
template <bool STATS>
__device__ __forceinline__ void slab_skip_cell(const DevParams& p, const VolDev& V, float3 dir, float sux, float suy, float suz, float fx, float fy, float fz,
                                               float tEnd, float step, float step_rcp, float& t, float& t_last, int& loop, bool& finished, RayStats& st) {
    float dt = 3.0e38f;
    {
        const float su[3] = {sux, suy, suz}, ff[3] = {fx, fy, fz}, dd[3] = {dir.x, dir.y, dir.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int ir = min(max((int)ff[a], 0), V.dims[a] - 1) - V.res_lo[a];       // resident voxel index, as ess_empty_cell_t
            const int c = ir >> 3;
            const float ur = su[a] - (float)V.res_lo[a];
            const float v = dd[a] * V.dimf[a];
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
            if (v > 0.f) { if (((c + 1) << 3) < V.res_dims[a]) dt = fminf(dt, ((float)((c + 1) << 3) - ur) * r); }
            else if (v < 0.f) { if (c > 0) dt = fminf(dt, ((float)(c << 3) - ur) * r); }
        }
    }
    const float t_exit = __fadd_rd(t, __fmul_rd(fmaxf(dt, 0.f), 0.999999f));
    if (t_exit > tEnd) { SLAB_STAT(st.skipped += (unsigned)floorf((tEnd - t) / step) + 1u); finished = true; return; }
    int n = 1; float tn;
    if (p.lattice_exact) {
        int k1 = max(loop + 1, __float2int_ru(t_exit * step_rcp));
        while ((float)k1 * step < t_exit) ++k1;
        while (k1 - 1 > loop && (float)(k1 - 1) * step >= t_exit) --k1;
        n = k1 - loop; tn = (float)k1 * step;
        t_last = (float)(k1 - 1) * step;
    } else {
        t_last = t;
        tn = xadd(t, step);
        while (tn < t_exit && tn <= tEnd) { t_last = tn; tn = xadd(tn, step); ++n; }
    }
    SLAB_STAT(st.skipped += n); loop += n; t = tn;
    finished = (t > tEnd) || loop >= 256 * 256;
}

// ---- tile scheduling: longest-processing-time first ---------------------------------------------------------------------
// The persistent CTAs pull 16x16 tiles from a counter.  A tile whose rays cross the whole volume costs ~1 ms; handed out
// last it leaves the other CTAs idle for that long (the tail grows relative to the frame when the image is split across
// GPUs).  slab_tile_cost_kernel estimates every tile's work as the summed ray length (inside this rank's shard, of the
// pixels this launch renders), slab_tile_order_kernel sorts the tiles by that estimate, most expensive first, and drops
// the tiles that own no pixel of this launch (another rank's band, no flagged pixel in the second sort-last pass).  Tile ids keep the Z-order numbering of the marcher (8x8 super-tiles).
static __global__ void __launch_bounds__(SLAB_THREADS) slab_tile_cost_kernel(const __grid_constant__ DevParams p, float* __restrict__ cost) {
    __shared__ float part[SLAB_THREADS / 32];
    const unsigned tile = blockIdx.x;
    const unsigned super_x = (unsigned)p.tiles_x >> 3;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned sup = tile >> 6, loc = tile & 63u;
    const unsigned tx = (sup % super_x) * 8u + compact1by1(loc), ty = (sup / super_x) * 8u + compact1by1(loc >> 1);
    float c = 0.f;
    const bool mine = !(p.band_mod > 1 && (int)((ty * SLAB_TILE_H / (unsigned)p.band_rows) % (unsigned)p.band_mod) != p.band_rem);
    const int px = (int)(tx * SLAB_TILE + (warp & 1u) * 8u + (lane & 7u)), py = (int)(ty * SLAB_TILE_H + (warp >> 1) * 4u + (lane >> 3));
    if (mine && px < p.W && py < p.H) {
        const size_t pix = (size_t)py * p.W + px;
        if (!p.flags || p.flags[pix] == 2) {
            const float4 e4 = p.entry[pix], x4 = p.exitp[pix];
            const float3 d = f3(x4.x - e4.x, x4.y - e4.y, x4.z - e4.z);
            const float tEnd = sqrtf(dot3(d, d));
            c = p.step;                             // every pixel this launch owns is written (background included): cost > 0
            if (tEnd > 0.f) {
                c = tEnd + 4.0f * p.step;
                const VolDev& V = p.vol[0];
                if (V.sharded) {
                    const ShardSeg seg = shard_segment(V, f3(e4.x, e4.y, e4.z), d * (1.0f / tEnd), tEnd, p.step);
                    c = seg.any ? fmaxf(seg.t_stop - (float)seg.k0 * p.step, 0.f) + 4.0f * p.step : p.step;
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if (lane == 0) part[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < SLAB_THREADS / 32; ++w) s += part[w];
        cost[tile] = s;
    }
}

// A/B switch (B200RC_TILE_ORDER=0): Z-order numbering as is; only the tiles that own no pixel are dropped
static __global__ void slab_tile_order_plain_kernel(const float* __restrict__ cost, unsigned ntiles, unsigned int* __restrict__ order, unsigned int* __restrict__ counter) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned n = 0;
        for (unsigned i = 0; i < ntiles; ++i) if (cost[i] > 0.f) order[n++] = i;
        counter[0] = 0u; counter[1] = n;
    }
}

#define SLAB_ORDER_BINS 1024
static __global__ void __launch_bounds__(1024) slab_tile_order_kernel(const float* __restrict__ cost, unsigned ntiles, unsigned int* __restrict__ order,
                                                               unsigned int* __restrict__ counter) {
    __shared__ unsigned hist[SLAB_ORDER_BINS];
    __shared__ float wmax[32];
    __shared__ unsigned carry;
    float m = 0.f;
    for (unsigned i = threadIdx.x; i < ntiles; i += blockDim.x) m = fmaxf(m, cost[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
    hist[threadIdx.x] = 0u;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    m = 0.f;
    for (int w = 0; w < 32; ++w) m = fmaxf(m, wmax[w]);
    const float scale = m > 0.f ? (float)(SLAB_ORDER_BINS - 1) / m : 0.f;
    // bin 0 = most expensive
    for (unsigned i = threadIdx.x; i < ntiles; i += blockDim.x) {
        const float c = cost[i];
        if (c > 0.f) atomicAdd(&hist[SLAB_ORDER_BINS - 1 - min((int)(c * scale), SLAB_ORDER_BINS - 1)], 1u);
    }
    __syncthreads();
    // exclusive scan of the 1024 bins: warp scans + carry
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned v = hist[threadIdx.x];
    unsigned inc = v;
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += t; }
    __shared__ unsigned wsum[32];
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned ws = wsum[lane], winc = ws;
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= (unsigned)o) winc += t; }
        wsum[lane] = winc - ws;
        if (lane == 31) carry = winc;
    }
    __syncthreads();
    hist[threadIdx.x] = wsum[warp] + inc - v;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < ntiles; i += blockDim.x) {
        const float c = cost[i];
        if (c > 0.f) order[atomicAdd(&hist[SLAB_ORDER_BINS - 1 - min((int)(c * scale), SLAB_ORDER_BINS - 1)], 1u)] = i;
    }
    if (threadIdx.x == 0) { counter[0] = 0u; counter[1] = carry; }
}

template <class C, bool STATS>
__global__ void __launch_bounds__(SLAB_THREADS, SLAB_MIN_CTAS) march_slab_kernel(const __grid_constant__ DevParams p, const SlabArgs sa) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SlabShared sh;
    constexpr bool CULL = (C::COMP == B200RC_COMP_MIP && C::CLS == B200RC_CLASS_TF);      // MIP: skip cells that cannot beat the running maximum
    __shared__ float sh_cmax[CULL ? SLAB_CMAX : 1];
    float* S = (float*)smem_raw;
    SmemTables sm;
    {   // 1D TF + opacity LUT behind the region (TF classification only; pre-integration reads its 2D table)
        sm.tf[0] = nullptr; sm.lut[0] = nullptr;
        const ClsDev& cd = p.cls[0];
        if (C::CLS == B200RC_CLASS_TF && cd.tf) {
            uchar4* dst = (uchar4*)(smem_raw + (size_t)sa.cap_floats * sizeof(float));
            for (int i = threadIdx.x; i < cd.tf_width; i += blockDim.x) dst[i] = cd.tf[i];
            float* l = (float*)(dst + cd.tf_width);
            for (int i = threadIdx.x; i < 256; i += blockDim.x) l[i] = cd.corr_lut[i];
            sm.tf[0] = dst; sm.lut[0] = l;
        }
        if (threadIdx.x == 0) { sh.zmin = INT_MAX; sh.leader = INT_MAX; for (int a = 0; a < 3; ++a) { sh.lo[a] = INT_MAX; sh.hi[a] = INT_MIN; } sh.dom_rcp = domain_rcp(cd.dom_lo, cd.dom_hi); sh.step_rcp = 1.0f / p.step; }
        __syncthreads();
    }
    const VolDev& V = p.vol[0];
    constexpr int cmode = C::CLS;
    constexpr int comp = C::COMP;
    constexpr bool pi_mode = (cmode == B200RC_CLASS_PREINT);
    const int smode = C::SHADE >= 0 ? C::SHADE : p.shading[0];
    const float step = p.step;
    __shared__ float2 vote_slope[SLAB_THREADS / 32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    RayStats st = {0u, 0u, 0u, 0u, 0ull};
    unsigned long long dbg_part = 0, dbg_rounds = 0;
#ifdef SLAB_PROFILE
    long long pf_sched = 0, pf_stage = 0, pf_march = 0, pf_wait = 0, pf_setup = 0, pf_rounds = 0, pf_bbox = 0, pf_wait2 = 0, pf_leader = 0, pf_t = clock64();
    long long pf_iters = 0, pf_lanes = 0;
#define PF_MARK(acc) { const long long now_ = clock64(); acc += now_ - pf_t; pf_t = now_; }
#else
#define PF_MARK(acc)
#endif
    const unsigned ntiles = (unsigned)p.tiles_x * (unsigned)p.tiles_y;      // 16x16 tiles, padded to 8x8 super-tiles
    const unsigned super_x = (unsigned)p.tiles_x >> 3;

    const unsigned n_active = min(__ldg(p.tile_counter + 1), ntiles);      // tiles with at least one ray to march (slab_tile_order_kernel)
    for (;;) {
        if (threadIdx.x == 0) {
            const unsigned k = atomicAdd(p.tile_counter, 1u);
            sh.tile = k < n_active ? __ldg(sa.order + k) : 0xffffffffu;
        }
        __syncthreads();
        const unsigned tile = sh.tile;
        if (tile >= ntiles) break;
        const unsigned sup = tile >> 6, loc = tile & 63u;
        const unsigned tx = (sup % super_x) * 8u + compact1by1(loc), ty = (sup / super_x) * 8u + compact1by1(loc >> 1);
        if (p.band_mod > 1 && (int)((ty * SLAB_TILE_H / (unsigned)p.band_rows) % (unsigned)p.band_mod) != p.band_rem) { __syncthreads(); continue; }   // another rank's band (uniform)
        // warp w covers the 8x4 sub-tile (w & 1, w >> 1) of the 16x16 tile
        const int px = (int)(tx * SLAB_TILE + (warp & 1u) * 8u + (lane & 7u)), py = (int)(ty * SLAB_TILE_H + (warp >> 1) * 4u + (lane >> 3));
        const bool inside = px < p.W && py < p.H;
        const size_t pix = inside ? (size_t)py * p.W + px : 0;
        float3 first = f3(0.f, 0.f, 0.f), dir = f3(0.f, 0.f, 1.f);
        float tEnd = 0.f;
        bool cast = false;
        const bool second_pass = p.flags != nullptr;
        const bool wanted = inside && (!second_pass || p.flags[pix] == 2);      // sort-last second pass: flagged pixels only
        const float prefixA = (second_pass && wanted) ? p.prefix[pix] : 0.f;
        if (wanted) {
            const float4 e4 = p.entry[pix], x4 = p.exitp[pix];
            first = f3(e4.x, e4.y, e4.z);
            const float3 last = f3(x4.x, x4.y, x4.z);
            cast = !(first.x == last.x && first.y == last.y && first.z == last.z);
            if (cast) {
                const float3 d = f3(xsub(last.x, first.x), xsub(last.y, first.y), xsub(last.z, first.z));
                tEnd = __fsqrt_rn(xadd(xadd(xmul(d.x, d.x), xmul(d.y, d.y)), xmul(d.z, d.z)));
                dir = f3(xdiv(d.x, tEnd), xdiv(d.y, tEnd), xdiv(d.z, tEnd));
                st.rays++;
                st.nominal += (unsigned long long)floorf(xdiv(tEnd, step)) + 1ull;
            }
        }
        float4 result = make_float4(0.f, 0.f, 0.f, prefixA);
        // pre-integration: lastI holds the TF texture coordinate of the previous sample's intensity (lastIntensity starts at 0)
        float t = 0.0f, t_last = 0.0f, tDepth = -1.0f, lastI = pi_mode ? rw_to_tex(p.cls[0].dom_lo, p.cls[0].dom_hi, 0.f) : 0.f, maxI = 0.f;
        int loop = 0;
        bool finished = !cast;
        bool prev_empty = false;       // pre-integration: the previous lattice sample lay in an empty cell
        bool lastI_stale = false;      // ... and lastIntensity has to be re-fetched at t_last before the next classification
        float tStop = tEnd;            // sort-last shard: march only [k0 * step, tStop]
        if (cast && V.sharded) {
            const ShardSeg seg = shard_segment(V, first, dir, tEnd, step);
            tStop = seg.t_stop;
            if (!seg.any) finished = true;
            else if (seg.k0 > 0) {
                loop = seg.k0;
                if (p.lattice_exact) { t = (float)loop * step; t_last = (float)(loop - 1) * step; }
                else for (int i = 0; i < loop; ++i) { t_last = t; t = xadd(t, step); }
                lastI_stale = true;
                finished = (t > tStop);
            }
        }
        const float vx = dir.x * V.dimf[0], vy = dir.y * V.dimf[1], vz = dir.z * V.dimf[2];
        // slab orientation: major axis + travel direction of the tile's first casting ray, uniform over the CTA
        int axis, DEPTH; float sgn;
        {
            const float ax_ = fabsf(vx), ay_ = fabsf(vy), az_ = fabsf(vz);
            const int my_axis = (ax_ >= ay_ && ax_ >= az_) ? 0 : (ay_ >= az_ ? 1 : 2);
            const unsigned m = __ballot_sync(0xffffffffu, cast);
            const int mine = my_axis | (pick3f(vx, vy, vz, my_axis) < 0.f ? 4 : 0);
            const int wa = m ? __shfl_sync(0xffffffffu, mine, __ffs(m) - 1) : 8;
            // lateral slopes of that ray (voxels per major-axis plane): they decide how deep a slab fits the region capacity
            const float vm = fmaxf(fabsf(pick3f(vx, vy, vz, my_axis)), 1e-20f);
            const float my_s1 = fabsf(pick3f(vx, vy, vz, (my_axis + 1) % 3)) / vm, my_s2 = fabsf(pick3f(vx, vy, vz, (my_axis + 2) % 3)) / vm;
            const float ws1 = m ? __shfl_sync(0xffffffffu, my_s1, __ffs(m) - 1) : 0.f, ws2 = m ? __shfl_sync(0xffffffffu, my_s2, __ffs(m) - 1) : 0.f;
            if (lane == 0) { sh.axis_vote[warp] = wa; vote_slope[warp] = make_float2(ws1, ws2); }
            __syncthreads();
            int a = 8;
            float2 sl = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < SLAB_THREADS / 32; ++w) if (a == 8 && sh.axis_vote[w] != 8) { a = sh.axis_vote[w]; sl = vote_slope[w]; }
            if (a == 8) a = 2;
            axis = a & 3;
            sgn = (a & 4) ? -1.f : 1.f;
            DEPTH = sa.depth;
            if (DEPTH <= 0) {       // per tile: the deepest slab (4..12 planes) whose region -- tile extent + drift of the rays + apron -- fits the capacity
                DEPTH = sa.dmin;
                for (int d = sa.dmax; d > sa.dmin; --d) {
                    const int n1 = sa.ext + (int)ceilf((float)d * sl.x), n2 = sa.ext + (int)ceilf((float)d * sl.y);
                    if ((long long)(n1 | 1) * (n2 | 1) * (d + 3) <= sa.cap_floats) { DEPTH = d; break; }
                }
            }
        }
        const float vmaj = fabsf(pick3f(vx, vy, vz, axis));
        // cached plan of the next sample (recomputed only after the ray moved)
        bool stale = true;
        int i0x = 0, i0y = 0, i0z = 0, zc = INT_MAX; float ux = 0.f, uy = 0.f, uz = 0.f;

        PF_MARK(pf_setup)
        // ---- rounds ---------------------------------------------------------------------------------------------------
        for (;;) {
            // (1) direction-normalised major index of this ray's next sample, after jumping over empty cells so that no
            //     slab is staged where the ray would only skip
            if (stale) {
                if (V.ess && (!pi_mode || prev_empty))
                    while (!finished && slab_skip_empty<CULL, STATS>(p, V, first, dir, tStop, step, t, t_last, loop, finished, st, maxI)) { if (pi_mode) lastI_stale = true; }
                zc = INT_MAX;
                if (!finished) {
                    const float3 sp = f3(xadd(first.x, xmul(t, dir.x)), xadd(first.y, xmul(t, dir.y)), xadd(first.z, xmul(t, dir.z)));
                    ux = xsub(xmul(sp.x, V.dimf[0]), 0.5f); uy = xsub(xmul(sp.y, V.dimf[1]), 0.5f); uz = xsub(xmul(sp.z, V.dimf[2]), 0.5f);
                    i0x = (int)floorf(ux); i0y = (int)floorf(uy); i0z = (int)floorf(uz);
                    const int im = pick3(i0x, i0y, i0z, axis);
                    zc = sgn > 0.f ? im : -im;
                }
                stale = false;
            }
            {
                const int wmin = __reduce_min_sync(0xffffffffu, zc);
                if (lane == 0 && wmin != INT_MAX) atomicMin(&sh.zmin, wmin);
            }
            PF_MARK(pf_sched)
            __syncthreads();
            PF_MARK(pf_wait)
            const int zmin = sh.zmin;
            if (zmin == INT_MAX) break;                     // every ray of the tile is done (uniform)
            // (2) bounding box of the participating segments
            const bool part = !finished && zc < zmin + DEPTH;
            {
                int lo0 = INT_MAX, lo1 = INT_MAX, lo2 = INT_MAX, hi0 = INT_MIN, hi1 = INT_MIN, hi2 = INT_MIN;
                if (part) {
                    // ray parameter at which the major coordinate leaves the slab
                    const float um = pick3f(ux, uy, uz, axis);
                    const float dist = (float)(zmin + DEPTH) - (sgn > 0.f ? um : -um);     // > 0, in voxels
                    const float dt = fminf(dist / fmaxf(vmaj, 1e-20f), tStop - t);
                    const int jx = __float2int_rd(ux + dt * vx), jy = __float2int_rd(uy + dt * vy), jz = __float2int_rd(uz + dt * vz);
                    lo0 = min(i0x, jx); hi0 = max(i0x, jx); lo1 = min(i0y, jy); hi1 = max(i0y, jy); lo2 = min(i0z, jz); hi2 = max(i0z, jz);
                }
                lo0 = __reduce_min_sync(0xffffffffu, lo0); lo1 = __reduce_min_sync(0xffffffffu, lo1); lo2 = __reduce_min_sync(0xffffffffu, lo2);
                hi0 = __reduce_max_sync(0xffffffffu, hi0); hi1 = __reduce_max_sync(0xffffffffu, hi1); hi2 = __reduce_max_sync(0xffffffffu, hi2);
                if (lane == 0 && lo0 != INT_MAX) {
                    atomicMin(&sh.lo[0], lo0); atomicMin(&sh.lo[1], lo1); atomicMin(&sh.lo[2], lo2);
                    atomicMax(&sh.hi[0], hi0); atomicMax(&sh.hi[1], hi1); atomicMax(&sh.hi[2], hi2);
                }
                {   // leader = smallest thread id among the rays at the slab front: one shared-memory atomic per warp, not per ray
                    const int cand = __reduce_min_sync(0xffffffffu, (!finished && zc == zmin) ? (int)threadIdx.x : INT_MAX);
                    if (lane == 0 && cand != INT_MAX) atomicMin(&sh.leader, cand);
                }
            }
            PF_MARK(pf_bbox)
            __syncthreads();
            PF_MARK(pf_wait2)
            // (3) the leader fixes the region (footprint apron -1..+2, one voxel of slack) and resets the scratch
            if ((int)threadIdx.x == sh.leader) {
                const int li[3] = {i0x, i0y, i0z};
                int lo[3], hi[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    lo[a] = sh.lo[a] - 2; hi[a] = sh.hi[a] + 3;
                    if (a == axis) {        // the slab itself: base planes [zmin, zmin + DEPTH) in travel direction
                        if (sgn > 0.f) { lo[a] = zmin - 1; hi[a] = zmin + DEPTH - 1 + 2; }
                        else { lo[a] = -(zmin + DEPTH - 1) - 1; hi[a] = -zmin + 2; }
                    }
                    sh.lo[a] = INT_MAX; sh.hi[a] = INT_MIN;
                }
                // capacity: shrink the wider lateral axis around the leader's footprint until the region fits
                const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
                const int nmaj = hi[axis] - lo[axis] + 1;
                for (int k = 0; k < 2; ++k) {       // no lateral extent beyond 64 voxels (occupancy table, odd pitch)
                    const int as = k ? a2 : a1;
                    if (hi[as] - lo[as] + 1 > 64) {
                        int l2 = li[as] - 32;
                        l2 = max(l2, lo[as]); l2 = min(l2, hi[as] - 63);
                        lo[as] = l2; hi[as] = l2 + 63;
                    }
                }
                for (int iter = 0; iter < 4; ++iter) {
                    const int n1 = hi[a1] - lo[a1] + 1, n2 = hi[a2] - lo[a2] + 1;
                    const int p1 = (a1 == 0) ? (n1 | 1) : n1, p2 = (a2 == 0) ? (n2 | 1) : n2, pm = (axis == 0) ? (nmaj | 1) : nmaj;
                    if ((long long)p1 * p2 * pm <= sa.cap_floats) break;
                    const int as = n1 >= n2 ? a1 : a2, other = n1 >= n2 ? p2 : p1;
                    int cap = sa.cap_floats / (other * pm);
                    if (as == 0) cap -= 1;      // the odd pitch (n | 1) must still fit
                    cap = max(cap, 8);
                    int l2 = li[as] - cap / 2;
                    l2 = max(l2, lo[as]); l2 = min(l2, hi[as] - cap + 1);
                    lo[as] = l2; hi[as] = l2 + cap - 1;
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) { sh.rlo[a] = lo[a]; sh.rn[a] = hi[a] - lo[a] + 1; }
                sh.pitch = (hi[0] - lo[0] + 1) | 1;
                sh.zmin = INT_MAX; sh.leader = INT_MAX;
            }
            __syncthreads();
            PF_MARK(pf_leader)
            // (4) stage voxels + occupancy bits
            slab_stage<C>(V, sh, S);
            if (V.ess) slab_stage_occ<CULL>(V, sh, sh_cmax, sh.occ);
            if (sa.prefetch && (C::DT ? C::DT : V.dtype) == B200RC_U16) {
                const int adv = sgn > 0.f ? DEPTH : -DEPTH;
                slab_prefetch_next(V, sh, axis == 0 ? adv : 0, axis == 1 ? adv : 0, axis == 2 ? adv : 0);
            }
            __syncthreads();
            PF_MARK(pf_stage)
            // (5) march while the footprint of the next sample lies inside the region.  The loop is warp-synchronous
            //     (one __any_sync per iteration) so that the lanes of a warp take their samples together.
            {
                const int rx = sh.rlo[0], ry = sh.rlo[1], rz = sh.rlo[2];
                const int nx = sh.rn[0], ny = sh.rn[1], nz = sh.rn[2];
                const int pitch = sh.pitch, slice = pitch * ny;
                const int ox = sh.occ_off[0], oy = sh.occ_off[1], oz = sh.occ_off[2], ony = sh.occ_ny;
                const bool cull_on = CULL && V.mip_cull && V.ess && sh.cm_ok;
                const int cm_nx = sh.cm_nx;
                bool run = part;
                if (STATS && p.collect_stats) { dbg_part += part ? 1u : 0u; dbg_rounds += 1u; }
                while (__any_sync(0xffffffffu, run)) {
#ifdef SLAB_PROFILE
                    pf_iters++; pf_lanes += __popc(__ballot_sync(0xffffffffu, run));
#endif
                    if (run) {
                        const float3 sp = f3(xadd(first.x, xmul(t, dir.x)), xadd(first.y, xmul(t, dir.y)), xadd(first.z, xmul(t, dir.z)));
                        const float sux = xsub(xmul(sp.x, V.dimf[0]), 0.5f), suy = xsub(xmul(sp.y, V.dimf[1]), 0.5f), suz = xsub(xmul(sp.z, V.dimf[2]), 0.5f);
                        const float fx = floorf(sux), fy = floorf(suy), fz = floorf(suz);
                        const int lx = (int)fx - rx, ly = (int)fy - ry, lz = (int)fz - rz;
                        // ---- is the 4x4x4 footprint staged? ----
                        if ((unsigned)(lx - 1) >= (unsigned)(nx - 3) || (unsigned)(ly - 1) >= (unsigned)(ny - 3) || (unsigned)(lz - 1) >= (unsigned)(nz - 3)) run = false;
                        else if (V.sharded && !shard_owns(V, (int)fx, (int)fy, (int)fz)) {
                            // another rank's sample: step over it (pre-integration re-fetches lastIntensity afterwards)
                            stale = true; prev_empty = false;
                            if (pi_mode) lastI_stale = true;
                            ++loop; t_last = t;
                            t = p.lattice_exact ? (float)loop * step : xadd(t, step);
                            finished = (t > tStop) || loop >= 256 * 256;
                            run = !finished;
                        } else {
                            // ---- empty-space skipping: occupancy bit of the sample's level-0 cell (staged) ----
                            bool empty = false;
                            if (V.ess) empty = !((sh.occ[((lz + oz) >> 3) * ony + ((ly + oy) >> 3)] >> ((lx + ox) >> 3)) & 1u);
                            if (CULL && cull_on && !empty) empty = mip_cell_culled(V, sh_cmax[(((lz + oz) >> 3) * ony + ((ly + oy) >> 3)) * cm_nx + ((lx + ox) >> 3)], maxI);
                            if (empty && (!pi_mode || prev_empty)) {
                                // Every lattice sample inside this cell classifies to alpha 0 without side effects.  With
                                // pre-integration that also needs the PREVIOUS sample to lie in an empty cell: the value
                                // ranges of the two (adjacent, overlapping) cells then form one interval on which the
                                // transfer function is transparent, so the table entry (cur, prev) is transparent too.
                                slab_skip_cell<STATS>(p, V, dir, sux, suy, suz, fx, fy, fz, tStop, step, sh.step_rcp, t, t_last, loop, finished, st);
                                if (pi_mode) lastI_stale = true;
                                stale = true;
                                run = !finished;
                            }
                            if (!(empty && (!pi_mode || prev_empty))) {
                                stale = true;
                                const float ax = xsub(sux, fx), ay = xsub(suy, fy), az = xsub(suz, fz);
                                const float* b = S + (lz * ny + ly) * pitch + lx;
                                if (pi_mode && lastI_stale) {
                                    // lastIntensity of the (skipped) sample just before this one: its 2x2x2 taps lie within one voxel
                                    // of this sample's, i.e. inside the staged footprint (sampling rate >= 1, checked by the launcher)
                                    const float3 q = f3(xadd(first.x, xmul(t_last, dir.x)), xadd(first.y, xmul(t_last, dir.y)), xadd(first.z, xmul(t_last, dir.z)));
                                    const float qx = xsub(xmul(q.x, V.dimf[0]), 0.5f), qy = xsub(xmul(q.y, V.dimf[1]), 0.5f), qz = xsub(xmul(q.z, V.dimf[2]), 0.5f);
                                    const float gx = floorf(qx), gy = floorf(qy), gz = floorf(qz);
                                    const int jx = min(max((int)gx - rx, 0), nx - 2), jy = min(max((int)gy - ry, 0), ny - 2), jz = min(max((int)gz - rz, 0), nz - 2);
                                    const float bx = xsub(qx, gx), by = xsub(qy, gy), bz = xsub(qz, gz);
                                    const float* c = S + (jz * ny + jy) * pitch + jx;
                                    const float d00 = xlerp(c[0], c[1], bx), d10 = xlerp(c[pitch], c[pitch + 1], bx);
                                    const float d01 = xlerp(c[slice], c[slice + 1], bx), d11 = xlerp(c[slice + pitch], c[slice + pitch + 1], bx);
                                    lastI = rw_to_tex_rcp(p.cls[0].dom_lo, p.cls[0].dom_hi, sh.dom_rcp, xadd(xmul(xlerp(xlerp(d00, d10, by), xlerp(d01, d11, by), bz), V.rwm_scale), V.rwm_offset));
                                    lastI_stale = false;
                                }
                                // ---- getVoxel: trilinear from the staged taps (GL float weights, identical to lerp_taps) ----
                                const float c000 = b[0], c100 = b[1], c010 = b[pitch], c110 = b[pitch + 1];
                                const float c001 = b[slice], c101 = b[slice + 1], c011 = b[slice + pitch], c111 = b[slice + pitch + 1];
                                const float c00 = xlerp(c000, c100, ax), c10 = xlerp(c010, c110, ax), c01 = xlerp(c001, c101, ax), c11 = xlerp(c011, c111, ax);
                                const float tex = xlerp(xlerp(c00, c10, ay), xlerp(c01, c11, ay), az);
                                const float intensity = xadd(xmul(tex, V.rwm_scale), V.rwm_offset);
                                SLAB_STAT(st.classified++);
                                int a8 = -1;
                                float4 color0;
                                if (pi_mode) {
                                    // applyTFpi: texture(table, vec2(intensity, previous intensity)); the previous sample's texture
                                    // coordinate is carried along instead of being mapped a second time (same value)
                                    const float s_cur = rw_to_tex_rcp(p.cls[0].dom_lo, p.cls[0].dom_hi, sh.dom_rcp, intensity);
                                    color0 = preint_lookup_unit(p.cls[0], s_cur, lastI);
                                    lastI = s_cur;
                                } else {        // applyTF (mod_transfunc.frag:71-78): sampler1D NEAREST = floor(s * w), clamp to edge
                                    const float s_cur = rw_to_tex_rcp(p.cls[0].dom_lo, p.cls[0].dom_hi, sh.dom_rcp, intensity);
                                    const int ti = min(max(__float2int_rd(xmul(s_cur, (float)p.cls[0].tf_width)), 0), p.cls[0].tf_width - 1);
                                    const uchar4 q = sm.tf[0][ti];
                                    a8 = q.w;
                                    color0 = u8tof4(q);
                                }
                                SLAB_STAT(if (color0.w > 0.f) st.shaded++);
                                if (color0.w > 0.f && compositing_takes_sample(comp, result, intensity, color0.w, maxI, p.iso)) {
                                    float3 gradient;
                                    if (t == 0.0f) gradient = fix_clip_border_gradient(p, sp, dir, px, py);
                                    else {
                                        // calcGradientA from the 32 distinct voxels its six offset fetches share (see grad_central_shared)
                                        auto L = [&](int i, int j, int k) -> float { return b[(i - 1) + (j - 1) * pitch + (k - 1) * slice]; };
                                        float dxq[2][2], dyq[2][2], dzq[2][2];
                                        const float cen[2][2][2] = {{{c000, c001}, {c010, c011}}, {{c100, c101}, {c110, c111}}};
#pragma unroll
                                        for (int j = 0; j < 2; ++j)
#pragma unroll
                                            for (int k = 0; k < 2; ++k) {
                                                const float A = L(0, j + 1, k + 1), B = cen[0][j][k], Cc = cen[1][j][k], D = L(3, j + 1, k + 1);
                                                dxq[j][k] = (A - Cc) + ax * ((B - D) - (A - Cc));
                                            }
#pragma unroll
                                        for (int i = 0; i < 2; ++i)
#pragma unroll
                                            for (int k = 0; k < 2; ++k) {
                                                const float A = L(i + 1, 0, k + 1), D = L(i + 1, 3, k + 1), B = cen[i][0][k], Cc = cen[i][1][k];
                                                dyq[i][k] = (A - Cc) + ay * ((B - D) - (A - Cc));
                                            }
#pragma unroll
                                        for (int i = 0; i < 2; ++i)
#pragma unroll
                                            for (int j = 0; j < 2; ++j) {
                                                const float A = L(i + 1, j + 1, 0), D = L(i + 1, j + 1, 3), B = cen[i][j][0], Cc = cen[i][j][1];
                                                dzq[i][j] = (A - Cc) + az * ((B - D) - (A - Cc));
                                            }
                                        auto bil = [](const float q[2][2], float s_, float tt) -> float {
                                            const float q0 = q[0][0] + s_ * (q[1][0] - q[0][0]), q1 = q[0][1] + s_ * (q[1][1] - q[0][1]);
                                            return q0 + tt * (q1 - q0);
                                        };
                                        gradient = f3(bil(dxq, ay, az), bil(dyq, ax, az), bil(dzq, ax, ay));
                                        gradient = gradient * f3(V.spacing_rcp[0] * 0.5f, V.spacing_rcp[1] * 0.5f, V.spacing_rcp[2] * 0.5f);
                                        gradient = gradient * V.rwm_scale;
                                    }
                                    const float3 posPhys = f3((sp.x * V.dimf[0] - 0.5f) * V.spacing[0] + V.offset[0], (sp.y * V.dimf[1] - 0.5f) * V.spacing[1] + V.offset[1],
                                                              (sp.z * V.dimf[2] - 0.5f) * V.spacing[2] + V.offset[2]);
                                    float4 color = color0;
                                    const float3 rgb = f3(color.x, color.y, color.z);
                                    const float3 shd = apply_shading(smode, p.light, gradient, posPhys, ld3(V.light_phys), ld3(V.cam_phys), rgb, rgb, f3(1.f, 1.f, 1.f));
                                    color.x = shd.x; color.y = shd.y; color.z = shd.z;
                                    result = apply_compositing(comp, 1, p.preint_comp, result, intensity, color, a8, sm.lut[0], sp, gradient, t, step, tDepth, maxI, p.gamma[0], p.iso);
                                }
                                if (STATS && p.collect_stats && p.touched) {
                                    const int cx = min(max((int)fx, 0), V.dims[0] - 1) / B200RC_STAT_CELL, cy = min(max((int)fy, 0), V.dims[1] - 1) / B200RC_STAT_CELL,
                                              cz = min(max((int)fz, 0), V.dims[2] - 1) / B200RC_STAT_CELL;
                                    p.touched[((size_t)cz * p.touched_dims[1] + cy) * p.touched_dims[0] + cx] = 1;
                                }
                                if (!pi_mode) lastI = intensity;
                                prev_empty = empty;
                                if (p.ert && result.w >= 0.95f) {
                                    if (p.ert_snap) result.w = 1.0f;
                                    finished = true;
                                }
                                ++loop;
                                t_last = t;
                                t = p.lattice_exact ? (float)loop * step : xadd(t, step);
                                finished = finished || (t > tStop) || loop >= 256 * 256;
                                run = !finished;
                            }
                        }
                    }
                }
                PF_MARK(pf_march)
#ifdef SLAB_PROFILE
                pf_rounds++;
#endif
            }
        }
        // ---- outputs ------------------------------------------------------------------------------------------------------
        if (wanted) {
            float depth = 1.0f;
            if (cast) depth = get_depth_value(tDepth, tEnd, p.entry_depth[pix], p.exit_depth[pix], p.near_d, p.far_d);
            if (second_pass) {      // back to the shard-local convention: C_local = dC / (1 - P), A_local = (A - P) / (1 - P)
                const float k = 1.0f / (1.0f - prefixA);
                result.x *= k; result.y *= k; result.z *= k; result.w = (result.w - prefixA) * k;
            }
            store_color(p, 0, pix, result);
            if (p.out_depth) p.out_depth[pix] = p.quantize_depth24 ? depth24_roundtrip(depth) : depth;
            if (p.out_key) p.out_key[pix] = cast ? ((comp == B200RC_COMP_MOP || comp == B200RC_COMP_DVR) ? result.w : maxI) : 0.0f;     // DVR: the partial image's alpha plane
        }
    }
#ifdef SLAB_PROFILE
    // phase clocks of lane 0 of EVERY warp (summed): the barrier waits are what the warps that arrive early spend
    if (p.collect_stats && p.stats && lane == 0) {
        atomicAdd(p.stats + 8, (unsigned long long)pf_sched); atomicAdd(p.stats + 9, (unsigned long long)pf_wait); atomicAdd(p.stats + 10, (unsigned long long)pf_bbox);
        atomicAdd(p.stats + 11, (unsigned long long)pf_wait2); atomicAdd(p.stats + 12, (unsigned long long)pf_leader); atomicAdd(p.stats + 13, (unsigned long long)pf_stage);
        atomicAdd(p.stats + 14, (unsigned long long)(pf_march + 0 * pf_setup)); atomicAdd(p.stats + 15, (unsigned long long)pf_rounds);
        atomicAdd(p.stats + 16, (unsigned long long)pf_iters); atomicAdd(p.stats + 17, (unsigned long long)pf_lanes);
    }
#endif
    if (p.collect_stats && p.stats) {
        unsigned long long vals[7] = {st.rays, st.classified, st.shaded, st.nominal, st.skipped, dbg_part, dbg_rounds};
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            unsigned long long v = vals[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0 && v) atomicAdd(p.stats + (i < 5 ? i : i + 1), v);      // [6] ray-rounds with work, [7] ray-rounds in total
        }
    }
}
