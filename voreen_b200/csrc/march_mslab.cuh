// march_mslab.cuh -- the shared-memory staging marcher for MultiVolumeRaycaster (rc_multivolume.frag:97-350,
// multivolumeraycaster.cpp:326-357) in its hot configuration (BASELINE config 4): up to four CO-REGISTERED volumes -- identical
// dimensions, spacing, offset and transformation, hence one world-space entry/exit pair (DevParams::eep_shared) and the SAME texture
// coordinate for every volume at every lattice sample -- one u16 channel each, linear filtering, clamp-to-edge, central-difference
// gradients, only image.output connected.
//
// Structure = march_slab.cuh: a CTA owns a 16x16-pixel tile and advances its rays in lock-step slabs; per slab it stages the
// sub-volume the rays traverse ONCE PER VOLUME (nv regions of the same shape, one after the other in shared memory, plus nv tables of
// occupancy bits) and every tap of every volume is a 32-bit shared-memory load at a constant offset from ONE base address: the
// lattice position, the footprint test, the trilinear weights and all tap addresses are computed once per lattice sample and shared
// by the volumes, where the per-ray kernel (march.cuh march_multi) repeats 64-bit brick addressing, clamps and conversions for the
// 8 + 32 taps of every volume.
//
// Semantics = march_multi, statement by statement: the world-space lattice t = fl(t + step) (accumulated, as the shader does),
// texture position vfirst + (t / tEnd) * vdir, the volumes classified, shaded (mvOpacityCorrection) and composited in slot order at
// every lattice sample, per-volume empty-space skipping with the same state (prev_empty / stale bits, lastIntensity re-fetched at the
// last skipped lattice sample).  Images are identical to the per-ray kernel up to the gradient's summation order (the 32-tap form of
// march_slab.cuh); alpha > 0 decisions, sample positions and the shaded-sample counter are bit-identical.  Only the split of the
// `skipped` counter between "tested" and "walked" samples differs (the jump bound comes from the staged level-0 cell here).
// Sort-last: shards (every volume bricked identically) and the exact-ERT second pass are handled as march_multi handles them -- a
// lattice sample belongs to the rank that owns its tap base index in volume 0, foreign samples only mark lastIntensity stale.
#pragma once
#include "march_slab.cuh"

#ifndef MSLAB_MIN_CTAS
#define MSLAB_MIN_CTAS (512 / SLAB_THREADS)
#endif

struct MSlabShared { unsigned int occ[B200RC_MAX_VOLUMES][SLAB_OCC_WORDS]; };

// one lattice step of the shader: t += samplingStepSize
#define MSLAB_ADVANCE() { t_prev = t; t = xadd(t, step); ++loop; }

// The shader's empty iterations `while (t < t_jump && t <= tEnd && loop < 65536) t += step`: returns their number.  When the lattice
// is exact (t_k = k * step for every k, checked by the launcher) the first index that ends the walk is found in closed form.
__device__ __forceinline__ unsigned mslab_walk(bool exact, float step, float step_rcp, float t_jump, float tEnd, float& t, float& t_prev, int& loop) {
    if (!exact) {
        unsigned n = 0;
        while (t < t_jump && t <= tEnd && loop < 256 * 256) { MSLAB_ADVANCE() ++n; }
        return n;
    }
    if (!(t < t_jump && t <= tEnd && loop < 256 * 256)) return 0u;
    // every index below stays <= 65537: the estimates are clamped before the exact fix-up loops run
    int kj = 256 * 256;                                                         // first k > loop with k * step >= t_jump
    if (t_jump <= tEnd) {                                                       // (beyond tEnd the end of the ray stops the walk first)
        kj = max(loop + 1, (int)fminf((float)__float2int_ru(fminf(t_jump * step_rcp, 70000.f)), 65536.f));
        while (kj < 256 * 256 && (float)kj * step < t_jump) ++kj;
        while (kj - 1 > loop && (float)(kj - 1) * step >= t_jump) --kj;
    }
    int ke = max(loop + 1, __float2int_rd(fminf(tEnd * step_rcp, 70000.f)) + 1); // first k > loop with k * step > tEnd
    ke = min(ke, 256 * 256);
    while (ke < 256 * 256 && (float)ke * step <= tEnd) ++ke;
    while (ke - 1 > loop && (float)(ke - 1) * step > tEnd) --ke;
    const int k1 = min(min(kj, ke), 256 * 256);
    const unsigned n = (unsigned)(k1 - loop);
    loop = k1; t = (float)k1 * step; t_prev = (float)(k1 - 1) * step;
    return n;
}

// Would EVERY volume skip the lattice sample at sp (march_multi's tests: the sample's cell is transparent for the volume's transfer
// function and, with pre-integration, the previous sample's was too)?  The cell index is common to the co-registered volumes; on
// success t_jump = the ray parameter at which the ray leaves the smallest empty cell found (64^3 or 8^3 voxels), rounded down.
__device__ __forceinline__ bool mslab_all_skip(const DevParams& p, int nv, unsigned pi_mask, unsigned prev_empty, float3 sp, float vx, float vy, float vz, float t, float& t_jump) {
    const VolDev& v = p.vol[0];
    const float gx = sp.x * v.dimf[0] - 0.5f, gy = sp.y * v.dimf[1] - 0.5f, gz = sp.z * v.dimf[2] - 0.5f;       // as ess_empty_cell_t
    const int ix = min(max(__float2int_rd(gx), 0), v.dims[0] - 1) - v.res_lo[0], iy = min(max(__float2int_rd(gy), 0), v.dims[1] - 1) - v.res_lo[1],
              iz = min(max(__float2int_rd(gz), 0), v.dims[2] - 1) - v.res_lo[2];
    if (v.sharded && ((unsigned)ix >= (unsigned)v.res_dims[0] || (unsigned)iy >= (unsigned)v.res_dims[1] || (unsigned)iz >= (unsigned)v.res_dims[2])) return false;
    int shift = 6;
    for (int k = 0; k < nv; ++k) {
        const VolDev& w = p.vol[k];
        if (!occ_bit(w.occ1, w.occ1_wpr, w.occ1_dims[1], ix >> 6, iy >> 6, iz >> 6)) {}
        else if (!occ_bit(w.occ0, w.occ0_wpr, w.occ0_dims[1], ix >> 3, iy >> 3, iz >> 3)) shift = 3;
        else return false;
        if ((pi_mask >> k & 1u) && !(prev_empty >> k & 1u)) return false;
    }
    const float S = (float)(1 << shift);
    const float ux = gx - (float)v.res_lo[0], uy = gy - (float)v.res_lo[1], uz = gz - (float)v.res_lo[2];
    float dt = 3.0e38f;
    {
        const float uu[3] = {ux, uy, uz}, vv[3] = {vx, vy, vz};
        const int ii[3] = {ix, iy, iz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int c = ii[a] >> shift;
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(vv[a]));
            if (vv[a] > 0.f) { if (((c + 1) << shift) < v.res_dims[a]) dt = fminf(dt, ((float)(c + 1) * S - uu[a]) * r); }
            else if (vv[a] < 0.f) { if (c > 0) dt = fminf(dt, ((float)c * S - uu[a]) * r); }
        }
    }
    t_jump = __fadd_rd(t, __fmul_rd(fmaxf(dt, 0.f), 0.999999f));
    return true;
}

template <class C>
__global__ void __launch_bounds__(SLAB_THREADS, MSLAB_MIN_CTAS) march_mslab_kernel(const __grid_constant__ DevParams p, const SlabArgs sa) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SlabShared sh;
    __shared__ MSlabShared msh;
    __shared__ float2 vote_slope[SLAB_THREADS / 32];
    float* S = (float*)smem_raw;                    // region of volume v: S + v * sa.cap_floats
    const VolDev& V = p.vol[0];                     // geometry (identical for all volumes: checked by the launcher)
    const int nv = p.nvol;
    const float step = p.step;
    const bool exact = sa.exact_lattice != 0;
    const float step_rcp = 1.0f / step;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    RayStats st = {0u, 0u, 0u, 0u, 0ull};
    // 1D transfer functions: the sample's alpha is one of 256 values, so mvOpacityCorrection (multivolumeraycaster.cpp:447-456) and the
    // opacity correction of the DVR / MIDA compositing that follows it are two 256-entry tables -- the same powf calls the per-ray kernel
    // evaluates per shaded sample, evaluated once per alpha value
    __shared__ float mv_lut[C::CLS == B200RC_CLASS_TF ? 256 : 1], oc_lut[C::CLS == B200RC_CLASS_TF ? 256 : 1];
    if (C::CLS == B200RC_CLASS_TF)
        for (int i = threadIdx.x; i < 256; i += SLAB_THREADS) {
            const float a = __fmul_rn((float)i, 1.0f / 255.0f);                     // u8tof4
            const float m = 1.0f - powf(1.0f - a, p.mv_factor);
            mv_lut[i] = m;
            oc_lut[i] = opacity_correct(m, -1, nullptr, step);
        }
    bool ess_all = true;
    unsigned pi_mask = 0u;                          // bit v: volume v classifies with pre-integration
    for (int v = 0; v < nv; ++v) {
        ess_all = ess_all && p.vol[v].ess;
        if ((C::CLS >= 0 ? C::CLS : p.classification[v]) == B200RC_CLASS_PREINT) pi_mask |= 1u << v;
    }
    if (threadIdx.x == 0) { sh.zmin = INT_MAX; sh.leader = INT_MAX; for (int a = 0; a < 3; ++a) { sh.lo[a] = INT_MAX; sh.hi[a] = INT_MIN; } }
    __syncthreads();
    const unsigned ntiles = (unsigned)p.tiles_x * (unsigned)p.tiles_y;
    const unsigned super_x = (unsigned)p.tiles_x >> 3;
    const unsigned n_active = min(__ldg(p.tile_counter + 1), ntiles);

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
        if (p.band_mod > 1 && (int)((ty * SLAB_TILE_H / (unsigned)p.band_rows) % (unsigned)p.band_mod) != p.band_rem) { __syncthreads(); continue; }
        const int px = (int)(tx * SLAB_TILE + (warp & 1u) * 8u + (lane & 7u)), py = (int)(ty * SLAB_TILE_H + (warp >> 1) * 4u + (lane >> 3));
        const bool inside = px < p.W && py < p.H;
        const size_t pix = inside ? (size_t)py * p.W + px : 0;
        const bool second_pass = p.flags != nullptr;
        const bool wanted = inside && (!second_pass || p.flags[pix] == 2);      // sort-last second pass: flagged pixels only
        const float prefixA = (second_pass && wanted) ? p.prefix[pix] : 0.f;
        float3 vfirst = f3(0.f, 0.f, 0.f), vdir = f3(0.f, 0.f, 1.f);
        float tEnd = 1.f;
        bool cast = false;
        if (wanted) {
            const float4 e4 = p.entry[pix], x4 = p.exitp[pix];
            cast = !(e4.x == x4.x && e4.y == x4.y && e4.z == x4.z);
            if (cast) {
                const float3 d = f3(xsub(x4.x, e4.x), xsub(x4.y, e4.y), xsub(x4.z, e4.z));
                tEnd = __fsqrt_rn(xadd(xadd(xmul(d.x, d.x), xmul(d.y, d.y)), xmul(d.z, d.z)));
                const float4 a = xmat_mul(V.w2t, make_float4(e4.x, e4.y, e4.z, 1.f));       // worldToTex (rc_multivolume.frag:119-122)
                const float4 b = xmat_mul(V.w2t, make_float4(x4.x, x4.y, x4.z, 1.f));
                vfirst = f3(a.x, a.y, a.z);
                vdir = f3(xsub(b.x, a.x), xsub(b.y, a.y), xsub(b.z, a.z));
                st.rays++;
                st.nominal += (unsigned long long)floorf(xdiv(tEnd, step)) + 1ull;
            }
        }
        float4 result = make_float4(0.f, 0.f, 0.f, prefixA);
        float lastI[B200RC_MAX_VOLUMES] = {0.f, 0.f, 0.f, 0.f};
        float t = 0.0f, t_prev = 0.0f, tDepth = -1.0f, maxI = 0.f;
        int loop = 0;
        bool finished = !cast;
        unsigned prev_empty = 0u, stale = 0u;        // march_multi's per-volume bits
        const float rcpT = 1.0f / tEnd;
        const unsigned all = (1u << nv) - 1u;
        // sort-last shard: march only the lattice samples near the owned range (march_multi: pre-roll of the shader's own t += step)
        const bool sharded = V.sharded != 0;
        float tStop = tEnd;
        if (cast && sharded) {
            const ShardSeg seg = shard_segment(V, vfirst, f3(vdir.x * rcpT, vdir.y * rcpT, vdir.z * rcpT), tEnd, step);
            if (!seg.any) finished = true;
            else {
                tStop = seg.t_stop;
                for (int i = 0; i < seg.k0 && t <= tEnd; ++i) MSLAB_ADVANCE()
                if (seg.k0 > 0) stale |= all;       // every volume's window is the whole ray (shared entry/exit)
                finished = (t > tStop);
            }
        }
        // slab orientation: major axis + travel direction of the tile's first casting ray (voxels per unit t, up to the common 1 / tEnd)
        const float vx = vdir.x * V.dimf[0] * rcpT, vy = vdir.y * V.dimf[1] * rcpT, vz = vdir.z * V.dimf[2] * rcpT;
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
            if (DEPTH <= 0) {       // deepest slab (4..12 planes) whose region -- tile extent + drift of the rays + apron -- fits the capacity
                DEPTH = sa.dmin;
                for (int d = sa.dmax; d > sa.dmin; --d) {
                    const int n1 = sa.ext + (int)ceilf((float)d * sl.x), n2 = sa.ext + (int)ceilf((float)d * sl.y);
                    if ((long long)(n1 | 1) * (n2 | 1) * (d + 3) <= sa.cap_floats) { DEPTH = d; break; }
                }
            }
        }
        const float vmaj = fabsf(pick3f(vx, vy, vz, axis));
        bool plan_stale = true;
        int i0x = 0, i0y = 0, i0z = 0, zc = INT_MAX; float ux = 0.f, uy = 0.f, uz = 0.f;

        for (;;) {
            // (1) next sample of this ray, after the lattice samples at which NO volume can contribute (all in transparent cells)
            if (plan_stale) {
                if (ess_all || sharded) {
                    while (!finished) {
                        const float realT = xdiv(t, tEnd);
                        const float3 sp = f3(xadd(vfirst.x, xmul(realT, vdir.x)), xadd(vfirst.y, xmul(realT, vdir.y)), xadd(vfirst.z, xmul(realT, vdir.z)));
                        if (sharded) {      // another rank's lattice sample: its owner sets lastIntensity of every volume
                            const int x0 = (int)floorf(xsub(xmul(sp.x, V.dimf[0]), 0.5f)), y0 = (int)floorf(xsub(xmul(sp.y, V.dimf[1]), 0.5f)), z0 = (int)floorf(xsub(xmul(sp.z, V.dimf[2]), 0.5f));
                            if (!shard_owns(V, x0, y0, z0)) {
                                stale |= all; prev_empty = 0u;
                                MSLAB_ADVANCE()
                                finished = (t > tStop) || loop >= 256 * 256;
                                continue;
                            }
                        }
                        // would every volume skip the sample at t?  (nothing is committed until all agree)
                        float t_jump = 3.0e38f;
                        if (!ess_all || !mslab_all_skip(p, nv, pi_mask, prev_empty, sp, vx, vy, vz, t, t_jump)) break;
                        stale |= pi_mask & all;
                        st.skipped += (unsigned)nv;
                        MSLAB_ADVANCE()
                        st.skipped += mslab_walk(exact, step, step_rcp, t_jump, tStop, t, t_prev, loop);
                        finished = (t > tStop) || loop >= 256 * 256;
                    }
                }
                zc = INT_MAX;
                if (!finished) {
                    const float realT = xdiv(t, tEnd);
                    const float3 sp = f3(xadd(vfirst.x, xmul(realT, vdir.x)), xadd(vfirst.y, xmul(realT, vdir.y)), xadd(vfirst.z, xmul(realT, vdir.z)));
                    ux = xsub(xmul(sp.x, V.dimf[0]), 0.5f); uy = xsub(xmul(sp.y, V.dimf[1]), 0.5f); uz = xsub(xmul(sp.z, V.dimf[2]), 0.5f);
                    i0x = (int)floorf(ux); i0y = (int)floorf(uy); i0z = (int)floorf(uz);
                    const int im = pick3(i0x, i0y, i0z, axis);
                    zc = sgn > 0.f ? im : -im;
                }
                plan_stale = false;
            }
            {
                const int wmin = __reduce_min_sync(0xffffffffu, zc);
                if (lane == 0 && wmin != INT_MAX) atomicMin(&sh.zmin, wmin);
            }
            __syncthreads();
            const int zmin = sh.zmin;
            if (zmin == INT_MAX) break;
            // (2) bounding box of the participating segments
            const bool part = !finished && zc < zmin + DEPTH;
            {
                int lo0 = INT_MAX, lo1 = INT_MAX, lo2 = INT_MAX, hi0 = INT_MIN, hi1 = INT_MIN, hi2 = INT_MIN;
                if (part) {
                    const float um = pick3f(ux, uy, uz, axis);
                    const float dist = (float)(zmin + DEPTH) - (sgn > 0.f ? um : -um);
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
                const int cand = __reduce_min_sync(0xffffffffu, (!finished && zc == zmin) ? (int)threadIdx.x : INT_MAX);
                if (lane == 0 && cand != INT_MAX) atomicMin(&sh.leader, cand);
            }
            __syncthreads();
            // (3) the leader fixes the region (footprint apron -1..+2) and resets the scratch; same rules as march_slab_kernel
            if ((int)threadIdx.x == sh.leader) {
                const int li[3] = {i0x, i0y, i0z};
                int lo[3], hi[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    lo[a] = sh.lo[a] - 2; hi[a] = sh.hi[a] + 3;
                    if (a == axis) {
                        if (sgn > 0.f) { lo[a] = zmin - 1; hi[a] = zmin + DEPTH - 1 + 2; }
                        else { lo[a] = -(zmin + DEPTH - 1) - 1; hi[a] = -zmin + 2; }
                    }
                    sh.lo[a] = INT_MAX; sh.hi[a] = INT_MIN;
                }
                const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
                const int nmaj = hi[axis] - lo[axis] + 1;
                for (int k = 0; k < 2; ++k) {
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
                    if (as == 0) cap -= 1;
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
            // (4) stage every volume's region + occupancy bits
            for (int v = 0; v < nv; ++v) {
                slab_stage<C>(p.vol[v], sh, S + (size_t)v * sa.cap_floats);
                if (p.vol[v].ess) slab_stage_occ<false>(p.vol[v], sh, nullptr, msh.occ[v]);
            }
            __syncthreads();
            // (5) march while the footprint of the next lattice sample lies inside the region
            {
                const int rx = sh.rlo[0], ry = sh.rlo[1], rz = sh.rlo[2];
                const int nx = sh.rn[0], ny = sh.rn[1], nz = sh.rn[2];
                const int pitch = sh.pitch, slice = pitch * ny;
                const int ox = sh.occ_off[0], oy = sh.occ_off[1], oz = sh.occ_off[2], ony = sh.occ_ny;
                bool run = part;
                while (__any_sync(0xffffffffu, run)) {
                    if (run) {
                        const float realT = xdiv(t, tEnd);
                        const float3 sp = f3(xadd(vfirst.x, xmul(realT, vdir.x)), xadd(vfirst.y, xmul(realT, vdir.y)), xadd(vfirst.z, xmul(realT, vdir.z)));
                        const float sux = xsub(xmul(sp.x, V.dimf[0]), 0.5f), suy = xsub(xmul(sp.y, V.dimf[1]), 0.5f), suz = xsub(xmul(sp.z, V.dimf[2]), 0.5f);
                        const float fx = floorf(sux), fy = floorf(suy), fz = floorf(suz);
                        const int lx = (int)fx - rx, ly = (int)fy - ry, lz = (int)fz - rz;
                        if ((unsigned)(lx - 1) >= (unsigned)(nx - 3) || (unsigned)(ly - 1) >= (unsigned)(ny - 3) || (unsigned)(lz - 1) >= (unsigned)(nz - 3)) run = false;
                        else if (sharded && !shard_owns(V, (int)fx, (int)fy, (int)fz)) {
                            plan_stale = true;
                            stale |= all; prev_empty = 0u;
                            MSLAB_ADVANCE()
                            finished = (t > tStop) || loop >= 256 * 256;
                            run = !finished;
                        } else {
                            plan_stale = true;
                            // per-volume emptiness of the sample's level-0 cell (staged bits), and march_multi's skip rule
                            unsigned empty = 0u;
                            {
                                const int w = ((lz + oz) >> 3) * ony + ((ly + oy) >> 3), b = (lx + ox) >> 3;
                                for (int v = 0; v < nv; ++v)
                                    if (p.vol[v].ess && !((msh.occ[v][w] >> b) & 1u)) empty |= 1u << v;
                            }
                            const unsigned skip = empty & (~pi_mask | prev_empty);      // empty, and (pre-integration) the previous sample was too
                            if (skip == all) {
                                // no volume takes this sample: the shader's iteration does nothing but t += step; walk the lattice up to
                                // the exit of the staged cell (rounded down: one more test at worst, never a sample too far)
                                stale |= pi_mask & all;
                                st.skipped += (unsigned)nv;
                                float dt = 3.0e38f;
                                {
                                    const float su[3] = {sux, suy, suz}, ff[3] = {fx, fy, fz}, vv[3] = {vx, vy, vz};
#pragma unroll
                                    for (int a = 0; a < 3; ++a) {
                                        const int ir = min(max((int)ff[a], 0), V.dims[a] - 1) - V.res_lo[a];
                                        const int c = ir >> 3;
                                        const float ur = su[a] - (float)V.res_lo[a];
                                        float r;
                                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(vv[a]));
                                        if (vv[a] > 0.f) { if (((c + 1) << 3) < V.res_dims[a]) dt = fminf(dt, ((float)((c + 1) << 3) - ur) * r); }
                                        else if (vv[a] < 0.f) { if (c > 0) dt = fminf(dt, ((float)(c << 3) - ur) * r); }
                                    }
                                }
                                const float t_jump = __fadd_rd(t, __fmul_rd(fmaxf(dt, 0.f), 0.999999f));
                                MSLAB_ADVANCE()
                                st.skipped += mslab_walk(exact, step, step_rcp, t_jump, tStop, t, t_prev, loop);
                                finished = (t > tStop) || loop >= 256 * 256;
                                run = !finished;
                            } else {
                                const float ax = xsub(sux, fx), ay = xsub(suy, fy), az = xsub(suz, fz);
                                const int boff = (lz * ny + ly) * pitch + lx;
                                // footprint of the previous lattice sample (lastIntensity after skipped samples): within one voxel of this one
                                int joff = 0; float bx = 0.f, by = 0.f, bz = 0.f;
                                if (stale & ~skip) {
                                    const float rp = xdiv(t_prev, tEnd);
                                    const float3 q = f3(xadd(vfirst.x, xmul(rp, vdir.x)), xadd(vfirst.y, xmul(rp, vdir.y)), xadd(vfirst.z, xmul(rp, vdir.z)));
                                    const float qx = xsub(xmul(q.x, V.dimf[0]), 0.5f), qy = xsub(xmul(q.y, V.dimf[1]), 0.5f), qz = xsub(xmul(q.z, V.dimf[2]), 0.5f);
                                    const float gx = floorf(qx), gy = floorf(qy), gz = floorf(qz);
                                    const int jx = min(max((int)gx - rx, 0), nx - 2), jy = min(max((int)gy - ry, 0), ny - 2), jz = min(max((int)gz - rz, 0), nz - 2);
                                    bx = xsub(qx, gx); by = xsub(qy, gy); bz = xsub(qz, gz);
                                    joff = (jz * ny + jy) * pitch + jx;
                                }
#pragma unroll
                                for (int v = 0; v < B200RC_MAX_VOLUMES; ++v) {
                                    if (v >= nv) break;
                                    const unsigned bit = 1u << v;
                                    const VolDev& Vv = p.vol[v];
                                    if (Vv.ess) {
                                        if (empty & bit) {
                                            if (skip & bit) { if (pi_mask & bit) stale |= bit; st.skipped++; continue; }
                                            prev_empty |= bit;
                                        } else prev_empty &= ~bit;
                                    }
                                    const float* b = S + (size_t)v * sa.cap_floats + boff;
                                    // getVoxel: trilinear from the staged taps (GL float weights), real-world mapping
                                    const float c000 = b[0], c100 = b[1], c010 = b[pitch], c110 = b[pitch + 1];
                                    const float c001 = b[slice], c101 = b[slice + 1], c011 = b[slice + pitch], c111 = b[slice + pitch + 1];
                                    const float c00 = xlerp(c000, c100, ax), c10 = xlerp(c010, c110, ax), c01 = xlerp(c001, c101, ax), c11 = xlerp(c011, c111, ax);
                                    const float tex = xlerp(xlerp(c00, c10, ay), xlerp(c01, c11, ay), az);
                                    const float intensity = xadd(xmul(tex, Vv.rwm_scale), Vv.rwm_offset);
                                    if (stale & bit) {
                                        const float* c = S + (size_t)v * sa.cap_floats + joff;
                                        const float d00 = xlerp(c[0], c[1], bx), d10 = xlerp(c[pitch], c[pitch + 1], bx);
                                        const float d01 = xlerp(c[slice], c[slice + 1], bx), d11 = xlerp(c[slice + pitch], c[slice + pitch + 1], bx);
                                        lastI[v] = xadd(xmul(xlerp(xlerp(d00, d10, by), xlerp(d01, d11, by), bz), Vv.rwm_scale), Vv.rwm_offset);
                                        stale &= ~bit;
                                    }
                                    st.classified++;
                                    int a8;
                                    const int cm = C::CLS >= 0 ? C::CLS : p.classification[v];
                                    const float4 color = classify<C>(cm, p.cls[v], p.cls[v].tf, f3(0.f, 0.f, 0.f), intensity, lastI[v], a8);
                                    if (color.w > 0.0f) {
                                        st.shaded++;
                                        // calcGradientA from the 32 distinct voxels its six offset fetches share (march_slab.cuh)
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
                                        float3 g = f3(bil(dxq, ay, az), bil(dyq, ax, az), bil(dzq, ax, ay));
                                        g = g * f3(Vv.spacing_rcp[0] * 0.5f, Vv.spacing_rcp[1] * 0.5f, Vv.spacing_rcp[2] * 0.5f);
                                        g = g * Vv.rwm_scale;
                                        const float3 posPhys = f3((sp.x * Vv.dimf[0] - 0.5f) * Vv.spacing[0] + Vv.offset[0], (sp.y * Vv.dimf[1] - 0.5f) * Vv.spacing[1] + Vv.offset[1],
                                                                  (sp.z * Vv.dimf[2] - 0.5f) * Vv.spacing[2] + Vv.offset[2]);
                                        const float3 rgb = f3(color.x, color.y, color.z);
                                        const float3 shd = apply_shading(C::SHADE >= 0 ? C::SHADE : p.shading[v], p.light, g, posPhys, ld3(Vv.light_phys), ld3(Vv.cam_phys), rgb, rgb, f3(1.f, 1.f, 1.f));
                                        // mvOpacityCorrection (multivolumeraycaster.cpp:447-456)
                                        const bool lut = (C::CLS == B200RC_CLASS_TF);
                                        const float4 cc = make_float4(shd.x, shd.y, shd.z, lut ? mv_lut[a8] : 1.0f - powf(1.0f - color.w, p.mv_factor));
                                        result = apply_compositing(p.comp[0], 1, p.preint_comp, result, intensity, cc, lut ? a8 : -1, lut ? oc_lut : nullptr, f3(0.f, 0.f, 0.f), g, t, step,
                                                                   tDepth, maxI, 0.0f, 0.0f);
                                    }
                                    lastI[v] = intensity;
                                }
                                if (p.ert && result.w >= 0.95f) {
                                    if (p.ert_snap) result.w = 1.0f;
                                    finished = true;
                                }
                                MSLAB_ADVANCE()
                                finished = finished || (t > tStop) || loop >= 256 * 256;
                                run = !finished;
                            }
                        }
                    }
                }
            }
        }
        // ---- outputs ------------------------------------------------------------------------------------------------------
        if (wanted) {
            float depth = 1.0f;
            if (cast) depth = get_depth_value(tDepth, tEnd, p.entry_depth[pix], p.exit_depth[pix], p.near_d, p.far_d);
            if (second_pass) {      // back to the shard-local convention the merge expects: C_local = dC / (1 - P), A_local = (A - P) / (1 - P)
                const float k = 1.0f / (1.0f - prefixA);
                result.x *= k; result.y *= k; result.z *= k; result.w = (result.w - prefixA) * k;
            }
            store_color(p, 0, pix, result);
            if (p.out_depth) p.out_depth[pix] = p.quantize_depth24 ? depth24_roundtrip(depth) : depth;
            if (p.out_key) p.out_key[pix] = result.w;       // sort-last DVR partial image: its alpha plane (first pass only)
        }
    }
    if (p.collect_stats && p.stats) {
        unsigned long long vals[5] = {st.rays, st.classified, st.shaded, st.nominal, st.skipped};
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            unsigned long long v = vals[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0 && v) atomicAdd(p.stats + i, v);
        }
    }
}
