// volume.cu -- brick-pool volume upload to HBM, cudaArray view for the texture units, the min/max macro-cell grid
// and the TF-dependent occupancy bits used for empty-space skipping, plus the synthetic benchmark volume.
//
// Replaces VolumeGL::generateTexture (V/src/core/datastructures/volume/volumegl.cpp:106-139,336-381) and
// StructureProxyGeometryBackgroundThread::computeRegionStructure
// (V/modules/base/processors/proxygeometry/optimizedproxygeometry.cpp:894-987).
//
// HBM layout: 8x8x8-voxel bricks (1 KiB for u16), bricks x-fastest over the resident range, channels interleaved per
// voxel (3 channels padded to 4).  A brick is 8 consecutive 128-byte lines, one 8x8 xy-slab each, so the 8 taps of a
// trilinear fetch touch two lines and a warp marching an 8x4 pixel tile touches a handful of lines per tap.
#include "ctx.hpp"
#include <cstring>
#include <cstdlib>
#include <cstdio>

// ------------------------------------------------------------------------------------------------------------------
// linear -> bricks
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void brick_rows_kernel(const T* __restrict__ lin, T* __restrict__ bricks, int dx, int dy, int dz, int nbx, int nby, int nbz,
                                  int nch_in, int nch_st, T pad_value) {
    // one thread per 8-voxel brick row; consecutive threads write consecutive rows of the pool (coalesced stores)
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrows = (long long)nbx * nby * nbz * 64;
    if (row >= nrows) return;
    const long long brick = row >> 6;
    const int within = (int)(row & 63);
    const int bx = (int)(brick % nbx), by = (int)((brick / nbx) % nby), bz = (int)(brick / ((long long)nbx * nby));
    const int y = min(by * 8 + (within & 7), dy - 1), z = min(bz * 8 + (within >> 3), dz - 1);
    const T* src = lin + ((long long)z * dy + y) * dx * nch_in;
    T* dst = bricks + row * 8 * nch_st;
    for (int i = 0; i < 8; ++i) {
        const int x = min(bx * 8 + i, dx - 1);
        for (int c = 0; c < nch_st; ++c) dst[i * nch_st + c] = (c < nch_in) ? src[(long long)x * nch_in + c] : pad_value;
    }
}

// vector path: 1 channel u16, dx % 8 == 0 -> one 16-byte load and one 16-byte store per thread
__global__ void brick_rows_u16x8_kernel(const uint4* __restrict__ lin, uint4* __restrict__ bricks, int dx8, int dy, int dz, int nbx, int nby, int nbz) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrows = (long long)nbx * nby * nbz * 64;
    if (row >= nrows) return;
    const long long brick = row >> 6;
    const int within = (int)(row & 63);
    const int bx = (int)(brick % nbx), by = (int)((brick / nbx) % nby), bz = (int)(brick / ((long long)nbx * nby));
    const int y = min(by * 8 + (within & 7), dy - 1), z = min(bz * 8 + (within >> 3), dz - 1);
    bricks[row] = __ldg(lin + ((long long)z * dy + y) * dx8 + bx);
}

// slab variants: `lin` holds planes [z0, z0 + nz) of the resident range; the launch covers the brick layers [bz0, bz0 + nbz_slab)
template <typename T>
__global__ void brick_rows_slab_kernel(const T* __restrict__ lin, T* __restrict__ bricks, int dx, int dy, int dz, int nbx, int nby, int z0, int nz,
                                       int bz0, int nbz_slab, int nch_in, int nch_st, T pad_value) {
    const long long lrow = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrows = (long long)nbx * nby * nbz_slab * 64;
    if (lrow >= nrows) return;
    const long long row = lrow + (long long)bz0 * nbx * nby * 64;
    const long long brick = row >> 6;
    const int within = (int)(row & 63);
    const int bx = (int)(brick % nbx), by = (int)((brick / nbx) % nby), bz = (int)(brick / ((long long)nbx * nby));
    const int y = min(by * 8 + (within & 7), dy - 1);
    const int z = min(min(bz * 8 + (within >> 3), dz - 1), z0 + nz - 1) - z0;
    const T* src = lin + ((long long)z * dy + y) * dx * nch_in;
    T* dst = bricks + row * 8 * nch_st;
    for (int i = 0; i < 8; ++i) {
        const int x = min(bx * 8 + i, dx - 1);
        for (int c = 0; c < nch_st; ++c) dst[i * nch_st + c] = (c < nch_in) ? src[(long long)x * nch_in + c] : pad_value;
    }
}
__global__ void brick_rows_slab_u16x8_kernel(const uint4* __restrict__ lin, uint4* __restrict__ bricks, int dx8, int dy, int dz, int nbx, int nby, int z0, int nz,
                                             int bz0, int nbz_slab) {
    const long long lrow = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrows = (long long)nbx * nby * nbz_slab * 64;
    if (lrow >= nrows) return;
    const long long row = lrow + (long long)bz0 * nbx * nby * 64;
    const long long brick = row >> 6;
    const int within = (int)(row & 63);
    const int bx = (int)(brick % nbx), by = (int)((brick / nbx) % nby), bz = (int)(brick / ((long long)nbx * nby));
    const int y = min(by * 8 + (within & 7), dy - 1);
    const int z = min(min(bz * 8 + (within >> 3), dz - 1), z0 + nz - 1) - z0;
    bricks[row] = __ldg(lin + ((long long)z * dy + y) * dx8 + bx);
}

template <typename T>
__global__ void debrick_kernel(const T* __restrict__ bricks, T* __restrict__ lin, int dx, int dy, int dz, int nbx, int nby, int nch_out, int nch_st) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)dx * dy * dz;
    if (i >= n) return;
    const int x = (int)(i % dx), y = (int)((i / dx) % dy), z = (int)(i / ((long long)dx * dy));
    const long long e = brick_off_x(x) + brick_off_y(y, nbx) + brick_off_z(z, nbx, nby);
    for (int c = 0; c < nch_out; ++c) lin[i * nch_out + c] = bricks[e * nch_st + c];
}

// bricks -> cudaArray through a surface (no linear staging copy of a 16 GiB volume)
template <typename T>
__global__ void bricks_to_surface_kernel(const T* __restrict__ bricks, cudaSurfaceObject_t surf, int dx, int dy, int dz, int nbx, int nby, int nch_st) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)dx * dy * dz;
    if (i >= n) return;
    const int x = (int)(i % dx), y = (int)((i / dx) % dy), z = (int)(i / ((long long)dx * dy));
    const long long e = (brick_off_x(x) + brick_off_y(y, nbx) + brick_off_z(z, nbx, nby)) * nch_st;
    const int bytes = (int)sizeof(T) * nch_st;
    if (nch_st == 1) surf3Dwrite(bricks[e], surf, x * bytes, y, z);
    else if (nch_st == 2) {
        if (sizeof(T) == 1) surf3Dwrite(make_uchar2((unsigned char)bricks[e], (unsigned char)bricks[e + 1]), surf, x * bytes, y, z);
        else if (sizeof(T) == 2) surf3Dwrite(make_ushort2((unsigned short)bricks[e], (unsigned short)bricks[e + 1]), surf, x * bytes, y, z);
        else surf3Dwrite(make_float2((float)bricks[e], (float)bricks[e + 1]), surf, x * bytes, y, z);
    } else {
        if (sizeof(T) == 1) surf3Dwrite(make_uchar4((unsigned char)bricks[e], (unsigned char)bricks[e + 1], (unsigned char)bricks[e + 2], (unsigned char)bricks[e + 3]), surf, x * bytes, y, z);
        else if (sizeof(T) == 2) surf3Dwrite(make_ushort4((unsigned short)bricks[e], (unsigned short)bricks[e + 1], (unsigned short)bricks[e + 2], (unsigned short)bricks[e + 3]), surf, x * bytes, y, z);
        else surf3Dwrite(make_float4((float)bricks[e], (float)bricks[e + 1], (float)bricks[e + 2], (float)bricks[e + 3]), surf, x * bytes, y, z);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// synthetic "CT-like" volume, written straight into the brick pool.  Integer noise + non-contracted fp32
// so that voreen_b200/synth.py reproduces it bit for bit on the CPU.
// ------------------------------------------------------------------------------------------------------------------
struct SynthSpheres { float cx[64], cy[64], cz[64], r2[64]; };

__host__ __device__ inline uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void synth_kernel(unsigned short* __restrict__ bricks, int gx, int gy, int gz, int lox, int loy, int loz, int rx, int ry, int rz,
                             int nbx, int nby, int nbz, float rcpx, float rcpy, float rcpz, float fall_scale, uint32_t seed,
                             const __grid_constant__ SynthSpheres sph) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)nbx * nby * nbz * 512;
    if (e >= n) return;
    const long long brick = e >> 9;
    const int within = (int)(e & 511);
    const int bx = (int)(brick % nbx), by = (int)((brick / nbx) % nby), bz = (int)(brick / ((long long)nbx * nby));
    const int x = min(bx * 8 + (within & 7), rx - 1) + lox, y = min(by * 8 + ((within >> 3) & 7), ry - 1) + loy, z = min(bz * 8 + (within >> 6), rz - 1) + loz;
    const float px = xsub(xmul(xadd((float)x, 0.5f), rcpx), 0.5f), py = xsub(xmul(xadd((float)y, 0.5f), rcpy), 0.5f), pz = xsub(xmul(xadd((float)z, 0.5f), rcpz), 0.5f);
    const float xx = xmul(px, px), yy = xmul(py, py), zz = xmul(pz, pz);
    const float r2 = xadd(xadd(xx, yy), zz);
    // soft ellipsoid "tissue": semi-axes (0.42, 0.36, 0.40), 0.35 inside, smooth ~4-voxel falloff
    const float q2 = xadd(xadd(xmul(xx, 5.6689343f), xmul(yy, 7.7160494f)), xmul(zz, 6.25f));
    const float q = __fsqrt_rn(q2);
    float fall = xmul(xmul(xsub(1.0f, q), fall_scale), 0.25f);
    fall = fminf(fmaxf(fall, 0.0f), 1.0f);
    const float s = xmul(xmul(fall, fall), xsub(3.0f, xmul(2.0f, fall)));
    float f = xmul(0.35f, s);
    // 64 small bright spheres
    bool hit = false;
    for (int i = 0; i < 64; ++i) {
        const float ax = xsub(px, sph.cx[i]), ay = xsub(py, sph.cy[i]), az = xsub(pz, sph.cz[i]);
        const float d2 = xadd(xadd(xmul(ax, ax), xmul(ay, ay)), xmul(az, az));
        hit = hit || (d2 <= sph.r2[i]);
    }
    if (hit) f = fmaxf(f, 0.6f);
    // solid sphere shell r in [0.30, 0.32]
    if (r2 >= 0.09f && r2 <= 0.1024f) f = fmaxf(f, 0.85f);
    int val = (int)floorf(xadd(xmul(f, 65535.0f), 0.5f));
    // additive noise ~ N(0, 0.01): Irwin-Hall sum of four 16-bit uniforms, integer arithmetic only
    const uint32_t h = hash32((uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)z * 83492791u ^ seed * 0x9e3779b9u);
    const uint32_t h2 = hash32(h ^ 0x68bc21ebu);
    const int sum = (int)(h & 0xffffu) + (int)(h >> 16) + (int)(h2 & 0xffffu) + (int)(h2 >> 16);
    val += ((sum - 131070) * 1135) >> 16;
    val = min(max(val, 0), 65535);
    bricks[e] = (unsigned short)val;
}

// ------------------------------------------------------------------------------------------------------------------
// min/max macro-cell grid (level 0: 8^3 voxels + apron)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void minmax_kernel(const T* __restrict__ bricks, float2* __restrict__ out, int rx, int ry, int rz, int nbx, int nby,
                              int ncx, int ncy, int ncz, int nch, int nch_st, float norm) {
    // one warp per cell; region = [c*8-1, c*8+9] per axis (taps of every sample whose base index lies in the cell,
    // plus one voxel of slack on both sides for float fuzz at the cell faces), clamped to the resident range
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long ncell = (long long)ncx * ncy * ncz;
    if (warp >= ncell) return;
    const int cx = (int)(warp % ncx), cy = (int)((warp / ncx) % ncy), cz = (int)(warp / ((long long)ncx * ncy));
    const int x0 = max(cx * 8 - 1, 0), x1 = min(cx * 8 + 9, rx - 1);
    const int y0 = max(cy * 8 - 1, 0), y1 = min(cy * 8 + 9, ry - 1);
    const int z0 = max(cz * 8 - 1, 0), z1 = min(cz * 8 + 9, rz - 1);
    const int ex = x1 - x0 + 1, ey = y1 - y0 + 1, ez = z1 - z0 + 1;
    const int total = ex * ey * ez;
    float mn[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f}, mx[4] = {-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = lane; i < total; i += 32) {
        const int x = x0 + i % ex, y = y0 + (i / ex) % ey, z = z0 + i / (ex * ey);
        const long long e = (brick_off_x(x) + brick_off_y(y, nbx) + brick_off_z(z, nbx, nby)) * nch_st;
        for (int c = 0; c < nch; ++c) {
            const float v = (float)bricks[e + c];
            mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v);
        }
    }
    for (int c = 0; c < nch; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        // same normalisation as the sampler: raw * (1/65535) -- monotone, so min/max commute with it
        if (lane == 0) out[warp * nch + c] = make_float2(__fmul_rn(mn[c], norm), __fmul_rn(mx[c], norm));
    }
}

// ---- streaming fast path: 1-channel u16 -------------------------------------------------------------------------------
// The cell region [c*8-1, c*8+9] per axis is brick c plus the LAST voxel plane of brick c-1 and the FIRST TWO planes of
// brick c+1, so min/max over it is separable into 27 per-brick partial results, one per neighbour offset d in {-1,0,+1}^3:
// per axis d=-1 reads the neighbour's sub-range H = {7}, d=0 reads A = {0..7}, d=+1 reads L = {0,1}.
//   pass A (HBM-bound, reads every voxel exactly once, fully coalesced-by-line): one thread per 128-byte brick line (one
//          xy-slab), 8 x LDG.128, packed 2xu16 SIMD min/max (__vminu2 / __vmaxu2) for the 9 (y,x) sub-ranges in registers,
//          then the 3 z sub-ranges over the 8 lanes of the brick with shuffles; a partial result is the 32-bit word
//          (min | (65535 - max) << 16) so that ONE __vminu2 folds both.  Output SoA: part[k][brick], k = (dz+1)*9+(dy+1)*3+(dx+1).
//   pass B: one thread per cell folds its 27 neighbours' words (coalesced: consecutive cells read consecutive words).
// Algorithmic bytes: 2 B read per voxel (+ 216 B per 1 KiB brick of partials written and read back, + 8 B per cell out).
__device__ __forceinline__ unsigned mm_pack(unsigned mn2, unsigned mx2) {        // horizontal fold of packed pairs -> (min | ~max << 16)
    const unsigned mn = min(mn2 & 0xffffu, mn2 >> 16), mx = max(mx2 & 0xffffu, mx2 >> 16);
    return mn | ((65535u - mx) << 16);
}

__global__ void __launch_bounds__(256) brick_partials_u16_kernel(const uint4* __restrict__ bricks, unsigned* __restrict__ part, long long nbricks) {
    const long long slab = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // 8 slabs per brick, lane & 7 = z inside the brick
    const bool live = slab < nbricks * 8;
    const uint4* src = bricks + (live ? slab : 0) * 8;
    uint4 r[8];
#pragma unroll
    for (int y = 0; y < 8; ++y) r[y] = __ldg(src + y);                               // 128 B in flight per thread
    // x sub-ranges: A = all four words, L = word 0 (x0,x1), H = high half of word 3 (x7); y accumulates over rows
    unsigned An = 0xffffffffu, Ax = 0u, Ln = 0xffffffffu, Lx = 0u, Hn = 0xffffffffu, Hx = 0u;
    unsigned yl[6] = {0, 0, 0, 0, 0, 0};
    unsigned rn = 0, rx = 0, r0 = 0, r3 = 0;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        rn = __vminu2(__vminu2(r[y].x, r[y].y), __vminu2(r[y].z, r[y].w));
        rx = __vmaxu2(__vmaxu2(r[y].x, r[y].y), __vmaxu2(r[y].z, r[y].w));
        r0 = r[y].x; r3 = r[y].w;
        An = __vminu2(An, rn); Ax = __vmaxu2(Ax, rx);
        Ln = __vminu2(Ln, r0); Lx = __vmaxu2(Lx, r0);
        Hn = __vminu2(Hn, r3); Hx = __vmaxu2(Hx, r3);
        if (y == 1) { yl[0] = An; yl[1] = Ax; yl[2] = Ln; yl[3] = Lx; yl[4] = Hn; yl[5] = Hx; }   // y sub-range L = rows {0,1}
    }
    const unsigned h3n = Hn >> 16 | (Hn & 0xffff0000u), h3x = Hx >> 16 | (Hx & 0xffff0000u);        // x7 only: duplicate the high half
    // v[sy*3 + sx], s = 0: H (d = -1), 1: A (d = 0), 2: L (d = +1)
    unsigned v[9];
    {   // y = H: row 7 alone (rn/rx/r0/r3 still hold row 7)
        const unsigned t3 = r3 >> 16 | (r3 & 0xffff0000u);
        v[0] = mm_pack(t3, t3); v[1] = mm_pack(rn, rx); v[2] = mm_pack(r0, r0);
    }
    v[3] = mm_pack(h3n, h3x); v[4] = mm_pack(An, Ax); v[5] = mm_pack(Ln, Lx);
    {
        const unsigned t3n = yl[4] >> 16 | (yl[4] & 0xffff0000u), t3x = yl[5] >> 16 | (yl[5] & 0xffff0000u);
        v[6] = mm_pack(t3n, t3x); v[7] = mm_pack(yl[0], yl[1]); v[8] = mm_pack(yl[2], yl[3]);
    }
    const unsigned lane = threadIdx.x & 31u, base = lane & ~7u, z = lane & 7u;
    const long long brick = slab >> 3;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const unsigned zh = __shfl_sync(0xffffffffu, v[k], base + 7);                                 // z = H: plane 7
        const unsigned zl = __vminu2(__shfl_sync(0xffffffffu, v[k], base), __shfl_sync(0xffffffffu, v[k], base + 1));   // z = L: planes {0,1}
        unsigned za = v[k];
        za = __vminu2(za, __shfl_xor_sync(0xffffffffu, za, 1));
        za = __vminu2(za, __shfl_xor_sync(0xffffffffu, za, 2));
        za = __vminu2(za, __shfl_xor_sync(0xffffffffu, za, 4));
        // three of the brick's eight lanes store (one word each per k): 4 consecutive bricks of a warp -> one 16-byte segment
        if (live && z < 3) part[(size_t)(z * 9 + k) * nbricks + brick] = (z == 0) ? zh : (z == 1 ? za : zl);
    }
}

// ---- pass A, asynchronous variant (the default) ---------------------------------------------------------------------------------
// ncu on the kernel above (profiles/r1_pass35_minmax_ncu_summary.txt): l1tex throughput 99 %, DRAM 46 % -- every LDG.128 of a warp
// touches 32 different 128-byte lines (one per thread), so the L1 tag stage, not HBM, sets the pace.  Here a warp copies its 4 KiB
// (32 slabs = 4 bricks) with COALESCED 16-byte cp.async (LDGSTS, L2 -> shared memory, no registers, 4 lines per instruction) into a
// slab-padded buffer (144-byte slab pitch: the later LDS.128 of 8 consecutive threads hit 8 different bank groups), two sub-chunks in
// flight per warp, and then runs the same register arithmetic on LDS.128 rows.  Everything is warp-local: no CTA barrier.  The 27
// partial words of 32 consecutive bricks are collected in shared memory and leave as 27 full 128-byte lines instead of 16-byte pieces.
#define MMA_SUB_SLABS 32                            // slabs (128 B each) per sub-chunk = one warp pass
#define MMA_SUBS 8                                  // sub-chunks per super-chunk: 32 bricks
#define MMA_STAGE_BYTES (MMA_SUB_SLABS * 144)
#ifndef MMA_STAGES
#define MMA_STAGES 2                                // stage buffers per warp: MMA_STAGES sub-chunks in flight while one is reduced from registers
#endif
#ifndef MMA_WARPS
#define MMA_WARPS 16
#endif
#define MMA_PW_CLASS (9 * 32 + 4)                   // words per z class in the partial staging (+4: the three storing lanes of a brick hit different banks)
#define MMA_WARP_BYTES (MMA_STAGES * MMA_STAGE_BYTES + 3 * MMA_PW_CLASS * 4)

__device__ __forceinline__ void mm_cp_async16(unsigned smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// same word as mm_pack with fewer instructions: the fold of the two halves stays in the SIMD unit, 65535 - x == ~x on 16 bits
__device__ __forceinline__ unsigned mm_pack2(unsigned mn2, unsigned mx2) {
    const unsigned mn = __vminu2(mn2, mn2 >> 16), mx = __vmaxu2(mx2, mx2 >> 16);      // low halves: min / max of both voxels
    return __byte_perm(mn, ~mx, 0x5410);
}

__global__ void __launch_bounds__(MMA_WARPS * 32) brick_partials_u16_async_kernel(const uint4* __restrict__ bricks, unsigned* __restrict__ part, long long nbricks) {
    extern __shared__ __align__(16) unsigned char mm_smem[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned char* wbase = mm_smem + (size_t)warp * MMA_WARP_BYTES;
    unsigned* pw = (unsigned*)(wbase + MMA_STAGES * MMA_STAGE_BYTES);
    const long long nslabs = nbricks * 8;
    const long long nsuper = (nbricks + 31) / 32;
    const long long gw = (long long)blockIdx.x * MMA_WARPS + warp, tw = (long long)gridDim.x * MMA_WARPS;
    if (gw >= nsuper) return;
    // The warp's work is ONE stream of sub-chunks g = 0 .. G-1 (super-chunk gw + (g >> 3) tw, sub-chunk g & 7 of it), so the copy
    // pipeline never drains at a super-chunk boundary: the partial lines of a finished super-chunk leave while the next one's
    // first sub-chunks are already in flight.
    const int G = (int)((nsuper - gw + tw - 1) / tw) * MMA_SUBS;
    // sub-chunk g -> stage g % MMA_STAGES: lane i copies uint4 (i + 32 j), j = 0..7, of the sub-chunk's 256 uint4 (coalesced: 512 B per
    // instruction); uint4 e = i + 32 j is row e & 7 of slab e >> 3, so j only adds constants to both addresses
    const unsigned sm_lane = (unsigned)__cvta_generic_to_shared(wbase) + (lane >> 3) * 144u + (lane & 7u) * 16u;
    auto prefetch = [&](int g) {
        if (g < G) {
            const long long sslab = ((gw + (long long)(g >> 3) * tw) * MMA_SUBS + (g & 7)) * MMA_SUB_SLABS;
            const uint4* src = bricks + sslab * 8 + lane;
            const unsigned dst = sm_lane + (unsigned)(g % MMA_STAGES) * MMA_STAGE_BYTES;
            if (sslab + MMA_SUB_SLABS <= nslabs) {
#pragma unroll
                for (int j = 0; j < 8; ++j) mm_cp_async16(dst + j * (4 * 144), src + j * 32);
            } else {                                                // the volume's last, partial sub-chunk (whole bricks: nslabs = 8 nbricks)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (sslab + (lane >> 3) + 4 * j < nslabs) mm_cp_async16(dst + j * (4 * 144), src + j * 32);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");     // one group per call (empty past the end): group g is sub-chunk g
    };
    // z classes over the brick's 8 lanes (lane & 7 = z) without indexed shuffles: after the first butterfly step lane 0 holds
    // min(z0, z1) = class L, after all three every lane holds class A, and lane 7's own word is class H -- lanes 7 / 2 / 0 store.
    const unsigned z = lane & 7u;
    const bool z_h = z == 7u, z_l = z == 0u, stores = z_h || z_l || z == 2u;
    unsigned* pws = pw + (z_h ? 0 : (z_l ? 2 : 1)) * MMA_PW_CLASS + (lane >> 3);
#pragma unroll
    for (int g = 0; g < MMA_STAGES; ++g) prefetch(g);
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
        const int s = g & 7;
        asm volatile("cp.async.wait_group %0;\n" ::"n"(MMA_STAGES - 1) : "memory");
        __syncwarp();
        // a dead slab (past the volume's end) reads stale shared memory: whole bricks are dead together and their words are never written out
        const uint4* src = (const uint4*)(wbase + (g % MMA_STAGES) * MMA_STAGE_BYTES + lane * 144);
        uint4 r[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) r[y] = src[y];
        __syncwarp();                                               // every lane has its rows in registers: the stage is free again
        prefetch(g + MMA_STAGES);
        unsigned An = 0xffffffffu, Ax = 0u, Ln = 0xffffffffu, Lx = 0u, Hn = 0xffffffffu, Hx = 0u;
        unsigned yl[6] = {0, 0, 0, 0, 0, 0};
        unsigned rn = 0, rx = 0, r0 = 0, r3 = 0;
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            rn = __vminu2(__vminu2(r[y].x, r[y].y), __vminu2(r[y].z, r[y].w));
            rx = __vmaxu2(__vmaxu2(r[y].x, r[y].y), __vmaxu2(r[y].z, r[y].w));
            r0 = r[y].x; r3 = r[y].w;
            An = __vminu2(An, rn); Ax = __vmaxu2(Ax, rx);
            Ln = __vminu2(Ln, r0); Lx = __vmaxu2(Lx, r0);
            Hn = __vminu2(Hn, r3); Hx = __vmaxu2(Hx, r3);
            if (y == 1) { yl[0] = An; yl[1] = Ax; yl[2] = Ln; yl[3] = Lx; yl[4] = Hn; yl[5] = Hx; }
        }
        // x = 7 is the high half of word 3: duplicate it into both halves (PRMT) so the generic fold applies
        unsigned v[9];
        {
            const unsigned t3 = __byte_perm(r3, 0, 0x3232);
            v[0] = mm_pack2(t3, t3); v[1] = mm_pack2(rn, rx); v[2] = mm_pack2(r0, r0);
        }
        v[3] = mm_pack2(__byte_perm(Hn, 0, 0x3232), __byte_perm(Hx, 0, 0x3232)); v[4] = mm_pack2(An, Ax); v[5] = mm_pack2(Ln, Lx);
        v[6] = mm_pack2(__byte_perm(yl[4], 0, 0x3232), __byte_perm(yl[5], 0, 0x3232)); v[7] = mm_pack2(yl[0], yl[1]); v[8] = mm_pack2(yl[2], yl[3]);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const unsigned s1 = __vminu2(v[k], __shfl_xor_sync(0xffffffffu, v[k], 1));
            unsigned za = __vminu2(s1, __shfl_xor_sync(0xffffffffu, s1, 2));
            za = __vminu2(za, __shfl_xor_sync(0xffffffffu, za, 4));
            v[k] = z_h ? v[k] : (z_l ? s1 : za);
        }
        if (stores) {
#pragma unroll
            for (int k = 0; k < 9; ++k) pws[k * 32 + s * 4] = v[k];
        }
        if (s == MMA_SUBS - 1) {
            __syncwarp();
            const long long brick0 = (gw + (long long)(g >> 3) * tw) * 32;
            const int nb = (int)min((long long)32, nbricks - brick0);
            if ((int)lane < nb) {
#pragma unroll
                for (int zc = 0; zc < 3; ++zc)
#pragma unroll
                    for (int k = 0; k < 9; ++k)                     // one full 128-byte line per stream
                        part[(size_t)(zc * 9 + k) * nbricks + brick0 + lane] = pw[zc * MMA_PW_CLASS + k * 32 + lane];
            }
            __syncwarp();
        }
    }
}

__global__ void cell_fold_u16_kernel(const unsigned* __restrict__ part, float2* __restrict__ out, int ncx, int ncy, int ncz, float norm) {
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ncell = (long long)ncx * ncy * ncz;
    if (cell >= ncell) return;
    const int cx = (int)(cell % ncx), cy = (int)((cell / ncx) % ncy), cz = (int)(cell / ((long long)ncx * ncy));
    // 27 independent loads (addresses clamped into the grid, out-of-grid neighbours replaced by the neutral word afterwards), so they
    // are all in flight before the first fold
    unsigned w[27];
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = min(max(cx + dx, 0), ncx - 1), y = min(max(cy + dy, 0), ncy - 1), z = min(max(cz + dz, 0), ncz - 1);
                const int k = (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1);
                w[k] = __ldg(part + (size_t)k * ncell + ((long long)z * ncy + y) * ncx + x);
            }
    unsigned acc = 0xffffffffu;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const bool in = (unsigned)(cx + dx) < (unsigned)ncx && (unsigned)(cy + dy) < (unsigned)ncy && (unsigned)(cz + dz) < (unsigned)ncz;
                acc = __vminu2(acc, in ? w[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)] : 0xffffffffu);
            }
    // same normalisation as the sampler: raw * (1/65535) -- monotone, so min/max commute with it
    out[cell] = make_float2(__fmul_rn((float)(acc & 0xffffu), norm), __fmul_rn((float)(65535u - (acc >> 16)), norm));
}

// reference-style region structure with arbitrary cell size (for b200rc_region_minmax / proxy geometry):
// min/max over [llf-1, urb+1] clamped, real-world units (optimizedproxygeometry.cpp:944-968)
template <typename T>
__global__ void region_minmax_kernel(const T* __restrict__ bricks, float* __restrict__ out, int dx, int dy, int dz, int nbx, int nby,
                                     int step, int ncx, int ncy, int ncz, int nch, int nch_st, float norm_div, float sc, float of) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long ncell = (long long)ncx * ncy * ncz;
    if (warp >= ncell) return;
    const int cx = (int)(warp % ncx), cy = (int)((warp / ncx) % ncy), cz = (int)(warp / ((long long)ncx * ncy));
    int lo[3] = {cx * step, cy * step, cz * step}, hi[3] = {(cx + 1) * step - 1, (cy + 1) * step - 1, (cz + 1) * step - 1};
    const int dim[3] = {dx, dy, dz};
    for (int a = 0; a < 3; ++a) {
        hi[a] = min(hi[a], dim[a] - 1);
        lo[a] = min(max(lo[a] - 1, 0), dim[a] - 1);
        hi[a] = min(max(hi[a] + 1, 0), dim[a] - 1);
    }
    const int ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1, ez = hi[2] - lo[2] + 1;
    const long long total = (long long)ex * ey * ez;
    float mn[4] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f, 3.402823466e38f};
    float mx[4] = {1.175494351e-38f, 1.175494351e-38f, 1.175494351e-38f, 1.175494351e-38f};   // numeric_limits<float>::min() (quirk kept)
    for (long long i = lane; i < total; i += 32) {
        const int x = lo[0] + (int)(i % ex), y = lo[1] + (int)((i / ex) % ey), z = lo[2] + (int)(i / ((long long)ex * ey));
        const long long e = (brick_off_x(x) + brick_off_y(y, nbx) + brick_off_z(z, nbx, nby)) * nch_st;
        for (int c = 0; c < nch; ++c) {
            float v = (float)bricks[e + c];
            if (norm_div > 0.f) v = __fdiv_rn(v, norm_div);             // VolumeRAM::getVoxelNormalized
            v = __fadd_rn(__fmul_rn(v, sc), of);                        // RealWorldMapping::normalizedToRealWorld
            mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v);
        }
    }
    for (int c = 0; c < nch; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if (lane == 0) { out[(warp * nch + c) * 2] = mn[c]; out[(warp * nch + c) * 2 + 1] = mx[c]; }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// occupancy bits for the current transfer function(s)
// ------------------------------------------------------------------------------------------------------------------
struct OccCls { const int* nz; int res; int linear; float lo, hi; };
struct OccArgs { OccCls c[4]; int nch; float rwm_scale, rwm_offset; int border_on; float border; };

__device__ __forceinline__ float occ_rw_to_tex(float lo, float hi, float v) {
    if (v <= lo) return 0.0f;
    if (v >= hi) return 1.0f;
    if (hi == lo) return 0.0f;
    return __fdiv_rn(__fsub_rn(v, lo), __fsub_rn(hi, lo));
}

// level-1 maximum (MIP culling): max over the 8^3 level-0 cells of a 64^3-voxel cell, channel 0
__global__ void cellmax1_kernel(const float2* __restrict__ minmax, float* __restrict__ max1, int ncx, int ncy, int ncz, int nch, int n1x, int n1y, int n1z) {
    const long long c1 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c1 >= (long long)n1x * n1y * n1z) return;
    const int x1 = (int)(c1 % n1x), y1 = (int)((c1 / n1x) % n1y), z1 = (int)(c1 / ((long long)n1x * n1y));
    float m = -3.0e38f;
    for (int z = z1 * 8; z < min(z1 * 8 + 8, ncz); ++z)
        for (int y = y1 * 8; y < min(y1 * 8 + 8, ncy); ++y)
            for (int x = x1 * 8; x < min(x1 * 8 + 8, ncx); ++x) m = fmaxf(m, minmax[(((long long)z * ncy + y) * ncx + x) * nch].y);
    max1[c1] = m;
}

__global__ void occupancy0_kernel(const float2* __restrict__ minmax, uint32_t* __restrict__ occ, int ncx, int ncy, int ncz, int wpr,
                                  const __grid_constant__ OccArgs a) {
    // thread per cell, x padded to 32 so that one ballot = one occupancy word
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int xpad = wpr * 32;
    const long long total = (long long)xpad * ncy * ncz;
    if (t >= total) return;
    const int cx = (int)(t % xpad), cy = (int)((t / xpad) % ncy), cz = (int)(t / ((long long)xpad * ncy));
    bool nonempty = false;
    if (cx < ncx) {
        const long long cell = ((long long)cz * ncy + cy) * ncx + cx;
        for (int c = 0; c < a.nch && !nonempty; ++c) {
            const OccCls& k = a.c[c];
            if (!k.nz) { nonempty = true; break; }
            float2 mm = minmax[cell * a.nch + c];
            // CLAMP_TO_BORDER: samples next to the volume faces blend in the border colour
            if (a.border_on && (cx == 0 || cy == 0 || cz == 0 || cx == ncx - 1 || cy == ncy - 1 || cz == ncz - 1)) {
                mm.x = fminf(mm.x, a.border); mm.y = fmaxf(mm.y, a.border);
            }
            const float vmin = __fadd_rn(__fmul_rn(mm.x, a.rwm_scale), a.rwm_offset), vmax = __fadd_rn(__fmul_rn(mm.y, a.rwm_scale), a.rwm_offset);
            const float s0 = occ_rw_to_tex(k.lo, k.hi, fminf(vmin, vmax)), s1 = occ_rw_to_tex(k.lo, k.hi, fmaxf(vmin, vmax));
            int i0, i1;
            if (k.linear) { i0 = __float2int_rd(s0 * (float)k.res - 0.5f) - 1; i1 = __float2int_rd(s1 * (float)k.res - 0.5f) + 2; }
            else { i0 = __float2int_rd(s0 * (float)k.res) - 1; i1 = __float2int_rd(s1 * (float)k.res) + 1; }
            i0 = min(max(i0, 0), k.res - 1); i1 = min(max(i1, 0), k.res - 1);
            nonempty = (k.nz[i1 + 1] - k.nz[i0]) != 0;
        }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, nonempty);
    if ((threadIdx.x & 31) == 0) occ[((long long)cz * ncy + cy) * wpr + (cx >> 5)] = word;
}

__global__ void occupancy1_kernel(const uint32_t* __restrict__ occ0, uint32_t* __restrict__ occ1, int n0x, int n0y, int n0z, int wpr0,
                                  int n1x, int n1y, int n1z, int wpr1) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int xpad = wpr1 * 32;
    const long long total = (long long)xpad * n1y * n1z;
    if (t >= total) return;
    const int cx = (int)(t % xpad), cy = (int)((t / xpad) % n1y), cz = (int)(t / ((long long)xpad * n1y));
    bool nonempty = false;
    if (cx < n1x) {
        for (int z = cz * 8; z < min(cz * 8 + 8, n0z) && !nonempty; ++z)
            for (int y = cy * 8; y < min(cy * 8 + 8, n0y) && !nonempty; ++y) {
                const uint32_t w = occ0[((long long)z * n0y + y) * wpr0 + ((cx * 8) >> 5)];
                nonempty = ((w >> ((cx * 8) & 31)) & 0xffu) != 0;
            }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, nonempty);
    if ((threadIdx.x & 31) == 0) occ1[((long long)cz * n1y + cy) * wpr1 + (cx >> 5)] = word;
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static size_t dtype_bytes(int dt) { return dt == B200RC_U8 ? 1 : (dt == B200RC_U16 ? 2 : 4); }

int rc_volume_release(b200rc_ctx* ctx, int slot) {
    VolumeSlot& v = ctx->vol[slot];
    if (v.tex) cudaDestroyTextureObject(v.tex);
    if (v.array) cudaFreeArray(v.array);
    cudaFree(v.bricks); cudaFree(v.minmax); cudaFree(v.max1); cudaFree(v.occ0); cudaFree(v.occ1); cudaFree(v.touched);
    v = VolumeSlot();
    rc_octree_release(ctx, slot);           // an octree describes the volume it was built from
    return B200RC_OK;
}

static int volume_prepare(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc) {
    if (slot < 0 || slot >= B200RC_MAX_VOLUMES || !desc) { ctx->set_error("bad volume slot/desc"); return B200RC_ERR_INVALID; }
    if (desc->channels < 1 || desc->channels > 4 || desc->dtype < B200RC_U8 || desc->dtype > B200RC_F32 ||
        desc->dims[0] < 1 || desc->dims[1] < 1 || desc->dims[2] < 1) { ctx->set_error("unsupported volume format"); return B200RC_ERR_INVALID; }
    rc_volume_release(ctx, slot);
    VolumeSlot& v = ctx->vol[slot];
    v.desc = *desc;
    v.sharded = false;
    for (int a = 0; a < 3; ++a) {
        v.res_lo[a] = 0; v.res_dims[a] = desc->dims[a]; v.own_lo[a] = 0; v.own_hi[a] = desc->dims[a];
    }
    const bool shard = desc->shard_hi[0] > desc->shard_lo[0] && desc->shard_hi[1] > desc->shard_lo[1] && desc->shard_hi[2] > desc->shard_lo[2];
    if (shard) {
        v.sharded = true;
        for (int a = 0; a < 3; ++a) {
            int lo = desc->shard_lo[a] - desc->shard_apron, hi = desc->shard_hi[a] + desc->shard_apron;
            lo = lo < 0 ? 0 : lo; hi = hi > desc->dims[a] ? desc->dims[a] : hi;
            v.res_lo[a] = lo; v.res_dims[a] = hi - lo;
            v.own_lo[a] = desc->shard_lo[a]; v.own_hi[a] = desc->shard_hi[a];
        }
    }
    v.stored_nch = desc->channels == 3 ? 4 : desc->channels;
    v.elem_bytes = dtype_bytes(desc->dtype);
    v.nbx = (v.res_dims[0] + 7) / 8; v.nby = (v.res_dims[1] + 7) / 8; v.nbz = (v.res_dims[2] + 7) / 8;
    v.brick_bytes = (size_t)v.nbx * v.nby * v.nbz * 512 * v.stored_nch * v.elem_bytes;
    if (cudaMalloc(&v.bricks, v.brick_bytes) != cudaSuccess) { cudaGetLastError(); ctx->set_error("out of device memory for the brick pool"); return B200RC_ERR_OOM; }
    return B200RC_OK;
}

static int volume_build_minmax(b200rc_ctx* ctx, int slot) {
    VolumeSlot& v = ctx->vol[slot];
    for (int a = 0; a < 3; ++a) v.cells[a] = (v.res_dims[a] + 7) / 8;
    const long long ncell = (long long)v.cells[0] * v.cells[1] * v.cells[2];
    RC_CHECK(cudaMalloc(&v.minmax, sizeof(float2) * ncell * v.desc.channels));
    const int threads = 256;
    const long long blocks = (ncell * 32 + threads - 1) / threads;
    const int nch = v.desc.channels;
    RC_CHECK(cudaEventRecord(ctx->ev[4], ctx->stream));
    static const bool slow_minmax = getenv("B200RC_MINMAX_SLOW") != nullptr;        // A/B switch for tests and measurements
    if (v.desc.dtype == B200RC_U16 && nch == 1 && !slow_minmax) {
        // streaming two-pass build (see brick_partials_u16_kernel); the partials are temporary (27 words per brick)
        unsigned* part = nullptr;
        if (cudaMalloc(&part, sizeof(unsigned) * 27 * (size_t)ncell) != cudaSuccess) { cudaGetLastError(); ctx->set_error("out of device memory for min/max partials"); return B200RC_ERR_OOM; }
        const long long slabs = ncell * 8;
        static const bool sync_partials = getenv("B200RC_MINMAX_SYNC") != nullptr;    // A/B: the register-staged kernel of round 1
        if (sync_partials) brick_partials_u16_kernel<<<(unsigned)((slabs + 255) / 256), 256, 0, ctx->stream>>>((const uint4*)v.bricks, part, ncell);
        else {
            const size_t smem = (size_t)MMA_WARPS * MMA_WARP_BYTES;
            RC_CHECK(cudaFuncSetAttribute(brick_partials_u16_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const long long nsuper = (ncell + 31) / 32;
            long long grid = (nsuper + MMA_WARPS - 1) / MMA_WARPS;
            if (grid > ctx->num_sms) grid = ctx->num_sms;       // persistent: one CTA of 16 warps per SM
            brick_partials_u16_async_kernel<<<(unsigned)grid, MMA_WARPS * 32, smem, ctx->stream>>>((const uint4*)v.bricks, part, ncell);
        }
        cell_fold_u16_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, ctx->stream>>>(part, v.minmax, v.cells[0], v.cells[1], v.cells[2], 1.0f / 65535.0f);
        ctx->launches++;
        RC_CHECK(cudaGetLastError());
        RC_CHECK(cudaEventRecord(ctx->ev[5], ctx->stream));
        RC_CHECK(cudaStreamSynchronize(ctx->stream));
        cudaFree(part);
    } else if (v.desc.dtype == B200RC_U16)
        minmax_kernel<unsigned short><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const unsigned short*)v.bricks, v.minmax, v.res_dims[0], v.res_dims[1], v.res_dims[2],
                                                                                    v.nbx, v.nby, v.cells[0], v.cells[1], v.cells[2], nch, v.stored_nch, 1.0f / 65535.0f);
    else if (v.desc.dtype == B200RC_U8)
        minmax_kernel<unsigned char><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const unsigned char*)v.bricks, v.minmax, v.res_dims[0], v.res_dims[1], v.res_dims[2],
                                                                                   v.nbx, v.nby, v.cells[0], v.cells[1], v.cells[2], nch, v.stored_nch, 1.0f / 255.0f);
    else
        minmax_kernel<float><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const float*)v.bricks, v.minmax, v.res_dims[0], v.res_dims[1], v.res_dims[2],
                                                                           v.nbx, v.nby, v.cells[0], v.cells[1], v.cells[2], nch, v.stored_nch, 1.0f);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    if (!(v.desc.dtype == B200RC_U16 && nch == 1 && !slow_minmax)) RC_CHECK(cudaEventRecord(ctx->ev[5], ctx->stream));
    RC_CHECK(cudaEventSynchronize(ctx->ev[5]));
    { float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); v.minmax_build_ms = ms; }
    v.occ0_wpr = (v.cells[0] + 31) / 32;
    for (int a = 0; a < 3; ++a) v.occ1_dims[a] = (v.cells[a] + 7) / 8;
    v.occ1_wpr = (v.occ1_dims[0] + 31) / 32;
    RC_CHECK(cudaMalloc(&v.occ0, sizeof(uint32_t) * (size_t)v.occ0_wpr * v.cells[1] * v.cells[2]));
    RC_CHECK(cudaMalloc(&v.occ1, sizeof(uint32_t) * (size_t)v.occ1_wpr * v.occ1_dims[1] * v.occ1_dims[2]));
    {
        const long long n1 = (long long)v.occ1_dims[0] * v.occ1_dims[1] * v.occ1_dims[2];
        RC_CHECK(cudaMalloc(&v.max1, sizeof(float) * n1));
        cellmax1_kernel<<<(unsigned)((n1 + 127) / 128), 128, 0, ctx->stream>>>(v.minmax, v.max1, v.cells[0], v.cells[1], v.cells[2], v.desc.channels,
                                                                              v.occ1_dims[0], v.occ1_dims[1], v.occ1_dims[2]);
        ctx->launches++;
        RC_CHECK(cudaGetLastError());
    }
    v.occ_key = 0;
    for (int a = 0; a < 3; ++a) v.touched_dims[a] = (v.desc.dims[a] + B200RC_STAT_CELL - 1) / B200RC_STAT_CELL;
    v.valid = true;
    return B200RC_OK;
}

int rc_volume_upload(b200rc_ctx* ctx, int slot, const void* data, const b200rc_volume_desc* desc) {
    if (!data) { ctx->set_error("null volume data"); return B200RC_ERR_INVALID; }
    int rc = volume_prepare(ctx, slot, desc);
    if (rc) return rc;
    VolumeSlot& v = ctx->vol[slot];
    // `data` holds the RESIDENT range (res_dims), x fastest, channels interleaved
    const size_t lin_bytes = (size_t)v.res_dims[0] * v.res_dims[1] * v.res_dims[2] * desc->channels * v.elem_bytes;
    const void* d_lin = data;
    void* staging = nullptr;
    if (!desc->data_on_device) {
        if (cudaMalloc(&staging, lin_bytes) != cudaSuccess) { cudaGetLastError(); ctx->set_error("out of device memory for upload staging"); return B200RC_ERR_OOM; }
        RC_CHECK(cudaMemcpyAsync(staging, data, lin_bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_lin = staging;
    }
    const long long nrows = (long long)v.nbx * v.nby * v.nbz * 64;
    const int threads = 256;
    const unsigned blocks = (unsigned)((nrows + threads - 1) / threads);
    const int dx = v.res_dims[0], dy = v.res_dims[1], dz = v.res_dims[2];
    if (desc->dtype == B200RC_U16 && desc->channels == 1 && dx % 8 == 0 && ((uintptr_t)d_lin % 16) == 0)
        brick_rows_u16x8_kernel<<<blocks, threads, 0, ctx->stream>>>((const uint4*)d_lin, (uint4*)v.bricks, dx / 8, dy, dz, v.nbx, v.nby, v.nbz);
    else if (desc->dtype == B200RC_U16)
        brick_rows_kernel<unsigned short><<<blocks, threads, 0, ctx->stream>>>((const unsigned short*)d_lin, (unsigned short*)v.bricks, dx, dy, dz, v.nbx, v.nby, v.nbz, desc->channels, v.stored_nch, (unsigned short)65535);
    else if (desc->dtype == B200RC_U8)
        brick_rows_kernel<unsigned char><<<blocks, threads, 0, ctx->stream>>>((const unsigned char*)d_lin, (unsigned char*)v.bricks, dx, dy, dz, v.nbx, v.nby, v.nbz, desc->channels, v.stored_nch, (unsigned char)255);
    else
        brick_rows_kernel<float><<<blocks, threads, 0, ctx->stream>>>((const float*)d_lin, (float*)v.bricks, dx, dy, dz, v.nbx, v.nby, v.nbz, desc->channels, v.stored_nch, 1.0f);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    if (staging) { RC_CHECK(cudaStreamSynchronize(ctx->stream)); cudaFree(staging); }
    return volume_build_minmax(ctx, slot);
}

// ---- streaming upload ------------------------------------------------------------------------------------------------------
int rc_volume_stream_begin(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc) {
    int rc = volume_prepare(ctx, slot, desc);
    if (rc) return rc;
    ctx->vol[slot].valid = false;           // not renderable before rc_volume_stream_end
    return B200RC_OK;
}

// d_lin: planes [z0, z0 + nz) of the resident range in DEVICE memory
static int volume_brick_slab(b200rc_ctx* ctx, VolumeSlot& v, int z0, int nz, const void* d_lin) {
    const int dx = v.res_dims[0], dy = v.res_dims[1], dz = v.res_dims[2];
    if (z0 < 0 || nz < 1 || z0 + nz > dz || (z0 & 7) || ((nz & 7) && z0 + nz != dz)) { ctx->set_error("upload slab: z0 / nz must be multiples of 8 (last slab excepted) inside the resident range"); return B200RC_ERR_INVALID; }
    const int bz0 = z0 >> 3, nbz_slab = (nz + 7) >> 3;
    const long long nrows = (long long)v.nbx * v.nby * nbz_slab * 64;
    const int threads = 256;
    const unsigned blocks = (unsigned)((nrows + threads - 1) / threads);
    const b200rc_volume_desc& d = v.desc;
    if (d.dtype == B200RC_U16 && d.channels == 1 && dx % 8 == 0 && ((uintptr_t)d_lin % 16) == 0)
        brick_rows_slab_u16x8_kernel<<<blocks, threads, 0, ctx->stream>>>((const uint4*)d_lin, (uint4*)v.bricks, dx / 8, dy, dz, v.nbx, v.nby, z0, nz, bz0, nbz_slab);
    else if (d.dtype == B200RC_U16)
        brick_rows_slab_kernel<unsigned short><<<blocks, threads, 0, ctx->stream>>>((const unsigned short*)d_lin, (unsigned short*)v.bricks, dx, dy, dz, v.nbx, v.nby, z0, nz, bz0, nbz_slab, d.channels, v.stored_nch, (unsigned short)65535);
    else if (d.dtype == B200RC_U8)
        brick_rows_slab_kernel<unsigned char><<<blocks, threads, 0, ctx->stream>>>((const unsigned char*)d_lin, (unsigned char*)v.bricks, dx, dy, dz, v.nbx, v.nby, z0, nz, bz0, nbz_slab, d.channels, v.stored_nch, (unsigned char)255);
    else
        brick_rows_slab_kernel<float><<<blocks, threads, 0, ctx->stream>>>((const float*)d_lin, (float*)v.bricks, dx, dy, dz, v.nbx, v.nby, z0, nz, bz0, nbz_slab, d.channels, v.stored_nch, 1.0f);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}

int rc_volume_stream_slab(b200rc_ctx* ctx, int slot, int z0, int nz, const void* data) {
    if (slot < 0 || slot >= B200RC_MAX_VOLUMES || !data) { ctx->set_error("upload slab: bad arguments"); return B200RC_ERR_INVALID; }
    VolumeSlot& v = ctx->vol[slot];
    if (!v.bricks) { ctx->set_error("upload slab: call b200rc_upload_volume_begin first"); return B200RC_ERR_NOT_READY; }
    if (v.desc.data_on_device) return volume_brick_slab(ctx, v, z0, nz, data);
    if (nz < 1 || z0 < 0 || z0 + nz > v.res_dims[2]) { ctx->set_error("upload slab: planes outside the resident range"); return B200RC_ERR_INVALID; }
    const size_t bytes = (size_t)v.res_dims[0] * v.res_dims[1] * nz * v.desc.channels * v.elem_bytes;
    void* staging = nullptr;
    if (cudaMalloc(&staging, bytes) != cudaSuccess) { cudaGetLastError(); ctx->set_error("out of device memory for upload staging"); return B200RC_ERR_OOM; }
    int rc = B200RC_OK;
    if (cudaMemcpyAsync(staging, data, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { ctx->set_error("upload slab: host to device copy failed"); rc = B200RC_ERR_CUDA; }
    if (!rc) rc = volume_brick_slab(ctx, v, z0, nz, staging);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(staging);
    return rc;
}

int rc_volume_stream_end(b200rc_ctx* ctx, int slot) {
    if (slot < 0 || slot >= B200RC_MAX_VOLUMES || !ctx->vol[slot].bricks) { ctx->set_error("upload end: no upload in progress"); return B200RC_ERR_NOT_READY; }
    return volume_build_minmax(ctx, slot);
}

// Raw file -> brick pool through two pinned buffers (see include/b200rc.h).
int rc_volume_load_raw(b200rc_ctx* ctx, int slot, const char* path, const b200rc_volume_desc* desc, uint64_t header_skip, int slab_planes) {
    if (!path || !desc) { ctx->set_error("load raw: bad arguments"); return B200RC_ERR_INVALID; }
    if (desc->data_on_device) { ctx->set_error("load raw: data_on_device makes no sense for a file"); return B200RC_ERR_INVALID; }
    FILE* f = fopen(path, "rb");
    if (!f) { ctx->set_error(std::string("Failed to open file '") + path + "' for reading."); return B200RC_ERR_INVALID; }
    int rc = rc_volume_stream_begin(ctx, slot, desc);
    if (rc) { fclose(f); return rc; }
    VolumeSlot& v = ctx->vol[slot];
    const size_t vox = (size_t)desc->channels * v.elem_bytes;
    const size_t row_bytes = (size_t)v.res_dims[0] * vox, plane_bytes = row_bytes * v.res_dims[1];
    // the file must hold the whole volume
    {
        fseeko(f, 0, SEEK_END);
        const unsigned long long have = (unsigned long long)ftello(f);
        const unsigned long long need = header_skip + (unsigned long long)desc->dims[0] * desc->dims[1] * desc->dims[2] * vox;
        if (have < need) { fclose(f); rc_volume_release(ctx, slot); ctx->set_error("load raw: file is shorter than the volume it should hold"); return B200RC_ERR_INVALID; }
    }
    int S = slab_planes > 0 ? (slab_planes + 7) / 8 * 8 : (int)(((size_t)64 << 20) / (plane_bytes ? plane_bytes : 1)) / 8 * 8;
    if (S < 8) S = 8;
    if (S > (v.res_dims[2] + 7) / 8 * 8) S = (v.res_dims[2] + 7) / 8 * 8;
    const size_t slab_bytes = plane_bytes * S;
    void* pinned[2] = {nullptr, nullptr}; void* dstage[2] = {nullptr, nullptr}; cudaEvent_t done[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
    auto cleanup = [&]() {
        cudaStreamSynchronize(ctx->stream);
        for (int k = 0; k < 2; ++k) { if (pinned[k]) cudaFreeHost(pinned[k]); if (dstage[k]) cudaFree(dstage[k]); if (done[k]) cudaEventDestroy(done[k]); }
        fclose(f);
    };
    for (int k = 0; k < 2; ++k) {
        if (cudaHostAlloc(&pinned[k], slab_bytes, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&dstage[k], slab_bytes) != cudaSuccess ||
            cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError(); cleanup(); rc_volume_release(ctx, slot); ctx->set_error("load raw: out of memory for the staging buffers"); return B200RC_ERR_OOM;
        }
    }
    const bool full_rows = v.res_dims[0] == desc->dims[0], full_planes = full_rows && v.res_dims[1] == desc->dims[1];
    int k = 0;
    for (int z0 = 0; z0 < v.res_dims[2] && !rc; z0 += S, k ^= 1) {
        const int nz = (z0 + S <= v.res_dims[2]) ? S : v.res_dims[2] - z0;
        if (used[k]) cudaEventSynchronize(done[k]);         // the copy out of this pinned buffer has finished
        unsigned char* dst = (unsigned char*)pinned[k];
        bool ok = true;
        if (full_planes) {
            const unsigned long long off = header_skip + (unsigned long long)(v.res_lo[2] + z0) * desc->dims[1] * desc->dims[0] * vox;
            ok = fseeko(f, (off_t)off, SEEK_SET) == 0 && fread(dst, 1, plane_bytes * nz, f) == plane_bytes * nz;
        } else {
            for (int z = 0; z < nz && ok; ++z)
                for (int y = 0; y < v.res_dims[1] && ok; ++y) {
                    const unsigned long long off = header_skip + ((((unsigned long long)(v.res_lo[2] + z0 + z) * desc->dims[1]) + (v.res_lo[1] + y)) * desc->dims[0] + v.res_lo[0]) * vox;
                    if (y == 0 || !full_rows) ok = fseeko(f, (off_t)off, SEEK_SET) == 0;
                    const size_t want = full_rows ? row_bytes * v.res_dims[1] : row_bytes;
                    ok = ok && fread(dst + ((size_t)z * v.res_dims[1] + y) * row_bytes, 1, want, f) == want;
                    if (full_rows) break;           // the whole plane's resident rows are contiguous
                }
        }
        if (!ok) { ctx->set_error("load raw: read failed"); rc = B200RC_ERR_INVALID; break; }
        if (cudaMemcpyAsync(dstage[k], pinned[k], plane_bytes * nz, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { ctx->set_error("load raw: host to device copy failed"); rc = B200RC_ERR_CUDA; break; }
        rc = volume_brick_slab(ctx, v, z0, nz, dstage[k]);
        cudaEventRecord(done[k], ctx->stream);
        used[k] = true;
    }
    cleanup();
    if (rc) { rc_volume_release(ctx, slot); return rc; }
    return volume_build_minmax(ctx, slot);
}

// the 64 spheres of the synthetic volume; voreen_b200/synth.py mirrors this arithmetic exactly
static void synth_spheres(uint64_t seed, SynthSpheres& s) {
    const uint32_t sd = (uint32_t)seed;
    for (int i = 0; i < 64; ++i) {
        const uint32_t h0 = hash32(sd * 0x9e3779b9u + (uint32_t)(i * 4 + 0)), h1 = hash32(sd * 0x9e3779b9u + (uint32_t)(i * 4 + 1));
        const uint32_t h2 = hash32(sd * 0x9e3779b9u + (uint32_t)(i * 4 + 2)), h3 = hash32(sd * 0x9e3779b9u + (uint32_t)(i * 4 + 3));
        s.cx[i] = ((float)(h0 & 0xffffu) * (1.0f / 65536.0f) - 0.5f) * 0.76f;
        s.cy[i] = ((float)(h1 & 0xffffu) * (1.0f / 65536.0f) - 0.5f) * 0.76f;
        s.cz[i] = ((float)(h2 & 0xffffu) * (1.0f / 65536.0f) - 0.5f) * 0.76f;
        const float r = 0.012f + (float)(h3 & 0xffffu) * (1.0f / 65536.0f) * 0.02f;
        s.r2[i] = r * r;
    }
}

int rc_volume_synth(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc, uint64_t seed) {
    if (!desc || desc->dtype != B200RC_U16 || desc->channels != 1) { ctx->set_error("synthetic volume is u16, 1 channel"); return B200RC_ERR_INVALID; }
    int rc = volume_prepare(ctx, slot, desc);
    if (rc) return rc;
    VolumeSlot& v = ctx->vol[slot];
    SynthSpheres sph;
    synth_spheres(seed, sph);
    const long long n = (long long)v.nbx * v.nby * v.nbz * 512;
    const int threads = 256;
    const int mind = desc->dims[0] < desc->dims[1] ? (desc->dims[0] < desc->dims[2] ? desc->dims[0] : desc->dims[2]) : (desc->dims[1] < desc->dims[2] ? desc->dims[1] : desc->dims[2]);
    const float fall_scale = 0.36f * (float)mind;
    synth_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, ctx->stream>>>(
        (unsigned short*)v.bricks, desc->dims[0], desc->dims[1], desc->dims[2], v.res_lo[0], v.res_lo[1], v.res_lo[2], v.res_dims[0], v.res_dims[1], v.res_dims[2],
        v.nbx, v.nby, v.nbz, 1.0f / (float)desc->dims[0], 1.0f / (float)desc->dims[1], 1.0f / (float)desc->dims[2], fall_scale, (uint32_t)seed, sph);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return volume_build_minmax(ctx, slot);
}

int rc_volume_download(b200rc_ctx* ctx, int slot, void* host_dst, size_t bytes) {
    VolumeSlot& v = ctx->vol[slot];
    if (!v.valid) { ctx->set_error("no volume in slot"); return B200RC_ERR_NOT_READY; }
    const long long n = (long long)v.res_dims[0] * v.res_dims[1] * v.res_dims[2];
    const size_t need = (size_t)n * v.desc.channels * v.elem_bytes;
    if (bytes < need) { ctx->set_error("download buffer too small"); return B200RC_ERR_INVALID; }
    void* lin = nullptr;
    RC_CHECK(cudaMalloc(&lin, need));
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    if (v.desc.dtype == B200RC_U16) debrick_kernel<unsigned short><<<blocks, threads, 0, ctx->stream>>>((const unsigned short*)v.bricks, (unsigned short*)lin, v.res_dims[0], v.res_dims[1], v.res_dims[2], v.nbx, v.nby, v.desc.channels, v.stored_nch);
    else if (v.desc.dtype == B200RC_U8) debrick_kernel<unsigned char><<<blocks, threads, 0, ctx->stream>>>((const unsigned char*)v.bricks, (unsigned char*)lin, v.res_dims[0], v.res_dims[1], v.res_dims[2], v.nbx, v.nby, v.desc.channels, v.stored_nch);
    else debrick_kernel<float><<<blocks, threads, 0, ctx->stream>>>((const float*)v.bricks, (float*)lin, v.res_dims[0], v.res_dims[1], v.res_dims[2], v.nbx, v.nby, v.desc.channels, v.stored_nch);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    RC_CHECK(cudaMemcpyAsync(host_dst, lin, need, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    cudaFree(lin);
    return B200RC_OK;
}

int rc_volume_ensure_texture(b200rc_ctx* ctx, int slot, int filter, int clamp, float border) {
    VolumeSlot& v = ctx->vol[slot];
    if (!v.valid) { ctx->set_error("no volume in slot"); return B200RC_ERR_NOT_READY; }
    if (!v.array) {
        cudaChannelFormatDesc fd;
        const int bits = (int)v.elem_bytes * 8;
        const cudaChannelFormatKind kind = v.desc.dtype == B200RC_F32 ? cudaChannelFormatKindFloat : cudaChannelFormatKindUnsigned;
        fd = cudaCreateChannelDesc(bits, v.stored_nch >= 2 ? bits : 0, v.stored_nch >= 4 ? bits : 0, v.stored_nch >= 4 ? bits : 0, kind);
        cudaExtent ext = make_cudaExtent((size_t)v.res_dims[0], (size_t)v.res_dims[1], (size_t)v.res_dims[2]);
        if (cudaMalloc3DArray(&v.array, &fd, ext, cudaArraySurfaceLoadStore) != cudaSuccess) {
            ctx->set_error(std::string("cudaMalloc3DArray: ") + cudaGetErrorString(cudaGetLastError()));
            v.array = nullptr;
            return B200RC_ERR_OOM;
        }
        cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypeArray; rd.res.array.array = v.array;
        cudaSurfaceObject_t surf = 0;
        RC_CHECK(cudaCreateSurfaceObject(&surf, &rd));
        const long long n = (long long)v.res_dims[0] * v.res_dims[1] * v.res_dims[2];
        const int threads = 256;
        const unsigned blocks = (unsigned)((n + threads - 1) / threads);
        if (v.desc.dtype == B200RC_U16) bricks_to_surface_kernel<unsigned short><<<blocks, threads, 0, ctx->stream>>>((const unsigned short*)v.bricks, surf, v.res_dims[0], v.res_dims[1], v.res_dims[2], v.nbx, v.nby, v.stored_nch);
        else if (v.desc.dtype == B200RC_U8) bricks_to_surface_kernel<unsigned char><<<blocks, threads, 0, ctx->stream>>>((const unsigned char*)v.bricks, surf, v.res_dims[0], v.res_dims[1], v.res_dims[2], v.nbx, v.nby, v.stored_nch);
        else bricks_to_surface_kernel<float><<<blocks, threads, 0, ctx->stream>>>((const float*)v.bricks, surf, v.res_dims[0], v.res_dims[1], v.res_dims[2], v.nbx, v.nby, v.stored_nch);
        ctx->launches++;
        RC_CHECK(cudaGetLastError());
        RC_CHECK(cudaStreamSynchronize(ctx->stream));
        cudaDestroySurfaceObject(surf);
    }
    if (v.tex && (v.tex_filter != filter || v.tex_clamp != clamp || v.tex_border != border)) { cudaDestroyTextureObject(v.tex); v.tex = 0; }
    if (!v.tex) {
        cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypeArray; rd.res.array.array = v.array;
        cudaTextureDesc td; memset(&td, 0, sizeof td);
        const cudaTextureAddressMode am = clamp == B200RC_CLAMP_TO_BORDER ? cudaAddressModeBorder : cudaAddressModeClamp;
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = am;
        td.filterMode = filter == B200RC_FILTER_NEAREST ? cudaFilterModePoint : cudaFilterModeLinear;
        td.readMode = v.desc.dtype == B200RC_F32 ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
        td.normalizedCoords = 0;
        td.borderColor[0] = td.borderColor[1] = td.borderColor[2] = td.borderColor[3] = border;
        RC_CHECK(cudaCreateTextureObject(&v.tex, &rd, &td, nullptr));
        v.tex_filter = filter; v.tex_clamp = clamp; v.tex_border = border;
    }
    return B200RC_OK;
}

int rc_volume_region_minmax(b200rc_ctx* ctx, int slot, int cell_size, int32_t cells[3], float* out, size_t out_floats) {
    VolumeSlot& v = ctx->vol[slot];
    if (!v.valid || v.sharded) { ctx->set_error("region_minmax needs a resident, unsharded volume"); return B200RC_ERR_NOT_READY; }
    if (cell_size < 1) cell_size = 1;
    for (int a = 0; a < 3; ++a) cells[a] = (v.desc.dims[a] + cell_size - 1) / cell_size;
    const long long ncell = (long long)cells[0] * cells[1] * cells[2];
    const size_t need = (size_t)ncell * v.desc.channels * 2;
    if (!out) return B200RC_OK;
    if (out_floats < need) { ctx->set_error("region_minmax output too small"); return B200RC_ERR_INVALID; }
    float* d_out = nullptr;
    RC_CHECK(cudaMalloc(&d_out, need * sizeof(float)));
    const int threads = 256;
    const unsigned blocks = (unsigned)((ncell * 32 + threads - 1) / threads);
    const int nch = v.desc.channels;
    if (v.desc.dtype == B200RC_U16)
        region_minmax_kernel<unsigned short><<<blocks, threads, 0, ctx->stream>>>((const unsigned short*)v.bricks, d_out, v.desc.dims[0], v.desc.dims[1], v.desc.dims[2], v.nbx, v.nby, cell_size, cells[0], cells[1], cells[2], nch, v.stored_nch, 65535.0f, v.desc.rwm_scale, v.desc.rwm_offset);
    else if (v.desc.dtype == B200RC_U8)
        region_minmax_kernel<unsigned char><<<blocks, threads, 0, ctx->stream>>>((const unsigned char*)v.bricks, d_out, v.desc.dims[0], v.desc.dims[1], v.desc.dims[2], v.nbx, v.nby, cell_size, cells[0], cells[1], cells[2], nch, v.stored_nch, 255.0f, v.desc.rwm_scale, v.desc.rwm_offset);
    else
        region_minmax_kernel<float><<<blocks, threads, 0, ctx->stream>>>((const float*)v.bricks, d_out, v.desc.dims[0], v.desc.dims[1], v.desc.dims[2], v.nbx, v.nby, cell_size, cells[0], cells[1], cells[2], nch, v.stored_nch, 0.0f, v.desc.rwm_scale, v.desc.rwm_offset);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    RC_CHECK(cudaMemcpyAsync(out, d_out, need * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_out);
    return B200RC_OK;
}

int rc_volume_build_occupancy(b200rc_ctx* ctx, int slot, const int* cls_modes, int n_channels, bool multi_slot, int clamp, float border) {
    VolumeSlot& v = ctx->vol[slot];
    // key = TF generations + modes
    unsigned long long key = 1469598103934665603ull;
    OccArgs a; memset(&a, 0, sizeof a);
    a.nch = n_channels; a.rwm_scale = v.desc.rwm_scale; a.rwm_offset = v.desc.rwm_offset;
    a.border_on = clamp == B200RC_CLAMP_TO_BORDER; a.border = border;
    if (a.border_on) { unsigned int bits; memcpy(&bits, &border, 4); key = (key ^ (0x100000000ull | bits)) * 1099511628211ull; }
    for (int c = 0; c < n_channels; ++c) {
        const ClassifierSlot& k = ctx->cls[multi_slot ? slot : c];
        const int mode = cls_modes[c];
        OccCls& o = a.c[c];
        o.lo = k.dom[0]; o.hi = k.dom[1];
        if (mode == B200RC_CLASS_TF) { o.nz = k.nz_tf; o.res = k.tf_width; o.linear = 0; }
        else if (mode == B200RC_CLASS_PREINT) { o.nz = k.nz_pi; o.res = k.pres; o.linear = 1; }
        else { o.nz = k.nz_2d; o.res = k.tf2d_w; o.linear = 1; }
        key = (key ^ k.generation) * 1099511628211ull;
        key = (key ^ (unsigned long long)(mode + 1)) * 1099511628211ull;
    }
    if (key == v.occ_key) return B200RC_OK;
    const int threads = 256;
    {
        const long long total = (long long)v.occ0_wpr * 32 * v.cells[1] * v.cells[2];
        occupancy0_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, ctx->stream>>>(v.minmax, v.occ0, v.cells[0], v.cells[1], v.cells[2], v.occ0_wpr, a);
        ctx->launches++;
        RC_CHECK(cudaGetLastError());
    }
    {
        const long long total = (long long)v.occ1_wpr * 32 * v.occ1_dims[1] * v.occ1_dims[2];
        occupancy1_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, ctx->stream>>>(v.occ0, v.occ1, v.cells[0], v.cells[1], v.cells[2], v.occ0_wpr,
                                                                                                 v.occ1_dims[0], v.occ1_dims[1], v.occ1_dims[2], v.occ1_wpr);
        ctx->launches++;
        RC_CHECK(cudaGetLastError());
    }
    v.occ_key = key;
    return B200RC_OK;
}
