// octree.cu -- the multi-resolution brick pool of the reference's large-volume path: VolumeOctree construction
// (V/src/core/datastructures/octree/volumeoctree.cpp:637-961,1186-1536) and the DVR traversal of SingleOctreeRaycasterCL
// (V/modules/opencl/cl/singleoctreeraycaster.cl:190-300, octree_config.cl:313-368,402-413,437-646, octree_dvr.cl:52-80).
//
// B200 layout: no pointer tree.  Level k (0 = root ... L-1 = full resolution) is a dense grid of 2^k nodes per axis; all levels are
// concatenated (offset (8^k - 1) / 7), a node is {avg, min, max : u16; flags : u8} in structure-of-arrays form and its brick --
// brick_dim^3 uint16 voxels, x fastest, as in the reference -- sits at node_index * brick_dim^3 of one pool allocation: 180 GB of
// HBM hold the complete pyramid of a 2048^3 volume (19.6 GB), so the request / in-use flag protocol with which the reference streams
// bricks into a small OpenCL buffer (singleoctreeraycastercl.cpp:1119-1556) has nothing left to do and is not reproduced.
// A node an optimised (homogeneous) ancestor replaced keeps its slot but loses its EXISTS flag (deleteSubTree).
#include "ctx.hpp"
#include <cstring>
#include <new>
#include <vector>

#define OCT_EXISTS 1u
#define OCT_IN_VOLUME 2u
#define OCT_HAS_BRICK 4u
#define OCT_HAS_CHILDREN 8u

struct OctreeDev {
    uint16_t* avg; uint16_t* mn; uint16_t* mx; uint8_t* flags; uint16_t* bricks;
    int levels, brick_dim, octree_dim;
    int dims[3];
};

__host__ __device__ inline size_t oct_level_offset(int k) { return (((size_t)1 << (3 * k)) - 1) / 7; }

// ---- level L-1: bricks cut from the volume (extractBrickFromTexture :1243-1340, createTreeNodeFromTexture :1186-1241) ----------
template <int DT>
__global__ void __launch_bounds__(256) octree_leaf_kernel(VolDev v, OctreeDev o, int thr, int optimize) {
    const int k = o.levels - 1, G = 1 << k, B = o.brick_dim;
    const int nx = blockIdx.x % G, ny = (blockIdx.x / G) % G, nz = blockIdx.x / (G * G);
    const size_t node = oct_level_offset(k) + ((size_t)nz * G + ny) * G + nx;
    uint16_t* brick = o.bricks + node * (size_t)B * B * B;
    const int lx = nx * B, ly = ny * B, lz = nz * B;
    const int sx = min(max(o.dims[0] - lx, 0), B), sy = min(max(o.dims[1] - ly, 0), B), sz = min(max(o.dims[2] - lz, 0), B);   // voxels inside the volume
    unsigned long long sum = 0; unsigned mn = 65535u, mx = 0u;
    for (int i = threadIdx.x; i < B * B * B; i += blockDim.x) {
        const int x = i % B, y = (i / B) % B, z = i / (B * B);
        unsigned val = 0;
        if (x < sx && y < sy && z < sz) {
            const long long e = brick_off_x(lx + x) + brick_off_y(ly + y, v.nbx) + brick_off_z(lz + z, v.nbx, v.nby);
            if (DT == B200RC_U16) val = ((const unsigned short*)v.bricks)[e];
            else val = (unsigned)((const unsigned char*)v.bricks)[e] << 8;          // convertVoxelValueToUInt16<uint8_t> (:171-173)
            sum += val; mn = min(mn, val); mx = max(mx, val);
        }
        brick[i] = (uint16_t)val;                                                   // outside the volume: zero (:1323-1333)
    }
    __shared__ unsigned long long s_sum[8]; __shared__ unsigned s_mn[8], s_mx[8];
    for (int d = 16; d > 0; d >>= 1) { sum += __shfl_down_sync(0xffffffffu, sum, d); mn = min(mn, __shfl_down_sync(0xffffffffu, mn, d)); mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d)); }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { sum += s_sum[w]; mn = min(mn, s_mn[w]); mx = max(mx, s_mx[w]); }
        const unsigned long long nsig = (unsigned long long)sx * sy * sz;
        unsigned f = OCT_EXISTS;
        if (nsig == 0) { o.avg[node] = 0; o.mn[node] = 0; o.mx[node] = 0; }        // dummy node outside the volume (:764-767)
        else {
            o.avg[node] = (uint16_t)(sum / nsig); o.mn[node] = (uint16_t)mn; o.mx[node] = (uint16_t)mx;
            f |= OCT_IN_VOLUME;
            const bool homogeneous = (int)(mx - mn) <= thr;
            if (!homogeneous || !optimize) f |= OCT_HAS_BRICK;
        }
        o.flags[node] = (uint8_t)f;
    }
}

// ---- level k from level k + 1 (createParentNodeConstChannels :1382-1536) ---------------------------------------------------------
__device__ __forceinline__ unsigned oct_halfsample(int fn, const unsigned v[8]) {
    if (fn == 0) { unsigned long long s = 0; for (int i = 0; i < 8; ++i) s += v[i]; return (unsigned)(s / 8); }     // HalfSampleMean (:188-202)
    unsigned r = v[0];
    for (int i = 1; i < 8; ++i) r = (fn == 1) ? max(r, v[i]) : min(r, v[i]);                                       // HalfSampleMax / Min
    return r;
}

__global__ void __launch_bounds__(256) octree_parent_kernel(OctreeDev o, int k, int thr, int optimize, int halfsample_fn) {
    const int G = 1 << k, B = o.brick_dim, H = B / 2, Gc = G * 2;
    const int nx = blockIdx.x % G, ny = (blockIdx.x / G) % G, nz = blockIdx.x / (G * G);
    const size_t node = oct_level_offset(k) + ((size_t)nz * G + ny) * G + nx;
    const size_t coff = oct_level_offset(k + 1);
    __shared__ int s_leaf;
    if (threadIdx.x == 0) {
        unsigned long long asum = 0; unsigned mn = 65535u, mx = 0u; int nonempty = 0;
        for (int c = 0; c < 8; ++c) {
            const size_t ci = coff + ((size_t)(nz * 2 + (c >> 2)) * Gc + (ny * 2 + ((c >> 1) & 1))) * Gc + (nx * 2 + (c & 1));
            if (!(o.flags[ci] & OCT_IN_VOLUME)) continue;
            ++nonempty; asum += o.avg[ci]; mn = min(mn, (unsigned)o.mn[ci]); mx = max(mx, (unsigned)o.mx[ci]);
        }
        unsigned f = OCT_EXISTS;
        bool leaf;
        if (nonempty > 0) {
            o.avg[node] = (uint16_t)(asum / (unsigned long long)nonempty); o.mn[node] = (uint16_t)mn; o.mx[node] = (uint16_t)mx;
            f |= OCT_IN_VOLUME;
            leaf = ((int)(mx - mn) <= thr) && optimize;
        } else { o.avg[node] = 0; o.mn[node] = 0; o.mx[node] = 0; leaf = true; }
        if (!leaf) f |= OCT_HAS_BRICK | OCT_HAS_CHILDREN;
        o.flags[node] = (uint8_t)f;
        s_leaf = leaf ? 1 : 0;
    }
    __syncthreads();
    if (s_leaf) return;
    // extent of the child-level volume inside this parent, in child-level voxels (:898-901)
    const int shift = o.levels - 1 - (k + 1);           // child level voxel = 2^shift full-resolution voxels
    int urb[3];
    {
        const int n[3] = {nx, ny, nz};
        for (int a = 0; a < 3; ++a) {
            const int cdim = (o.dims[a] + (1 << shift) - 1) >> shift;          // ceil(volumeDim / 2^shift)
            const int llf = min(n[a] * B * 2, cdim);
            urb[a] = min(llf + 2 * B, cdim) - llf;
        }
    }
    uint16_t* brick = o.bricks + node * (size_t)B * B * B;
    for (int i = threadIdx.x; i < B * B * B; i += blockDim.x) {
        const int px = i % B, py = (i / B) % B, pz = i / (B * B);
        const int cx = px / H, cy = py / H, cz = pz / H;                        // which child
        const size_t ci = coff + ((size_t)(nz * 2 + cz) * Gc + (ny * 2 + cy)) * Gc + (nx * 2 + cx);
        unsigned val;
        if (o.flags[ci] & OCT_HAS_BRICK) {
            const uint16_t* cb = o.bricks + ci * (size_t)B * B * B;
            // voxels of the child inside the volume; samples beyond replicate the last one (:1460-1514)
            const int ux = min(min(cx * B, urb[0]) + B, urb[0]) - min(cx * B, urb[0]);
            const int uy = min(min(cy * B, urb[1]) + B, urb[1]) - min(cy * B, urb[1]);
            const int uz = min(min(cz * B, urb[2]) + B, urb[2]) - min(cz * B, urb[2]);
            const int x = (px - cx * H) * 2, y = (py - cy * H) * 2, z = (pz - cz * H) * 2;
            const int xl = min(x, ux - 1), xh = min(x + 1, ux - 1), yl = min(y, uy - 1), yh = min(y + 1, uy - 1), zl = min(z, uz - 1), zh = min(z + 1, uz - 1);
            const unsigned v8[8] = {cb[(zl * B + yl) * B + xl], cb[(zl * B + yl) * B + xh], cb[(zl * B + yh) * B + xl], cb[(zl * B + yh) * B + xh],
                                    cb[(zh * B + yl) * B + xl], cb[(zh * B + yl) * B + xh], cb[(zh * B + yh) * B + xl], cb[(zh * B + yh) * B + xh]};
            val = oct_halfsample(halfsample_fn, v8);
        } else val = o.avg[ci];                                                 // brickless child: its average (:1523-1530)
        brick[i] = (uint16_t)val;
    }
}

// deleteSubTree (:1428-1436): what hangs below a node without children does not exist
__global__ void octree_prune_kernel(OctreeDev o, int k) {
    const int G = 1 << k;
    const size_t n = (size_t)G * G * G, i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % G), y = (int)((i / G) % G), z = (int)(i / ((size_t)G * G));
    const int Gp = G / 2;
    const size_t pi = oct_level_offset(k - 1) + ((size_t)(z / 2) * Gp + (y / 2)) * Gp + (x / 2);
    const uint8_t pf = o.flags[pi];
    if (!((pf & OCT_EXISTS) && (pf & OCT_HAS_CHILDREN))) o.flags[oct_level_offset(k) + i] &= (uint8_t)~OCT_EXISTS;
}

// ---- the DVR traversal (singleoctreeraycaster.cl:190-300 with octree_dvr.cl:52-80; one channel) -----------------------------------
struct OctRender {
    float step; int lod; int filter_linear; int comp;      // comp: B200RC_COMP_DVR / _MIP / _MOP (octree_dvr.cl / octree_mip.cl / octree_mop.cl)
    const uchar4* tf; int tf_width; float dom_lo, dom_hi, rwm_scale, rwm_offset;
    const float4* entry; const float4* exitp; float4* out; float* out_firsthit;
    int W, H; float scale[3];       // volume texture space -> octree space: dims / octreeDim (copyimage_octreecoarseness / rescale of entry/exit)
};

__device__ __forceinline__ float4 oct_tf_linear(const OctRender& r, float s) {      // read_imagef, CLK_FILTER_LINEAR, CLAMP_TO_EDGE, normalised coordinate
    const float u = xsub(xmul(s, (float)r.tf_width), 0.5f);
    const float fu = floorf(u);
    const float a = xsub(u, fu);
    const int x0 = min(max((int)fu, 0), r.tf_width - 1), x1 = min(max((int)fu + 1, 0), r.tf_width - 1);
    const uchar4 q0 = r.tf[x0], q1 = r.tf[x1];       // CL_UNORM_INT8: c / 255
    const float4 c0 = make_float4(xdiv((float)q0.x, 255.f), xdiv((float)q0.y, 255.f), xdiv((float)q0.z, 255.f), xdiv((float)q0.w, 255.f));
    const float4 c1 = make_float4(xdiv((float)q1.x, 255.f), xdiv((float)q1.y, 255.f), xdiv((float)q1.z, 255.f), xdiv((float)q1.w, 255.f));
    return make_float4(xlerp(c0.x, c1.x, a), xlerp(c0.y, c1.y, a), xlerp(c0.z, c1.z, a), xlerp(c0.w, c1.w, a));
}

__global__ void __launch_bounds__(128) octree_march_kernel(OctreeDev o, OctRender r) {
    const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (px >= r.W || py >= r.H) return;
    const size_t pix = (size_t)py * r.W + px;
    const float4 e4 = r.entry[pix], x4 = r.exitp[pix];
    const float3 entry = f3(xmul(e4.x, r.scale[0]), xmul(e4.y, r.scale[1]), xmul(e4.z, r.scale[2]));
    const float3 exitp = f3(xmul(x4.x, r.scale[0]), xmul(x4.y, r.scale[1]), xmul(x4.z, r.scale[2]));
    float4 color = make_float4(0.f, 0.f, 0.f, 0.f);
    float firsthit = 1.0f;
    float3 dir = f3(xsub(exitp.x, entry.x), xsub(exitp.y, entry.y), xsub(exitp.z, entry.z));
    const float tEnd = __fsqrt_rn(xadd(xadd(xmul(dir.x, dir.x), xmul(dir.y, dir.y)), xmul(dir.z, dir.z)));
    const bool moving = tEnd > 0.f;
    if (moving) dir = f3(xdiv(dir.x, tEnd), xdiv(dir.y, tEnd), xdiv(dir.z, tEnd));
    if (e4.w > 0.f || x4.w > 0.f) {         // covered pixel (entry == exit still marches its single sample, as the kernel does)
        const int B = o.brick_dim;
        float param = 0.f, tEndNode = -1.f;
        float3 nllf = f3(0.f, 0.f, 0.f), ndim_rcp = f3(1.f, 1.f, 1.f);
        size_t node = 0; bool has_brick = false;
        // MIP / MOP ray state (singleoctreeraycaster.cl:360-368): pending intensity 0, pending first hit 1; MOP starts from the transfer
        // function's colour at intensity 0 (:232-237).  In full-frame mode the "current" copies only mirror the pending ones.
        float pend_i = 0.f, pend_fh = 1.0f;
        float4 pend_c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r.comp == B200RC_COMP_MOP) pend_c = oct_tf_linear(r, xdiv(xsub(xadd(xmul(0.f, r.rwm_scale), r.rwm_offset), r.dom_lo), xsub(r.dom_hi, r.dom_lo)));
        while (param <= tEnd) {
            const float3 s = f3(xadd(entry.x, xmul(param, dir.x)), xadd(entry.y, xmul(param, dir.y)), xadd(entry.z, xmul(param, dir.z)));
            if (param > xadd(tEndNode, 1.e-6f)) {
                // getNodeAtSamplePos (octree_config.cl:461-506): descend from the root towards the requested level while children exist
                int k = 0; int ix = 0, iy = 0, iz = 0;
                float3 llf = f3(0.f, 0.f, 0.f); float half = 1.0f;
                node = 0;
                while (k < r.lod && (o.flags[node] & OCT_HAS_CHILDREN)) {
                    half = half * 0.5f;
                    const int cx = s.x >= llf.x + half ? 1 : 0, cy = s.y >= llf.y + half ? 1 : 0, cz = s.z >= llf.z + half ? 1 : 0;
                    ix = ix * 2 + cx; iy = iy * 2 + cy; iz = iz * 2 + cz;
                    llf = f3(llf.x + half * (float)cx, llf.y + half * (float)cy, llf.z + half * (float)cz);
                    ++k;
                    const int G = 1 << k;
                    node = oct_level_offset(k) + ((size_t)iz * G + iy) * G + ix;
                }
                nllf = llf;
                const float3 urb = f3(llf.x + half, llf.y + half, llf.z + half);
                has_brick = (o.flags[node] & OCT_HAS_BRICK) != 0;
                ndim_rcp = f3(xdiv(1.f, xsub(urb.x, llf.x)), xdiv(1.f, xsub(urb.y, llf.y)), xdiv(1.f, xsub(urb.z, llf.z)));
                // fetchNextRayNode (:626-636) over computeNodeExitPoint (:437-456): ray parameter of the last sample inside the node
                if (moving) {
                    const float ex = dir.x >= 0.f ? urb.x : llf.x, ey = dir.y >= 0.f ? urb.y : llf.y, ez = dir.z >= 0.f ? urb.z : llf.z;
                    const float tx = dir.x != 0.f ? xdiv(xsub(ex, s.x), dir.x) : 1e6f, ty = dir.y != 0.f ? xdiv(xsub(ey, s.y), dir.y) : 1e6f, tz = dir.z != 0.f ? xdiv(xsub(ez, s.z), dir.z) : 1e6f;
                    const float tmin = xsub(fminf(tx, fminf(ty, tz)), 1e-6f);
                    const float3 nexit = f3(xadd(s.x, xmul(tmin, dir.x)), xadd(s.y, xmul(tmin, dir.y)), xadd(s.z, xmul(tmin, dir.z)));
                    float toff = fminf(xdiv(xsub(nexit.x, s.x), dir.x), fminf(xdiv(xsub(nexit.y, s.y), dir.y), xdiv(xsub(nexit.z, s.z), dir.z)));    // NaN (0/0) drops out of fmin like OpenCL's
                    toff = xmul(floorf(xdiv(toff, r.step)), r.step);
                    tEndNode = xadd(param, toff);
                } else tEndNode = param;
            }
            float intensity;
            if (s.x < 0.f || s.x > 1.f || s.y < 0.f || s.y > 1.f || s.z < 0.f || s.z > 1.f) intensity = 0.f;
            else if (has_brick) {
                const uint16_t* brick = o.bricks + node * (size_t)B * B * B;
                const float3 q = f3(xmul(xsub(s.x, nllf.x), ndim_rcp.x), xmul(xsub(s.y, nllf.y), ndim_rcp.y), xmul(xsub(s.z, nllf.z), ndim_rcp.z));
                if (r.filter_linear) {      // filterBrickLinear (octree_config.cl:323-368)
                    const float cxf = xmul(q.x, (float)(B - 1)), cyf = xmul(q.y, (float)(B - 1)), czf = xmul(q.z, (float)(B - 1));
                    const float fx = floorf(cxf), fy = floorf(cyf), fz = floorf(czf);
                    const int lx = min(max((int)fx, 0), B - 1), ly = min(max((int)fy, 0), B - 1), lz = min(max((int)fz, 0), B - 1);
                    const int hx = min((int)fx + 1, B - 1), hy = min((int)fy + 1, B - 1), hz = min((int)fz + 1, B - 1);
                    const float pxw = xsub(cxf, (float)lx), pyw = xsub(cyf, (float)ly), pzw = xsub(czf, (float)lz);
                    auto V = [&](int x, int y, int z) -> float { return (float)brick[((size_t)z * B + y) * B + x]; };
                    const float ox = xsub(1.f, pxw), oy = xsub(1.f, pyw), oz = xsub(1.f, pzw);
                    float acc = xmul(xmul(xmul(V(lx, ly, lz), ox), oy), oz);
                    acc = xadd(acc, xmul(xmul(xmul(V(lx, ly, hz), ox), oy), pzw));
                    acc = xadd(acc, xmul(xmul(xmul(V(lx, hy, lz), ox), pyw), oz));
                    acc = xadd(acc, xmul(xmul(xmul(V(lx, hy, hz), ox), pyw), pzw));
                    acc = xadd(acc, xmul(xmul(xmul(V(hx, ly, lz), pxw), oy), oz));
                    acc = xadd(acc, xmul(xmul(xmul(V(hx, ly, hz), pxw), oy), pzw));
                    acc = xadd(acc, xmul(xmul(xmul(V(hx, hy, lz), pxw), pyw), oz));
                    acc = xadd(acc, xmul(xmul(xmul(V(hx, hy, hz), pxw), pyw), pzw));
                    intensity = xmul(acc, 1.f / 65536.f);
                } else {                    // filterBrickNearest (:313-321)
                    const int cx = min(max((int)rintf(xsub(xmul(q.x, (float)B), 0.5f)), 0), B - 1), cy = min(max((int)rintf(xsub(xmul(q.y, (float)B), 0.5f)), 0), B - 1),
                              cz = min(max((int)rintf(xsub(xmul(q.z, (float)B), 0.5f)), 0), B - 1);
                    intensity = xmul((float)brick[((size_t)cz * B + cy) * B + cx], 1.f / 65536.f);
                }
            } else intensity = fminf(fmaxf(xdiv(xmul((float)o.avg[node], 1.f), 65535.f), 0.f), 1.f);       // getNodeAvgValue0 (:150-155), 16 bits for one channel
            // applyChannelTransFunc (:402-413) + applyTFandCombineColors (octree_dvr.cl:52-80)
            if (r.comp == B200RC_COMP_MIP) {            // octree_mip.cl:58-69: the transfer function is applied once, to the maximum
                if (pend_i < intensity) { pend_i = intensity; pend_fh = fminf(pend_fh, xdiv(param, tEnd)); }
                param = xadd(param, r.step);
                continue;
            }
            const float rw = xadd(xmul(intensity, r.rwm_scale), r.rwm_offset);
            const float4 c = oct_tf_linear(r, xdiv(xsub(rw, r.dom_lo), xsub(r.dom_hi, r.dom_lo)));
            if (r.comp == B200RC_COMP_MOP) {            // octree_mop.cl:58-73: the sample with the largest opacity
                if (pend_c.w < c.w) { pend_c = c; pend_i = intensity; pend_fh = fminf(pend_fh, xdiv(param, tEnd)); }
                param = xadd(param, r.step);
                continue;
            }
            if (c.w > 0.f) {
                const float3 rgb = f3(fminf(c.x, 1.f), fminf(c.y, 1.f), fminf(c.z, 1.f));
                const float a = 1.f - powf(1.f - c.w, r.step * 200.f);
                const float k = (1.f - color.w) * a;
                color.x += k * rgb.x; color.y += k * rgb.y; color.z += k * rgb.z; color.w += k;
                firsthit = fminf(xdiv(param, tEnd), firsthit);
            }
            if (color.w >= 0.95f) { color.w = 1.f; param = xadd(tEnd, r.step); }
            param = xadd(param, r.step);
        }
        if (r.comp != B200RC_COMP_DVR) {                // postRaycastingLoop (octree_mip.cl:71-90, octree_mop.cl:75-96)
            if (r.comp == B200RC_COMP_MIP) pend_c = oct_tf_linear(r, xdiv(xsub(xadd(xmul(pend_i, r.rwm_scale), r.rwm_offset), r.dom_lo), xsub(r.dom_hi, r.dom_lo)));
            if (pend_c.w > 0.f) color = make_float4(fminf(pend_c.x, 1.f), fminf(pend_c.y, 1.f), fminf(pend_c.z, 1.f), fminf(pend_c.w, 1.f));
            firsthit = pend_fh;
        }
    }
    r.out[pix] = color;
    if (r.out_firsthit) r.out_firsthit[pix] = firsthit;
}

// ---- host -------------------------------------------------------------------------------------------------------------------------
struct OctreeSlot {
    OctreeDev dev{}; bool valid = false; int thr = 0; bool optimize = false; double build_ms = 0.0;
    std::vector<uint8_t> flags_host;
};
static OctreeSlot* rc_octree_slot(b200rc_ctx* ctx, int slot) {
    if (slot < 0 || slot >= B200RC_MAX_VOLUMES) return nullptr;
    if (!ctx->octree[slot]) ctx->octree[slot] = new (std::nothrow) OctreeSlot();
    return ctx->octree[slot];
}

static void octree_free(OctreeDev& d) {
    cudaFree(d.avg); cudaFree(d.mn); cudaFree(d.mx); cudaFree(d.flags); cudaFree(d.bricks);
    d.avg = d.mn = d.mx = nullptr; d.flags = nullptr; d.bricks = nullptr;
}

int rc_octree_release(b200rc_ctx* ctx, int slot) {
    if (slot < 0 || slot >= B200RC_MAX_VOLUMES || !ctx->octree[slot]) return B200RC_OK;
    OctreeSlot* s = ctx->octree[slot];
    if (s->valid) { octree_free(s->dev); s->valid = false; }
    return B200RC_OK;
}
void rc_octree_destroy_all(b200rc_ctx* ctx) {
    for (int i = 0; i < B200RC_MAX_VOLUMES; ++i) { rc_octree_release(ctx, i); delete ctx->octree[i]; ctx->octree[i] = nullptr; }
}

int rc_octree_build(b200rc_ctx* ctx, int slot, int brick_dim, float homogeneity_threshold, int halfsample_fn, const VolDev& v, const VolumeSlot& vs) {
    if (brick_dim < 2 || brick_dim > 128 || (brick_dim & (brick_dim - 1))) { ctx->set_error("octree: brick size must be a power of two in [2, 128]"); return B200RC_ERR_INVALID; }
    if (vs.desc.channels != 1 || (vs.desc.dtype != B200RC_U8 && vs.desc.dtype != B200RC_U16) || vs.sharded) { ctx->set_error("octree: needs an unsharded one-channel uint8 / uint16 volume"); return B200RC_ERR_INVALID; }
    if (halfsample_fn < 0 || halfsample_fn > 2) { ctx->set_error("octree: half-sample function must be mean (0), max (1) or min (2)"); return B200RC_ERR_INVALID; }
    int maxdim = vs.desc.dims[0] > vs.desc.dims[1] ? vs.desc.dims[0] : vs.desc.dims[1];
    if (vs.desc.dims[2] > maxdim) maxdim = vs.desc.dims[2];
    int odim = 1; while (odim < maxdim) odim <<= 1;                    // octree dimension: next power of two (cubic)
    if (brick_dim > odim) { ctx->set_error("octree: brick size larger than the volume"); return B200RC_ERR_INVALID; }
    int levels = 1; while ((brick_dim << (levels - 1)) < odim) ++levels;
    OctreeSlot* s = rc_octree_slot(ctx, slot);
    if (!s) { ctx->set_error("octree: bad slot"); return B200RC_ERR_INVALID; }
    rc_octree_release(ctx, slot);
    OctreeDev& d = s->dev;
    d.levels = levels; d.brick_dim = brick_dim; d.octree_dim = odim;
    for (int a = 0; a < 3; ++a) d.dims[a] = vs.desc.dims[a];
    const size_t nodes = oct_level_offset(levels);
    const size_t bvox = (size_t)brick_dim * brick_dim * brick_dim;
    if (cudaMalloc(&d.avg, nodes * 2) != cudaSuccess || cudaMalloc(&d.mn, nodes * 2) != cudaSuccess || cudaMalloc(&d.mx, nodes * 2) != cudaSuccess ||
        cudaMalloc(&d.flags, nodes) != cudaSuccess || cudaMalloc(&d.bricks, nodes * bvox * 2) != cudaSuccess) {
        cudaGetLastError(); octree_free(d); ctx->set_error("octree: out of device memory"); return B200RC_ERR_OOM;
    }
    // VolumeOctree ctor (:242-276): optimisation iff threshold >= 0; threshold in uint16 units
    s->optimize = homogeneity_threshold >= 0.f;
    int t = (int)floorf(homogeneity_threshold * 65535.f + 0.5f);
    s->thr = t < 0 ? 0 : (t > 65535 ? 65535 : t);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, ctx->stream);
    const int G = 1 << (levels - 1);
    if (vs.desc.dtype == B200RC_U16) octree_leaf_kernel<B200RC_U16><<<(unsigned)((size_t)G * G * G), 256, 0, ctx->stream>>>(v, d, s->thr, s->optimize ? 1 : 0);
    else octree_leaf_kernel<B200RC_U8><<<(unsigned)((size_t)G * G * G), 256, 0, ctx->stream>>>(v, d, s->thr, s->optimize ? 1 : 0);
    ctx->launches++;
    for (int k = levels - 2; k >= 0; --k) {
        const int Gk = 1 << k;
        octree_parent_kernel<<<(unsigned)((size_t)Gk * Gk * Gk), 256, 0, ctx->stream>>>(d, k, s->thr, s->optimize ? 1 : 0, halfsample_fn);
        ctx->launches++;
    }
    for (int k = 1; k < levels; ++k) {
        const size_t n = (size_t)1 << (3 * k);
        octree_prune_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d, k);
        ctx->launches++;
    }
    cudaEventRecord(e1, ctx->stream);
    RC_CHECK(cudaGetLastError());
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); s->build_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    s->flags_host.resize(nodes);
    RC_CHECK(cudaMemcpy(s->flags_host.data(), d.flags, nodes, cudaMemcpyDeviceToHost));
    s->valid = true;
    return B200RC_OK;
}

int rc_octree_info(b200rc_ctx* ctx, int slot, b200rc_octree_info* out) {
    OctreeSlot* s = rc_octree_slot(ctx, slot);
    if (!s || !s->valid) { ctx->set_error("octree: not built"); return B200RC_ERR_NOT_READY; }
    memset(out, 0, sizeof *out);
    out->octree_dim = s->dev.octree_dim; out->brick_dim = s->dev.brick_dim; out->num_levels = s->dev.levels; out->build_ms = s->build_ms;
    // VolumeOctreeNode::getNodeCount / getNumBricks / getDepth (volumeoctreebase.cpp:102-145) over the nodes that exist
    for (int k = 0; k < s->dev.levels; ++k) {
        const size_t off = oct_level_offset(k), n = (size_t)1 << (3 * k);
        for (size_t i = 0; i < n; ++i) {
            const uint8_t f = s->flags_host[off + i];
            if (!(f & OCT_EXISTS)) continue;
            out->num_nodes++;
            if (f & OCT_HAS_BRICK) out->num_bricks++;
            if (k + 1 > out->actual_depth) out->actual_depth = k + 1;
        }
    }
    return B200RC_OK;
}

int rc_octree_download(b200rc_ctx* ctx, int slot, int level, uint16_t* avg, uint16_t* mn, uint16_t* mx, uint8_t* flags, uint16_t* bricks) {
    OctreeSlot* s = rc_octree_slot(ctx, slot);
    if (!s || !s->valid || level < 0 || level >= s->dev.levels) { ctx->set_error("octree: no such level"); return B200RC_ERR_INVALID; }
    const size_t off = oct_level_offset(level), n = (size_t)1 << (3 * level), bvox = (size_t)s->dev.brick_dim * s->dev.brick_dim * s->dev.brick_dim;
    if (avg) RC_CHECK(cudaMemcpyAsync(avg, s->dev.avg + off, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (mn) RC_CHECK(cudaMemcpyAsync(mn, s->dev.mn + off, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (mx) RC_CHECK(cudaMemcpyAsync(mx, s->dev.mx + off, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (flags) RC_CHECK(cudaMemcpyAsync(flags, s->dev.flags + off, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (bricks) RC_CHECK(cudaMemcpyAsync(bricks, s->dev.bricks + off * bvox, n * bvox * 2, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200RC_OK;
}

int rc_octree_render(b200rc_ctx* ctx, int slot, const b200rc_octree_render_params* rp, const float4* entry, const float4* exitp, int W, int H, float4* out, float* out_firsthit) {
    OctreeSlot* s = rc_octree_slot(ctx, slot);
    if (!s || !s->valid) { ctx->set_error("octree: not built"); return B200RC_ERR_NOT_READY; }
    const ClassifierSlot& k = ctx->cls[0];
    if (!k.tf) { ctx->set_error("octree render: no 1D transfer function"); return B200RC_ERR_NOT_READY; }
    if (!(rp->sampling_step > 0.f)) { ctx->set_error("octree render: sampling step must be positive"); return B200RC_ERR_INVALID; }
    OctRender r;
    r.step = rp->sampling_step;
    r.lod = rp->node_level < 0 ? s->dev.levels - 1 : (rp->node_level > s->dev.levels - 1 ? s->dev.levels - 1 : rp->node_level);
    r.filter_linear = rp->filter == B200RC_FILTER_LINEAR ? 1 : 0;
    r.comp = rp->compositing;
    if (r.comp != B200RC_COMP_DVR && r.comp != B200RC_COMP_MIP && r.comp != B200RC_COMP_MOP) { ctx->set_error("octree render: compositing must be dvr, mip or mop"); return B200RC_ERR_INVALID; }
    r.tf = k.tf; r.tf_width = k.tf_width; r.dom_lo = k.dom[0]; r.dom_hi = k.dom[1];
    r.rwm_scale = ctx->vol[slot].desc.rwm_scale; r.rwm_offset = ctx->vol[slot].desc.rwm_offset;
    r.entry = entry; r.exitp = exitp; r.out = out; r.out_firsthit = out_firsthit; r.W = W; r.H = H;
    for (int a = 0; a < 3; ++a) r.scale[a] = (float)s->dev.dims[a] / (float)s->dev.octree_dim;
    dim3 grid((unsigned)((W + 15) / 16), (unsigned)((H + 7) / 8));
    octree_march_kernel<<<grid, 128, 0, ctx->stream>>>(s->dev, r);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}
