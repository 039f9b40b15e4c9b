// composite.cu -- sort-last merge of partial images (premultiplied front-to-back over, depth = min) and the final
// early-ray-termination snap + output conversion.  New work with no counterpart in the reference (SURVEY section 5,
// 8e); the over operator matches what the raycaster itself leaves in its accumulators
// (V/modules/core/glsl/modules/mod_compositing.frag:53-54: premultiplied colours).
#include "ctx.hpp"

__global__ void composite_kernel(int mode, const float4* __restrict__ fc, const float* __restrict__ fd, const float4* __restrict__ bc,
                                 const float* __restrict__ bd, float4* __restrict__ dc, float* __restrict__ dd, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 F = fc[i], B = bc[i];
    const float fdp = fd ? fd[i] : 1.f, bdp = bd ? bd[i] : 1.f;
    float4 D; float dep;
    const bool ert = (mode & B200RC_COMPOSITE_ERT) != 0;
    mode &= 0xff;
    if (mode == B200RC_COMP_DVR) {
        const float k = (ert && F.w >= 0.95f) ? 0.0f : 1.0f - F.w;      // a terminated ray sees nothing behind it
        D = make_float4(F.x + k * B.x, F.y + k * B.y, F.z + k * B.z, F.w + k * B.w);
        dep = fminf(fdp, bdp);
    } else if (mode == B200RC_COMP_MIP || mode == B200RC_COMP_MOP) {
        const bool takeB = bdp > fdp;           // key (max intensity / max alpha) travels in the depth slot
        D = takeB ? B : F; dep = takeB ? bdp : fdp;
    } else {
        const bool takeB = !(F.w > 0.f);        // first hit: the front image wins wherever it has one
        D = takeB ? B : F; dep = takeB ? bdp : fdp;
    }
    dc[i] = D;
    if (dd) dd[i] = dep;
}

__global__ void finalize_kernel(const float4* __restrict__ rgba, size_t n, int out_format, int snap, void* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 r = rgba[i];
    if (snap && r.w >= 0.95f) r.w = 1.0f;       // earlyRayTermination (mod_raysetup.frag:65-74)
    if (out_format == B200RC_OUT_RGBA16) ((ushort4*)out)[i] = make_ushort4(to_unorm16(r.x), to_unorm16(r.y), to_unorm16(r.z), to_unorm16(r.w));
    else ((float4*)out)[i] = r;
}

// rescaleRendering: bilinear (GL_LINEAR, CLAMP_TO_EDGE) copy of colour + depth to another size (copyimage.frag:37-55)
__global__ void rescale_kernel(const void* __restrict__ src, const float* __restrict__ sdepth, int sw, int sh, int format,
                               void* __restrict__ dst, float* __restrict__ ddepth, int dw, int dh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    // fragCoord = gl_FragCoord.xy * screenDimRCP_;  u = s * size - 0.5 (GL 4.x 8.14)
    const float fx = xmul(xadd((float)x, 0.5f), xdiv(1.0f, (float)dw)), fy = xmul(xadd((float)y, 0.5f), xdiv(1.0f, (float)dh));
    const float u = xsub(xmul(fx, (float)sw), 0.5f), v = xsub(xmul(fy, (float)sh), 0.5f);
    const float fu = floorf(u), fv = floorf(v);
    const float a = xsub(u, fu), b = xsub(v, fv);
    const int x0 = min(max((int)fu, 0), sw - 1), x1 = min(max((int)fu + 1, 0), sw - 1);
    const int y0 = min(max((int)fv, 0), sh - 1), y1 = min(max((int)fv + 1, 0), sh - 1);
    auto fetch = [&](int px, int py) -> float4 {
        const size_t i = (size_t)py * sw + px;
        if (format == B200RC_OUT_RGBA16) {
            const ushort4 q = ((const ushort4*)src)[i];
            return make_float4(xdiv((float)q.x, 65535.f), xdiv((float)q.y, 65535.f), xdiv((float)q.z, 65535.f), xdiv((float)q.w, 65535.f));
        }
        return ((const float4*)src)[i];
    };
    const float4 c00 = fetch(x0, y0), c10 = fetch(x1, y0), c01 = fetch(x0, y1), c11 = fetch(x1, y1);
    float4 r;
    r.x = xlerp(xlerp(c00.x, c10.x, a), xlerp(c01.x, c11.x, a), b); r.y = xlerp(xlerp(c00.y, c10.y, a), xlerp(c01.y, c11.y, a), b);
    r.z = xlerp(xlerp(c00.z, c10.z, a), xlerp(c01.z, c11.z, a), b); r.w = xlerp(xlerp(c00.w, c10.w, a), xlerp(c01.w, c11.w, a), b);
    const size_t o = (size_t)y * dw + x;
    if (format == B200RC_OUT_RGBA16) ((ushort4*)dst)[o] = make_ushort4(to_unorm16(r.x), to_unorm16(r.y), to_unorm16(r.z), to_unorm16(r.w));
    else ((float4*)dst)[o] = r;
    if (sdepth && ddepth) {
        const float d00 = sdepth[(size_t)y0 * sw + x0], d10 = sdepth[(size_t)y0 * sw + x1], d01 = sdepth[(size_t)y1 * sw + x0], d11 = sdepth[(size_t)y1 * sw + x1];
        ddepth[o] = depth24_roundtrip(xlerp(xlerp(d00, d10, a), xlerp(d01, d11, a), b));
    }
}

int rc_rescale_launch(b200rc_ctx* ctx, const void* src, const float* sdepth, int sw, int sh, int format, void* dst, float* ddepth, int dw, int dh) {
    dim3 block(32, 8), grid((unsigned)((dw + 31) / 32), (unsigned)((dh + 7) / 8));
    rescale_kernel<<<grid, block, 0, ctx->stream>>>(src, sdepth, sw, sh, format, dst, ddepth, dw, dh);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}

// max-key merge (sort-last MIP / MOP): the partial image with the larger key wins, the front one on ties -- the reference keeps the
// FIRST sample whose intensity (alpha) exceeds the running maximum (mod_compositing.frag:85-106), and every sample of the front
// shard precedes every sample of the back shard on the ray
__global__ void composite_keyed_kernel(const float4* __restrict__ fc, const float* __restrict__ fk, const float* __restrict__ fd,
                                       const float4* __restrict__ bc, const float* __restrict__ bk, const float* __restrict__ bd,
                                       float4* __restrict__ dc, float* __restrict__ dk, float* __restrict__ dd, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float kf = fk[i], kb = bk[i];
    const bool takeB = kb > kf;
    const float4 c = takeB ? bc[i] : fc[i];
    const float d = (fd && bd) ? (takeB ? bd[i] : fd[i]) : 1.0f;
    dc[i] = c; dk[i] = takeB ? kb : kf;
    if (dd) dd[i] = d;
}

int rc_composite_keyed_launch(b200rc_ctx* ctx, const float* fc, const float* fk, const float* fd, const float* bc, const float* bk, const float* bd,
                              float* dc, float* dk, float* dd, size_t n) {
    if (!n) return B200RC_OK;
    composite_keyed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const float4*)fc, fk, fd, (const float4*)bc, bk, bd, (float4*)dc, dk, dd, n);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}

int rc_composite_launch(b200rc_ctx* ctx, int mode, const float* fc, const float* fd, const float* bc, const float* bd, float* dc, float* dd, size_t n) {
    if (!n) return B200RC_OK;
    composite_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(mode, (const float4*)fc, fd, (const float4*)bc, bd, (float4*)dc, dd, n);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}

int rc_finalize_launch(b200rc_ctx* ctx, const float* rgba, const float* depth, size_t n, int out_format, int snap, void* out) {
    (void)depth;
    if (!n) return B200RC_OK;
    finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const float4*)rgba, n, out_format, snap ? 1 : 0, out);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}

// ---- exact early ray termination across shards (sort-last second pass planning) ------------------------------------
// alphas: world x n (alpha of every rank's first-pass partial image).  front_mask: ranks whose brick lies in front of
// this rank's along every ray that crosses both (axis-aligned grid: on the camera's side of each differing split plane).
//   P >= 0.95                         -> the ray terminated in front of this shard: its contribution is erased
//   0 < P, P + (1-P) A_local >= 0.95  -> the global 0.95 crossing lies inside this shard's segment: re-march (flag 2)
//   otherwise                         -> the first-pass result stands (P == 0: local and global thresholds coincide)
// the alpha planes of the ranks: one gathered array (rank major) or one pointer per rank (planes in peer memory, read over NVLink)
struct AlphaPlanes { const float* gathered; const float* plane[32]; };

__global__ void ert_plan_kernel(const __grid_constant__ AlphaPlanes ap, int world, unsigned front_mask, float4* __restrict__ color,
                                float* __restrict__ depth, size_t n, float* __restrict__ prefix, unsigned char* __restrict__ flags,
                                unsigned long long* __restrict__ counter, unsigned int* __restrict__ list, unsigned int* __restrict__ list_n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned flag = 0;
    if (i < n) {
        float T = 1.0f;
        for (int r = 0; r < world; ++r)
            if ((front_mask >> r) & 1u) T *= 1.0f - (ap.gathered ? ap.gathered[(size_t)r * n + i] : ap.plane[r][i]);
        const float P = 1.0f - T;
        const float A = color[i].w;
        if (P >= 0.95f) {
            if (A != 0.f) { color[i] = make_float4(0.f, 0.f, 0.f, 0.f); depth[i] = 1.0f; }
            flag = 1;
        } else if (P > 0.f && A > 0.f && P + (1.0f - P) * A >= 0.95f) flag = 2;
        prefix[i] = P;
        flags[i] = (unsigned char)flag;
    }
    if (list) {         // compact the pixels to re-march (warp-aggregated append: one atomic per warp, row neighbours stay together)
        const unsigned m = __ballot_sync(0xffffffffu, flag == 2);
        if (m) {
            const unsigned lane = threadIdx.x & 31u;
            const int leader = __ffs(m) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(list_n, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (flag == 2) list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned int)i;
        }
    }
    if (counter) {
        const unsigned c = __reduce_add_sync(0xffffffffu, flag == 2 ? 1u : 0u);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(counter, (unsigned long long)c);
    }
}

int rc_ert_plan_launch(b200rc_ctx* ctx, const float* alphas, const float* const* planes, int world, unsigned front_mask, float* color, float* depth, size_t n,
                       float* prefix, unsigned char* flags, unsigned long long* counter) {
    if (!n) return B200RC_OK;
    AlphaPlanes ap;
    ap.gathered = alphas;
    for (int r = 0; r < 32; ++r) ap.plane[r] = (!alphas && planes && r < world) ? planes[r] : nullptr;
    ctx->plan_flags = nullptr; ctx->plan_n = 0;
    unsigned int* list = nullptr;
    if (n < 0xffffffffull) {
        if (ctx->remarch_capacity < n) {
            cudaFree(ctx->d_remarch_list); ctx->d_remarch_list = nullptr; ctx->remarch_capacity = 0;
            RC_CHECK(cudaMalloc(&ctx->d_remarch_list, sizeof(unsigned int) * n));
            ctx->remarch_capacity = n;
        }
        list = ctx->d_remarch_list;
        RC_CHECK(cudaMemsetAsync(ctx->d_tile_counter + 2, 0, sizeof(unsigned int), ctx->stream));
    }
    ert_plan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ap, world, front_mask, (float4*)color, depth, n, prefix, flags, counter,
                                                                         list, ctx->d_tile_counter + 2);
    if (list) { ctx->plan_flags = flags; ctx->plan_n = n; }
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}
