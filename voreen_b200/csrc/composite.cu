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

int rc_composite_launch(b200rc_ctx* ctx, int mode, const float* fc, const float* fd, const float* bc, const float* bd, float* dc, float* dd, size_t n) {
    if (!n) return B200RC_OK;
    composite_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(mode, (const float4*)fc, fd, (const float4*)bc, bd, (float4*)dc, dd, n);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}

int rc_finalize_launch(b200rc_ctx* ctx, const float* rgba, const float* depth, size_t n, int out_format, void* out) {
    (void)depth;
    if (!n) return B200RC_OK;
    finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const float4*)rgba, n, out_format, 1, out);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}
