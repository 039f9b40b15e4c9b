// post.cu -- the image-space steps that follow the raycaster in a standard workspace, so that the frame can stay
// in HBM up to the canvas: voreen::Compositor (V/modules/base/glsl/image/compositor.frag:41-140, all 14 modes of
// V/modules/base/processors/image/compositor.cpp:47-60) and voreen::Background (V/modules/base/glsl/image/background.frag:35-57
// with the "monochrome" and "gradient" background generators of V/modules/base/processors/image/background.cpp:262-300).
// Both shaders fetch their inputs with textureLookup2D* at the fragment's own pixel (equal-sized render targets; mod_sampler2d.frag),
// i.e. they are per-pixel maps.  Inputs and outputs are what the RenderPorts hold: RGBA16 unorm (or RGBA32F) colour + float depth.
// Every arithmetic step is an explicitly rounded operation in the shader's evaluation order, so the oracle's restatement and this
// kernel agree bit for bit.
#include "ctx.hpp"

__device__ __forceinline__ float4 post_fetch(const void* img, int format, size_t i) {
    if (format == B200RC_OUT_RGBA16) {
        const ushort4 q = ((const ushort4*)img)[i];
        return make_float4(xdiv((float)q.x, 65535.f), xdiv((float)q.y, 65535.f), xdiv((float)q.z, 65535.f), xdiv((float)q.w, 65535.f));
    }
    return ((const float4*)img)[i];
}
__device__ __forceinline__ void post_store(void* img, int format, size_t i, float4 c) {
    if (format == B200RC_OUT_RGBA16) ((ushort4*)img)[i] = make_ushort4(to_unorm16(c.x), to_unorm16(c.y), to_unorm16(c.z), to_unorm16(c.w));
    else ((float4*)img)[i] = c;
}
__device__ __forceinline__ float post_one_minus(float a) { return xsub(1.0f, a); }

__global__ void compositor_kernel(int mode, float wf, float w0, float w1, int add_depth, const void* __restrict__ c0p, const float* __restrict__ d0p,
                                  const void* __restrict__ c1p, const float* __restrict__ d1p, int format, void* __restrict__ dc, float* __restrict__ dd, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = post_fetch(c0p, format, i), b = post_fetch(c1p, format, i);
    const float da = d0p ? d0p[i] : 1.0f, db = d1p ? d1p[i] : 1.0f;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    float fd = 1.0f;
    switch (mode) {
    case B200RC_COMPOSITOR_DEPTH_TEST:
        if (da < db) { f = a; fd = da; } else { f = b; fd = db; }
        break;
    case B200RC_COMPOSITOR_ALPHA_COMPOSITING:
        if (db <= da) {
            const float k = xmul(a.w, post_one_minus(b.w));      // color0.rgb * color0.a * (1 - color1.a): left to right
            f.x = xadd(b.x, xmul(xmul(a.x, a.w), post_one_minus(b.w))); f.y = xadd(b.y, xmul(xmul(a.y, a.w), post_one_minus(b.w)));
            f.z = xadd(b.z, xmul(xmul(a.z, a.w), post_one_minus(b.w))); f.w = xadd(b.w, k);
        } else {
            const float k = xmul(b.w, post_one_minus(a.w));
            f.x = xadd(a.x, xmul(xmul(b.x, b.w), post_one_minus(a.w))); f.y = xadd(a.y, xmul(xmul(b.y, b.w), post_one_minus(a.w)));
            f.z = xadd(a.z, xmul(xmul(b.z, b.w), post_one_minus(a.w))); f.w = xadd(a.w, k);
        }
        fd = fminf(da, db);
        break;
    case B200RC_COMPOSITOR_ALPHA_BLENDING:
        f.x = xadd(xmul(a.w, a.x), xmul(post_one_minus(a.w), b.x)); f.y = xadd(xmul(a.w, a.y), xmul(post_one_minus(a.w), b.y));
        f.z = xadd(xmul(a.w, a.z), xmul(post_one_minus(a.w), b.z)); f.w = xadd(a.w, xmul(b.w, post_one_minus(a.w)));
        fd = a.w > 0.f ? da : db;
        break;
    case B200RC_COMPOSITOR_ALPHA_BLENDING_B_OVER_A:
        f.x = xadd(xmul(b.w, b.x), xmul(post_one_minus(b.w), a.x)); f.y = xadd(xmul(b.w, b.y), xmul(post_one_minus(b.w), a.y));
        f.z = xadd(xmul(b.w, b.z), xmul(post_one_minus(b.w), a.z)); f.w = xadd(b.w, xmul(a.w, post_one_minus(b.w)));
        fd = b.w > 0.f ? db : da;
        break;
    case B200RC_COMPOSITOR_WEIGHTED_AVERAGE:
        fd = (wf == 0.0f) ? db : (wf == 1.0f ? da : fminf(da, db));
        f.x = xadd(xmul(wf, a.x), xmul(post_one_minus(wf), b.x)); f.y = xadd(xmul(wf, a.y), xmul(post_one_minus(wf), b.y));
        f.z = xadd(xmul(wf, a.z), xmul(post_one_minus(wf), b.z)); f.w = xadd(xmul(wf, a.w), xmul(post_one_minus(wf), b.w));
        break;
    case B200RC_COMPOSITOR_TAKE_FIRST: case B200RC_COMPOSITOR_TAKE_SECOND_IF_READY: f = a; fd = da; break;
    case B200RC_COMPOSITOR_TAKE_SECOND: f = b; fd = db; break;
    case B200RC_COMPOSITOR_FIRST_HAS_PRIORITY: if (a.w > 0.f) { f = a; fd = da; } else { f = b; fd = db; } break;
    case B200RC_COMPOSITOR_SECOND_HAS_PRIORITY: if (b.w > 0.f) { f = b; fd = db; } else { f = a; fd = da; } break;
    case B200RC_COMPOSITOR_MAXIMUM_ALPHA:
        fd = fminf(da, db);
        f.x = xadd(xmul(a.x, a.w), xmul(b.x, b.w)); f.y = xadd(xmul(a.y, a.w), xmul(b.y, b.w)); f.z = xadd(xmul(a.z, a.w), xmul(b.z, b.w));
        f.w = fmaxf(a.w, b.w);
        break;
    case B200RC_COMPOSITOR_DIFFERENCE:
        fd = fabsf(xsub(da, db));
        f.x = xsub(1.0f, fabsf(xsub(a.x, b.x))); f.y = xsub(1.0f, fabsf(xsub(a.y, b.y))); f.z = xsub(1.0f, fabsf(xsub(a.z, b.z)));
        f.w = fmaxf(a.w, b.w);
        break;
    case B200RC_COMPOSITOR_DIFFERENCE_ALPHA_MUL:
        fd = fminf(da, db);
        f.x = fabsf(xsub(xmul(a.x, a.w), xmul(b.x, b.w))); f.y = fabsf(xsub(xmul(a.y, a.w), xmul(b.y, b.w))); f.z = fabsf(xsub(xmul(a.z, a.w), xmul(b.z, b.w)));
        f.w = 1.0f;
        break;
    case B200RC_COMPOSITOR_ADD:
        f.x = xadd(xmul(w0, a.x), xmul(w1, b.x)); f.y = xadd(xmul(w0, a.y), xmul(w1, b.y)); f.z = xadd(xmul(w0, a.z), xmul(w1, b.z)); f.w = xadd(xmul(w0, a.w), xmul(w1, b.w));
        fd = add_depth ? xadd(xmul(w0, da), xmul(w1, db)) : da;
        break;
    default: break;
    }
    post_store(dc, format, i, f);
    if (dd) dd[i] = fd;
}

// Background: the generated background (render target image.tmp: RGBA16, depth cleared to 1) under the input image.
__global__ void background_kernel(int bg_mode, float4 col1, float4 col2, float sin_a, float cos_a, float scale_rcp, int blend_mode,
                                  const void* __restrict__ src, const float* __restrict__ sdepth, int has_input, int format, void* __restrict__ dc, float* __restrict__ dd, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * W + x;
    float4 bg = col1;
    if (bg_mode == B200RC_BACKGROUND_GRADIENT) {
        // quad (-1,-1)..(1,1), first colour on the local edge y = -1, second on y = +1, under rotate(angle, z) * scale(s): the colour at
        // a pixel centre is linear in the quad-local y = (-sin a * X + cos a * Y) / s (background.cpp:271-292)
        const float X = xsub(xmul(xadd((float)x, 0.5f), xdiv(2.0f, (float)W)), 1.0f), Y = xsub(xmul(xadd((float)y, 0.5f), xdiv(2.0f, (float)H)), 1.0f);
        const float ly = xmul(xadd(xmul(-sin_a, X), xmul(cos_a, Y)), scale_rcp);
        const float w = fminf(fmaxf(xadd(xmul(0.5f, ly), 0.5f), 0.0f), 1.0f);
        bg = make_float4(xlerp(col1.x, col2.x, w), xlerp(col1.y, col2.y, w), xlerp(col1.z, col2.z, w), xlerp(col1.w, col2.w, w));
    }
    if (format == B200RC_OUT_RGBA16) bg = make_float4(unorm16_roundtrip(bg.x), unorm16_roundtrip(bg.y), unorm16_roundtrip(bg.z), unorm16_roundtrip(bg.w));
    float4 f = bg; float fd = 1.0f;
    if (has_input) {
        const float4 c = post_fetch(src, format, i);
        const float d = sdepth ? sdepth[i] : 1.0f;
        const float om = post_one_minus(c.w);
        if (blend_mode == B200RC_BACKGROUND_BLEND_ALPHA) {     // color1.rgb * color1.a + color0.rgb * color0.a * (1 - color1.a)
            f.x = xadd(xmul(c.x, c.w), xmul(xmul(bg.x, bg.w), om)); f.y = xadd(xmul(c.y, c.w), xmul(xmul(bg.y, bg.w), om));
            f.z = xadd(xmul(c.z, c.w), xmul(xmul(bg.z, bg.w), om)); f.w = xadd(c.w, xmul(bg.w, om));
        } else {                                                // color1 + color0 * (1 - color1.a), clamped to 1
            f.x = fminf(xadd(c.x, xmul(bg.x, om)), 1.0f); f.y = fminf(xadd(c.y, xmul(bg.y, om)), 1.0f);
            f.z = fminf(xadd(c.z, xmul(bg.z, om)), 1.0f); f.w = fminf(xadd(c.w, xmul(bg.w, om)), 1.0f);
        }
        fd = fminf(1.0f, d);
    }
    post_store(dc, format, i, f);
    if (dd) dd[i] = fd;
}

int rc_compositor_launch(b200rc_ctx* ctx, int mode, float wf, float w0, float w1, int add_depth, const void* c0, const float* d0, const void* c1, const float* d1,
                         int format, void* dc, float* dd, size_t n) {
    if (!n) return B200RC_OK;
    compositor_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(mode, wf, w0, w1, add_depth, c0, d0, c1, d1, format, dc, dd, n);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}

int rc_background_launch(b200rc_ctx* ctx, const b200rc_background_params* bp, const void* src, const float* sdepth, int format, void* dc, float* dd, int W, int H) {
    // MatStack.rotate(angle) / scale(1 + (45 - |angle % 90 - 45|) / 45 * 0.82842712...) (background.cpp:275-279)
    const int ang = bp->angle_deg;
    const float rad = (float)ang * 0.017453292519943295f;       // tgt::deg2rad
    const float s = 1.0f + (float)(45 - abs(ang % 90 - 45)) / 45.0f * 0.8284271247461900976033774484194f;
    dim3 block(32, 8), grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8));
    background_kernel<<<grid, block, 0, ctx->stream>>>(bp->mode, make_float4(bp->color1[0], bp->color1[1], bp->color1[2], bp->color1[3]),
                                                       make_float4(bp->color2[0], bp->color2[1], bp->color2[2], bp->color2[3]), sinf(rad), cosf(rad), 1.0f / s,
                                                       bp->blend_mode, src, sdepth, src ? 1 : 0, format, dc, dd, W, H);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}
