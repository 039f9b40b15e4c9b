// eep.cu -- entry/exit-point generation on the device.
//
// Replaces MeshEntryExitPoints::process (V/modules/base/processors/entryexitpoints/meshentryexitpoints.cpp:188-273)
// and its shaders eep_simple.vert/.frag, eep_insidevolume.frag:44-83, eep_jitter.frag:38-78 for proxy geometries made
// of axis-aligned cubes in texture space (OptimizedProxyGeometry modes boundingbox / minboundingbox / visiblebricks,
// optimizedproxygeometry.cpp:420-491,1014-1116).  Instead of rasterising the cube faces, every pixel intersects the
// ray through its centre with the cubes: the interpolated vertex colour GL would produce IS that intersection point.
// Nearest front face / nearest back face / farthest back face, near- and far-plane clipping, the camera-inside-volume
// fill and the GL_RGBA16 / GL_RGBA16F / GL_DEPTH_COMPONENT24 render-target formats are reproduced.
//
// Every operation here is explicitly rounded (no contraction): the ray count (pixels with entry != exit after
// quantisation) must be bit-exact against the oracle (BASELINE.json north star).
#include "ctx.hpp"
#include <cuda_fp16.h>
#include <cstring>
#include <cfloat>
#include <vector>

struct EepArgs {
    float viewInv[16], projInv[16], modelInv[16], model[16], mvp[16];
    float bb[6], wbb[6];
    int W, H, n_boxes;
    int inside, world;
    float rcpW, rcpH;
};

__device__ __forceinline__ float half_roundtrip(float v) { return __half2float(__float2half_rn(v)); }

__global__ void eep_kernel(const __grid_constant__ EepArgs a, const float* __restrict__ boxes,
                           float4* __restrict__ entry, float* __restrict__ entry_depth, float4* __restrict__ exitp, float* __restrict__ exit_depth) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= a.W || py >= a.H) return;
    const size_t pix = (size_t)py * a.W + px;
    const float sx = xmul(xadd((float)px, 0.5f), a.rcpW), sy = xmul(xadd((float)py, 0.5f), a.rcpH);
    const float nx = xsub(xmul(2.0f, sx), 1.0f), ny = xsub(xmul(2.0f, sy), 1.0f);
    // unproject the near (z=-1) and far (z=+1) NDC points as eep_insidevolume.frag:57-66 does
    const float4 tn = xmat_mul(a.viewInv, xmat_mul(a.projInv, make_float4(nx, ny, -1.0f, 1.0f)));
    const float4 tf = xmat_mul(a.viewInv, xmat_mul(a.projInv, make_float4(nx, ny, 1.0f, 1.0f)));
    const float4 rn = make_float4(xdiv(tn.x, tn.w), xdiv(tn.y, tn.w), xdiv(tn.z, tn.w), xdiv(tn.w, tn.w));
    const float4 rf = make_float4(xdiv(tf.x, tf.w), xdiv(tf.y, tf.w), xdiv(tf.z, tf.w), xdiv(tf.w, tf.w));
    const float4 on = xmat_mul(a.modelInv, rn), of = xmat_mul(a.modelInv, rf);
    const float o[3] = {on.x, on.y, on.z};
    const float d[3] = {xsub(of.x, on.x), xsub(of.y, on.y), xsub(of.z, on.z)};
    float tFront = INFINITY, tBack1 = INFINITY, tBackFar = -INFINITY;
    int aFront = -1, aBackFar = -1; float pFront = 0.f, pBackFar = 0.f;
    for (int b = 0; b < a.n_boxes; ++b) {
        const float* box = boxes + 6 * b;
        float tmin = -INFINITY, tmax = INFINITY; int hit = 1, ain = -1, aout = -1; float pin = 0.f, pout = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float lo = box[k], hi = box[3 + k];
            if (d[k] == 0.f) { if (o[k] < lo || o[k] > hi) hit = 0; continue; }
            const float inv = xdiv(1.0f, d[k]);
            const float t1 = xmul(xsub(lo, o[k]), inv), t2 = xmul(xsub(hi, o[k]), inv);
            float tn_ = t1, tf_ = t2, pn = lo, pf = hi;
            if (t1 > t2) { tn_ = t2; tf_ = t1; pn = hi; pf = lo; }
            if (tn_ > tmin) { tmin = tn_; ain = k; pin = pn; }
            if (tf_ < tmax) { tmax = tf_; aout = k; pout = pf; }
        }
        if (!(tmin <= tmax)) hit = 0;
        if (!hit) continue;
        if (tmin >= 0.f && tmin <= 1.f && tmin < tFront) { tFront = tmin; aFront = ain; pFront = pin; }
        if (tmax >= 0.f && tmax <= 1.f) {
            if (tmax < tBack1) tBack1 = tmax;
            if (tmax > tBackFar) { tBackFar = tmax; aBackFar = aout; pBackFar = pout; }
        }
    }
    const bool haveFront = isfinite(tFront), haveBack = isfinite(tBackFar) && tBackFar >= 0.f, haveBack1 = isfinite(tBack1);
    float pF[3] = {0.f, 0.f, 0.f}, pX[3] = {0.f, 0.f, 0.f};
    float dF = 1.f, dX = 0.f;
    if (haveFront) {
#pragma unroll
        for (int k = 0; k < 3; ++k) pF[k] = (k == aFront) ? pFront : xadd(o[k], xmul(d[k], tFront));
        const float4 c = xmat_mul(a.mvp, make_float4(pF[0], pF[1], pF[2], 1.f));
        dF = xadd(xmul(0.5f, xdiv(c.z, c.w)), 0.5f);
    }
    if (haveBack) {
#pragma unroll
        for (int k = 0; k < 3; ++k) pX[k] = (k == aBackFar) ? pBackFar : xadd(o[k], xmul(d[k], tBackFar));
        const float4 c = xmat_mul(a.mvp, make_float4(pX[0], pX[1], pX[2], 1.f));
        dX = xadd(xmul(0.5f, xdiv(c.z, c.w)), 0.5f);
    }
    if (a.world) {   // colour = world position (eep_simple.vert:35-38)
        if (haveFront) { const float4 w = xmat_mul(a.model, make_float4(pF[0], pF[1], pF[2], 1.f)); pF[0] = w.x; pF[1] = w.y; pF[2] = w.z; }
        if (haveBack) { const float4 w = xmat_mul(a.model, make_float4(pX[0], pX[1], pX[2], 1.f)); pX[0] = w.x; pX[1] = w.y; pX[2] = w.z; }
    }
    float4 ecol = make_float4(0.f, 0.f, 0.f, 0.f), xcol = make_float4(0.f, 0.f, 0.f, 0.f);
    float edep = 1.0f, xdep = 0.0f;
    if (a.inside == B200RC_INSIDE_GPU) {        // fillEntryPoints
        if (haveBack1) {
            const float r[3] = {a.world ? rn.x : on.x, a.world ? rn.y : on.y, a.world ? rn.z : on.z};
            const float* bb = a.world ? a.wbb : a.bb;
            const bool outside = r[0] > bb[3] || r[1] > bb[4] || r[2] > bb[5] || r[0] < bb[0] || r[1] < bb[1] || r[2] < bb[2];
            if (outside) { if (haveFront) { ecol = make_float4(pF[0], pF[1], pF[2], 1.f); edep = dF; } }
            else { ecol = make_float4(r[0], r[1], r[2], 1.0f); edep = 0.0f; }
        }
    } else if (haveFront) { ecol = make_float4(pF[0], pF[1], pF[2], 1.f); edep = dF; }
    if (haveBack) { xcol = make_float4(pX[0], pX[1], pX[2], 1.f); xdep = dX; }
    if (!a.world) {
        ecol = make_float4(unorm16_roundtrip(ecol.x), unorm16_roundtrip(ecol.y), unorm16_roundtrip(ecol.z), unorm16_roundtrip(ecol.w));
        xcol = make_float4(unorm16_roundtrip(xcol.x), unorm16_roundtrip(xcol.y), unorm16_roundtrip(xcol.z), unorm16_roundtrip(xcol.w));
    } else {
        ecol = make_float4(half_roundtrip(ecol.x), half_roundtrip(ecol.y), half_roundtrip(ecol.z), half_roundtrip(ecol.w));
        xcol = make_float4(half_roundtrip(xcol.x), half_roundtrip(xcol.y), half_roundtrip(xcol.z), half_roundtrip(xcol.w));
    }
    entry[pix] = ecol; exitp[pix] = xcol;
    entry_depth[pix] = depth24_roundtrip(edep); exit_depth[pix] = depth24_roundtrip(xdep);
}

// eep_jitter.frag:38-78
__global__ void eep_jitter_kernel(int W, int H, float stepLen, int world, float4* __restrict__ entry, float* __restrict__ entry_depth, const float4* __restrict__ exitp) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= W || py >= H) return;
    const size_t pix = (size_t)py * W + px;
    float4 fr = entry[pix]; const float4 bk = exitp[pix];
    float fragDepth;
    if (bk.w == 0.0f || fr.w == 0.0f) fragDepth = 1.0f;
    else {
        const float dx = xsub(bk.x, fr.x), dy = xsub(bk.y, fr.y), dz = xsub(bk.z, fr.z);
        const float rayLength = __fsqrt_rn(xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz)));
        fragDepth = entry_depth[pix];
        if (!(rayLength <= stepLen)) {
            const float x = (float)px + 0.5f, y = (float)py + 0.5f;
            float permY = xadd(xmul(xmul(62.0f, y), y), y); permY = xsub(permY, xmul(961.0f, floorf(xdiv(permY, 961.0f))));
            const float arg = xadd(x, permY);
            float perm = xadd(xmul(xmul(62.0f, arg), arg), arg); perm = xsub(perm, xmul(961.0f, floorf(xdiv(perm, 961.0f))));
            const float jitterValue = xdiv(perm, 961.0f);
            const float s = xmul(jitterValue, stepLen);
            const float nx = xadd(fr.x, xmul(xdiv(dx, rayLength), s)), ny = xadd(fr.y, xmul(xdiv(dy, rayLength), s)), nz = xadd(fr.z, xmul(xdiv(dz, rayLength), s));
            const float al = xsub(1.0f, s);
            if (!world) fr = make_float4(unorm16_roundtrip(nx), unorm16_roundtrip(ny), unorm16_roundtrip(nz), unorm16_roundtrip(al));
            else fr = make_float4(half_roundtrip(nx), half_roundtrip(ny), half_roundtrip(nz), half_roundtrip(al));
            entry[pix] = fr;
        }
    }
    entry_depth[pix] = depth24_roundtrip(fragDepth);
}

int rc_eep_launch(b200rc_ctx* ctx, const b200rc_eep_params* eep, const b200rc_camera* cam, const float* boxes, int n_boxes,
                  const float model[16], int W, int H, float4* entry, float* entry_depth, float4* exitp, float* exit_depth) {
    EepArgs a; memset(&a, 0, sizeof a);
    if (!rc_mat4_invert(cam->view, a.viewInv) || !rc_mat4_invert(cam->proj, a.projInv) || !rc_mat4_invert(model, a.modelInv)) {
        ctx->set_error("singular camera / model matrix"); return B200RC_ERR_INVALID;
    }
    float mv[16];
    rc_mat4_mul(cam->view, model, mv);
    rc_mat4_mul(cam->proj, mv, a.mvp);
    memcpy(a.model, model, sizeof a.model);
    for (int k = 0; k < 3; ++k) { a.bb[k] = FLT_MAX; a.bb[3 + k] = -FLT_MAX; }
    for (int b = 0; b < n_boxes; ++b)
        for (int k = 0; k < 3; ++k) { a.bb[k] = fminf(a.bb[k], boxes[6 * b + k]); a.bb[3 + k] = fmaxf(a.bb[3 + k], boxes[6 * b + 3 + k]); }
    for (int k = 0; k < 3; ++k) { a.wbb[k] = FLT_MAX; a.wbb[3 + k] = -FLT_MAX; }
    for (int c = 0; c < 8; ++c) {       // Geometry::getBoundingBox(true): transformed corners
        const float p[4] = {a.bb[(c & 1) ? 3 : 0], a.bb[(c & 2) ? 4 : 1], a.bb[(c & 4) ? 5 : 2], 1.f};
        for (int k = 0; k < 3; ++k) {
            const float w = ((model[k * 4] * p[0] + model[k * 4 + 1] * p[1]) + model[k * 4 + 2] * p[2]) + model[k * 4 + 3] * p[3];
            a.wbb[k] = fminf(a.wbb[k], w); a.wbb[3 + k] = fmaxf(a.wbb[3 + k], w);
        }
    }
    a.W = W; a.H = H; a.n_boxes = n_boxes; a.inside = eep->inside_technique; a.world = eep->world_coords;
    a.rcpW = 1.0f / (float)W; a.rcpH = 1.0f / (float)H;
    if (n_boxes > ctx->boxes_capacity) {
        cudaFree(ctx->d_boxes); ctx->d_boxes = nullptr;
        RC_CHECK(cudaMalloc(&ctx->d_boxes, sizeof(float) * 6 * (size_t)(n_boxes > 16 ? n_boxes : 16)));
        ctx->boxes_capacity = n_boxes > 16 ? n_boxes : 16;
    }
    if (n_boxes > 0) RC_CHECK(cudaMemcpyAsync(ctx->d_boxes, boxes, sizeof(float) * 6 * (size_t)n_boxes, cudaMemcpyHostToDevice, ctx->stream));
    dim3 block(16, 8), grid((W + 15) / 16, (H + 7) / 8);
    eep_kernel<<<grid, block, 0, ctx->stream>>>(a, ctx->d_boxes, entry, entry_depth, exitp, exit_depth);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    if (eep->jitter) {
        eep_jitter_kernel<<<grid, block, 0, ctx->stream>>>(W, H, eep->jitter_step_length, eep->world_coords, entry, entry_depth, exitp);
        ctx->launches++;
        RC_CHECK(cudaGetLastError());
    }
    return B200RC_OK;
}
