// preint.cu -- transfer-function state on the device: 1D TF texels, opacity-correction LUT, pre-integration tables
// (uploaded or built here), 2D TFs, and the "any alpha > 0" prefix sums the empty-space classifier queries.
//
// Replaces TransFuncBase::getTexture()->bind / TransFunc1D::setUniform
// (V/src/core/datastructures/transfunc/1d/transfunc1d.cpp:184-194), ClassificationModes::bindTexture
// (V/src/core/utils/classificationmodes.cpp:56-73), PreIntegrationTable::computeTable / computeTableGPU
// (V/src/core/datastructures/transfunc/1d/preintegrationtable.cpp:67-254) and preintegration.frag:35-87.
#include "ctx.hpp"
#include <cstring>
#include <vector>

// ---- pre-integration table built on the device ------------------------------------------------------------------
// mode B200RC_PI_GPU: preintegration.frag -- fragment (x = sf, y = sb); TF fetched with texture(tf, s/res) NEAREST
//                     (no half-texel offset), diagonal fetched at (x+0.5)/res, result stored as GL_RGBA16.
// mode B200RC_PI_CPU: computeTable(!useIntegral) -- TF entry s read directly (the caller provides a TF texture of
//                     exactly `res` texels sampled at i/(res-1), see host mirror), fp32 table, clamp.
// One thread per table entry; the inner loop is the reference's O(|sb - sf|) walk with the 0.95 early out.
__global__ void preint_build_kernel(const uchar4* __restrict__ tf, int tfw, int res, float step, int mode, float4* __restrict__ table) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= res || y >= res) return;
    const float k255 = 1.0f / 255.0f;
    const float rcp = 1.0f / (float)res;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](int s, bool centre) -> float4 {
        int i;
        if (mode == B200RC_PI_GPU) {
            const float coord = centre ? __fmul_rn(__fadd_rn((float)s, 0.5f), rcp) : __fmul_rn((float)s, rcp);
            i = __float2int_rd(__fmul_rn(coord, (float)tfw));
        } else i = s;
        i = min(max(i, 0), tfw - 1);
        const uchar4 q = tf[i];
        // rounded here, never contracted into the 1 - alpha below (255 * k255 must be exactly 1.0: pow(-6e-8, fraction) is NaN)
        return make_float4(__fmul_rn(q.x, k255), __fmul_rn(q.y, k255), __fmul_rn(q.z, k255), __fmul_rn(q.w, k255));
    };
    if (x == y) {
        c = fetch(x, true);
        if (mode != B200RC_PI_GPU) { c.x = fminf(fmaxf(c.x, 0.f), 1.f); c.y = fminf(fmaxf(c.y, 0.f), 1.f); c.z = fminf(fmaxf(c.z, 0.f), 1.f); c.w = fminf(fmaxf(c.w, 0.f), 1.f); }
        c.w = 1.0f - powf(1.0f - c.w, step * 200.0f);
    } else {
        const int sf = x, sb = y;
        const float scale = 1.0f / (fabsf((float)(sb - sf)) + 1.0f);
        const int incr = sb > sf ? 1 : -1;
        for (int s = sf; (incr * s) <= (incr * sb) && c.w < 0.95f; s += incr) {
            float4 cur = fetch(s, false);
            if (cur.w > 0.0f) {
                cur.w = 1.0f - powf(1.0f - cur.w, step * 200.0f * scale);
                const float k = (1.0f - c.w) * cur.w;
                c.x = c.x + k * cur.x; c.y = c.y + k * cur.y; c.z = c.z + k * cur.z;
                c.w = c.w + (1.0f - c.w) * cur.w;
            }
        }
        const float d = fmaxf(c.w, 0.001f);
        c.x /= d; c.y /= d; c.z /= d;
        c.x = fminf(fmaxf(c.x, 0.f), 1.f); c.y = fminf(fmaxf(c.y, 0.f), 1.f); c.z = fminf(fmaxf(c.z, 0.f), 1.f); c.w = fminf(fmaxf(c.w, 0.f), 1.f);
    }
    if (mode == B200RC_PI_GPU) c = make_float4(unorm16_roundtrip(c.x), unorm16_roundtrip(c.y), unorm16_roundtrip(c.z), unorm16_roundtrip(c.w));
    table[(size_t)y * res + x] = c;
}

// integral-function variant (preintegrationtable.cpp:146-201): integ[] built by one thread (res <= 4096), then a map
__global__ void preint_integral_kernel(const uchar4* __restrict__ tf, int res, float4* __restrict__ integ) {
    if (blockIdx.x || threadIdx.x) return;
    const float k255 = 1.0f / 255.0f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < res; ++i) {
        const uchar4 q = tf[i];
        const float4 c = make_float4(__fmul_rn(q.x, k255), __fmul_rn(q.y, k255), __fmul_rn(q.z, k255), __fmul_rn(q.w, k255));
        if (c.w > 0.0f) {
            acc.x = __fadd_rn(acc.x, __fmul_rn(c.x, c.w)); acc.y = __fadd_rn(acc.y, __fmul_rn(c.y, c.w));
            acc.z = __fadd_rn(acc.z, __fmul_rn(c.z, c.w)); acc.w = __fadd_rn(acc.w, c.w);
        }
        integ[i] = acc;
    }
}
__global__ void preint_approx_kernel(const uchar4* __restrict__ tf, const float4* __restrict__ integ, int res, float step, float4* __restrict__ table) {
    const int sf = blockIdx.x * blockDim.x + threadIdx.x, sb = blockIdx.y * blockDim.y + threadIdx.y;
    if (sf >= res || sb >= res) return;
    const int smin = min(sb, sf), smax = max(sb, sf);
    float4 col;
    if (smax != smin) {
        const float factor = step * 200.f / (float)(smax - smin);
        const float4 a = integ[smax], b = integ[smin];
        col.x = (a.x - b.x) * factor; col.y = (a.y - b.y) * factor; col.z = (a.z - b.z) * factor;
        col.w = 1.f - expf(-(a.w - b.w) * factor);
        const float d = fmaxf(col.w, 0.001f);
        col.x /= d; col.y /= d; col.z /= d;
    } else {
        const uchar4 q = tf[smin];
        const float k255 = 1.0f / 255.0f;
        col = make_float4(__fmul_rn(q.x, k255), __fmul_rn(q.y, k255), __fmul_rn(q.z, k255), __fmul_rn(q.w, k255));     // rounded: 255 -> exactly 1.0
        col.w = 1.f - powf(__fsub_rn(1.f, col.w), step * 200.0f);
    }
    table[(size_t)sb * res + sf] = make_float4(fminf(fmaxf(col.x, 0.f), 1.f), fminf(fmaxf(col.y, 0.f), 1.f), fminf(fmaxf(col.z, 0.f), 1.f), fminf(fmaxf(col.w, 0.f), 1.f));
}

// nz[i] = 1 iff the table holds alpha > 0 on the diagonal at i or next to it.  A pre-integration table entry (sf, sb)
// composites the TF entries between sf and sb, so "every diagonal entry in [lo, hi] is zero" <=> "every entry of
// [lo, hi]^2 is zero" (preintegrationtable.cpp:105-133); the off-diagonal neighbours make that robust against the
// opacity-correction pow() underflowing on the diagonal only.
__global__ void preint_nz_kernel(const float4* __restrict__ table, int res, int* __restrict__ nz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= res) return;
    int f = table[(size_t)i * res + i].w > 0.f;
    if (i + 1 < res) f |= (table[(size_t)i * res + i + 1].w > 0.f) || (table[(size_t)(i + 1) * res + i].w > 0.f);
    if (i > 0) f |= (table[(size_t)i * res + i - 1].w > 0.f) || (table[(size_t)(i - 1) * res + i].w > 0.f);
    nz[i] = f;
}
__global__ void exclusive_scan_small_kernel(const int* __restrict__ in, int n, int* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    int acc = 0;
    for (int i = 0; i < n; ++i) { out[i] = acc; acc += in[i]; }
    out[n] = acc;
}

__global__ void corr_lut_kernel(float* __restrict__ lut, float step) {
    const int i = threadIdx.x;
    // mod_compositing.frag:40,51.  The texel alpha is rounded to float before the subtraction (255 -> exactly 1.0): contracted into an
    // fma, 1 - 255 * (1/255.f) is -5.9e-8 and pow(negative, fraction) is NaN -- a fully opaque TF entry must give alpha 1
    if (i < 256) lut[i] = 1.0f - powf(__fsub_rn(1.0f, __fmul_rn((float)i, 1.0f / 255.0f)), step * 200.0f);
}

// ---- host ---------------------------------------------------------------------------------------------------------
static int upload_prefix(b200rc_ctx* ctx, int** dst, const std::vector<int>& flags) {
    std::vector<int> pre(flags.size() + 1, 0);
    for (size_t i = 0; i < flags.size(); ++i) pre[i + 1] = pre[i] + (flags[i] ? 1 : 0);
    cudaFree(*dst); *dst = nullptr;
    RC_CHECK(cudaMalloc(dst, pre.size() * sizeof(int)));
    RC_CHECK(cudaMemcpyAsync(*dst, pre.data(), pre.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200RC_OK;
}

int rc_tf_set_1d(b200rc_ctx* ctx, int chan, const uint8_t* rgba8, int width, float lo, float hi) {
    if (chan < 0 || chan >= B200RC_MAX_CHANNELS || !rgba8 || width < 1 || width > 65536) { ctx->set_error("bad 1D transfer function"); return B200RC_ERR_INVALID; }
    ClassifierSlot& k = ctx->cls[chan];
    // unchanged texels + domain: nothing to do (the occupancy grids, the opacity LUT and a table built from this TF stay valid)
    if (k.tf && k.tf_width == width && k.dom[0] == lo && k.dom[1] == hi && k.tf_host.size() == (size_t)width * 4 &&
        memcmp(k.tf_host.data(), rgba8, (size_t)width * 4) == 0) return B200RC_OK;
    if (k.tf_width != width) { cudaFree(k.tf); k.tf = nullptr; RC_CHECK(cudaMalloc(&k.tf, (size_t)width * 4)); }
    k.tf_width = width; k.dom[0] = lo; k.dom[1] = hi;
    k.tf_gen++;
    k.tf_host.assign(rgba8, rgba8 + (size_t)width * 4);
    RC_CHECK(cudaMemcpyAsync(k.tf, k.tf_host.data(), (size_t)width * 4, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int> flags(width);
    for (int i = 0; i < width; ++i) flags[i] = rgba8[4 * i + 3] != 0;
    int rc = upload_prefix(ctx, &k.nz_tf, flags);
    if (rc) return rc;
    if (!k.corr_lut) RC_CHECK(cudaMalloc(&k.corr_lut, 256 * sizeof(float)));
    k.lut_step = -1.f;
    k.generation++;
    return B200RC_OK;
}

int rc_tf_update_lut(b200rc_ctx* ctx, int chan, float step) {
    ClassifierSlot& k = ctx->cls[chan];
    if (!k.corr_lut) RC_CHECK(cudaMalloc(&k.corr_lut, 256 * sizeof(float)));
    if (k.lut_step == step) return B200RC_OK;
    corr_lut_kernel<<<1, 256, 0, ctx->stream>>>(k.corr_lut, step);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    k.lut_step = step;
    return B200RC_OK;
}

// texture view of the table for B200RC_SAMPLER_TEX (hardware bilinear, clamp to edge, normalised coordinates): created on first
// use, the array is kept while the resolution stays the same
int rc_tf_ensure_preint_texture(b200rc_ctx* ctx, int chan) {
    ClassifierSlot& k = ctx->cls[chan];
    if (!k.preint) return B200RC_OK;
    const int res = k.pres;
    if (k.preint_tex && !k.preint_tex_dirty && k.preint_array_res == res) return B200RC_OK;
    if (!k.preint_array || k.preint_array_res != res) {
        if (k.preint_tex) { cudaDestroyTextureObject(k.preint_tex); k.preint_tex = 0; }
        if (k.preint_array) { cudaFreeArray(k.preint_array); k.preint_array = nullptr; }
        cudaChannelFormatDesc fd = cudaCreateChannelDesc<float4>();
        RC_CHECK(cudaMallocArray(&k.preint_array, &fd, res, res));
        k.preint_array_res = res;
    }
    RC_CHECK(cudaMemcpy2DToArrayAsync(k.preint_array, 0, 0, k.preint, (size_t)res * sizeof(float4), (size_t)res * sizeof(float4), res, cudaMemcpyDeviceToDevice, ctx->stream));
    if (!k.preint_tex) {
        cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypeArray; rd.res.array.array = k.preint_array;
        cudaTextureDesc td; memset(&td, 0, sizeof td);
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
        RC_CHECK(cudaCreateTextureObject(&k.preint_tex, &rd, &td, nullptr));
    }
    k.preint_tex_dirty = false;
    return B200RC_OK;
}

static int preint_finish(b200rc_ctx* ctx, ClassifierSlot& k) {
    const int res = k.pres;
    k.preint_tex_dirty = true;
    // occupancy helper: "any alpha > 0" prefix sums over the table's diagonal band (scratch flags live behind the sums)
    if (!k.nz_pi || k.nz_pi_res != res) {
        cudaFree(k.nz_pi); k.nz_pi = nullptr;
        RC_CHECK(cudaMalloc(&k.nz_pi, (2 * (size_t)res + 1) * sizeof(int)));
        k.nz_pi_res = res;
    }
    int* d_flags = k.nz_pi + res + 1;
    preint_nz_kernel<<<(res + 127) / 128, 128, 0, ctx->stream>>>(k.preint, res, d_flags);
    exclusive_scan_small_kernel<<<1, 1, 0, ctx->stream>>>(d_flags, res, k.nz_pi);
    ctx->launches += 2;
    RC_CHECK(cudaGetLastError());
    k.generation++;
    return B200RC_OK;
}

static unsigned long long hash_words(const void* data, size_t bytes) {      // 4 interleaved FNV-1a lanes over 64-bit words
    const unsigned long long* w = (const unsigned long long*)data;
    const size_t n = bytes / 8;
    unsigned long long h[4] = {0xcbf29ce484222325ull, 0x84222325cbf29ce4ull, 0x9ce484222325cbf2ull, 0x2325cbf29ce48422ull};
    size_t i = 0;
    for (; i + 4 <= n; i += 4) for (int l = 0; l < 4; ++l) h[l] = (h[l] ^ w[i + l]) * 0x100000001b3ull;
    for (; i < n; ++i) h[0] = (h[0] ^ w[i]) * 0x100000001b3ull;
    return h[0] ^ (h[1] * 3) ^ (h[2] * 5) ^ (h[3] * 7) ^ (unsigned long long)bytes;
}

int rc_tf_set_preint(b200rc_ctx* ctx, int chan, const float* rgba32f, int res) {
    if (chan < 0 || chan >= B200RC_MAX_CHANNELS || !rgba32f || res < 2 || res > 4096) { ctx->set_error("bad pre-integration table"); return B200RC_ERR_INVALID; }
    ClassifierSlot& k = ctx->cls[chan];
    const unsigned long long h = hash_words(rgba32f, (size_t)res * res * sizeof(float4));
    if (k.preint && k.pres == res && k.pi_uploaded && k.pi_hash == h) return B200RC_OK;      // the table already resident
    if (k.pres != res) { cudaFree(k.preint); k.preint = nullptr; RC_CHECK(cudaMalloc(&k.preint, (size_t)res * res * sizeof(float4))); }
    k.pres = res;
    RC_CHECK(cudaMemcpyAsync(k.preint, rgba32f, (size_t)res * res * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    RC_CHECK(cudaStreamSynchronize(ctx->stream));       // the caller's buffer may be pageable / reused
    k.pi_uploaded = true; k.pi_hash = h; k.pi_built_mode = -1;
    return preint_finish(ctx, k);
}

int rc_tf_build_preint(b200rc_ctx* ctx, int chan, int res, float step, int mode) {
    if (chan < 0 || chan >= B200RC_MAX_CHANNELS) { ctx->set_error("bad channel"); return B200RC_ERR_INVALID; }
    ClassifierSlot& k = ctx->cls[chan];
    if (!k.tf) { ctx->set_error("build_preint: set the 1D transfer function first"); return B200RC_ERR_NOT_READY; }
    if (res <= 1) res = 256;            // PreIntegrationTable ctor (preintegrationtable.cpp:42-45)
    if (step <= 0.f) step = 1.f;
    if (res > 4096) { ctx->set_error("pre-integration resolution too large"); return B200RC_ERR_INVALID; }
    if (mode != B200RC_PI_GPU && k.tf_width != res) { ctx->set_error("PI_CPU/PI_APPROX need a TF texture of exactly `res` texels"); return B200RC_ERR_INVALID; }
    if (k.preint && k.pres == res && !k.pi_uploaded && k.pi_built_mode == mode && k.pi_built_step == step && k.pi_built_gen == k.tf_gen) return B200RC_OK;
    k.pi_uploaded = false; k.pi_built_mode = mode; k.pi_built_step = step; k.pi_built_gen = k.tf_gen;
    if (k.pres != res) { cudaFree(k.preint); k.preint = nullptr; RC_CHECK(cudaMalloc(&k.preint, (size_t)res * res * sizeof(float4))); }
    k.pres = res;
    dim3 block(16, 16), grid((res + 15) / 16, (res + 15) / 16);
    if (mode == B200RC_PI_APPROX) {
        float4* integ = nullptr;
        RC_CHECK(cudaMallocAsync(&integ, res * sizeof(float4), ctx->stream));
        preint_integral_kernel<<<1, 32, 0, ctx->stream>>>(k.tf, res, integ);
        preint_approx_kernel<<<grid, block, 0, ctx->stream>>>(k.tf, integ, res, step, k.preint);
        ctx->launches += 2;
        RC_CHECK(cudaGetLastError());
        RC_CHECK(cudaFreeAsync(integ, ctx->stream));
    } else {
        preint_build_kernel<<<grid, block, 0, ctx->stream>>>(k.tf, k.tf_width, res, step, mode, k.preint);
        ctx->launches++;
        RC_CHECK(cudaGetLastError());
    }
    return preint_finish(ctx, k);
}

int rc_tf_set_2d(b200rc_ctx* ctx, int chan, const uint8_t* rgba8, int w, int h, float lox, float hix, float loy, float hiy) {
    if (chan < 0 || chan >= B200RC_MAX_CHANNELS || !rgba8 || w < 1 || h < 1) { ctx->set_error("bad 2D transfer function"); return B200RC_ERR_INVALID; }
    ClassifierSlot& k = ctx->cls[chan];
    if (k.tf2d && k.tf2d_w == w && k.tf2d_h == h && k.dom[0] == lox && k.dom[1] == hix && k.domy[0] == loy && k.domy[1] == hiy &&
        k.tf2d_host.size() == (size_t)w * h * 4 && memcmp(k.tf2d_host.data(), rgba8, (size_t)w * h * 4) == 0) return B200RC_OK;
    k.tf2d_host.assign(rgba8, rgba8 + (size_t)w * h * 4);
    cudaFree(k.tf2d); k.tf2d = nullptr;
    RC_CHECK(cudaMalloc(&k.tf2d, (size_t)w * h * 4));
    RC_CHECK(cudaMemcpyAsync(k.tf2d, rgba8, (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream));
    k.tf2d_w = w; k.tf2d_h = h; k.dom[0] = lox; k.dom[1] = hix; k.domy[0] = loy; k.domy[1] = hiy;
    std::vector<int> flags(w, 0);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) flags[x] |= rgba8[4 * ((size_t)y * w + x) + 3] != 0;
    int rc = upload_prefix(ctx, &k.nz_2d, flags);
    if (rc) return rc;
    k.generation++;
    return B200RC_OK;
}
