// march_launch.cuh -- shared launch helper for the marcher instantiations.
#pragma once
#include "ctx.hpp"
#include "march.cuh"

template <class C>
int rc_launch_march_cfg(b200rc_ctx* ctx, const DevParams& p) {
    // dynamic shared memory: TF texels + opacity LUT per classifier that has a 1D TF staged
    size_t smem = 0;
    const int ncls = C::MULTI ? p.nvol : C::NCH;
    for (int c = 0; c < ncls; ++c)
        if (p.cls[c].tf) smem += (size_t)p.cls[c].tf_width * 4 + 1024;
    auto kern = march_kernel<C>;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) { ctx->set_error("transfer functions do not fit in shared memory"); return B200RC_ERR_INVALID; }
        RC_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int per_sm = 0;
    RC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, B200RC_MARCH_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long tiles = (long long)p.tiles_x * p.tiles_y;
    long long grid = (long long)ctx->num_sms * per_sm;       // persistent: one wave, multiple of the SM count
    const long long need = (tiles + (B200RC_MARCH_THREADS / 32) - 1) / (B200RC_MARCH_THREADS / 32);
    if (grid > need && !p.pixel_list) grid = need;
    if (grid < 1) grid = 1;
    RC_CHECK(cudaMemsetAsync(p.tile_counter, 0, sizeof(unsigned int), ctx->stream));
    kern<<<(unsigned)grid, B200RC_MARCH_THREADS, smem, ctx->stream>>>(p);
    ctx->launches++;
    RC_CHECK(cudaGetLastError());
    return B200RC_OK;
}
