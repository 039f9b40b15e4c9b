// march_generic.cu -- run-time-switched instantiations of the marcher: every property permutation of
// SingleVolumeRaycaster / MultiVolumeRaycaster (5 gradients x 3 classifications x 10 shadings x 7^3 compositing
// triples x 1-4 channels x channel shift), the CUDA analogue of the reference's shader-header machinery
// (V/src/core/processors/volumeraycaster.cpp:193-272).
#include "march_launch.cuh"

template <int S, int N> using GenericCfg = MarchCfg<S, N, 0, -1, -1, -1, -1, false, false>;
template <int S> using GenericMultiCfg = MarchCfg<S, 1, 0, -1, -1, -1, -1, false, true>;

int rc_march_generic(b200rc_ctx* ctx, const DevParams& p, int sampler, int nch) {
    if (p.multi) {
        if (sampler == B200RC_SAMPLER_TEX) return rc_launch_march_cfg<GenericMultiCfg<B200RC_SAMPLER_TEX>>(ctx, p);
        return rc_launch_march_cfg<GenericMultiCfg<B200RC_SAMPLER_EXACT>>(ctx, p);
    }
    if (sampler == B200RC_SAMPLER_TEX) {
        switch (nch) {
            case 1: return rc_launch_march_cfg<GenericCfg<B200RC_SAMPLER_TEX, 1>>(ctx, p);
            case 2: return rc_launch_march_cfg<GenericCfg<B200RC_SAMPLER_TEX, 2>>(ctx, p);
            case 3: return rc_launch_march_cfg<GenericCfg<B200RC_SAMPLER_TEX, 3>>(ctx, p);
            default: return rc_launch_march_cfg<GenericCfg<B200RC_SAMPLER_TEX, 4>>(ctx, p);
        }
    }
    switch (nch) {
        case 1: return rc_launch_march_cfg<GenericCfg<B200RC_SAMPLER_EXACT, 1>>(ctx, p);
        case 2: return rc_launch_march_cfg<GenericCfg<B200RC_SAMPLER_EXACT, 2>>(ctx, p);
        case 3: return rc_launch_march_cfg<GenericCfg<B200RC_SAMPLER_EXACT, 3>>(ctx, p);
        default: return rc_launch_march_cfg<GenericCfg<B200RC_SAMPLER_EXACT, 4>>(ctx, p);
    }
}
