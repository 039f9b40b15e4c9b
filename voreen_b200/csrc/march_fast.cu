// march_fast.cu -- specialised instantiations of the marcher for the hot configurations (1 channel u16, one connected
// output, central-difference gradients, Phong): same source as the generic kernel with the switches folded at compile
// time, so results are identical by construction.
#include "march_launch.cuh"

template <int S, int DT, int CLS, int COMP> using FastCfg = MarchCfg<S, 1, DT, CLS, B200RC_GRAD_CENTRAL, B200RC_SHADE_PHONG, COMP, true, false>;

#define FAST_CASE(S, DT, CLS, COMP)                                                            \
    if (sampler == S && cls == CLS && comp == COMP) { *rc_out = rc_launch_march_cfg<FastCfg<S, DT, CLS, COMP>>(ctx, p); return true; }

bool rc_march_fast(b200rc_ctx* ctx, const DevParams& p, int sampler, int nch, int dtype, int* rc_out) {
    if (p.multi || nch != 1 || p.channel_shift) return false;
    if (!p.connected[0] || p.connected[1] || p.connected[2]) return false;
    if (p.gradient != B200RC_GRAD_CENTRAL || p.shading[0] != B200RC_SHADE_PHONG) return false;
    if (sampler == B200RC_SAMPLER_EXACT && dtype != B200RC_U16) return false;
    const int cls = p.classification[0], comp = p.comp[0];
    FAST_CASE(B200RC_SAMPLER_EXACT, B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_DVR)
    FAST_CASE(B200RC_SAMPLER_EXACT, B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_MIP)
    FAST_CASE(B200RC_SAMPLER_EXACT, B200RC_U16, B200RC_CLASS_TF, B200RC_COMP_ISO)
    FAST_CASE(B200RC_SAMPLER_EXACT, B200RC_U16, B200RC_CLASS_PREINT, B200RC_COMP_DVR)
    FAST_CASE(B200RC_SAMPLER_EXACT, B200RC_U16, B200RC_CLASS_PREINT, B200RC_COMP_ISO)
    FAST_CASE(B200RC_SAMPLER_TEX, 0, B200RC_CLASS_TF, B200RC_COMP_DVR)
    FAST_CASE(B200RC_SAMPLER_TEX, 0, B200RC_CLASS_TF, B200RC_COMP_MIP)
    FAST_CASE(B200RC_SAMPLER_TEX, 0, B200RC_CLASS_TF, B200RC_COMP_ISO)
    FAST_CASE(B200RC_SAMPLER_TEX, 0, B200RC_CLASS_PREINT, B200RC_COMP_DVR)
    FAST_CASE(B200RC_SAMPLER_TEX, 0, B200RC_CLASS_PREINT, B200RC_COMP_ISO)
    return false;
}
