// march_multi.cu -- specialised instantiations of the multi-volume marcher (rc_multivolume.frag semantics, march.cuh
// march_multi) for the hot MultiVolumeRaycaster configuration: every volume one u16 channel, the same classification
// mode for all volumes, central-difference gradients, Phong, only image.output connected.  Same source as the generic
// kernel with the switches folded, so results are identical by construction.
#include "march_launch.cuh"

template <int CLS> using FastMultiCfg = MarchCfg<B200RC_SAMPLER_EXACT, 1, B200RC_U16, CLS, B200RC_GRAD_CENTRAL, B200RC_SHADE_PHONG, -1, true, true>;

bool rc_march_fast_multi(b200rc_ctx* ctx, const DevParams& p, int sampler, int* rc_out) {
    if (!p.multi || sampler != B200RC_SAMPLER_EXACT) return false;
    if (!p.connected[0] || p.connected[1] || p.connected[2]) return false;
    if (p.gradient != B200RC_GRAD_CENTRAL) return false;
    const int cls = p.classification[0];
    for (int v = 0; v < p.nvol; ++v) {
        if (p.vol[v].dtype != B200RC_U16 || p.vol[v].nch != 1) return false;
        if (p.classification[v] != cls || p.shading[v] != B200RC_SHADE_PHONG) return false;
    }
    if (cls == B200RC_CLASS_TF) { *rc_out = rc_launch_march_cfg<FastMultiCfg<B200RC_CLASS_TF>>(ctx, p); return true; }
    if (cls == B200RC_CLASS_PREINT) { *rc_out = rc_launch_march_cfg<FastMultiCfg<B200RC_CLASS_PREINT>>(ctx, p); return true; }
    return false;
}
