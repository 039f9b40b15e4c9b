"""GPU parity tests added in round 2 (all through the C ABI):

* the sort-last merge kernels themselves (b200rc_composite in every mode, b200rc_finalize with and without the ERT snap)
  against the oracle's restatement / numpy on random and adversarial partial images;
* BASELINE.json's configurations at FULL SIZE against the oracle: crops of the C2 / C3 / C4 frames are marched by the oracle
  on a CPU-synthesized window of the same procedural volume (identical bytes) and must agree with the GPU frame to
  <= 1/255 on >= 99.9 % of the pixels with equal ray, nominal-sample and shaded-sample counts;
* the real multi-process sort-last / sort-first path (torchrun, one rank per GPU) when the box has >= 2 GPUs;
* shards whose border voxels are opaque (the outer half voxel belongs to the border shard), sharded MIP end to end
  (no alpha snap), and the rejection of key planes for multi-channel volumes.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import orc
import scenes
from voreen_b200 import abi

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------------------------------
# merge kernels
# ---------------------------------------------------------------------------------------------------------------------
def _partial_images(n, seed):
    """Random premultiplied partial images + the adversarial values: alpha exactly at / around the 0.95 ERT threshold, empty
    pixels, fully opaque pixels, equal keys, depth 1.0 (no hit), negative zero."""
    rng = np.random.default_rng(seed)
    def one():
        a = rng.random(n, dtype=np.float32)
        a[rng.random(n) < 0.2] = 0.0
        c = rng.random((n, 3), dtype=np.float32) * a[:, None]
        d = np.where(a > 0, rng.random(n, dtype=np.float32), np.float32(1.0)).astype(np.float32)
        img = np.concatenate([c, a[:, None]], axis=1).astype(np.float32)
        special = np.array([0.95, np.nextafter(np.float32(0.95), np.float32(0)), np.nextafter(np.float32(0.95), np.float32(1)), 1.0, 0.0, -0.0, 0.9499999, 0.5], np.float32)
        k = min(n, 4096)
        img[:k, 3] = special[rng.integers(0, len(special), k)]
        return img, d
    f, b = one(), one()
    b[1][: n // 8] = f[1][: n // 8]         # equal keys / depths: the front image must win the tie
    return f, b


@pytest.mark.parametrize("mode", ["dvr+ert", "dvr", "iso", "fhp", "fhn", "mip", "mop"])
def test_composite_kernel_against_the_oracle(ctx, mode):
    m = (abi.COMP["dvr"] | abi.COMPOSITE_ERT) if mode == "dvr+ert" else abi.COMP[mode]
    for n, seed in [(1, 1), (255, 2), (96 * 88, 3), (1 << 20, 4)]:
        front, back = _partial_images(n, seed)
        dc, dd = scenes.product_composite(ctx, m, front, back)
        oc = np.empty_like(front[0]); od = np.empty_like(front[1])
        orc.lib().orc_composite(m, orc.fp(front[0]), orc.fp(front[1]), orc.fp(back[0]), orc.fp(back[1]), orc.fp(oc), orc.fp(od), n)
        if mode.startswith("dvr"):
            # F + (1 - Fa) * B: the device contracts the multiply-add, the CPU may not -> one rounding apart at most
            assert float(np.abs(dc - oc).max()) <= 2e-7, (mode, n)
            if mode == "dvr+ert":       # a terminated front ray hides everything behind it, exactly
                t = front[0][:, 3] >= np.float32(0.95)
                assert np.array_equal(dc[t], front[0][t])
        else:
            assert np.array_equal(dc, oc), (mode, n)
        assert np.array_equal(dd, od), (mode, n)
        # in-place use (dst aliases front), as the binary swap does
        import torch
        dev = torch.device("cuda", 0)
        fc, fd, bc, bd = (torch.from_numpy(a).to(dev) for a in (front[0], front[1], back[0], back[1]))
        ctx.composite(m, fc.data_ptr(), fd.data_ptr(), bc.data_ptr(), bd.data_ptr(), fc.data_ptr(), fd.data_ptr(), n)
        ctx.synchronize()
        assert np.array_equal(fc.cpu().numpy(), dc) and np.array_equal(fd.cpu().numpy(), dd)


@pytest.mark.parametrize("snap", [True, False])
@pytest.mark.parametrize("fmt", ["rgba16", "rgba32f"])
def test_finalize_kernel(ctx, snap, fmt):
    f = abi.OUT_RGBA16 if fmt == "rgba16" else abi.OUT_RGBA32F
    for n, seed in [(1, 5), (1000, 6), (1 << 18, 7)]:
        img, _ = _partial_images(n, seed)[0]
        img = img.copy()
        img[: min(n, 64), 0] = np.linspace(-0.5, 1.5, min(n, 64), dtype=np.float32)       # clamps of the unorm16 store
        if n > 100:
            img[100, 1] = np.nan
        got = scenes.product_finalize(ctx, img, f, snap=snap)
        want = scenes.numpy_finalize(img, f, snap=snap)
        if f == abi.OUT_RGBA32F:
            assert np.array_equal(got, want, equal_nan=True)
        else:
            assert np.array_equal(got, want)
        a = img[:, 3]
        touched = (a >= np.float32(0.95)) & (a < 1)
        if touched.any() and f == abi.OUT_RGBA32F:
            assert np.all(got[touched, 3] == (1.0 if snap else a[touched]))


# ---------------------------------------------------------------------------------------------------------------------
# full-size configurations against the oracle, crop by crop
# ---------------------------------------------------------------------------------------------------------------------
def _window_for(crops, dims):
    lo = np.array([1e9] * 3); hi = np.array([-1e9] * 3)
    for (e, ed, x, xd) in crops:
        m = np.any(e[..., :3] != x[..., :3], axis=2)
        if not m.any():
            continue
        pts = np.concatenate([e[m][:, :3], x[m][:, :3]])
        lo = np.minimum(lo, pts.min(axis=0)); hi = np.maximum(hi, pts.max(axis=0))
    if lo[0] > hi[0]:
        return (0, 0, 0), (8, 8, 8)
    d = np.array(dims, np.float64)
    vlo = np.clip(np.floor(lo * d - 0.5).astype(int) - 4, 0, dims)
    vhi = np.clip(np.ceil(hi * d - 0.5).astype(int) + 5, 0, dims)
    return tuple(int(v) for v in vlo), tuple(int(v) for v in vhi)


CROPS = [(0.5, 0.5), (0.31, 0.62), (0.2, 0.47)]     # centre, off-centre, across the silhouette (fractions of the image)


@pytest.mark.parametrize("n,img,cls,comp", [(512, 1024, "tf", "dvr"), (2048, 2048, "preint", "dvr"), (2048, 2048, "tf", "mip"), (2048, 2048, "tf", "iso")])
def test_full_size_crops_against_the_oracle(n, img, cls, comp):
    """C2 / C3: the GPU renders the whole frame; 64x64 crops of it are (a) re-rendered on the GPU from the cropped entry/exit images
    (b200rc_render) -- bit-identical to the frame -- and (b) marched by the oracle over a CPU-synthesized window of the volume."""
    import voreen_b200 as vb
    from voreen_b200 import host
    from voreen_b200 import raycaster as rc
    dims = (n, n, n)
    step = float(np.float32(1.0) / (np.float32(n) * np.float32(2.0)))
    edge = 64
    c = vb.Context(0)
    try:
        c.synth_volume(0, vb.make_volume_desc(dims), 1234)
        tf = orc.make_tf(scenes.CT_KEYS, domain=(0.0, 65535.0))
        tf_tex = orc.tf1d_texture(tf, 4096)
        preint = None
        if cls == "preint":
            c.set_tf1d(0, host.tf1d_texture(scenes.CT_KEYS, 1024), (0.0, 65535.0))
            c.build_preint(0, 1024, step, abi.PI_CPU)
            preint = orc.preint_cpu(tf, 1024, step)
        c.set_tf1d(0, host.tf1d_texture(scenes.CT_KEYS, 4096), (0.0, 65535.0))
        p = vb.default_params(step, classification=abi.CLASS_PREINT if cls == "preint" else abi.CLASS_TF, empty_space_skipping=2 if comp == "mip" else 1,
                              out_format=abi.OUT_RGBA16, compositing=[abi.COMP[comp]] * 3, iso_value=0.5)
        eep = vb.default_eep()
        full_desc = rc.make_volume_desc(dims)
        checked = 0
        for k in (3, 20):
            cam = host.make_camera(scenes.orbit_eye(k), W=img, H=img)
            p.collect_stats = 0
            outs, depth = c.render_camera(p, eep, cam, img, img)
            frame = np.asarray(outs[0]).copy(); fdepth = depth.copy()
            e, ed, x, xd = c.entry_exit(eep, cam, img, img)
            crops = []
            for fx, fy in CROPS:
                x0, y0 = int(fx * img) - edge // 2, int(fy * img) - edge // 2
                sl = (slice(y0, y0 + edge), slice(x0, x0 + edge))
                crops.append((sl, tuple(np.ascontiguousarray(a[sl]) for a in (e, ed, x, xd))))
            vlo, vhi = _window_for([cr[1] for cr in crops], dims)
            rd = tuple(vhi[a] - vlo[a] for a in range(3))
            data = orc.synth_volume(dims, 1234, vlo, rd)
            wdesc = rc.make_volume_desc(dims, shard_lo=vlo, shard_hi=vhi, shard_apron=0)
            ov = orc.make_volume(data, wdesc)
            ocls = [orc.make_classifier(tf_tex, (0.0, 65535.0), preint=preint)]
            for sl, (ce, ced, cx, cxd) in crops:
                # (a) the same rays through b200rc_render with the cropped entry/exit images: the frame's pixels, bit for bit
                p.collect_stats = 1
                g_outs, g_depth = c.render(p, cam, edge, edge, ce, ced, cx, cxd)
                st = c.stats()
                assert np.array_equal(np.asarray(g_outs[0]), frame[sl]) and np.array_equal(g_depth, fdepth[sl])
                # (b) the oracle (no empty-space skipping, float output)
                po = rc.default_params(step, classification=p.vol[0].classification, empty_space_skipping=0, out_format=abi.OUT_RGBA32F,
                                       compositing=[abi.COMP[comp]] * 3, iso_value=0.5)
                po.collect_stats = 1
                o_outs, o_depth, o_st = orc.raycast([ov], ocls, po, cam, edge, edge, ce, ced, cx, cxd)
                # both sides as GL_RGBA16 (what the outport holds); 1/255 = 257 units
                got = np.asarray(g_outs[0]).view(np.uint16).astype(np.int64)
                want = scenes.numpy_finalize(np.asarray(o_outs[0], np.float32).reshape(-1, 4), abi.OUT_RGBA16, snap=False).reshape(got.shape).astype(np.int64)
                d = np.abs(got - want).max(axis=2) / 65535.0
                frac = float((d <= 1.0 / 255.0).mean())
                dfrac = float((np.abs(o_depth - g_depth) <= 1e-4).mean())
                print(f"C{'2' if n == 512 else '3'} {n}^3 @{img}^2 {cls} {comp} view {k} crop {sl[1].start},{sl[0].start}: within 1/255 {frac:.5f} max {d.max():.2e} depth {dfrac:.5f} "
                      f"rays {st.rays_cast}/{o_st.rays_cast} nominal {st.samples_nominal}/{o_st.samples_nominal} shaded {st.samples_shaded}/{o_st.samples_shaded} "
                      f"classified+skipped {st.samples_classified + st.samples_skipped}/{o_st.samples_classified}")
                assert st.rays_cast == o_st.rays_cast and st.samples_nominal == o_st.samples_nominal
                if comp != "mip":       # MIP culling (mode 2) classifies fewer samples by design; every other mode takes the oracle's decisions
                    assert st.samples_shaded == o_st.samples_shaded
                    assert st.samples_classified + st.samples_skipped == o_st.samples_classified
                assert frac >= 0.999 and dfrac >= 0.999
                checked += int(o_st.rays_cast)
        assert checked > 5000
    finally:
        c.close()


def test_c4_multivolume_crops_against_the_oracle():
    """C4: MultiVolumeRaycaster, 4 x 1024^3 u16 with per-volume TFs, DVR at 2048^2 -- crops against the oracle as above."""
    import voreen_b200 as vb
    from voreen_b200 import host
    from voreen_b200 import raycaster as rc
    n, img, nvol, edge = 1024, 2048, 4, 48
    dims = (n, n, n)
    keys = [scenes.CT_KEYS, [(0.0, (0, 0, 0, 0)), (0.3, (0, 255, 0, 0)), (1.0, (0, 255, 0, 120))],
            [(0.0, (0, 0, 0, 0)), (0.4, (0, 0, 255, 0)), (1.0, (0, 0, 255, 90))], [(0.0, (0, 0, 0, 0)), (0.5, (255, 0, 255, 0)), (1.0, (255, 0, 255, 60))]]
    step_world = float(np.float32(2.0) / np.float32(n) / np.float32(2.0))      # multivolumeraycaster.cpp:326-344
    step_tex = float(np.float32(1.0) / (np.float32(n) * np.float32(2.0)))
    c = vb.Context(0)
    try:
        tex = []
        for v in range(nvol):
            c.synth_volume(v, vb.make_volume_desc(dims), 1234 + v)
            tex.append(orc.tf1d_texture(orc.make_tf(keys[v], domain=(0.0, 65535.0)), 4096))
            c.set_tf1d(v, host.tf1d_texture(keys[v], 4096), (0.0, 65535.0))
        eep = vb.default_eep(world_coords=1)
        checked = 0
        for k in (3, 20):
            cam = host.make_camera(scenes.orbit_eye(k), W=img, H=img)
            e, ed, x, xd = c.entry_exit(eep, cam, img, img)
            p = vb.default_params(step_world, empty_space_skipping=1, out_format=abi.OUT_RGBA16)
            p.num_volumes = nvol; p.multi_volume = 1
            p.mv_opacity_correction_factor = step_tex / step_world
            outs, depth = c.render(p, cam, img, img, np.stack([e] * (nvol + 1)), ed, np.stack([x] * (nvol + 1)), xd)
            frame = np.asarray(outs[0]).copy(); fdepth = depth.copy()
            crops = []
            for fx, fy in CROPS[:2]:
                x0, y0 = int(fx * img) - edge // 2, int(fy * img) - edge // 2
                sl = (slice(y0, y0 + edge), slice(x0, x0 + edge))
                crops.append((sl, tuple(np.ascontiguousarray(a[sl]) for a in (e, ed, x, xd))))
            # entry/exit are WORLD coordinates here: texture = (world + 1) / 2 for the [-1, 1]^3 cube of make_volume_desc
            tex_crops = [((ce[..., :3] + 1.0) * 0.5, None, (cx[..., :3] + 1.0) * 0.5, None) for _, (ce, ced, cx, cxd) in crops]
            vlo, vhi = _window_for([(np.concatenate([a, np.zeros(a.shape[:2] + (1,), a.dtype)], axis=2), None,
                                     np.concatenate([b, np.zeros(b.shape[:2] + (1,), b.dtype)], axis=2), None) for a, _, b, _ in tex_crops], dims)
            rd = tuple(vhi[a] - vlo[a] for a in range(3))
            wdesc = rc.make_volume_desc(dims, shard_lo=vlo, shard_hi=vhi, shard_apron=0)
            ovs = [orc.make_volume(orc.synth_volume(dims, 1234 + v, vlo, rd), wdesc) for v in range(nvol)]
            ocls = [orc.make_classifier(tex[v], (0.0, 65535.0)) for v in range(nvol)]
            for sl, (ce, ced, cx, cxd) in crops:
                p.collect_stats = 1
                g_outs, g_depth = c.render(p, cam, edge, edge, np.stack([ce] * (nvol + 1)), ced, np.stack([cx] * (nvol + 1)), cxd)
                st = c.stats()
                p.collect_stats = 0
                assert np.array_equal(np.asarray(g_outs[0]), frame[sl]) and np.array_equal(g_depth, fdepth[sl])
                po = rc.default_params(step_world, empty_space_skipping=0, out_format=abi.OUT_RGBA32F)
                po.num_volumes = nvol; po.multi_volume = 1; po.collect_stats = 1
                po.mv_opacity_correction_factor = step_tex / step_world
                o_outs, o_depth, o_st = orc.raycast(ovs, ocls, po, cam, edge, edge, np.stack([ce] * (nvol + 1)), ced, np.stack([cx] * (nvol + 1)), cxd)
                # both sides as GL_RGBA16 (what the outport holds); 1/255 = 257 units
                got = np.asarray(g_outs[0]).view(np.uint16).astype(np.int64)
                want = scenes.numpy_finalize(np.asarray(o_outs[0], np.float32).reshape(-1, 4), abi.OUT_RGBA16, snap=False).reshape(got.shape).astype(np.int64)
                d = np.abs(got - want).max(axis=2) / 65535.0
                frac = float((d <= 1.0 / 255.0).mean())
                dfrac = float((np.abs(o_depth - g_depth) <= 1e-4).mean())
                print(f"C4 4x{n}^3 @{img}^2 view {k} crop {sl[1].start},{sl[0].start}: within 1/255 {frac:.5f} max {d.max():.2e} depth {dfrac:.5f} "
                      f"rays {st.rays_cast}/{o_st.rays_cast} shaded {st.samples_shaded}/{o_st.samples_shaded}")
                assert st.rays_cast == o_st.rays_cast and st.samples_shaded == o_st.samples_shaded
                assert frac >= 0.999 and dfrac >= 0.999
                checked += int(o_st.rays_cast)
        assert checked > 3000
    finally:
        c.close()


# ---------------------------------------------------------------------------------------------------------------------
# the real multi-process path (one rank per GPU)
# ---------------------------------------------------------------------------------------------------------------------
def _device_count():
    import voreen_b200 as vb
    return int(vb.abi.load_library().b200rc_device_count())


@pytest.mark.parametrize("mode,transport,comp", [("sort-last", "p2p", "dvr"), ("sort-last", "nccl", "dvr"), ("sort-last", "p2p", "iso"),
                                                 ("sort-last", "p2p", "mip"), ("sort-first", "p2p", "dvr"), ("sort-last", "p2p", "multi4"),
                                                 ("sort-first", "p2p", "multi4")])
def test_multi_gpu_partitions_match_the_single_gpu_image(mode, transport, comp):
    """torchrun, one rank per GPU (2, 4 or 8): sharded march + exact-ERT second pass + binary swap (b200rc_composite reading the
    partner's half over NVLink, b200rc_finalize storing into rank 0's image) against the single-GPU frame rendered on rank 0."""
    ndev = _device_count()
    if ndev < 2:
        pytest.skip(f"needs >= 2 GPUs (this box has {ndev})")
    world = 8 if ndev >= 8 else (4 if ndev >= 4 else 2)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", SL_TRANSPORT=transport, OMP_NUM_THREADS="1")
    env.pop("SL_NOREF", None)
    env.pop("SL_MULTI", None)
    if comp.startswith("multi"):        # config C4: MultiVolumeRaycaster, every volume sharded the same way
        env["SL_MULTI"] = comp[5:]
    elif comp != "dvr":
        env["SL_COMP"] = comp
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(HERE, "sortlast_gpu_worker.py"), "256", "512", "tf" if comp == "mip" else "preint", mode]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" OK") >= 6 and "FAIL" not in r.stdout


# ---------------------------------------------------------------------------------------------------------------------
# shards: opaque border voxels, MIP end to end, key-plane misuse
# ---------------------------------------------------------------------------------------------------------------------
def _shard_upload(ctx, vol, dims, r, world):
    from voreen_b200 import raycaster as rc
    from voreen_b200 import sortlast
    lo, hi = sortlast.shard_range(dims, r, world)
    rl = [max(0, lo[a] - sortlast.APRON) for a in range(3)]; rh = [min(dims[a], hi[a] + sortlast.APRON) for a in range(3)]
    ctx.upload_volume(0, np.ascontiguousarray(vol[rl[2]:rh[2], rl[1]:rh[1], rl[0]:rh[0]]), desc=rc.make_volume_desc(dims, shard_lo=lo, shard_hi=hi, shard_apron=sortlast.APRON))


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_rendering_with_opaque_border_voxels(ctx, world):
    """The samples of the outer half voxel (tap base index -1 on the low faces) belong to the border shards: a volume whose
    border voxels are opaque must merge to the unsharded image (translucent TF: no early termination, the partition alone)."""
    from voreen_b200 import sortlast
    dims = (48, 40, 56)
    rng = np.random.default_rng(9)
    vol = (rng.random(dims[::-1]) * 20000 + 30000).astype(np.uint16)          # everything visible, border included
    keys = [(0.0, (255, 140, 60, 1)), (1.00, (255, 255, 255, 1))]      # alpha 1/255 everywhere: the longest ray stays below the 0.95 ERT threshold
    for eye in [(1.4, 0.9, 2.9), (-2.2, -1.1, -1.9)]:
        s = scenes.Scene(vol, keys=keys, W=96, H=88, eye=eye, out_format=abi.OUT_RGBA32F)
        s.params.collect_stats = 1
        ref_outs, ref_depth = s.cuda_render(ctx)
        ref_st = ctx.stats()
        cam_tex = sortlast.camera_texture_position(s.cam, s.desc)
        s.params.partial_image = 1
        parts, classified = {}, 0
        for r in range(world):
            _shard_upload(ctx, vol, dims, r, world)
            outs, depth = ctx.render_camera(s.params, s.eep, s.cam, s.W, s.H)
            classified += ctx.stats().samples_classified
            parts[r] = (np.asarray(outs[0]).reshape(-1, 4).copy(), depth.reshape(-1).copy())
        s.params.partial_image = 0
        assert classified == ref_st.samples_classified, "the shards' sample sets must partition the single-GPU sample set"
        imgs = parts
        for rnd in range(world.bit_length() - 1):
            nxt = {}
            for r in sorted(imgs):
                if (r >> rnd) & 1:
                    continue
                a, b = imgs[r], imgs[r | (1 << rnd)]
                front, back = (a, b) if sortlast.my_group_in_front(r, rnd, world, dims, cam_tex) else (b, a)
                nxt[r] = scenes.product_composite(ctx, abi.COMP["dvr"] | abi.COMPOSITE_ERT, front, back)
            imgs = nxt
        merged = scenes.product_finalize(ctx, imgs[0][0], abi.OUT_RGBA32F, snap=True).reshape(s.H, s.W, 4)
        frac, mx = scenes.parity_report(ref_outs[0], merged)
        print(f"opaque border x{world} eye {eye}: within 1/255 {frac:.5f} max {mx:.2e} classified {classified}")
        assert frac >= 0.9999 and mx < 2e-3
        assert float((np.abs(ref_depth.reshape(-1) - imgs[0][1]) <= 1e-4).mean()) >= 0.999
        ctx.upload_volume(0, vol, desc=s.desc)


def test_sharded_mip_end_to_end_keeps_unsnapped_alpha(ctx):
    """MIP / MOP disable early ray termination (volumeraycaster.cpp:225-233): the sharded image must not snap alpha in [0.95, 1)."""
    import torch
    from voreen_b200 import sortlast
    dims, world = (64, 56, 72), 2
    vol = scenes.ct_like_volume(0, dims=dims)
    keys = [(0.0, (0, 0, 0, 0)), (0.25, (0, 0, 0, 0)), (0.30, (255, 140, 60, 200)), (0.60, (255, 240, 220, 246)), (1.00, (255, 255, 255, 250))]       # alpha 0.965 .. 0.98 above 0.6
    s = scenes.Scene(vol, keys=keys, W=96, H=88, eye=(1.4, 0.9, 2.9), out_format=abi.OUT_RGBA32F, compositing=[abi.COMP["mip"]] * 3)
    ref_outs, ref_depth = s.cuda_render(ctx)
    ref = np.asarray(ref_outs[0]).reshape(-1, 4)
    assert ((ref[:, 3] >= 0.95) & (ref[:, 3] < 1.0)).sum() > 20, "the scene must contain alphas the snap would change"
    cam_tex = sortlast.camera_texture_position(s.cam, s.desc)
    dev = torch.device("cuda", 0)
    npix = s.W * s.H
    s.params.partial_image = 1
    col = [torch.empty((npix, 4), dtype=torch.float32, device=dev) for _ in range(world)]
    key = [torch.empty((npix,), dtype=torch.float32, device=dev) for _ in range(world)]
    dep = [torch.empty((npix,), dtype=torch.float32, device=dev) for _ in range(world)]
    for r in range(world):
        _shard_upload(ctx, vol, dims, r, world)
        ctx.render_camera_ptrs(s.params, s.eep, s.cam, s.W, s.H, [col[r].data_ptr(), key[r].data_ptr(), 0], dep[r].data_ptr(), device=True)
    s.params.partial_image = 0
    f, b = (0, 1) if sortlast.my_group_in_front(0, 0, world, dims, cam_tex) else (1, 0)
    ctx.composite_keyed(col[f].data_ptr(), key[f].data_ptr(), dep[f].data_ptr(), col[b].data_ptr(), key[b].data_ptr(), dep[b].data_ptr(),
                        col[0].data_ptr(), key[0].data_ptr(), dep[0].data_ptr(), npix)
    out = torch.empty((npix, 4), dtype=torch.float32, device=dev)
    ctx.finalize(col[0].data_ptr(), 0, npix, abi.OUT_RGBA32F, out.data_ptr(), snap=False)        # what SortLast.render does for mip / mop
    ctx.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    ctx.finalize(col[0].data_ptr(), 0, npix, abi.OUT_RGBA32F, out.data_ptr(), snap=True)
    ctx.synchronize()
    assert not np.array_equal(out.cpu().numpy(), ref)
    ctx.upload_volume(0, vol, desc=s.desc)


def test_key_plane_rejected_for_multichannel_and_extra_outputs(ctx):
    import torch
    vol = np.stack([scenes.ct_like_volume(24), scenes.ct_like_volume(24, seed=7)], axis=3)
    s = scenes.Scene(vol, W=48, H=48, out_format=abi.OUT_RGBA32F, compositing=[abi.COMP["mip"]] * 3)
    s.upload(ctx)
    dev = torch.device("cuda", 0)
    col = torch.empty((48 * 48, 4), dtype=torch.float32, device=dev); key = torch.empty((48 * 48,), dtype=torch.float32, device=dev)
    dep = torch.empty((48 * 48,), dtype=torch.float32, device=dev)
    s.params.partial_image = 1
    with pytest.raises(abi.B200RCError):
        ctx.render_camera_ptrs(s.params, s.eep, s.cam, 48, 48, [col.data_ptr(), key.data_ptr(), 0], dep.data_ptr(), device=True)
    s1 = scenes.Scene(scenes.ct_like_volume(24), W=48, H=48, out_format=abi.OUT_RGBA32F, compositing=[abi.COMP["mip"]] * 3, output_connected=[1, 1, 0])
    s1.upload(ctx)
    s1.params.partial_image = 1
    with pytest.raises(abi.B200RCError):       # a partial image renders image.output only: the slot of output 1 is the key plane
        ctx.render_camera_ptrs(s1.params, s1.eep, s1.cam, 48, 48, [col.data_ptr(), key.data_ptr(), 0], dep.data_ptr(), device=True)


def test_unchanged_transfer_functions_are_not_rebuilt(ctx):
    """A drop-in processor hands its TFs over every frame: identical texels / tables must not bump the occupancy generation
    (no rebuild, no allocation), a changed texel must."""
    vol = scenes.ct_like_volume(32)
    s = scenes.Scene(vol, W=64, H=64).use_preint("cpu")
    a, ad = s.cuda_render(ctx)
    launches0 = ctx.stats().kernel_launches
    ctx.set_tf1d(0, scenes.host.tf1d_texture(s.keys_per_ch[0], s.tf_width), s.domain)
    ctx.set_preint(0, s.preint[0])
    assert ctx.stats().kernel_launches == launches0, "identical TF / table: nothing may be launched"
    b, bd = s.cuda_render(ctx, upload=False)
    assert np.array_equal(np.asarray(a[0]), np.asarray(b[0]))
    t = scenes.host.tf1d_texture(s.keys_per_ch[0], s.tf_width).copy()
    t[s.tf_width // 2, 3] ^= 0x40
    ctx.set_tf1d(0, t, s.domain)
    pre = s.preint[0].copy(); pre[250:900, 250:900, 3] = 0.9
    ctx.set_preint(0, pre)
    assert ctx.stats().kernel_launches > launches0
    c, cd = s.cuda_render(ctx, upload=False)
    assert not np.array_equal(np.asarray(a[0]), np.asarray(c[0]))


def test_device_preintegration_builders_against_the_reference_computeTable(ctx):
    """b200rc_build_preint (PI_CPU / PI_APPROX, preint.cu) against tables printed by the REFERENCE'S OWN PreIntegrationTable::computeTable
    (tests/golden/reference_pins.npz, generated from the reference's sources compiled where they lie)."""
    import json
    from voreen_b200 import host
    z = np.load(os.path.join(HERE, "golden", "reference_pins.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    tfs = {c["name"]: c for c in meta["tf_cases"]}
    done = 0
    for pc in meta["pi_cases"]:
        case = tfs[pc["tf"]]
        if case["threshold"] != [0.0, 1.0] or case["alpha_mode"] != 1:
            continue        # computeTable zeroes i >= back_start, the texture x > back_start: only the full threshold makes the TF texels the table's input
        keys = [(k[0], tuple(k[1])) if k[2] is None else (k[0], tuple(k[1]), tuple(k[2])) for k in case["keys"]]
        ctx.set_tf1d(0, host.tf1d_texture(keys, pc["res"], gamma=case["gamma"]), tuple(case["domain"]))
        ctx.build_preint(0, pc["res"], pc["step"], abi.PI_APPROX if pc["integral"] else abi.PI_CPU)
        got = ctx.download_preint(0, pc["res"])
        want = z[f"pi_{pc['index']}_table"]
        err = float(np.abs(got - want).max())
        print(f"device pre-integration table vs reference computeTable: tf {pc['tf']} res {pc['res']} step {pc['step']:.5f} integral {pc['integral']}: max |d| {err:.2e}")
        # powf / expf differ in the last bits between libm and the device; the integral variant divides by max(alpha, 0.001)
        assert err <= (1e-4 if pc["integral"] else 5e-6)
        assert np.array_equal(got[..., 3] > 0, want[..., 3] > 0), "alpha > 0 pattern (what empty-space skipping and the shaded-sample count depend on)"
        done += 1
    assert done >= 5


# ---------------------------------------------------------------------------------------------------------------------
# config 4: MultiVolumeRaycaster, sharded (sort-last), on one device shard by shard
# ---------------------------------------------------------------------------------------------------------------------
MV_KEYS = [scenes.CT_KEYS, [(0.0, (0, 0, 0, 0)), (0.3, (0, 255, 0, 0)), (1.0, (0, 255, 0, 120))],
           [(0.0, (0, 0, 0, 0)), (0.4, (0, 0, 255, 0)), (1.0, (0, 0, 255, 90))], [(0.0, (0, 0, 0, 0)), (0.5, (255, 0, 255, 0)), (1.0, (255, 0, 255, 60))]]


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("cls", ["tf", "preint"])
def test_sharded_multivolume_single_device(ctx, world, cls):
    """Four co-registered volumes bricked identically (config C4): every shard marches the lattice samples it owns -- all volumes'
    contributions at them, in the shader's order -- into a partial image; exact-ERT second pass; merge with the product's kernels.
    Against (a) the unsharded b200rc_render_camera_device frame, (b) that frame against the oracle's MultiVolumeRaycaster."""
    import torch
    from voreen_b200 import host, sortlast
    from voreen_b200 import raycaster as rc
    nvol, dims = 4, (48, 40, 56)
    vols = [scenes.ct_like_volume(0, dims=dims, seed=300 + v) for v in range(nvol)]
    desc = rc.make_volume_desc(dims)
    W, H = 96, 88
    cam = orc.make_camera((1.4, 0.9, 2.9), W=W, H=H)
    eep = rc.default_eep(world_coords=1)
    step_world = float(np.float32(2.0) * np.float32(max(d / max(dims) for d in dims)) / np.float32(max(dims)) / np.float32(2.0))
    step_tex = float(np.float32(1.0) / (np.float32(max(dims)) * np.float32(2.0)))
    tfs = [orc.make_tf(MV_KEYS[v], domain=(0.0, 65535.0)) for v in range(nvol)]
    tex = [orc.tf1d_texture(t, 4096) for t in tfs]
    pre = [orc.preint_cpu(t, 1024, step_tex) for t in tfs] if cls == "preint" else None
    p = rc.default_params(step_world, out_format=abi.OUT_RGBA32F)
    p.num_volumes = nvol; p.multi_volume = 1
    p.mv_opacity_correction_factor = step_tex / step_world
    if pre:
        for v in range(nvol):
            p.vol[v].classification = abi.CLASS_PREINT
    dev = torch.device("cuda", 0)
    npix = W * H

    def load(shard):
        for v in range(nvol):
            if shard is None:
                ctx.upload_volume(v, vols[v], desc=desc)
            else:
                lo, hi = sortlast.shard_range(dims, shard, world)
                rl = [max(0, lo[a] - sortlast.APRON) for a in range(3)]; rh = [min(dims[a], hi[a] + sortlast.APRON) for a in range(3)]
                ctx.upload_volume(v, np.ascontiguousarray(vols[v][rl[2]:rh[2], rl[1]:rh[1], rl[0]:rh[0]]),
                                  desc=rc.make_volume_desc(dims, shard_lo=lo, shard_hi=hi, shard_apron=sortlast.APRON))
            ctx.set_tf1d(v, tex[v], (0.0, 65535.0))
            if pre:
                ctx.set_preint(v, pre[v])

    def render_dev(col, dep):
        ctx.render_camera_ptrs(p, eep, cam, W, H, [col.data_ptr(), 0, 0], dep.data_ptr(), device=True)

    # unsharded frame (in-kernel world-space entry/exit shared by the four volumes)
    load(None)
    col0 = torch.empty((npix, 4), dtype=torch.float32, device=dev); dep0 = torch.empty((npix,), dtype=torch.float32, device=dev)
    p.collect_stats = 1
    render_dev(col0, dep0)
    ctx.synchronize()
    ref_st = ctx.stats()
    ref, ref_d = col0.cpu().numpy().reshape(H, W, 4), dep0.cpu().numpy().reshape(H, W)
    # ... which is the oracle's frame for the same entry/exit images
    e, ed, x, xd = ctx.entry_exit(eep, cam, W, H)
    ovs = [orc.make_volume(vols[v], desc) for v in range(nvol)]
    ocls = [orc.make_classifier(tex[v], (0.0, 65535.0), preint=pre[v] if pre else None) for v in range(nvol)]
    o_outs, o_depth, o_st = orc.raycast(ovs, ocls, p, cam, W, H, np.stack([e] * (nvol + 1)), ed, np.stack([x] * (nvol + 1)), xd)
    frac_o, mx_o = scenes.parity_report(o_outs[0], ref)
    assert frac_o >= 0.999 and ref_st.rays_cast == o_st.rays_cast and ref_st.samples_shaded == o_st.samples_shaded, (frac_o, mx_o)

    cam_tex = sortlast.camera_texture_position(cam, desc)
    p.partial_image = 1
    col = [torch.empty((npix, 4), dtype=torch.float32, device=dev) for _ in range(world)]
    dep = [torch.empty((npix,), dtype=torch.float32, device=dev) for _ in range(world)]
    shaded = 0
    for r in range(world):
        load(r)
        render_dev(col[r], dep[r])
        shaded += ctx.stats().samples_shaded
    alphas = torch.cat([c[:, 3].contiguous() for c in col]).contiguous()
    prefix = torch.empty((npix,), dtype=torch.float32, device=dev); flags = torch.empty((npix,), dtype=torch.uint8, device=dev)
    remarched = 0
    for r in range(world):
        load(r)
        render_dev(col[r], dep[r])      # restores the entry/exit images + first-pass result of this shard
        remarched += ctx.ert_plan(alphas.data_ptr(), world, sortlast.front_mask(r, world, dims, cam_tex), col[r].data_ptr(), dep[r].data_ptr(), npix,
                                  prefix.data_ptr(), flags.data_ptr(), count=True)
        ctx.remarch_device(p, cam, W, H, prefix.data_ptr(), flags.data_ptr(), [col[r].data_ptr(), 0, 0], dep[r].data_ptr())
    ctx.synchronize()
    p.partial_image = 0
    imgs = {r: (col[r].cpu().numpy().copy(), dep[r].cpu().numpy().copy()) for r in range(world)}
    for rnd in range(world.bit_length() - 1):
        nxt = {}
        for r in sorted(imgs):
            if (r >> rnd) & 1:
                continue
            a, b = imgs[r], imgs[r | (1 << rnd)]
            front, back = (a, b) if sortlast.my_group_in_front(r, rnd, world, dims, cam_tex) else (b, a)
            nxt[r] = scenes.product_composite(ctx, abi.COMP["dvr"] | abi.COMPOSITE_ERT, front, back)
        imgs = nxt
    merged = scenes.product_finalize(ctx, imgs[0][0], abi.OUT_RGBA32F, snap=True).reshape(H, W, 4)
    frac, mx = scenes.parity_report(ref, merged)
    dfrac = float((np.abs(ref_d.reshape(-1) - imgs[0][1]) <= 1e-4).mean())
    print(f"sharded multi-volume x{world} {cls}: within 1/255 {frac:.5f} max {mx:.4f} depth {dfrac:.5f} re-marched {remarched} shaded (first pass) {shaded}/{ref_st.samples_shaded}; "
          f"unsharded vs oracle {frac_o:.5f}")
    assert frac >= 0.999 and dfrac >= 0.999, (frac, mx)
    assert shaded >= ref_st.samples_shaded          # local early termination can only march further than the global one
    for v in range(1, nvol):
        ctx.lib.b200rc_release_volume(ctx.h, v)


# ---------------------------------------------------------------------------------------------------------------------
# image-space post chain: Compositor and Background against the oracle's restatement of the shaders
# ---------------------------------------------------------------------------------------------------------------------
def _render_pair(ctx, dtype):
    """Two real renderings (DVR and first-hit normals of the same scene) + random-alpha noise rows, as RGBA16 or RGBA32F."""
    vol = scenes.ct_like_volume(32)
    fmt = abi.OUT_RGBA16 if dtype == np.uint16 else abi.OUT_RGBA32F
    a = scenes.Scene(vol, W=72, H=56, out_format=fmt)
    b = scenes.Scene(vol, W=72, H=56, out_format=fmt, compositing=[abi.COMP["fhn"]] * 3, eye=(1.0, 1.3, 3.0))
    (ao, ad), (bo, bd) = a.cuda_render(ctx), b.cuda_render(ctx)
    c0, c1 = np.asarray(ao[0]).copy(), np.asarray(bo[0]).copy()
    rng = np.random.default_rng(3)
    if dtype == np.uint16:
        c0 = c0.view(np.uint16); c1 = c1.view(np.uint16)
        c0[:6] = rng.integers(0, 65536, c0[:6].shape); c1[:6] = rng.integers(0, 65536, c1[:6].shape)
        c0[6, :, 3] = 0; c1[7, :, 3] = 65535
    else:
        c0[:6] = rng.random(c0[:6].shape); c1[:6] = rng.random(c1[:6].shape)
        c0[6, :, 3] = 0; c1[7, :, 3] = 1
    ad = ad.copy(); bd = bd.copy()
    ad[:6] = rng.random(ad[:6].shape); bd[:6] = rng.random(bd[:6].shape); bd[3] = ad[3]
    return c0, ad, c1, bd


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_compositor_all_modes_against_the_oracle(ctx, dtype):
    c0, d0, c1, d1 = _render_pair(ctx, dtype)
    for mode in abi.COMPOSITOR:
        for kw in ([dict(weighting_factor=w) for w in (0.0, 0.3, 1.0)] if mode == "weighted-average" else
                   [dict(weight_first=0.7, weight_second=0.6, add_depth=ad_) for ad_ in (False, True)] if mode == "add" else [{}]):
            got, gd = ctx.image_compositor(mode, c0, d0, c1, d1, **kw)
            want, wd = orc.image_compositor(mode, c0, d0, c1, d1, **kw)
            assert np.array_equal(got, want, equal_nan=True) and np.array_equal(gd, wd), (mode, kw)
    # spot checks of the semantics themselves (compositor.frag:56-69): depth test picks the nearer fragment, alpha compositing is the over operator
    got, gd = ctx.image_compositor("depth-test", c0, d0, c1, d1)
    near0 = d0 < d1
    assert np.array_equal(got[near0], c0[near0]) and np.array_equal(got[~near0], c1[~near0]) and np.array_equal(gd, np.minimum(d0, d1))


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_background_against_the_oracle(ctx, dtype):
    c0, d0, _, _ = _render_pair(ctx, dtype)
    H, W = d0.shape
    for mode in ("monochrome", "gradient"):
        for blend in ("alpha-blending", "add"):
            for angle in (0, 60, 135, 359):
                kw = dict(mode=mode, blend=blend, color1=(1.0, 0.9, 0.2, 1.0), color2=(0.1, 0.392, 0.7, 0.8), angle=angle)
                got, gd = ctx.image_background(W, H, c0, d0, **kw)
                want, wd = orc.image_background(W, H, c0, d0, **kw)
                assert np.array_equal(got, want, equal_nan=True) and np.array_equal(gd, wd), kw
            got, gd = ctx.image_background(W, H, None, None, mode=mode, blend=blend, dtype=dtype)      # "image.input" unconnected: the background alone
            want, wd = orc.image_background(W, H, None, None, mode=mode, blend=blend, dtype=dtype)
            assert np.array_equal(got, want) and np.all(gd == 1.0)
    # semantics: an opaque pixel hides the background, an empty one shows it (background.frag:43-46)
    got, _ = ctx.image_background(W, H, c0, d0, mode="monochrome", color1=(0.25, 0.5, 0.75, 1.0))
    a = c0[..., 3].astype(np.float64) / (65535.0 if dtype == np.uint16 else 1.0)
    empty = a == 0
    bgc = np.array([0.25, 0.5, 0.75, 1.0])
    ref = (np.floor(bgc * 65535 + 0.5).astype(np.uint16) if dtype == np.uint16 else bgc.astype(np.float32))
    assert empty.any() and np.all(got[empty] == ref)


# ---------------------------------------------------------------------------------------------------------------------
# SegmentationRaycaster: rc_segmentation.frag
# ---------------------------------------------------------------------------------------------------------------------
def _segmentation_tf(segment_keys, width=256):
    """segmentationraycaster.cpp:381-415: every segment's 1D TF (width 256 here) resampled into rows 3 i .. 3 i + 2 of a 768-row texture;
    unset segments use the default TF."""
    tex = np.zeros((768, width, 4), np.uint8)
    dom = np.zeros((256, 2), np.float32)
    default = segment_keys[None]
    for i in range(256):
        keys, d = segment_keys.get(i, default)
        src = orc.tf1d_texture(orc.make_tf(keys, domain=d), 256)
        fac = np.float32(1.0) * np.float32(256) / np.float32(width)
        idx = np.minimum((fac * np.arange(width, dtype=np.float32)).astype(np.int32), 255)        # resampleTexture1D
        tex[3 * i: 3 * i + 3] = src[idx]
        dom[i] = d
    return tex, dom


@pytest.mark.parametrize("apply_seg", [0, 1])
@pytest.mark.parametrize("seg_dtype", [np.uint8, np.uint16])
def test_segmentation_raycaster(ctx, apply_seg, seg_dtype):
    from voreen_b200 import raycaster as rc
    n = 40
    vol = scenes.ct_like_volume(n)
    # segmentation: ids by octant-ish regions + a sphere, incl. id 0 and ids whose normalisation lands just below an integer
    z, y, x = np.meshgrid(*(np.arange(n),) * 3, indexing="ij")
    seg = ((x > n // 2).astype(np.int32) + 2 * (y > n // 3) + 4 * (z > n // 2)) % 5
    seg[(x - 20) ** 2 + (y - 18) ** 2 + (z - 22) ** 2 < 64] = 7
    seg = seg.astype(seg_dtype)
    s = scenes.Scene(vol, W=96, H=88, compositing=[abi.COMP["dvr"], abi.COMP["mip"], abi.COMP["fhn"]], output_connected=[1, 1, 1])
    s.params.raycaster = abi.RAYCASTER_SEGMENTATION
    s.params.apply_segmentation = apply_seg
    s.params.segmentation_bit_depth = 8 if seg_dtype == np.uint8 else 16
    s.params.empty_space_skipping = 1
    seg_keys = {None: (scenes.CT_KEYS, (0.0, 65535.0)),
                1: ([(0.0, (0, 0, 0, 0)), (0.3, (0, 255, 0, 0)), (1.0, (0, 255, 0, 160))], (0.0, 40000.0)),
                3: ([(0.0, (0, 0, 0, 0)), (0.2, (0, 0, 255, 30)), (1.0, (90, 0, 255, 255))], (10000.0, 65535.0)),
                7: ([(0.0, (255, 0, 0, 0)), (0.1, (255, 0, 0, 120)), (1.0, (255, 255, 0, 255))], (0.0, 65535.0))}
    tex, dom = _segmentation_tf(seg_keys)
    # 16-bit ids 1..7 normalise to ~1e-4: scale 65535 brings them back (rc_segmentation.frag:72-76)
    seg_desc = rc.make_volume_desc((n, n, n), dtype=abi.U8 if seg_dtype == np.uint8 else abi.U16)
    ctx.upload_volume(1, seg, desc=seg_desc)
    ctx.set_segmentation_tf(tex, dom)
    s.upload(ctx)
    s.params.collect_stats = 1
    c_outs, c_depth = ctx.render_camera(s.params, s.eep, s.cam, s.W, s.H)
    st = ctx.stats()
    # oracle
    ov = orc.make_volume(s.vol, s.desc); osg = orc.make_volume(seg, seg_desc)
    e, ed, x_, xd = s.oracle_eep()
    cls = [orc.make_classifier(s.tf_tex[0], s.domain), orc.make_classifier(tf2d=tex, seg_domains=dom)]
    po = s.params
    ess = po.empty_space_skipping; po.empty_space_skipping = 0
    o_outs, o_depth, o_st = orc.raycast([ov, osg], cls, po, s.cam, s.W, s.H, e, ed, x_, xd)
    po.empty_space_skipping = ess
    for o in range(3):
        frac, mx = scenes.parity_report(o_outs[o], c_outs[o])
        print(f"segmentation raycaster apply={apply_seg} {np.dtype(seg_dtype).name} out{o}: within 1/255 {frac:.5f} max {mx:.2e} shaded {st.samples_shaded}/{o_st.samples_shaded}")
        assert frac >= 0.999, (o, frac, mx)
    assert st.rays_cast == o_st.rays_cast and st.samples_shaded == o_st.samples_shaded and o_st.samples_shaded > 1000
    assert float((np.abs(o_depth - c_depth) <= 1e-4).mean()) >= 0.999
    if apply_seg:       # the per-segment TFs must actually differ from the plain classification
        s.params.apply_segmentation = 0
        plain, _ = ctx.render_camera(s.params, s.eep, s.cam, s.W, s.H)
        assert float(np.abs(np.asarray(plain[0]) - np.asarray(c_outs[0])).max()) > 0.05
    ctx.lib.b200rc_release_volume(ctx.h, 1)


# ---------------------------------------------------------------------------------------------------------------------
# TransFunc2DPrimitives end to end: primitives -> texture (host mirror) -> TFT_2DPRIMITIVES classification in the marcher
# ---------------------------------------------------------------------------------------------------------------------
def test_tf2d_primitives_end_to_end(ctx):
    """A 2D transfer function painted from a quad, a triangle and a banana (b200host_tf2d_primitives_texture), rendered with
    applyTF(tf, tex2D, intensity, |gradient|) (mod_transfunc.frag:80-95), against the oracle marching the oracle's own texture."""
    from voreen_b200 import host
    vol = scenes.ct_like_volume(48)
    s = scenes.Scene(vol, W=96, H=96, gradient=abi.GRAD["central-differences"], classification=abi.CLASS_TF2D)
    prims = [{"type": "quad", "points": [(0.30, 0.0), (0.62, 0.0), (0.62, 0.5), (0.30, 0.5)], "colors": (255, 160, 60, 40), "fuzziness": 0.6},
             {"type": "banana", "points": [(0.35, 0.30), (0.65, 0.95), (0.65, 0.55), (0.98, 0.30)], "colors": (80, 200, 255, 230), "fuzziness": 0.7},
             {"type": "triangle", "points": [(0.60, 0.0), (1.0, 0.0), (0.85, 0.35)], "colors": [(255, 255, 255, 255), (255, 255, 200, 255), (255, 200, 200, 120)], "fuzziness": 0.8}]
    f32, u8 = host.tf2d_primitives_texture(prims, 128, 64)
    want = orc.tf2d_primitives_texture(prims, 128, 64)
    assert np.array_equal(f32, want)
    o_u8 = np.clip(np.floor(want * np.float32(255.0) + np.float32(0.5)), 0, 255).astype(np.uint8)
    assert np.array_equal(u8, o_u8) and int((u8[..., 3] > 0).sum()) > 1000
    dom_x, dom_y = (0.0, s.desc.rwm_scale), (0.0, 0.6 * s.desc.rwm_scale)
    ov = orc.make_volume(s.vol, s.desc)
    e, ed, x, xd = s.oracle_eep()
    s.params.collect_stats = 1
    o_outs, o_depth, o_st = orc.raycast([ov], [orc.make_classifier(tf2d=o_u8, domain=dom_x, domain_y=dom_y)], s.params, s.cam, s.W, s.H, e, ed, x, xd)
    ctx.upload_volume(0, s.vol, desc=s.desc)
    ctx.set_tf2d(0, u8, dom_x, dom_y)
    c_outs, c_depth = ctx.render_camera(s.params, s.eep, s.cam, s.W, s.H)
    st = ctx.stats()
    frac, mx = scenes.parity_report(o_outs[0], c_outs[0])
    print(f"tf2d primitives: within-1/255 {frac:.5f} max|d| {mx:.2e} shaded {st.samples_shaded}/{o_st.samples_shaded}")
    assert o_st.samples_shaded > 1000 and st.rays_cast == o_st.rays_cast and frac >= 0.999
    assert abs(int(st.samples_shaded) - int(o_st.samples_shaded)) <= 0.002 * o_st.samples_shaded


# ---------------------------------------------------------------------------------------------------------------------
# config 4 on the staging marcher (march_mslab.cuh): against the per-ray multi-volume kernel and the oracle
# ---------------------------------------------------------------------------------------------------------------------
def _ctx_with_marcher(marcher):
    import voreen_b200 as vb
    old = os.environ.get("B200RC_MARCHER")
    if marcher == "ray":
        os.environ["B200RC_MARCHER"] = "ray"        # read once, when the context is created
    try:
        return vb.Context(0)
    finally:
        if old is None:
            os.environ.pop("B200RC_MARCHER", None)
        else:
            os.environ["B200RC_MARCHER"] = old


@pytest.mark.parametrize("cls", ["tf", "preint"])
@pytest.mark.parametrize("nvol,dims,img,comp", [(4, (48, 40, 56), (96, 88), "dvr"), (2, (33, 47, 29), (70, 61), "dvr"), (3, (64, 64, 64), (128, 128), "mip"),
                                                (4, (40, 40, 40), (64, 64), "fhn"), (1, (40, 48, 36), (80, 72), "dvr")])
def test_multivolume_staging_marcher(cls, nvol, dims, img, comp):
    """Co-registered volumes through b200rc_render_camera: the staging marcher (one shared-memory region per volume) against the
    per-ray kernel of the same library (B200RC_MARCHER=ray) -- identical sample decisions (rays, classified, shaded), images within
    the gradient's rounding -- against the oracle's MultiVolumeRaycaster, and with empty-space skipping off (bit-identical)."""
    import torch
    from voreen_b200 import raycaster as rc
    W, H = img
    vols = [scenes.ct_like_volume(0, dims=dims, seed=500 + v) for v in range(nvol)]
    desc = rc.make_volume_desc(dims)
    cam = orc.make_camera((1.3, 1.0, 2.8), W=W, H=H)
    eep = rc.default_eep(world_coords=1)
    step_world = float(np.float32(2.0) * np.float32(max(d / max(dims) for d in dims)) / np.float32(max(dims)) / np.float32(2.0))
    step_tex = float(np.float32(1.0) / (np.float32(max(dims)) * np.float32(2.0)))
    tfs = [orc.make_tf(MV_KEYS[v], domain=(0.0, 65535.0)) for v in range(nvol)]
    tex = [orc.tf1d_texture(t, 4096) for t in tfs]
    pre = [orc.preint_cpu(t, 1024, step_tex) for t in tfs] if cls == "preint" else None

    def params(ess):
        p = rc.default_params(step_world, out_format=abi.OUT_RGBA32F, empty_space_skipping=ess, compositing=[abi.COMP[comp]] * 3)
        p.num_volumes = nvol; p.multi_volume = 1
        p.mv_opacity_correction_factor = step_tex / step_world
        p.collect_stats = 1
        if pre:
            for v in range(nvol):
                p.vol[v].classification = abi.CLASS_PREINT
        return p

    def render(c, ess):
        for v in range(nvol):
            c.upload_volume(v, vols[v], desc=desc)
            c.set_tf1d(v, tex[v], (0.0, 65535.0))
            if pre:
                c.set_preint(v, pre[v])
        outs, depth = c.render_camera(params(ess), eep, cam, W, H)
        return np.asarray(outs[0]).copy(), np.asarray(depth).copy(), c.stats()

    slab, ray = _ctx_with_marcher("slab"), _ctx_with_marcher("ray")
    try:
        a, ad, sa_ = render(slab, 1)
        a0, ad0, sa0 = render(slab, 0)
        b, bd, sb = render(ray, 1)
        launches = slab.stats().kernel_launches
        assert np.array_equal(a, a0) and np.array_equal(ad, ad0), "empty-space skipping must not change a pixel"
        assert (sa_.rays_cast, sa_.samples_classified, sa_.samples_shaded, sa_.samples_nominal) == (sb.rays_cast, sb.samples_classified, sb.samples_shaded, sb.samples_nominal)
        assert sa0.samples_shaded == sa_.samples_shaded and sa0.samples_classified > sa_.samples_classified and sa_.samples_skipped > 0
        frac, mx = scenes.parity_report(b, a)
        dfrac = float((np.abs(ad - bd) <= 1e-6).mean())
        print(f"multi-volume staging marcher x{nvol} {dims} {cls} {comp}: vs per-ray within 1/255 {frac:.5f} max {mx:.2e} depth-equal {dfrac:.5f}; "
              f"classified {sa_.samples_classified} shaded {sa_.samples_shaded} skipped {sa_.samples_skipped} (per-ray {sb.samples_skipped}) launches {launches}")
        assert frac >= (0.999 if comp == "fhn" else 0.9999) and (comp != "dvr" or mx < 2e-2) and dfrac >= 0.9999
        # the oracle's MultiVolumeRaycaster on the same entry/exit images
        e, ed, x, xd = slab.entry_exit(eep, cam, W, H)
        ovs = [orc.make_volume(vols[v], desc) for v in range(nvol)]
        ocls = [orc.make_classifier(tex[v], (0.0, 65535.0), preint=pre[v] if pre else None) for v in range(nvol)]
        o_outs, o_depth, o_st = orc.raycast(ovs, ocls, params(0), cam, W, H, np.stack([e] * (nvol + 1)), ed, np.stack([x] * (nvol + 1)), xd)
        frac_o, mx_o = scenes.parity_report(o_outs[0], a.reshape(H, W, 4))
        assert frac_o >= 0.999 and o_st.samples_shaded == sa_.samples_shaded and o_st.rays_cast == sa_.rays_cast, (frac_o, mx_o)
    finally:
        slab.close(); ray.close()


# ---------------------------------------------------------------------------------------------------------------------
# exact-ERT plan without the all-gather: alpha planes by pointer; DVR partial images export their alpha plane
# ---------------------------------------------------------------------------------------------------------------------
def test_ert_plan_over_plane_pointers_equals_the_gathered_plan(ctx):
    import torch
    dev = torch.device("cuda", 0)
    world, n = 4, 70001
    g = torch.Generator(device="cpu").manual_seed(5)
    alphas = torch.rand((world, n), generator=g)
    alphas[:, ::7] = 0.0; alphas[1, ::5] = 0.97; alphas[2, 1::11] = 1.0
    alphas = alphas.to(dev).contiguous()
    color = torch.rand((n, 4), generator=g).to(dev); color[::3] = 0.0
    depth = torch.rand((n,), generator=g).to(dev)
    for me, mask in ((0, 0b1110), (3, 0b0011), (2, 0b0000)):
        color[:, 3] = alphas[me]
        outs = []
        for planes in (False, True):
            c, d = color.clone(), depth.clone()
            prefix = torch.empty((n,), dtype=torch.float32, device=dev); flags = torch.empty((n,), dtype=torch.uint8, device=dev)
            if planes:
                ptrs = [alphas[r].data_ptr() if (mask >> r) & 1 else 0 for r in range(world)]
                cnt = ctx.ert_plan_planes(ptrs, mask, c.data_ptr(), d.data_ptr(), n, prefix.data_ptr(), flags.data_ptr(), count=True)
            else:
                cnt = ctx.ert_plan(alphas.data_ptr(), world, mask, c.data_ptr(), d.data_ptr(), n, prefix.data_ptr(), flags.data_ptr(), count=True)
            ctx.synchronize()
            outs.append((cnt, c.cpu().numpy(), d.cpu().numpy(), prefix.cpu().numpy(), flags.cpu().numpy()))
        assert outs[0][0] == outs[1][0]
        for a, b in zip(outs[0][1:], outs[1][1:]):
            assert np.array_equal(a, b)
        if mask:
            assert outs[0][0] > 0 and int((outs[0][4] == 1).sum()) > 0
    with pytest.raises(abi.B200RCError):        # a rank in front without a plane
        ctx.ert_plan_planes([0, 0, 0, 0], 0b0010, color.data_ptr(), depth.data_ptr(), n, prefix.data_ptr(), flags.data_ptr())


@pytest.mark.parametrize("kind", ["slab", "ray", "multi"])
def test_dvr_partial_image_exports_its_alpha_plane(kind):
    import torch
    from voreen_b200 import raycaster as rc
    from voreen_b200 import sortlast
    dev = torch.device("cuda", 0)
    c = _ctx_with_marcher("ray" if kind == "ray" else "slab")
    try:
        dims, W, H = (48, 40, 56), 96, 88
        nvol = 2 if kind == "multi" else 1
        lo, hi = sortlast.shard_range(dims, 1, 2)
        rl = [max(0, lo[a] - sortlast.APRON) for a in range(3)]; rh = [min(dims[a], hi[a] + sortlast.APRON) for a in range(3)]
        step = float(np.float32(1.0) / (np.float32(max(dims)) * np.float32(2.0)))
        for v in range(nvol):
            vol = scenes.ct_like_volume(0, dims=dims, seed=900 + v)
            c.upload_volume(v, np.ascontiguousarray(vol[rl[2]:rh[2], rl[1]:rh[1], rl[0]:rh[0]]), desc=rc.make_volume_desc(dims, shard_lo=lo, shard_hi=hi, shard_apron=sortlast.APRON))
            c.set_tf1d(v, orc.tf1d_texture(orc.make_tf(MV_KEYS[v], domain=(0.0, 65535.0)), 4096), (0.0, 65535.0))
        p = rc.default_params(step if nvol == 1 else float(np.float32(2.0) * np.float32(max(d / max(dims) for d in dims)) / np.float32(max(dims)) / np.float32(2.0)), out_format=abi.OUT_RGBA32F)
        p.partial_image = 1
        if nvol > 1:
            p.num_volumes = nvol; p.multi_volume = 1; p.mv_opacity_correction_factor = 0.5
        eep = rc.default_eep(world_coords=1) if nvol > 1 else rc.default_eep()
        cam = orc.make_camera((1.4, 0.9, 2.9), W=W, H=H)
        col = torch.zeros((W * H, 4), dtype=torch.float32, device=dev); dep = torch.zeros((W * H,), dtype=torch.float32, device=dev)
        key = torch.full((W * H,), -1.0, dtype=torch.float32, device=dev)
        c.render_camera_ptrs(p, eep, cam, W, H, [col.data_ptr(), key.data_ptr(), 0], dep.data_ptr(), device=True)
        c.synchronize()
        a = col[:, 3].cpu().numpy(); k = key.cpu().numpy()
        assert np.array_equal(a, k) and float(a.max()) > 0.3
    finally:
        c.close()
