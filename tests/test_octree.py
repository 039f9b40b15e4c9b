"""Octree / multi-resolution brick pool.

CPU: the numpy restatement of VolumeOctree (oracle/octree_ref.py) against the reference's own known answers
(V/apps/tests/octreetest/octreetest.cpp:1041-2112): level counts, actual tree depth, node / brick counts of complete and optimised
trees, level textures against the mip map, min / max values, for 128^3 and the non-power-of-two 125 x 96 x 115 and 35 x 96 x 115 cases.
GPU: the product's construction kernels node by node and brick by brick against that restatement, and its DVR traversal against the
oracle's restatement of singleoctreeraycaster.cl.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import octree_ref as oref  # noqa: E402

import orc  # noqa: E402
import scenes  # noqa: E402
from voreen_b200 import abi  # noqa: E402


def node_count(levels):            # getNodeCount (octreetest.cpp:103-110): complete tree
    return sum(8 ** l for l in range(levels))


def check_against_mip_map(tree, vol):
    mm = oref.mip_map(vol, tree.num_levels)
    for l in range(tree.num_levels):
        assert np.array_equal(tree.level_volume(l), mm[l]), f"level {l}"


# ---- reference known answers, scaled to 64^3 where the test only states structure (the 128^3 numbers are run once below) --------------
@pytest.mark.parametrize("brick,levels", [(64, 1), (32, 2), (16, 3), (8, 4), (4, 5)])
def test_complete_tree_known_answers(brick, levels):
    """OctreeConstruction_CompleteTree_* (octreetest.cpp:1041-1280): homogeneity threshold -1 => complete tree, every node a brick."""
    rng = np.random.default_rng(brick)
    noise = rng.integers(0, 65536, (64, 64, 64)).astype(np.uint16)
    bricked = oref.bricked_volume((64, 64, 64), (32, 32, 32))
    for vol in (noise, bricked):
        t = oref.Octree(vol, brick, -1.0)
        assert t.octree_dim == 64 and t.num_levels == levels and t.root.depth() == levels
        assert t.root.node_count() == node_count(levels) and t.root.num_bricks() == node_count(levels)
        check_against_mip_map(t, vol)
        assert (t.root.mn, t.root.mx) == (int(vol.min()), int(vol.max()))            # checkMinMaxValues


def test_optimized_tree_known_answers_128():
    """OctreeConstruction_OptimizedTree_SingleBuffer_Dim128_Brick64 / _Brick16 (octreetest.cpp:1413-1482), at the reference's own size."""
    noise = oref.noise_volume((32, 32, 32))            # noise: no homogeneous regions => complete tree (any size shows it)
    t = oref.Octree(noise, 8, 0.0)
    assert t.num_levels == 3 and t.root.depth() == 3 and t.root.node_count() == node_count(3)
    # completely homogeneous volume => only the root node
    hom = oref.bricked_volume((128, 128, 128), (128, 128, 128))
    t = oref.Octree(hom, 64, 0.0)
    assert t.num_levels == 2 and t.root.depth() == 1 and t.root.node_count() == 1
    check_against_mip_map(t, hom)
    # homogeneous regions 4 times larger than the brick size => tree depth reduced by 2
    b64 = oref.bricked_volume((128, 128, 128), (64, 64, 64))
    t = oref.Octree(b64, 16, 0.0)
    assert t.num_levels == 4 and t.root.depth() == 2 and t.root.node_count() == node_count(2)
    check_against_mip_map(t, b64)
    assert (t.root.mn, t.root.mx) == (int(b64.min()), int(b64.max()))


def test_nonzero_homogeneity_threshold_known_answers():
    """OctreeConstruction_OptimizedTree_NonZeroHomogeneityThreshold_Dim128 (octreetest.cpp:1625-1653): 64-bricked volume + regular
    noise of range 20 (relative inhomogeneity 20 / 65535 ~ 0.0003): threshold 0.00025 keeps the full tree, 0.00035 cuts a level."""
    base = oref.bricked_volume((128, 128, 128), (64, 64, 64)).astype(np.int64)
    flat = base.reshape(-1).copy()
    even = np.arange(flat.size) % 2 == 0                              # addRegularNoise (octreetest.cpp:296-316)
    flat[even] = np.maximum(flat[even] - 10, 0)
    flat[~even] = np.where(flat[~even] <= 65535 - 10, flat[~even] + 10, 65535)
    noisy = flat.reshape(base.shape).astype(np.uint16)
    full = oref.Octree(noisy, 32, 0.00025)
    assert full.num_levels == 3 and full.root.depth() == 3 and full.root.node_count() == node_count(3)
    check_against_mip_map(full, noisy)
    opt = oref.Octree(noisy, 32, 0.00035)
    assert opt.num_levels == 3 and opt.root.depth() == 2 and opt.root.node_count() == node_count(2)


@pytest.mark.parametrize("dims,brick,levels", [((125, 96, 115), 128, 1), ((125, 96, 115), 16, 4), ((35, 96, 115), 16, 4)])
def test_npot_known_answers(dims, brick, levels):
    """OctreeConstruction_NPOT_* (octreetest.cpp:1846-2112): octree dimension 128, level count, depth, level textures == mip map."""
    rng = np.random.default_rng(7)
    vol = rng.integers(0, 65536, dims[::-1]).astype(np.uint16)
    for thr in (-1.0, 0.0):
        t = oref.Octree(vol, brick, thr)
        assert t.octree_dim == 128 and t.num_levels == levels and t.root.depth() == levels
        check_against_mip_map(t, vol)


def test_invalid_brick_sizes_are_rejected():
    """OctreeConstruction_InvalidBrickSize (octreetest.cpp:1015-1025)."""
    vol = np.zeros((16, 16, 16), np.uint16)
    for b in (0, 1, 7, 10, 100, 256):
        with pytest.raises(AssertionError):
            oref.Octree(vol, b)


# ---- the product -----------------------------------------------------------------------------------------------------------------------
def _compare_with_oracle(ctx, vol, brick, thr, fn="mean"):
    from voreen_b200 import raycaster as rc
    dims = vol.shape[::-1]
    desc = rc.make_volume_desc(dims, dtype=abi.U16 if vol.dtype == np.uint16 else abi.U8)
    ctx.upload_volume(0, vol, desc=desc)
    info = ctx.octree_build(0, brick, thr, fn)
    t = oref.Octree(vol, brick, thr, fn)
    assert (info.octree_dim, info.num_levels, info.actual_depth) == (t.octree_dim, t.num_levels, t.root.depth())
    assert (info.num_nodes, info.num_bricks) == (t.root.node_count(), t.root.num_bricks())
    dense = t.dense()
    for k in range(t.num_levels):
        avg, mn, mx, fl, br = ctx.octree_level(0, k, brick)
        want = dense[k]
        exists = (want["flags"] & 1) != 0
        assert np.array_equal((fl & 1) != 0, exists), f"level {k}: node existence"
        assert np.array_equal(fl[exists], want["flags"][exists]), f"level {k}: flags"
        for name, got in (("avg", avg), ("mn", mn), ("mx", mx)):
            assert np.array_equal(got[exists], want[name][exists]), f"level {k}: {name}"
        G = 1 << k
        for (x, y, z), b in want["bricks"].items():
            assert np.array_equal(br[(z * G + y) * G + x], b), f"level {k}: brick {(x, y, z)}"
    return info, t


@pytest.mark.gpu
@pytest.mark.parametrize("dims,brick,thr", [((64, 64, 64), 16, -1.0), ((64, 64, 64), 8, 0.0), ((61, 48, 57), 8, -1.0), ((61, 48, 57), 16, 0.001), ((35, 48, 57), 8, 0.0),
                                            ((32, 32, 32), 2, -1.0)])
def test_octree_build_matches_the_restatement(ctx, dims, brick, thr):
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 65536, dims[::-1]).astype(np.uint16)
    info, t = _compare_with_oracle(ctx, noise, brick, thr)
    # a volume with homogeneous regions (and a little noise in one of them): the optimisation prunes
    vol = np.full(dims[::-1], 12345, np.uint16)
    vol[: dims[2] // 2, :, : dims[0] // 2] = 40000
    vol[dims[2] // 2:, dims[1] // 2:, :] += rng.integers(0, 30, vol[dims[2] // 2:, dims[1] // 2:, :].shape).astype(np.uint16)
    info, t = _compare_with_oracle(ctx, vol, brick, thr)
    if thr >= 0:
        assert info.num_nodes < node_count(info.num_levels)
    print(f"octree {dims} brick {brick} thr {thr}: levels {info.num_levels} depth {info.actual_depth} nodes {info.num_nodes} bricks {info.num_bricks} build {info.build_ms:.3f} ms")


@pytest.mark.gpu
@pytest.mark.parametrize("fn", ["max", "min"])
def test_octree_halfsample_functions_and_uint8(ctx, fn):
    rng = np.random.default_rng(4)
    _compare_with_oracle(ctx, rng.integers(0, 65536, (40, 48, 44)).astype(np.uint16), 8, -1.0, fn)
    _compare_with_oracle(ctx, rng.integers(0, 256, (40, 48, 44)).astype(np.uint8), 8, 0.0, "mean")


@pytest.mark.gpu
@pytest.mark.parametrize("thr", [-1.0, 0.002])
@pytest.mark.parametrize("linear", [True, False])
def test_octree_dvr_traversal_against_the_oracle(ctx, thr, linear):
    """SingleOctreeRaycasterCL DVR (singleoctreeraycaster.cl:190-300): every level of detail, brick filter linear / nearest."""
    from voreen_b200 import host
    from voreen_b200 import raycaster as rc
    dims = (61, 48, 57)
    vol = scenes.ct_like_volume(0, dims=dims)
    desc = rc.make_volume_desc(dims)
    ctx.upload_volume(0, vol, desc=desc)
    brick = 8
    info = ctx.octree_build(0, brick, thr)
    t = oref.Octree(vol, brick, thr)
    avg, mn, mx, flags, pool = oref.pool_arrays(t)
    tex = host.tf1d_texture(scenes.CT_KEYS, 4096)
    ctx.set_tf1d(0, tex, (0.0, 65535.0))
    W, H = 96, 80
    cam = orc.make_camera((1.3, 1.0, 2.8), W=W, H=H)
    eep = rc.default_eep()
    e, ed, x, xd = ctx.entry_exit(eep, cam, W, H)
    step = 1.0 / (float(sum(dims) // 3) * 1.0)                    # avgDim * samplingRate (singleoctreeraycastercl.cpp:731-734)
    for lod in range(info.num_levels):
        got, gfh = ctx.octree_render(0, step, lod, eep, cam, W, H, linear=linear)
        want, wfh = orc.octree_raycast(avg, flags, pool, t.num_levels, brick, dims, t.octree_dim, tex, (0.0, 65535.0), (desc.rwm_scale, desc.rwm_offset),
                                       step, lod, linear, e, x, W, H)
        frac, mxd = scenes.parity_report(want, got)
        print(f"octree DVR thr {thr} linear {linear} level {lod}: within 1/255 {frac:.5f} max {mxd:.2e} covered {(want[..., 3] > 0).mean():.3f}")
        assert frac >= 0.999, (lod, frac, mxd)
        assert float((np.abs(gfh - wfh) <= 1e-5).mean()) >= 0.999
        assert (want[..., 3] > 0).mean() > 0.1
    # coarser levels really are coarser: the full-resolution image differs from the root-level one
    a, _ = ctx.octree_render(0, step, info.num_levels - 1, eep, cam, W, H, linear=linear)
    b, _ = ctx.octree_render(0, step, 0, eep, cam, W, H, linear=linear)
    assert float(np.abs(a - b).max()) > 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("clip", [False, True])
@pytest.mark.parametrize("dims,brick,thr", [((64, 64, 64), 8, 0.0), ((61, 48, 57), 8, 0.002), ((40, 40, 40), 4, -1.0)])
def test_volumeoctree_proxy_geometry(ctx, dims, brick, thr, clip):
    """OptimizedProxyGeometry "volumeoctree" (optimizedproxygeometry.cpp:1620-1737): the host mirror over the library's octree against
    the restatement over the pointer tree -- same cubes, in the reference's traversal order; then the entry/exit images they give."""
    import ctypes as C
    from voreen_b200 import host
    from voreen_b200 import raycaster as rc
    vol = scenes.ct_like_volume(0, dims=dims)
    vol[:, :, : dims[0] // 4] = 0                     # a transparent slab: homogeneous, pruned by the optimisation when thr >= 0
    desc = rc.make_volume_desc(dims)
    keys = [(0.0, (0, 0, 0, 0)), (0.3, (0, 0, 0, 0)), (0.35, (255, 120, 40, 3)), (0.7, (255, 255, 255, 120)), (1.0, (255, 255, 255, 255))]
    dom, thr_tf, gamma, amode = (0.0, 65535.0), (0.0, 1.0), 1.0, 1
    eep = rc.default_eep(proxy_mode=abi.PROXY_VOLUME_OCTREE, proxy_threshold=1.0)
    if clip:
        eep.enable_clipping = 1
        eep.clip_llf[:] = [3.0, 0.0, 5.0]; eep.clip_urb[:] = [30.0, 39.0, 33.0]
    ctx.upload_volume(0, vol, desc=desc)
    with pytest.raises(abi.B200RCError):              # no octree yet (a new volume drops the octree of the previous one)
        host.optimized_proxy_geometry(ctx, 0, desc, eep, [(keys, dom, thr_tf, gamma, amode)])
    info = ctx.octree_build(0, brick, thr, "mean")
    got = host.optimized_proxy_geometry(ctx, 0, desc, eep, [(keys, dom, thr_tf, gamma, amode)])
    tree = oref.Octree(vol, brick, thr, "mean")
    tf = orc.make_tf(keys, domain=dom, threshold=thr_tf, gamma=gamma, alpha_mode=amode)
    table = orc.preint_cpu(tf, 256, 1.0)              # getPreIntegrationTable(1.f, 256) (:1637)
    cl = np.zeros(4, np.float32)

    def region_empty(mn, mx):                           # :1668-1676 + isRegionEmptyPi (:878-880)
        f = []
        for v in (mn, mx):
            rw = np.float32(np.float32(v) / np.float32(65535.0)) * np.float32(desc.rwm_scale) + np.float32(desc.rwm_offset)
            f.append(0.0 if rw < dom[0] else (1.0 if rw > dom[1] else float(np.float32(rw - np.float32(dom[0])) / np.float32(dom[1] - dom[0]))))
        orc.lib().orc_preint_classify(orc.fp(table), 256, f[0], f[1], orc.fp(cl))
        return bool(cl[3] <= np.float32(0.0001) * np.float32(eep.proxy_threshold))
    want = oref.proxy_boxes(tree, region_empty, clip=(list(eep.clip_llf), list(eep.clip_urb)) if clip else None)
    print(f"volumeoctree proxy geometry {dims} brick {brick} thr {thr} clip={clip}: {len(got)} cubes (restatement {len(want)}), tree nodes {info.num_nodes}")
    assert len(got) == len(want) and len(got) >= 2
    assert np.array_equal(got, want)
    if thr >= 0:                                        # leaves of different sizes: the pruned slab is one big (transparent) leaf, opaque leaves differ in extent
        ext = np.round((got[:, 3] - got[:, 0]) * dims[0])
        assert len(np.unique(ext)) >= 1
    cam = orc.make_camera((1.4, 0.8, 2.9), W=96, H=96)
    eep2 = rc.default_eep(proxy_mode=eep.proxy_mode, enable_clipping=eep.enable_clipping, clip_llf=list(eep.clip_llf), clip_urb=list(eep.clip_urb), boxes=got)
    e, ed, x, xd = ctx.entry_exit(eep2, cam, 96, 96)
    oe, oed, ox, oxd = orc.entry_exit(want, orc.texture_to_world(desc), eep, cam, 96, 96)
    assert np.array_equal(e, oe) and np.array_equal(x, ox) and np.array_equal(ed, oed) and np.array_equal(xd, oxd)
    assert int((e[..., 3] > 0).sum()) > 500


@pytest.mark.gpu
@pytest.mark.parametrize("comp", ["mip", "mop"])
@pytest.mark.parametrize("thr", [-1.0, 0.002])
def test_octree_mip_mop_against_the_oracle(ctx, comp, thr):
    """SingleOctreeRaycasterCL compositing modes MIP / MOP (octree_mip.cl, octree_mop.cl): the transfer function applied to the
    maximum intensity, resp. the sample of maximum opacity; first-hit output = the ray parameter where that maximum was first raised."""
    from voreen_b200 import host
    from voreen_b200 import raycaster as rc
    dims = (61, 48, 57)
    vol = scenes.ct_like_volume(0, dims=dims)
    desc = rc.make_volume_desc(dims)
    ctx.upload_volume(0, vol, desc=desc)
    brick = 8
    info = ctx.octree_build(0, brick, thr)
    t = oref.Octree(vol, brick, thr)
    avg, mn, mx, flags, pool = oref.pool_arrays(t)
    tex = host.tf1d_texture(scenes.CT_KEYS, 4096)
    ctx.set_tf1d(0, tex, (0.0, 65535.0))
    W, H = 96, 80
    cam = orc.make_camera((1.3, 1.0, 2.8), W=W, H=H)
    eep = rc.default_eep()
    e, ed, x, xd = ctx.entry_exit(eep, cam, W, H)
    step = 1.0 / float(sum(dims) // 3)
    dvr, _ = ctx.octree_render(0, step, info.num_levels - 1, eep, cam, W, H)
    for lod in (info.num_levels - 1, 1):
        for linear in (True, False):
            got, gfh = ctx.octree_render(0, step, lod, eep, cam, W, H, linear=linear, compositing=comp)
            want, wfh = orc.octree_raycast(avg, flags, pool, t.num_levels, brick, dims, t.octree_dim, tex, (0.0, 65535.0), (desc.rwm_scale, desc.rwm_offset),
                                           step, lod, linear, e, x, W, H, compositing=abi.COMP[comp])
            frac, mxd = scenes.parity_report(want, got)
            print(f"octree {comp} thr {thr} linear {linear} level {lod}: within 1/255 {frac:.5f} max {mxd:.2e} covered {(want[..., 3] > 0).mean():.3f}")
            assert frac >= 0.999, (lod, frac, mxd)
            assert float((np.abs(gfh - wfh) <= 1e-5).mean()) >= 0.999
            assert (want[..., 3] > 0).mean() > 0.1 and float(want.max()) <= 1.0
    full, _ = ctx.octree_render(0, step, info.num_levels - 1, eep, cam, W, H, compositing=comp)
    assert float(np.abs(full - dvr).max()) > 0.05         # a different compositing mode, not DVR again
    with pytest.raises(abi.B200RCError):
        ctx.octree_render(0, step, 0, eep, cam, W, H, compositing="iso")
