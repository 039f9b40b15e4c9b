"""Sort-last host logic on CPU: brick partition, visibility order, binary swap over gloo (world size 2 and 4)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from voreen_b200 import sortlast

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_partition(world):
    for dims in [(64, 64, 64), (41, 37, 29), (2048, 2048, 2048), (100, 40, 24)]:
        owned = np.zeros(dims[::-1], np.int32) if max(dims) <= 128 else None
        vox = 0
        for r in range(world):
            lo, hi = sortlast.shard_range(dims, r, world)
            assert all(0 <= lo[a] < hi[a] <= dims[a] for a in range(3))
            assert all(lo[a] % 8 == 0 for a in range(3))        # brick-aligned split planes
            vox += (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2])
            if owned is not None:
                owned[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]] += 1
        assert vox == dims[0] * dims[1] * dims[2]
        if owned is not None:
            assert owned.min() == 1 and owned.max() == 1        # every voxel's base index is owned exactly once
    assert len(set(sortlast.final_ranges(4096, world))) == world
    covered = sorted(sortlast.final_ranges(4096, world))
    assert covered[0][0] == 0 and covered[-1][1] == 4096 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_visibility_order():
    dims = (64, 64, 64)
    # 2 ranks split along y at texture coordinate (32 + 0.5) / 64
    assert sortlast.my_group_in_front(0, 0, 2, dims, (0.5, -1.0, 0.5))
    assert not sortlast.my_group_in_front(1, 0, 2, dims, (0.5, -1.0, 0.5))
    assert sortlast.my_group_in_front(1, 0, 2, dims, (0.5, 2.0, 0.5))
    # 8 ranks: bit 0 <-> y, bit 1 <-> z, bit 2 <-> x
    assert sortlast.shard_coords(5, 8) == (1, 1, 0)
    assert sortlast.my_group_in_front(5, 0, 8, dims, (0.0, 2.0, 0.0)) and not sortlast.my_group_in_front(5, 1, 8, dims, (0.0, 0.0, 2.0))


@pytest.mark.parametrize("world", [2, 4])
def test_binary_swap_gloo(world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + world * 7 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "sortlast_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert "SORTLAST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_binary_swap_peer_memory_schedule(world):
    """The peer-memory variant (no send/recv: every round blends the partner's copy of the kept range in place).  All ranks
    are driven round by round in one process over ordinary tensors standing in for the peer-mapped buffers; the merged image
    must equal the sequential front-to-back composite, and no round may read a range another rank writes in that round."""
    import torch
    sys.path.insert(0, HERE)
    from sortlast_worker import over, partial_image
    npix, dims = 64 * 48 + 24, (64, 48, 80)
    _, axes = sortlast.shard_grid(world)
    for cam_tex in [(-1.0, -1.0, -1.0), (2.0, 0.3, 2.0), (0.2, 3.0, -2.0)]:
        bufs = {r: tuple(t.clone() for t in partial_image(r, npix)) for r in range(world)}
        for rnd in range(world.bit_length() - 1):
            written = {}
            for rank in range(world):
                sched = sortlast.swap_schedule(npix, rank, world)[rnd]
                written[rank] = sched[1]
            for rank in range(world):
                partner, keep = sortlast.swap_schedule(npix, rank, world)[rnd]
                pk = written[partner]
                assert keep[1] <= pk[0] or pk[1] <= keep[0]           # what I read of the partner is not what it writes
                lo, hi = sortlast.binary_swap_p2p_round(rnd, lambda r: bufs[r], rank, world,
                                                        lambda r_, rank=rank: sortlast.my_group_in_front(rank, r_, world, dims, cam_tex), over)
                assert (lo, hi) == keep
        got_c = torch.zeros((npix, 4)); got_d = torch.zeros((npix,))
        for r, (a, b) in enumerate(sortlast.final_ranges(npix, world)):
            assert sortlast.swap_schedule(npix, r, world)[-1][1] == (a, b)
            got_c[a:b] = bufs[r][0][a:b]; got_d[a:b] = bufs[r][1][a:b]

        def key(r):
            return [0 if (((r >> bit) & 1) == 0) == (cam_tex[a] < sortlast.split_plane(dims, a)) else 1 for bit, a in reversed(list(enumerate(axes)))]
        order = sorted(range(world), key=key)
        acc_c, acc_d = (t.clone() for t in partial_image(order[0], npix))
        for r in order[1:]:
            img = partial_image(r, npix)
            over(acc_c, acc_d, img[0], img[1], acc_c, acc_d)
        assert float((got_c - acc_c).abs().max()) <= 1e-5 and float((got_d - acc_d).abs().max()) == 0.0
