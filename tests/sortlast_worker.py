"""Worker of tests/test_sortlast.py: one gloo rank of the binary-swap host logic on CPU tensors.

Every rank fabricates the partial image of "its brick" (seeded by rank), the ranks merge them with
voreen_b200.sortlast.binary_swap, rank 0 collects the result and compares it with the sequential front-to-back
composite of all partial images in visibility order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voreen_b200 import sortlast  # noqa: E402


def partial_image(rank, npix):
    g = torch.Generator().manual_seed(1000 + rank)
    a = torch.rand((npix,), generator=g) * 0.7
    a[torch.rand((npix,), generator=g) < 0.3] = 0.0
    rgb = torch.rand((npix, 3), generator=g) * a[:, None]            # premultiplied
    depth = torch.where(a > 0, torch.rand((npix,), generator=g), torch.ones(npix))
    return torch.cat([rgb, a[:, None]], dim=1).contiguous(), depth.contiguous()


def over(fc, fd, bc, bd, dc, dd):
    k = 1.0 - fc[:, 3:4]
    c = fc + k * bc
    d = torch.minimum(fd, bd)
    dc.copy_(c); dd.copy_(d)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    npix = 64 * 48
    dims = (64, 48, 80)
    ok = True
    for cam_tex in [(-1.0, -1.0, -1.0), (2.0, 0.3, 2.0), (0.2, 3.0, -2.0), (0.7, 0.7, 0.7)]:
        color, depth = partial_image(rank, npix)
        lo, hi = sortlast.binary_swap(color, depth, rank, world, lambda r: sortlast.my_group_in_front(rank, r, world, dims, cam_tex), over, dist=dist)
        assert (lo, hi) == sortlast.final_ranges(npix, world)[rank]
        # gather the owned ranges on rank 0
        parts_c = [torch.zeros((hi - lo, 4)) for _ in range(world)] if rank == 0 else None
        parts_d = [torch.zeros((hi - lo,)) for _ in range(world)] if rank == 0 else None
        dist.gather(color[lo:hi].contiguous(), parts_c, dst=0)
        dist.gather(depth[lo:hi].contiguous(), parts_d, dst=0)
        if rank == 0:
            got_c = torch.zeros((npix, 4)); got_d = torch.zeros((npix,))
            for r, (a, b) in enumerate(sortlast.final_ranges(npix, world)):
                got_c[a:b] = parts_c[r]; got_d[a:b] = parts_d[r]
            # reference: sort all bricks front to back (per axis: the side of the split plane the camera is on comes first)
            _, axes = sortlast.shard_grid(world)

            def key(r):
                k = 0
                for bit, a in enumerate(axes):
                    low_front = cam_tex[a] < sortlast.split_plane(dims, a)
                    mine_low = ((r >> bit) & 1) == 0
                    k |= (0 if mine_low == low_front else 1) << bit
                return k
            # the over operator is associative: any order consistent with every plane works; merge like the swap tree does
            imgs = {r: partial_image(r, npix) for r in range(world)}
            order = sorted(range(world), key=lambda r: [((key(r) >> b) & 1) for b in reversed(range(len(axes)))])
            acc_c, acc_d = imgs[order[0]]
            acc_c, acc_d = acc_c.clone(), acc_d.clone()
            for r in order[1:]:
                over(acc_c, acc_d, imgs[r][0], imgs[r][1], acc_c, acc_d)
            err = float((got_c - acc_c).abs().max()); errd = float((got_d - acc_d).abs().max())
            if err > 1e-5 or errd > 0:
                print(f"MISMATCH cam {cam_tex}: colour {err} depth {errd}")
                ok = False
    if rank == 0:
        print("SORTLAST_OK" if ok else "SORTLAST_FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
