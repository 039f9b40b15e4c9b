"""Sort-last multi-GPU rendering: one process per GPU, each holding one brick (shard) of the volume.

New work with no counterpart in the reference.  Per frame every rank
  1. marches exactly the lattice samples its shard owns (b200rc_params.partial_image, volume.cu shard_lo/hi: the
     shards' sample sets partition the single-GPU sample set, so merging partial images reproduces the single-GPU
     accumulators up to the ERT caveat below) into a premultiplied RGBA32F + depth partial image of the full viewport;
  2. merges the partial images by BINARY SWAP: log2(P) rounds, in round r partner = rank xor 2^r, the current pixel
     range is halved, one half is sent and the partner's matching half received and blended with the over operator
     in visibility order (the two groups of bricks are separated by one axis-aligned plane, so the order is the same
     for every pixel: the group on the camera's side of the plane is in front);
  3. finalises its 1/P of the image (ERT alpha snap + RGBA16 conversion) and sends it to rank 0.

The exchange uses torch.distributed point-to-point operations (NCCL over NVLink on GPUs, gloo in the CPU tests); the
blend and finalise steps are the CUDA kernels behind b200rc_composite / b200rc_finalize (or an injected function in the
host-logic tests).  No collective touches the volume: bricks never move.

Early ray termination: in the first pass each rank terminates a ray on its LOCAL alpha >= 0.95 (safe: the global alpha is
then >= 0.95 as well).  That alone differs from the single-GPU image by up to (1 - 0.95) * colour where the global 0.95
crossing lies inside a brick that has something in front of it.  The second pass (exact_ert, default on) removes the
difference: an all-gather of the alpha planes (4 B/pixel/rank) gives every rank the alpha in front of its brick; pixels
behind a terminated ray are erased, pixels whose crossing lies inside the brick are re-marched with the accumulator started
at that prefix, so mod_raysetup.frag:65-74 fires on the global alpha exactly where the single-GPU march stops.
"""
from __future__ import annotations

import math

import numpy as np

from . import abi
from . import raycaster as rc

APRON = 8       # voxels; a multiple of 8 keeps the brick grid of the shard aligned with the global one


def shard_grid(world):
    """(gx, gy, gz) bricks per axis and the axis each rank bit splits (bit r <-> axes[r]).

    y is split first: for turntable views (the reference's own benchmark rotates the camera about +y,
    V/resource/voreencore/scripts/benchmark.py:30-47) the two y halves always lie side by side on screen, so early ray
    termination cannot starve one of them; z and x follow."""
    if world & (world - 1) or world < 1 or world > 8:
        raise ValueError("sort-last needs 1, 2, 4 or 8 ranks")
    nbits = world.bit_length() - 1
    axes = {0: [], 1: [1], 2: [1, 2], 3: [1, 2, 0]}[nbits]
    grid = [1, 1, 1]
    for a in axes:
        grid[a] *= 2
    return tuple(grid), axes


def shard_coords(rank, world):
    grid, axes = shard_grid(world)
    c = [0, 0, 0]
    for r, a in enumerate(axes):
        c[a] = (rank >> r) & 1
    return tuple(c)


def shard_range(dims, rank, world):
    """Owned voxel range [lo, hi) of a rank: the volume is halved along each split axis at a multiple of 8 voxels."""
    grid, _ = shard_grid(world)
    c = shard_coords(rank, world)
    lo, hi = [], []
    for a in range(3):
        if grid[a] == 1:
            lo.append(0); hi.append(dims[a])
        else:
            mid = (dims[a] // 2 + 7) // 8 * 8
            if mid >= dims[a]:      # the upper half would be empty (and an all-zero shard range means "unsharded" to the library)
                raise ValueError(f"axis {a} with {dims[a]} voxels is too small to split at a multiple of 8 voxels")
            lo.append(0 if c[a] == 0 else mid); hi.append(mid if c[a] == 0 else dims[a])
    return tuple(lo), tuple(hi)


def shard_desc(dims, rank, world, **desc_kw):
    lo, hi = shard_range(dims, rank, world)
    return rc.make_volume_desc(dims, shard_lo=lo, shard_hi=hi, shard_apron=APRON if world > 1 else 0, **desc_kw) if world > 1 else rc.make_volume_desc(dims, **desc_kw)


def camera_texture_position(cam, desc):
    """Camera position in the volume's texture space (worldToTexture * position)."""
    d = np.array(desc.dims[:], np.float64); sp = np.array(desc.spacing[:], np.float64); off = np.array(desc.offset[:], np.float64)
    p2w = np.array(desc.phys_to_world[:], np.float64).reshape(4, 4)
    w = np.array(list(cam.position[:]) + [1.0])
    phys = np.linalg.solve(p2w, w)[:3]
    vox = (phys - off) / sp
    return (vox + 0.5) / d


def split_plane(dims, axis):
    """Texture coordinate of the plane separating the two halves along `axis`: base indices < mid | >= mid."""
    mid = (dims[axis] // 2 + 7) // 8 * 8
    return (mid + 0.5) / dims[axis]


def my_group_in_front(rank, rnd, world, dims, cam_tex):
    """Round `rnd` merges the groups {bit rnd == 0} and {bit rnd == 1}, separated by the split plane of axes[rnd]."""
    _, axes = shard_grid(world)
    a = axes[rnd]
    low_side_front = cam_tex[a] < split_plane(dims, a)
    mine_low = ((rank >> rnd) & 1) == 0
    return mine_low == low_side_front


def front_mask(rank, world, dims, cam_tex):
    """Bit f set: brick f lies in front of brick `rank` along every ray that crosses both, i.e. f differs from `rank` and
    sits on the camera's side of every split plane in which they differ (rays are monotone in each axis)."""
    _, axes = shard_grid(world)
    mine = shard_coords(rank, world)
    mask = 0
    for f in range(world):
        if f == rank:
            continue
        c = shard_coords(f, world)
        ok = True
        for a in set(axes):
            if c[a] != mine[a]:
                cam_low = cam_tex[a] < split_plane(dims, a)
                if (c[a] == 0) != cam_low:
                    ok = False
        if ok:
            mask |= 1 << f
    return mask


def final_ranges(npix, world):
    """Pixel range [lo, hi) every rank owns after the last round (deterministic, known to all ranks)."""
    out = []
    nbits = world.bit_length() - 1
    for rank in range(world):
        lo, hi = 0, npix
        for r in range(nbits):
            mid = (lo + hi) // 2
            if (rank >> r) & 1 == 0:
                hi = mid
            else:
                lo = mid
        out.append((lo, hi))
    return out


def binary_swap(color, depth, rank, world, in_front, blend, dist=None, group=None):
    """In-place binary swap over `color` (npix, 4) and `depth` (npix,) torch tensors of this rank's partial image.

    in_front(rnd) -> bool: is this rank's group in front of the partner's in round rnd.
    blend(front_c, front_d, back_c, back_d, dst_c, dst_d): dst = front OVER back (may alias front or back).
    Returns (lo, hi): the pixel range whose merged result this rank now holds in color[lo:hi], depth[lo:hi]."""
    import torch
    if dist is None:
        import torch.distributed as dist
    npix = color.shape[0]
    lo, hi = 0, npix
    nbits = world.bit_length() - 1
    for r in range(nbits):
        partner = rank ^ (1 << r)
        mid = (lo + hi) // 2
        keep = (lo, mid) if (rank >> r) & 1 == 0 else (mid, hi)
        send = (mid, hi) if (rank >> r) & 1 == 0 else (lo, mid)
        n = keep[1] - keep[0]
        rc_ = torch.empty((n, 4), dtype=color.dtype, device=color.device)
        rd_ = torch.empty((n,), dtype=depth.dtype, device=depth.device)
        ops = [dist.P2POp(dist.isend, color[send[0]:send[1]], partner, group), dist.P2POp(dist.isend, depth[send[0]:send[1]], partner, group),
               dist.P2POp(dist.irecv, rc_, partner, group), dist.P2POp(dist.irecv, rd_, partner, group)]
        if rank > partner:      # both sides post in the same global order (send/recv pairs match up)
            ops = ops[2:] + ops[:2]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        mc, md = color[keep[0]:keep[1]], depth[keep[0]:keep[1]]
        if in_front(r):
            blend(mc, md, rc_, rd_, mc, md)
        else:
            blend(rc_, rd_, mc, md, mc, md)
        lo, hi = keep
    return lo, hi


def swap_schedule(npix, rank, world):
    """[(partner, (keep_lo, keep_hi))] per round of the binary swap for `rank` (the range halves every round)."""
    out = []
    lo, hi = 0, npix
    for r in range(world.bit_length() - 1):
        mid = (lo + hi) // 2
        lo, hi = (lo, mid) if (rank >> r) & 1 == 0 else (mid, hi)
        out.append((rank ^ (1 << r), (lo, hi)))
    return out


def binary_swap_p2p_round(rnd, views, rank, world, in_front, blend):
    """One round of the binary swap over PEER MEMORY: nothing is sent.  views(r) -> (color, depth) tensors that map rank r's
    partial image (peer-mapped over NVLink for r != rank).  The blend kernel reads the partner's copy of the range this
    rank keeps straight from the partner's HBM and blends it into the local copy in place -- transfer and over operator are
    one kernel.  The caller synchronises the ranks before every round (the partner's range must be final) and after the
    last one.  No hazard inside a round: a rank writes only the range it keeps, its partner reads only the other half."""
    npix = views(rank)[0].shape[0]
    partner, (a, b) = swap_schedule(npix, rank, world)[rnd]
    mc, md = views(rank)
    pc, pd = views(partner)
    mc, md, pc, pd = mc[a:b], md[a:b], pc[a:b], pd[a:b]
    if in_front(rnd):
        blend(mc, md, pc, pd, mc, md)
    else:
        blend(pc, pd, mc, md, mc, md)
    return a, b


def binary_swap_p2p(views, rank, world, in_front, blend, barrier):
    lo, hi = 0, views(rank)[0].shape[0]
    for rnd in range(world.bit_length() - 1):
        barrier()
        lo, hi = binary_swap_p2p_round(rnd, views, rank, world, in_front, blend)
    return lo, hi


class PeerMemory:
    """A buffer of `nbytes` on every rank, each mapped into every other rank's address space (torch symmetric memory:
    cuMem allocations exchanged at rendezvous, NVLink/NVSwitch peer access).  ptr(r) is the device address of rank r's
    buffer as seen from this process; kernels launched here may load from and store to it.  barrier() is a device-side
    barrier enqueued on the current stream (signal pads in peer memory; no host round trip, no NCCL kernel)."""

    def __init__(self, nbytes, device):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.torch = torch
        self.nbytes = int(nbytes)
        self.local = symm.empty((self.nbytes,), dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.local, dist.group.WORLD)
        self.world, self.rank = self.hdl.world_size, self.hdl.rank
        self.ptrs = [int(x) for x in self.hdl.buffer_ptrs]

    def ptr(self, r, offset=0):
        return self.ptrs[r] + int(offset)

    def view(self, r, offset, shape, dtype):
        """Tensor view of rank r's buffer at byte `offset`."""
        esz = self.torch.empty((), dtype=dtype).element_size()
        if r == self.rank:
            n = int(np.prod(shape))
            return self.local[offset:offset + n * esz].view(dtype).view(*shape)
        return self.hdl.get_buffer(r, tuple(shape), dtype, offset // esz)

    def barrier(self):
        self.hdl.barrier(channel=0)


class SortLast:
    """One rank of the sort-last renderer (GPU)."""

    def __init__(self, ctx, dims, rank, world, W, H, device, stream, transport="auto"):
        import torch
        self.torch = torch
        self.ctx, self.dims, self.rank, self.world, self.W, self.H = ctx, tuple(dims), rank, world, W, H
        self.device, self.stream = device, stream
        self.npix = W * H
        # transport "p2p": the partial image lives in peer-mapped memory, every binary-swap round is ONE kernel that reads the
        # partner's half over NVLink and blends it in place, the finalise kernel stores each rank's 1/P straight into rank 0's
        # output image.  transport "nccl": send/recv of the halves + a separate blend kernel (also what the gloo CPU tests run).
        self.transport = "nccl"
        self.peer = None
        if world > 1 and transport in ("auto", "p2p"):
            try:
                # [colour npix x 16 B | depth npix x 4 B | final RGBA16 npix x 8 B | final depth npix x 4 B | MIP/MOP key npix x 4 B]
                self.peer = PeerMemory(self.npix * 36, device)
                self.transport = "p2p"
            except Exception as e:      # symmetric memory unavailable on this box: NCCL point-to-point still works
                if transport == "p2p":
                    raise
                import sys
                print(f"[sortlast] peer memory unavailable ({type(e).__name__}: {e}); using NCCL send/recv", file=sys.stderr)
        if self.peer is not None:
            self.color = self.peer.view(rank, 0, (self.npix, 4), torch.float32)
            self.depth = self.peer.view(rank, self.npix * 16, (self.npix,), torch.float32)
            self._views = {r: (self.peer.view(r, 0, (self.npix, 4), torch.float32), self.peer.view(r, self.npix * 16, (self.npix,), torch.float32))
                           for r in range(world)}
            # max-key merges (MIP / MOP) swap on the key plane; the depth plane rides along at a fixed offset inside every rank's buffer
            self.key = self.peer.view(rank, self.npix * 32, (self.npix,), torch.float32)
            self._key_views = {r: (self._views[r][0], self.peer.view(r, self.npix * 32, (self.npix,), torch.float32)) for r in range(world)}
        else:
            self.color = torch.empty((self.npix, 4), dtype=torch.float32, device=device)
            self.depth = torch.empty((self.npix,), dtype=torch.float32, device=device)
        self.ranges = final_ranges(self.npix, world)
        self.desc = shard_desc(self.dims, rank, world)
        self.exact_ert = True       # second pass: early ray termination on the global alpha (images match the single-GPU result)
        import os as _os
        self.peer_plan = _os.environ.get("B200RC_SORTLAST_PLAN", "planes") != "gather"      # A/B: alpha planes read from peer memory vs NCCL all-gather
        self.alpha = torch.empty((self.npix,), dtype=torch.float32, device=device)
        self.alphas = torch.empty((world * self.npix,), dtype=torch.float32, device=device)
        self.prefix = torch.empty((self.npix,), dtype=torch.float32, device=device)
        self.flags = torch.empty((self.npix,), dtype=torch.uint8, device=device)
        self.last_remarch_ms = 0.0
        self._merge_mode = abi.COMP["dvr"] | abi.COMPOSITE_ERT
        self.profile = False        # True: per-phase CUDA-event timings + flagged-pixel count of the last frame in self.phases
        self.phases = {}

    def shard_desc(self):
        return self.desc

    def _blend_keyed(self, fc, fk, bc, bk, dc, dk):
        # fk / bk / dk are slices of key planes; the matching depth slices sit (16 n - 32 n) bytes before them in the same buffers
        off = self.npix * 16 - self.npix * 32
        self.ctx.composite_keyed(fc.data_ptr(), fk.data_ptr(), fk.data_ptr() + off, bc.data_ptr(), bk.data_ptr(), bk.data_ptr() + off,
                                 dc.data_ptr(), dk.data_ptr(), dk.data_ptr() + off, fc.shape[0])

    def _blend(self, fc, fd, bc, bd, dc, dd):
        self.ctx.composite(self._merge_mode, fc.data_ptr(), fd.data_ptr(), bc.data_ptr(), bd.data_ptr(), dc.data_ptr(), dd.data_ptr(), fc.shape[0])

    def render(self, params, eep, cam, out_rgba16, out_depth):
        """Renders one frame; rank 0 receives the final RGBA16 image + depth in out_rgba16 / out_depth (device tensors).
        Returns this rank's march-kernel time in ms."""
        torch = self.torch
        import torch.distributed as dist
        p = params
        # DVR: over operator + exact early ray termination (second pass).  ISO / FHP / FHN: the front-most partial image that has a hit
        # wins (the first sample along the ray that satisfies the predicate lies in the front-most shard that has one) -- exact without
        # a second pass.  MIP / MOP / MIDA on a sharded volume are not supported (the marcher does not export the key plane).
        comp = int(p.compositing[0])
        if comp == abi.COMP["dvr"]:
            self._merge_mode, need_ert_pass = abi.COMP["dvr"] | abi.COMPOSITE_ERT, True
        elif comp in (abi.COMP["iso"], abi.COMP["fhp"], abi.COMP["fhn"]):
            self._merge_mode, need_ert_pass = comp, False
        elif comp in (abi.COMP["mip"], abi.COMP["mop"]) and self.peer is not None:
            self._merge_mode, need_ert_pass = comp, False      # max-key merge over peer memory (key plane exported by the marcher)
        else:
            raise abi.B200RCError("sort-last rendering supports dvr, iso, fhp, fhn (and mip, mop over peer memory); use the image partition (SortFirst) otherwise")
        keyed = comp in (abi.COMP["mip"], abi.COMP["mop"])
        if keyed and self.desc.channels > 1:
            # multi-channel MIP composites every channel on its own and adds the results (rc_singlevolume.frag:272-309): one key per
            # pixel cannot pick the summed colour shard by shard
            raise abi.B200RCError("sort-last mip / mop needs a single-channel volume; use the image partition (SortFirst) otherwise")
        # only image.output is rendered: the partial image buffers hold one colour plane (+ the key plane in the slot of output 1)
        saved = (p.partial_image, p.out_format, tuple(p.output_connected))
        p.partial_image, p.out_format = 1, abi.OUT_RGBA32F
        p.output_connected[0], p.output_connected[1], p.output_connected[2] = 1, 0, 0
        snap = comp not in (abi.COMP["mip"], abi.COMP["mop"], abi.COMP["mida"])     # USE_EARLY_TERMINATION (volumeraycaster.cpp:225-233)
        prof = self.profile
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if prof else None
        with torch.cuda.stream(self.stream):
            if prof:
                ev[0].record(self.stream)
            # peer transport + DVR: the marcher also writes the partial image's alpha plane (4 B / pixel) into the key plane in peer memory;
            # the ranks behind this one read it from there for the exact-ERT plan -- no strided copy, no all-gather
            alpha_planes = self.peer is not None and self.exact_ert and need_ert_pass and self.world > 1 and self.peer_plan and not int(p.raycaster)
            self.ctx.render_camera_ptrs(p, eep, cam, self.W, self.H, [self.color.data_ptr(), self.key.data_ptr() if (keyed or alpha_planes) else 0, 0],
                                        self.depth.data_ptr(), device=True)
            march_ms = self.ctx.stats().last_march_ms
            if prof:
                self.phases = {"first_pass_ms": march_ms}
                ev[1].record(self.stream)
            cam_tex = camera_texture_position(cam, self.desc)
            if self.exact_ert and need_ert_pass and self.world > 1:
                # every rank learns the alpha of every partial image, erases what a terminated ray hides and re-marches the
                # pixels whose ray crosses the global 0.95 threshold inside its own shard
                fm = front_mask(self.rank, self.world, self.dims, cam_tex)
                if alpha_planes:
                    self.peer.barrier()         # every rank's first pass (and with it its alpha plane) is complete
                    planes = [self.peer.ptr(r, self.npix * 32) if (fm >> r) & 1 else 0 for r in range(self.world)]
                    n_flag = self.ctx.ert_plan_planes(planes, fm, self.color.data_ptr(), self.depth.data_ptr(), self.npix, self.prefix.data_ptr(),
                                                      self.flags.data_ptr(), count=prof)
                else:
                    self.alpha.copy_(self.color[:, 3])
                    dist.all_gather_into_tensor(self.alphas, self.alpha)
                    n_flag = self.ctx.ert_plan(self.alphas.data_ptr(), self.world, fm, self.color.data_ptr(), self.depth.data_ptr(), self.npix,
                                               self.prefix.data_ptr(), self.flags.data_ptr(), count=prof)
                if prof:
                    self.phases["flagged_pixels"] = n_flag
                    ev[2].record(self.stream)
                self.ctx.remarch_device(p, cam, self.W, self.H, self.prefix.data_ptr(), self.flags.data_ptr(), [self.color.data_ptr(), 0, 0], self.depth.data_ptr())
                self.last_remarch_ms = self.ctx.stats().last_march_ms
                march_ms += self.last_remarch_ms
                if prof:
                    self.phases["remarch_ms"] = self.last_remarch_ms
            if prof:
                ev[3].record(self.stream)
            p.partial_image, p.out_format = saved[0], saved[1]
            for o in range(3):
                p.output_connected[o] = saved[2][o]
            flat16 = out_rgba16.view(-1, 4)
            flatd = out_depth.view(-1)
            if self.peer is not None:
                in_front = lambda r: my_group_in_front(self.rank, r, self.world, self.dims, cam_tex)     # noqa: E731
                if keyed:
                    lo, hi = binary_swap_p2p(lambda r: self._key_views[r], self.rank, self.world, in_front, self._blend_keyed, self.peer.barrier)
                else:
                    lo, hi = binary_swap_p2p(lambda r: self._views[r], self.rank, self.world, in_front, self._blend, self.peer.barrier)
                n = hi - lo
                # finalise my 1/P (alpha snap + RGBA16) with the stores going to rank 0's image in peer memory
                o16, od = self.npix * 20, self.npix * 28
                self.ctx.finalize(self.color[lo:hi].data_ptr(), self.depth[lo:hi].data_ptr(), n, abi.OUT_RGBA16, self.peer.ptr(0, o16 + lo * 8), snap=snap)
                self.peer.view(0, od, (self.npix,), torch.float32)[lo:hi].copy_(self.depth[lo:hi])
                self.peer.barrier()         # every 1/P has landed; every read of this frame's partial images is done
                if self.rank == 0:
                    flat16.copy_(self.peer.view(0, o16, (self.npix, 4), torch.int16))
                    flatd.copy_(self.peer.view(0, od, (self.npix,), torch.float32))
            else:
                lo, hi = binary_swap(self.color, self.depth, self.rank, self.world,
                                     lambda r: my_group_in_front(self.rank, r, self.world, self.dims, cam_tex), self._blend)
                # finalise my 1/P of the image (alpha snap + RGBA16) and hand it to rank 0
                n = hi - lo
                if self.rank == 0:
                    self.ctx.finalize(self.color[lo:hi].data_ptr(), self.depth[lo:hi].data_ptr(), n, abi.OUT_RGBA16, flat16[lo:hi].data_ptr(), snap=snap)
                    flatd[lo:hi].copy_(self.depth[lo:hi])
                    ops = []
                    for r in range(1, self.world):
                        a, b = self.ranges[r]
                        ops.append(dist.P2POp(dist.irecv, flat16[a:b].view(torch.uint8), r))       # NCCL has no int16: ship bytes
                        ops.append(dist.P2POp(dist.irecv, flatd[a:b], r))
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
                else:
                    part16 = torch.empty((n, 4), dtype=torch.int16, device=self.device)
                    self.ctx.finalize(self.color[lo:hi].data_ptr(), self.depth[lo:hi].data_ptr(), n, abi.OUT_RGBA16, part16.data_ptr(), snap=snap)
                    ops = [dist.P2POp(dist.isend, part16.view(torch.uint8), 0), dist.P2POp(dist.isend, self.depth[lo:hi], 0)]
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
            if prof:
                ev[4].record(self.stream)
                torch.cuda.synchronize(self.device)
                self.phases["render_ms"] = ev[0].elapsed_time(ev[1])
                if self.exact_ert and need_ert_pass and self.world > 1:
                    self.phases["allgather_plan_ms"] = ev[1].elapsed_time(ev[2])
                    self.phases["remarch_total_ms"] = ev[2].elapsed_time(ev[3])
                self.phases["swap_gather_ms"] = ev[3].elapsed_time(ev[4])
                self.phases["frame_ms"] = ev[0].elapsed_time(ev[4])
        return march_ms


# ---------------------------------------------------------------------------------------------------------------------
# sort-first: the volume is REPLICATED (it fits in one GPU's 180 GB), the image is partitioned
# ---------------------------------------------------------------------------------------------------------------------
BAND_ROWS = 16      # b200rc_params.image_band_modulus works on 16-pixel-row bands


def band_owner_rows(H, rank, world):
    """Pixel rows rank `rank` renders: bands b = row // 16 with b % world == rank."""
    rows = np.arange(H)
    return rows[(rows // BAND_ROWS) % world == rank]


class SortFirst:
    """One rank of the image-space partition: every rank holds the whole volume and renders the 16-row bands
    b % world == rank (interleaved, so the load is balanced whatever the camera sees); rank 0 gathers the bands.
    No partial images, no over-compositing, no occlusion lost between ranks: the pixels are exactly the single-GPU ones."""

    def __init__(self, ctx, rank, world, W, H, device, stream, transport="auto"):
        import torch
        self.torch = torch
        self.ctx, self.rank, self.world, self.W, self.H = ctx, rank, world, W, H
        self.device, self.stream = device, stream
        # transport "p2p": the marcher of every rank stores its bands straight into rank 0's image through peer-mapped memory
        # (the RGBA16 / depth stores of the ray-marching kernel ARE the transfer; one device-side barrier ends the frame).
        # transport "nccl": render locally, pack the bands, dist.gather, unpack on rank 0.
        self.transport, self.peer, self.frame_no = "nccl", None, 0
        self.npix = W * H
        if world > 1 and transport in ("auto", "p2p"):
            try:
                self.peer = PeerMemory(2 * self.npix * 12, device)      # two frames of [RGBA16 | depth] (double-buffered)
                self.transport = "p2p"
            except Exception as e:
                if transport == "p2p":
                    raise
                import sys
                print(f"[sortfirst] peer memory unavailable ({type(e).__name__}: {e}); using NCCL gather", file=sys.stderr)
        if self.peer is not None:
            return
        if H % (BAND_ROWS * world):
            raise ValueError(f"image height {H} must be a multiple of {BAND_ROWS} x {world}")
        self.groups = H // (BAND_ROWS * world)
        self.full16 = torch.empty((H, W, 4), dtype=torch.int16, device=device)
        self.fulld = torch.empty((H, W), dtype=torch.float32, device=device)
        # my bands packed: colour (8 B/pixel) followed by depth (4 B/pixel), as bytes (NCCL has no int16)
        self.npix_mine = self.groups * BAND_ROWS * W
        self.pack = torch.empty((self.npix_mine * 12,), dtype=torch.uint8, device=device)
        self.gathered = torch.empty((world, self.npix_mine * 12), dtype=torch.uint8, device=device) if rank == 0 else None

    def render(self, params, eep, cam, out_rgba16, out_depth):
        """Rank 0 receives the complete RGBA16 image + depth in out_rgba16 (H, W, 4) int16 / out_depth (H, W) float32."""
        torch = self.torch
        import torch.distributed as dist
        p = params
        saved = (p.image_band_modulus, p.image_band_remainder)
        p.image_band_modulus, p.image_band_remainder = self.world, self.rank
        if self.peer is not None:
            # Frame k lands in half k & 1 of rank 0's buffer.  A rank can run at most one frame ahead of rank 0 (the barrier),
            # and rank 0 enqueues its copy-out of frame k before it enters the barrier of frame k + 1, so the writes of frame
            # k + 2 into the same half never overtake that copy.
            base = (self.frame_no & 1) * self.npix * 12
            self.frame_no += 1
            with torch.cuda.stream(self.stream):
                self.ctx.render_camera_ptrs(p, eep, cam, self.W, self.H, [self.peer.ptr(0, base), 0, 0], self.peer.ptr(0, base + self.npix * 8), device=True)
                march_ms = self.ctx.stats().last_march_ms
                p.image_band_modulus, p.image_band_remainder = saved
                self.peer.barrier()
                if self.rank == 0:
                    out_rgba16.view(-1, 4).copy_(self.peer.view(0, base, (self.npix, 4), torch.int16))
                    out_depth.view(-1).copy_(self.peer.view(0, base + self.npix * 8, (self.npix,), torch.float32))
            return march_ms
        G, B, W, N = self.groups, BAND_ROWS, self.W, self.world
        with torch.cuda.stream(self.stream):
            self.ctx.render_camera_ptrs(p, eep, cam, self.W, self.H, [self.full16.data_ptr(), 0, 0], self.fulld.data_ptr(), device=True)
            march_ms = self.ctx.stats().last_march_ms
            p.image_band_modulus, p.image_band_remainder = saved
            mine16 = self.full16.view(G, N, B, W, 4)[:, self.rank]
            mined = self.fulld.view(G, N, B, W)[:, self.rank]
            nc = self.npix_mine * 8
            self.pack[:nc].view(torch.int16).view(G, B, W, 4).copy_(mine16)
            self.pack[nc:].view(torch.float32).view(G, B, W).copy_(mined)
            dist.gather(self.pack, list(self.gathered.unbind(0)) if self.rank == 0 else None, dst=0)
            if self.rank == 0:
                g = self.gathered
                out_rgba16.view(G, N, B, W, 4).copy_(g[:, :nc].view(torch.int16).view(N, G, B, W, 4).permute(1, 0, 2, 3, 4))
                out_depth.view(G, N, B, W).copy_(g[:, nc:].view(torch.float32).view(N, G, B, W).permute(1, 0, 2, 3))
        return march_ms
