#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the sort-last image (sharded march + exact-ERT second
pass + binary swap over NCCL) against the single-GPU image of the same volume rendered on rank 0.
Bar: <= 1/255 per channel on >= 99.9 % of pixels, depth within 1e-4.  Prints one line per view; exit code 1 on failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # tests/ -> repo root
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definition: TF keys, orbit)
import voreen_b200 as vb  # noqa: E402
from voreen_b200 import abi, host, sortlast  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    img = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    cls = sys.argv[3] if len(sys.argv) > 3 else "preint"
    mode = sys.argv[4] if len(sys.argv) > 4 else "sort-last"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = vb.Context(local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    dims = (size,) * 3
    step = float(np.float32(1.0) / (np.float32(size) * np.float32(2.0)))
    params = vb.default_params(step, classification=abi.CLASS_PREINT if cls == "preint" else abi.CLASS_TF, out_format=abi.OUT_RGBA16)
    # SL_MULTI=n: MultiVolumeRaycaster over n co-registered volumes bricked identically (config C4): world-space entry/exit shared by
    # the volumes, world-space sampling step, per-volume transfer functions, mvOpacityCorrection
    nvol = int(os.environ.get("SL_MULTI", "1"))
    if nvol > 1:
        step_world = float(np.float32(2.0) / np.float32(size) / np.float32(2.0))
        params = vb.default_params(step_world, classification=abi.CLASS_PREINT if cls == "preint" else abi.CLASS_TF, out_format=abi.OUT_RGBA16)
        params.num_volumes = nvol; params.multi_volume = 1
        params.mv_opacity_correction_factor = step / step_world
        for v in range(nvol):
            params.vol[v].classification = abi.CLASS_PREINT if cls == "preint" else abi.CLASS_TF
    if os.environ.get("SL_COMP"):         # e.g. iso / fhn / fhp: merged by "front-most hit wins"
        params.compositing[:] = [abi.COMP[os.environ["SL_COMP"]]] * 3
    eep = vb.default_eep(world_coords=1) if nvol > 1 else vb.default_eep()
    views = [0, 5, 9, 14, 22, 31]
    cams = {k: host.make_camera(bench.orbit_eye(k), W=img, H=img) for k in views}

    mv_keys = [bench.CT_KEYS, [(0.0, (0, 0, 0, 0)), (0.3, (0, 255, 0, 0)), (1.0, (0, 255, 0, 120))],
               [(0.0, (0, 0, 0, 0)), (0.4, (0, 0, 255, 0)), (1.0, (0, 0, 255, 90))], [(0.0, (0, 0, 0, 0)), (0.5, (255, 0, 255, 0)), (1.0, (255, 0, 255, 60))]]

    def load_tf():
        for v in range(nvol):
            if cls == "preint":
                ctx.set_tf1d(v, host.tf1d_texture(mv_keys[v], 1024), (0.0, 65535.0))
                ctx.build_preint(v, 1024, step, abi.PI_CPU)
            ctx.set_tf1d(v, host.tf1d_texture(mv_keys[v], 4096), (0.0, 65535.0))

    def synth(desc):
        for v in range(nvol):
            ctx.synth_volume(v, desc, 1234 + 77 * v)

    out = torch.empty((img, img, 4), dtype=torch.int16, device=dev)
    depth = torch.empty((img, img), dtype=torch.float32, device=dev)
    ref = {}
    if rank == 0 and not os.environ.get("SL_NOREF"):
        synth(vb.make_volume_desc(dims))
        load_tf()
        for k in views:
            ctx.render_camera_ptrs(params, eep, cams[k], img, img, [out.data_ptr(), 0, 0], depth.data_ptr(), device=True)
            ctx.synchronize()
            ref[k] = (out.cpu().numpy().view(np.uint16).astype(np.float64) / 65535.0, depth.cpu().numpy().copy())
    dist.barrier()
    if mode == "sort-first":
        sl = sortlast.SortFirst(ctx, rank, world, img, img, device=dev, stream=stream, transport=os.environ.get("SL_TRANSPORT", "auto"))
        sl.phases = {}
        if rank != 0 or os.environ.get("SL_NOREF"):
            synth(vb.make_volume_desc(dims))
            load_tf()
    else:
        sl = sortlast.SortLast(ctx, dims, rank, world, img, img, device=dev, stream=stream, transport=os.environ.get("SL_TRANSPORT", "auto"))
        synth(sl.shard_desc())
        load_tf()
        sl.profile = True
    ok = True
    for k in views:
        ms = sl.render(params, eep, cams[k], out, depth)          # first call of a view: warm-up
        ms = sl.render(params, eep, cams[k], out, depth)
        torch.cuda.synchronize(dev)
        # whole-frame time (march + merge + gather), max over ranks, 5 frames back to back
        prof_saved = getattr(sl, "profile", False)
        sl.profile = False
        dist.barrier(); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            sl.render(params, eep, cams[k], out, depth)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        tt = torch.tensor([e0.elapsed_time(e1) / 5.0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        frame_ms = float(tt.item())
        sl.profile = prof_saved
        ms = sl.render(params, eep, cams[k], out, depth)
        torch.cuda.synchronize(dev)
        ph = [None] * world
        dist.all_gather_object(ph, {a: round(float(b), 2) for a, b in sl.phases.items()})
        if rank == 0 and os.environ.get("SL_PHASES"):
            for r_, d_ in enumerate(ph):
                print(f"   view {k} rank {r_}: {d_}", flush=True)
        if rank == 0 and ref:
            got = out.cpu().numpy().view(np.uint16).astype(np.float64) / 65535.0
            d = np.abs(got - ref[k][0]).max(axis=2)
            frac = float((d <= 1.0 / 255.0).mean())
            dfrac = float((np.abs(depth.cpu().numpy() - ref[k][1]) <= 1e-4).mean())
            good = frac >= 0.999 and dfrac >= 0.999
            ok = ok and good
            print(f"{mode} x{world} {nvol} x {size}^3 @{img}^2 {cls} view {k}: within-1/255 {frac:.6f} max|d| {d.max():.4f} depth-ok {dfrac:.6f} "
                  f"rank0 march {ms:.2f} ms frame {frame_ms:.2f} ms transport {sl.transport} {'OK' if good else 'FAIL'}", flush=True)
        elif rank == 0:     # SL_NOREF: timing only
            print(f"{mode} x{world} {nvol} x {size}^3 @{img}^2 {cls} view {k}: rank0 march {ms:.2f} ms frame {frame_ms:.2f} ms transport {sl.transport} phases {ph[0]}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
