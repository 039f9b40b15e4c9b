#!/usr/bin/env python
"""bench.py -- the measurement contract for the direct-volume-ray-casting hot path.

One "step" = one frame of the BASELINE.json workload ("2048^3 u16 DVR @2048^2": synthetic CT-like volume, pre-integrated
classification, empty-space skipping, central-difference gradients + Phong, sampling rate 2, early ray termination) at
one view of the 36-view orbit of SURVEY.md 8d (step k renders view k mod 36).

value    = reference-equivalent Gsamples/s: the samples the reference's fragment shader executes for these frames
           (classified + skipped-by-ESS, i.e. with ERT but without empty-space skipping, which never changes a pixel)
           divided by the device time of the timed region; inputs (volume, TF, camera) resident in HBM.
e2e      = the same metric through the drop-in call a SingleVolumeRaycaster processor makes: b200rc_render with HOST
           (pinned) entry/exit buffers in and HOST RGBA16 + depth out, copies inside the timed region.
roofline = HBM: algorithmic bytes (voxels of the touched 16^3 macro-cells x 2 B + image I/O) / march-kernel
           time, against MEASURED_PEAKS.json.
N > 1    = strong scaling, sort-last: the same volume split into N bricks (one per rank/GPU), partial images merged by
           binary-swap over compositing (voreen_b200/sortlast.py).

`--impl reference` times the CPU restatement of the reference's GLSL path (oracle/, OpenMP on all host cores) on a
bounded sample (a centred crop of each frame) of the same workload.  The reference itself is GLSL behind an OpenGL
driver + Qt and cannot be built here, so `kind` is "port".
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_VIEWS = 36
METRIC = "dvr_reference_equivalent_gsamples_per_s"
UNIT = "Gsamples/s"


# ---------------------------------------------------------------------------------------------------------------------
# workload definition
# ---------------------------------------------------------------------------------------------------------------------
CT_KEYS = [(0.00, (0, 0, 0, 0)), (0.25, (0, 0, 0, 0)), (0.30, (255, 140, 60, 8)), (0.55, (255, 230, 200, 40)),
           (0.80, (255, 255, 255, 200)), (1.00, (255, 255, 255, 255))]


def orbit_eye(k, n=N_VIEWS, radius=3.5):
    th = 2.0 * math.pi * (k % n) / n
    return (radius * math.sin(th), radius * 0.3, radius * math.cos(th))


def workload_name(args):
    img = f"{args.image}^2" if args.image_h in (0, args.image) else f"{args.image}x{args.image_h}"
    return (f"synthetic CT-like {args.size}^3 u16, {args.classification} classification, ESS, central-diff + Phong, "
            f"DVR, rate {args.rate:g}, {img}, {N_VIEWS}-view orbit")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def mark(self):
        """Start of the window whose samples count (nvidia-smi itself needs ~0.1 s to come up, so it is started during the warm-up)."""
        self.t_mark = time.perf_counter()

    def samples(self):
        return sum(1 for t, _ in self.lines if t >= getattr(self, "t_mark", 0.0))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.kill()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t_, ln in self.lines:
            if t_ < getattr(self, "t_mark", 0.0):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU leg: the oracle on a bounded sample (centred crop) of the same frames
# ---------------------------------------------------------------------------------------------------------------------
class CpuSample:
    """Holds what the oracle needs to march a centred crop of the workload's frames: a window of the synthetic volume
    (generated on the CPU, identical bytes to the device generator), the TF / pre-integration table, the cameras."""

    def __init__(self, args, dims):
        import orc
        from voreen_b200 import abi
        from voreen_b200 import raycaster as rc
        self.orc, self.abi, self.rc = orc, abi, rc
        self.args, self.dims = args, tuple(dims)
        self.W, self.H = args.image, (args.image_h or args.image)
        self.step = float(np.float32(1.0) / (np.float32(max(dims)) * np.float32(args.rate)))
        self.cores = os.cpu_count() or 1
        orc.lib().orc_set_threads(self.cores)
        tf = orc.make_tf(CT_KEYS, domain=(0.0, 65535.0))
        self.tf_tex = orc.tf1d_texture(tf, 4096)
        self.preint = None
        if args.classification == "preint":
            tf1k = orc.make_tf(CT_KEYS, domain=(0.0, 65535.0))
            self.preint = orc.preint_cpu(tf1k, 1024, self.step)
        self.params = rc.default_params(self.step, classification=abi.CLASS_PREINT if self.preint is not None else abi.CLASS_TF,
                                        empty_space_skipping=0, out_format=abi.OUT_RGBA32F)
        self.params.collect_stats = 1
        self.full_desc = rc.make_volume_desc(self.dims)
        self.eep = rc.default_eep()
        self.window = {}

    def crop_eep(self, view, cw, ch):
        orc = self.orc
        cam = orc.make_camera(orbit_eye(view), W=self.W, H=self.H)
        vol = orc.Volume(); vol.desc = self.full_desc
        boxes = orc.bbox_boxes(vol, self.eep)
        key = ("eep", view)
        if key not in self.window:
            self.window[key] = orc.entry_exit(boxes, orc.texture_to_world(self.full_desc), self.eep, cam, self.W, self.H)
        e, ed, x, xd = self.window[key]
        y0, x0 = (self.H - ch) // 2, (self.W - cw) // 2
        sl = (slice(y0, y0 + ch), slice(x0, x0 + cw))
        return cam, tuple(np.ascontiguousarray(a[sl]) for a in (e, ed, x, xd))

    def volume_window(self, crops):
        """Synthesize only the voxels the crop's rays can touch (+ apron for gradients)."""
        lo = np.array([1e9] * 3); hi = np.array([-1e9] * 3)
        for (e, ed, x, xd) in crops:
            m = np.any(e[..., :3] != x[..., :3], axis=2)
            if not m.any():
                continue
            pts = np.concatenate([e[m][:, :3], x[m][:, :3]])
            lo = np.minimum(lo, pts.min(axis=0)); hi = np.maximum(hi, pts.max(axis=0))
        d = np.array(self.dims, np.float64)
        vlo = np.clip(np.floor(lo * d - 0.5).astype(int) - 4, 0, self.dims)
        vhi = np.clip(np.ceil(hi * d - 0.5).astype(int) + 5, 0, self.dims)
        key = ("vol", tuple(vlo), tuple(vhi))
        if key not in self.window:
            rd = tuple(int(v) for v in (vhi - vlo))
            data = self.orc.synth_volume(self.dims, self.args.seed, tuple(int(v) for v in vlo), rd)
            desc = self.rc.make_volume_desc(self.dims, shard_lo=tuple(int(v) for v in vlo), shard_hi=tuple(int(v) for v in vhi), shard_apron=0)
            self.window = {k: v for k, v in self.window.items() if k[0] != "vol"}
            self.window[key] = (data, desc)
        return self.window[key]

    def march(self, cam, crop, vol):
        orc = self.orc
        data, desc = vol
        ov = orc.make_volume(data, desc)
        cls = [orc.make_classifier(self.tf_tex, (0.0, 65535.0), preint=self.preint)]
        e, ed, x, xd = crop
        t0 = time.perf_counter()
        outs, depth, st = orc.raycast([ov], cls, self.params, cam, e.shape[1], e.shape[0], e, ed, x, xd)
        dt = time.perf_counter() - t0
        return dt, int(st.samples_classified)

    def calibrate(self, target_s):
        """Pick the crop edge so that one step costs about target_s seconds of CPU work."""
        edge = 32
        cam, crop = self.crop_eep(0, edge, edge)
        vol = self.volume_window([crop])
        dt, ns = self.march(cam, crop, vol)
        per_px = max(dt, 1e-4) / (edge * edge)
        edge2 = int(max(16, min(self.W, math.sqrt(target_s / per_px))) // 8 * 8)
        return max(16, edge2)


def run_reference(args, rank, world):
    """`--impl reference`: the oracle (CPU port of the reference's GLSL path) on the host cores, rank 0 only."""
    if rank != 0:
        return
    dims = (args.size,) * 3
    cs = CpuSample(args, dims)
    total_steps = args.steps + args.warmup
    edge = cs.calibrate(target_s=min(8.0, max(0.5, 150.0 / max(1, total_steps))))
    crops = []
    views = [k % N_VIEWS for k in range(total_steps)]
    for v in sorted(set(views)):
        crops.append((v,) + cs.crop_eep(v, edge, edge))
    vol = cs.volume_window([c[2] for c in crops])
    by_view = {c[0]: (c[1], c[2]) for c in crops}
    for k in range(args.warmup):
        cs.march(*by_view[views[k]], vol)
    t, samples = 0.0, 0
    for k in range(args.warmup, total_steps):
        dt, ns = cs.march(*by_view[views[k]], vol)
        t += dt; samples += ns
    value = samples / t / 1e9
    sample = f"centred {edge}x{edge}-pixel crop of each {args.image}^2 frame (all samples of those rays, ERT on, no ESS), volume window synthesized on the CPU"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args), "l2": "CPU run"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cs.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU leg
# ---------------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import copy
    import torch
    import torch.distributed as dist
    import voreen_b200 as vb

    # NCCL and friends print to stdout ("NCCL version ..."): keep fd 1 for the one JSON line, send everything else to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = vb.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)

    line = measure(args, rank, world, local_rank, ctx, dev, stream, args.parallelism, light=False)
    # config C4 (MultiVolumeRaycaster: 4 co-registered 1024^3 u16 volumes, per-volume transfer functions, DVR at 2048^2), one GPU
    if world == 1 and line is not None and not args.no_c4 and args.size == 2048:
        try:
            line["c4_multivolume"] = measure_c4(ctx, dev, steps=min(args.steps, 8))
        except Exception as e:      # the headline line must not depend on it
            line["c4_multivolume"] = {"error": f"{type(e).__name__}: {e}"}
    # N > 1: the partition the north star names -- the SHARDED volume, sort-last with binary-swap compositing and the exact-ERT second
    # pass -- measured in the same run, next to the headline (which, for a volume that fits one GPU, is the image partition)
    if world > 1 and not args.no_sortlast_record:
        def summary(l):
            return None if l is None else {"ms_per_frame": l["ms_per_frame"], "value": l["value"], "unit": l["unit"], "gsamples_classified_per_s": l["gsamples_classified_per_s"],
                                           "workload": l["config"]["workload"], "parallelism": l["config"]["parallelism"], "transport": l["config"]["transport"],
                                           "exact_ert": True, "phases": l["sortlast_phases"], "steps": l["steps"]}
        headline_is_sortlast = line is not None and line["config"]["parallelism"].startswith("sort-last")
        flag = torch.tensor([1 if headline_is_sortlast else 0], device=dev)
        dist.broadcast(flag, 0)
        if int(flag.item()) == 0:
            a2 = copy.copy(args); a2.steps = min(args.steps, 12); a2.warmup = 3
            l2 = measure(a2, rank, world, local_rank, ctx, dev, stream, "sort-last", light=True)
            if line is not None:
                line["sortlast"] = summary(l2)
        if world == 8 and not args.no_c5 and args.size == 2048:
            # config C5: 4096^3 u16 (128 GiB, 16 GiB + apron per GPU) at 3840 x 2160, sort-last across the 8 GPUs
            a5 = copy.copy(args); a5.size = 4096; a5.image = 3840; a5.image_h = 2160; a5.steps = min(args.steps, 8); a5.warmup = 3
            l5 = measure(a5, rank, world, local_rank, ctx, dev, stream, "sort-last", light=True)
            if line is not None:
                line["sortlast_c5"] = summary(l5)
    if rank == 0:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


C4_KEYS = [CT_KEYS, [(0.0, (0, 0, 0, 0)), (0.3, (0, 255, 0, 0)), (1.0, (0, 255, 0, 120))],
           [(0.0, (0, 0, 0, 0)), (0.4, (0, 0, 255, 0)), (1.0, (0, 0, 255, 90))], [(0.0, (0, 0, 0, 0)), (0.5, (255, 0, 255, 0)), (1.0, (255, 0, 255, 60))]]


def measure_c4(ctx, dev, steps=8, n=1024, img=2048, nvol=4):
    """BASELINE config 4 through b200rc_render_camera_device: device time per frame (CUDA events of the library around its kernels),
    TF classification and pre-integration; multi-volume staging marcher (march_mslab.cuh)."""
    import torch
    import voreen_b200 as vb
    from voreen_b200 import abi, host
    desc = vb.make_volume_desc((n, n, n))
    step_world = float(np.float32(2.0) / np.float32(n) / np.float32(2.0))          # multivolumeraycaster.cpp:326-344, sampling rate 2
    step_tex = float(np.float32(1.0) / (np.float32(n) * np.float32(2.0)))
    for v in range(nvol):
        ctx.synth_volume(v, desc, 1234 + v)
    out_rgba = torch.empty((img, img, 4), dtype=torch.int16, device=dev)
    out_depth = torch.empty((img, img), dtype=torch.float32, device=dev)
    eep = vb.default_eep(world_coords=1)
    res = {"workload": f"{nvol} x synthetic CT-like {n}^3 u16, per-volume transfer functions, ESS, central-diff + Phong, DVR, rate 2, {img}^2, orbit views",
           "kernel": "march_mslab_kernel"}
    try:
        for cls in ("tf", "preint"):
            for v in range(nvol):
                if cls == "preint":
                    ctx.set_tf1d(v, host.tf1d_texture(C4_KEYS[v], 1024), (0.0, 65535.0))
                    ctx.build_preint(v, 1024, step_tex, abi.PI_CPU)
                ctx.set_tf1d(v, host.tf1d_texture(C4_KEYS[v], 4096), (0.0, 65535.0))
            p = vb.default_params(step_world, empty_space_skipping=1, out_format=abi.OUT_RGBA16)
            p.num_volumes = nvol; p.multi_volume = 1
            p.mv_opacity_correction_factor = step_tex / step_world
            for v in range(nvol):
                p.vol[v].classification = abi.CLASS_PREINT if cls == "preint" else abi.CLASS_TF
            ms = []
            for k in range(3 + steps):
                cam = host.make_camera(orbit_eye((k * 5) % N_VIEWS), W=img, H=img)
                ctx.render_camera_ptrs(p, eep, cam, img, img, [out_rgba.data_ptr(), 0, 0], out_depth.data_ptr(), device=True)
                if k >= 3:
                    ms.append(ctx.stats().last_render_ms)
            res[f"ms_per_frame_{cls}"] = sum(ms) / len(ms)
    finally:
        for v in range(1, nvol):
            ctx.lib.b200rc_release_volume(ctx.h, v)
    return res


def measure(args, rank, world, local_rank, ctx, dev, stream, parallelism, light):
    """One measurement of the workload `args` describes; returns the bench line (rank 0) or None.  light: no e2e / CPU baseline / clocks."""
    import torch
    import torch.distributed as dist
    import voreen_b200 as vb
    from voreen_b200 import abi, host

    dims = (args.size,) * 3
    W, H = args.image, (args.image_h or args.image)
    step = float(np.float32(1.0) / (np.float32(max(dims)) * np.float32(args.rate)))
    use_pi = args.classification == "preint"
    sampler = abi.SAMPLER_TEX if args.sampler == "tex" else abi.SAMPLER_EXACT

    # N > 1: "sort-last" shards the volume (what a volume larger than one GPU's HBM needs); "sort-first" replicates it and
    # partitions the image (what a volume that fits should do: no occlusion is lost between ranks, no compositing).
    # "auto" = sort-first when the whole volume + grids fit in 60 % of this GPU's memory, else sort-last.
    sl = sf = None
    mode = "single GPU"
    need = int(np.prod(dims)) * 2 * 1.25
    fits = need < 0.6 * torch.cuda.get_device_properties(dev).total_memory and not args.treat_as_sharded_only
    if world > 1:
        from voreen_b200 import sortlast
        mode = parallelism
        if mode == "auto":
            mode = "sort-first" if fits else "sort-last"
        if mode == "sort-last":
            sl = sortlast.SortLast(ctx, dims, rank, world, W, H, device=dev, stream=stream, transport=args.transport)
            desc = sl.shard_desc()
        else:
            sf = sortlast.SortFirst(ctx, rank, world, W, H, device=dev, stream=stream, transport=args.transport)
            desc = vb.make_volume_desc(dims)
    else:
        desc = vb.make_volume_desc(dims)
    if not fits and sl is None:
        raise SystemExit(f"a {args.size}^3 u16 volume does not fit one GPU: run with --gpus N --parallelism sort-last")
    # A volume that only fits sharded (C5: 4096^3 = 128 GiB) cannot be rendered unsharded for the reference-equivalent sample
    # count; rank 0 measures the ratio (reference-equivalent / nominal) per view on the half-resolution rendering of the same
    # procedural volume (same geometry, same optical depth per unit length) and applies it to the exact nominal count.
    stats_dims = dims if fits else tuple(d // 2 for d in dims)
    stats_step = float(np.float32(1.0) / (np.float32(max(stats_dims)) * np.float32(args.rate)))

    def setup_classifier(st_):
        if use_pi:
            ctx.set_tf1d(0, host.tf1d_texture(CT_KEYS, 1024), (0.0, 65535.0))
            ctx.build_preint(0, 1024, st_, abi.PI_CPU)      # PreIntegrationTable::computeTable on the device
        ctx.set_tf1d(0, host.tf1d_texture(CT_KEYS, 4096), (0.0, 65535.0))

    t0 = time.perf_counter()
    first_desc = desc
    if sl is not None:
        first_desc = vb.make_volume_desc(stats_dims) if rank == 0 else None
    mm_ms, mm_vox = 0.0, 1
    if first_desc is not None:
        ctx.synth_volume(0, first_desc, args.seed)
        ctx.synchronize()
        # the empty-space-skipping min/max build (brick_partials_u16_async_kernel + cell_fold_u16_kernel): HBM streaming, one read of
        # the volume; timed by the library with CUDA events on its stream during the upload above
        _, mm_ms = ctx.download_minmax(0)
        mm_vox = int(np.prod(stats_dims if sl is not None else dims))
    t_synth = time.perf_counter() - t0
    mm_bytes = mm_vox * 2 + (mm_vox // 512) * (27 * 4 * 2 + 8)

    setup_classifier(stats_step if sl is not None else step)
    params = vb.default_params(step, classification=abi.CLASS_PREINT if use_pi else abi.CLASS_TF, sampler=sampler,
                               empty_space_skipping=1, out_format=abi.OUT_RGBA16)
    eep = vb.default_eep()
    cams = [host.make_camera(orbit_eye(k), W=W, H=H) for k in range(N_VIEWS)]

    px = W * H
    out_rgba = torch.empty((H, W, 4), dtype=torch.int16, device=dev)
    out_depth = torch.empty((H, W), dtype=torch.float32, device=dev)

    def frame(k):
        if sl is not None:
            return sl.render(params, eep, cams[k % N_VIEWS], out_rgba, out_depth)
        if sf is not None:
            return sf.render(params, eep, cams[k % N_VIEWS], out_rgba, out_depth)
        ctx.render_camera_ptrs(params, eep, cams[k % N_VIEWS], W, H, [out_rgba.data_ptr(), 0, 0], out_depth.data_ptr(), device=True)
        return ctx.stats().last_march_ms

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- statistics pass (outside the timed region): reference-equivalent samples and touched macro-cells per view ----
    # N > 1: the sample counts that define the metric are those of the UNSHARDED frame (what the reference executes), so
    # rank 0 first renders the whole volume once per view, then every rank loads its shard.
    shaded_by_view = {}

    def stats_pass(render):
        params.collect_stats = 1
        out = []
        for k in range(N_VIEWS):
            render(k)
            st = ctx.stats()
            out.append((st.samples_classified + st.samples_skipped, st.samples_classified, st.cells_touched, st.samples_nominal))
            shaded_by_view[k % N_VIEWS] = st.samples_shaded
        params.collect_stats = 0
        return [list(c) for c in zip(*out)]

    def frame_unsharded(k):
        ctx.render_camera_ptrs(params, eep, cams[k % N_VIEWS], W, H, [out_rgba.data_ptr(), 0, 0], out_depth.data_ptr(), device=True)

    if world > 1:
        ref_samples = classified = nominal = [0.0] * N_VIEWS
        if rank == 0:
            params.sampling_step = stats_step
            ref_samples, classified, _, nominal = stats_pass(frame_unsharded)
            params.sampling_step = step
        t = torch.tensor([ref_samples, classified, nominal], dtype=torch.float64, device=dev)
        dist.broadcast(t, 0)
        ref_samples, classified, nominal = (t[i].tolist() for i in range(3))
        if sl is not None:
            ctx.synth_volume(0, desc, args.seed)        # now the shard
            setup_classifier(step)
            ctx.synchronize()
            sl.exact_ert = False                         # statistics of the FIRST pass (the re-march resets the counters)
        _, _, cells16, nominal_run = stats_pass(frame)   # this rank's touched macro-cells (roofline of its own kernel)
        if sl is not None:
            sl.exact_ert = True
            if not fits:                                 # nominal is exact (every rank counts every cast ray); the rest is scaled
                ref_samples = [n1 * r / max(n0, 1.0) for n1, r, n0 in zip(nominal_run, ref_samples, nominal)]
                classified = [n1 * c / max(n0, 1.0) for n1, c, n0 in zip(nominal_run, classified, nominal)]
                nominal = [float(n1) for n1 in nominal_run]
    else:
        ref_samples, classified, cells16, nominal = stats_pass(frame)

    # ---- timed region -------------------------------------------------------------------------------------------------
    clocks = ClockSampler(local_rank) if (rank == 0 and not light) else None
    for k in range(args.warmup):
        frame(k)
    barrier()
    if clocks:
        clocks.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.stats().kernel_launches
    march_ms = []
    ev0.record(stream)
    for k in range(args.steps):
        march_ms.append(frame(args.warmup + k))
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.stats().kernel_launches - launches0
    # a timed region shorter than nvidia-smi's sampling period: keep the same load running (untimed) until two samples exist
    extra = torch.zeros(1, dtype=torch.int32, device=dev)
    t_x = time.perf_counter()
    while True:
        if clocks:
            extra[0] = 1 if (clocks.samples() < 2 and time.perf_counter() - t_x < 2.0) else 0
        if world > 1:
            dist.broadcast(extra, 0)
        if int(extra.item()) == 0:
            break
        for k in range(4):
            frame(k)
        torch.cuda.synchronize(dev)
    clk = clocks.stop() if clocks else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    # sort-last: per-phase device times of a few frames outside the timed region (max over ranks of each phase)
    sl_phases = None
    if sl is not None:
        sl.profile = True
        acc = {}
        pv = [0, 9, 18, 27]
        for k in pv:
            frame(k)
            mine = {a: float(b) for a, b in sl.phases.items()}
            allp = [None] * world
            dist.all_gather_object(allp, mine)
            for a in mine:
                acc.setdefault(a, []).append(max(d_.get(a, 0.0) for d_ in allp))
        sl.profile = False
        sl_phases = {a: sum(v) / len(v) for a, v in acc.items()}
        sl_phases["note"] = f"views {pv}, max over ranks per phase, mean over views; phases overlap across ranks, so they need not add up to frame_ms"
    steps_views = [(args.warmup + k) % N_VIEWS for k in range(args.steps)]
    S = sum(ref_samples[v] for v in steps_views)
    value = S / (ms_total * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (march) on this rank -----------------------------------------------------------
    peak, peak_src = peaks()
    img_bytes = px * (8 + 4) + px * 2 * 8          # RGBA16 + depth written, entry/exit read
    alg_bytes = [cells16[v] * 4096 * 2 + img_bytes for v in steps_views]
    ach = [b / (m * 1e-3) / 1e9 for b, m in zip(alg_bytes, march_ms)]
    achieved = sum(alg_bytes) / (sum(march_ms) * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "kernel": "march_slab_kernel", "kernel_ms_avg": sum(march_ms) / len(march_ms),
                "kernel_share_of_step": sum(march_ms) / ms_total,
                "alg_bytes_per_launch_avg": sum(alg_bytes) / len(alg_bytes), "frac_min": min(ach) / peak, "frac_max": max(ach) / peak,
                "whole_volume_bytes": int(np.prod(dims)) * 2 + img_bytes,
                "whole_volume_frac": (int(np.prod(dims)) * 2 / (world if sl is not None else 1) + img_bytes) / (sum(march_ms) / len(march_ms) * 1e-3) / 1e9 / peak}
    # second candidate bound: the tap rate.  The slab marcher reads 8 staged taps per classified sample and 24 more per
    # shaded one from shared memory; peak = SMs x 32 banks x SM clock (4-byte words, no bank conflicts)
    if world == 1 and shaded_by_view:
        taps = sum(classified[v] * 8 + shaded_by_view.get(v, 0) * 24 for v in steps_views)
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_clock_hz = (clk["sm_mhz"] if clk and clk.get("sm_mhz") else 1965.0) * 1e6
        tap_peak = sm_count * 32 * sm_clock_hz
        roofline["shared_memory_taps"] = {"achieved_gtaps_per_s": taps / (sum(march_ms) * 1e-3) / 1e9, "peak_gtaps_per_s": tap_peak / 1e9,
                                          "frac": taps / (sum(march_ms) * 1e-3) / tap_peak,
                                          "note": "32-bit shared-memory loads of the trilinear + central-difference footprints; data-dependent gathers replay ~2.4x on bank conflicts (ncu)"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("march_slab_kernel_dram_bytes_per_launch")
        except Exception:
            pass

    # ---- e2e: what the drop-in PROCESSOR does per frame, through the host mirror of SingleVolumeRaycaster (b200host_svrc_process):
    #      property translation, transfer-function hand-over (texture regenerated from the keys and handed to the library, which
    #      recognises unchanged texels / tables), entry/exit images from HOST memory in, RGBA16 + depth to HOST memory out --------
    e2e = None
    if world == 1 and not args.no_e2e and not light:
        import ctypes as C
        n_e2e_views = 6
        ev = [k * (N_VIEWS // n_e2e_views) for k in range(n_e2e_views)]
        hostbuf = []
        for v in ev:
            e, ed, x, xd = ctx.entry_exit(eep, cams[v], W, H)          # what MeshEntryExitPoints hands to the raycaster
            hostbuf.append(tuple(torch.from_numpy(a).pin_memory() for a in (e, ed, x, xd)))
        h_rgba = torch.empty((H, W, 4), dtype=torch.int16).pin_memory()
        h_depth = torch.empty((H, W), dtype=torch.float32).pin_memory()
        pf = C.POINTER(C.c_float)
        sv = host.SingleVolumeRaycaster(local_rank)
        try:
            sv.set_synthetic_volume(desc, args.seed)
            for pid, val in [("samplingRate", args.rate), ("classification", "pre-integrated-cpu" if use_pi else "transfer-function"), ("gradient", "central-differences"),
                             ("shading", "phong"), ("compositing", "dvr"), ("emptySpaceSkipping", True), ("outputFormat", "rgba16")]:
                sv.set_property(pid, val)
            keys = host.make_keys(CT_KEYS)
            thr = (C.c_float * 2)(0.0, 1.0)

            def e2e_frame(i):
                e, ed, x, xd = hostbuf[i % n_e2e_views]
                sv.set_camera(orbit_eye(ev[i % n_e2e_views]))
                sv._check(sv.lib.b200host_svrc_set_transfer_function(sv.h, 0, keys, len(CT_KEYS), 0.0, 65535.0, thr, 1.0), "setTransferFunction")
                sv._check(sv.lib.b200host_svrc_set_entry_exit(sv.h, W, H, C.cast(e.data_ptr(), pf), C.cast(ed.data_ptr(), pf), C.cast(x.data_ptr(), pf), C.cast(xd.data_ptr(), pf)), "setEntryExit")
                ptrs = (C.c_void_p * 3)(h_rgba.data_ptr(), None, None)
                sv._check(sv.lib.b200host_svrc_process(sv.h, W, H, C.byref(ptrs), C.cast(h_depth.data_ptr(), pf)), "process")
            for i in range(max(3, args.warmup)):
                e2e_frame(i)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for i in range(args.steps):
                e2e_frame(i)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
        finally:
            sv.close()
        S2 = sum(ref_samples[ev[i % n_e2e_views]] for i in range(args.steps))
        e2e = {"value": S2 / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": px * (16 + 16 + 4 + 4), "d2h_bytes_per_step": px * (8 + 4),
               "ms_per_step": 1e3 * dt / args.steps,
               "call": "b200host_svrc_process: the SingleVolumeRaycaster mirror (property translation, TF hand-over, host entry/exit in, host RGBA16 + depth out, pinned)"}

    if world > 1 and not args.no_e2e and not light:
        # N ranks: the public call is SortFirst / SortLast.render(params, eep, camera) -- the camera goes in (entry/exit points are
        # generated on every rank from it, MeshEntryExitPoints in-kernel), the finished RGBA16 + depth frame comes back to rank 0's
        # pinned HOST memory every step.  Wall clock between barriers, max over ranks.
        h_rgba = torch.empty((H, W, 4), dtype=torch.int16).pin_memory() if rank == 0 else None
        h_depth = torch.empty((H, W), dtype=torch.float32).pin_memory() if rank == 0 else None

        def e2e_frame_n(i):
            frame(i)
            if rank == 0:
                with torch.cuda.stream(stream):
                    h_rgba.copy_(out_rgba, non_blocking=True)
                    h_depth.copy_(out_depth, non_blocking=True)
                stream.synchronize()
        for i in range(3):
            e2e_frame_n(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_frame_n(args.warmup + i)
        barrier()
        dt_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        dt = float(dt_t.item())
        import ctypes as C
        e2e = {"value": S / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(cams[0]) + C.sizeof(params) + C.sizeof(eep), "d2h_bytes_per_step": px * (8 + 4),
               "ms_per_step": 1e3 * dt / args.steps,
               "call": f"{'SortLast' if sl is not None else 'SortFirst'}.render on {world} ranks: camera + parameter structs in, finished RGBA16 + depth frame to rank 0's pinned host memory out"}

    # ---- CPU baseline: the oracle on a bounded sample of the same frames --------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not light:
        cs = CpuSample(args, dims)
        edge = cs.calibrate(target_s=4.0)
        t, ns = 0.0, 0
        views = [0, 9, 18, 27]
        crops = [(v,) + cs.crop_eep(v, edge, edge) for v in views]
        vol = cs.volume_window([c[2] for c in crops])
        for c in crops:
            dt, n_ = cs.march(c[1], c[2], vol)
            t += dt; ns += n_
        cpu = {"value": ns / t / 1e9, "unit": UNIT, "cores": cs.cores, "kind": "port",
               "sample": f"centred {edge}x{edge}-pixel crops of views {views} of the same frames ({t:.1f} s of oracle work, ERT on, no ESS)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "sampler": args.sampler, "parallelism": f"{mode} x{world}" if world > 1 else "single GPU",
                       "transport": (sl or sf).transport if world > 1 else None,
                       "l2": f"volume {int(np.prod(dims)) * 2 / 2**30:.1f} GiB >> 126 MB L2, a different view every step",
                       "value_is": "reference-equivalent samples/s: the samples the reference's shader executes for these frames (classified + skipped by ESS); "
                                   "the samples this kernel actually classifies are in gsamples_classified_per_s (SURVEY 8d's Gsamples/s)",
                       "samples": "exact (unsharded rendering of the same frames)" if fits else
                                  f"nominal exact; reference-equivalent = nominal x ratio measured on the {stats_dims[0]}^3 rendering of the same procedural volume"},
            "ms_per_frame": ms_total / args.steps,
            "gsamples_classified_per_s": sum(classified[v] for v in steps_views) / (ms_total * 1e-3) / 1e9,
            "gsamples_nominal_per_s": sum(nominal[v] for v in steps_views) / (ms_total * 1e-3) / 1e9,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "synth_s": t_synth, "sortlast_phases": sl_phases,
            "minmax_build": {"kernel": "brick_partials_u16_async_kernel + cell_fold_u16_kernel", "ms": mm_ms, "alg_bytes": mm_bytes,
                             "achieved": mm_bytes / (mm_ms * 1e-3) / 1e9 if mm_ms > 0 else None, "peak": peak, "unit": "GB/s",
                             "frac": (mm_bytes / (mm_ms * 1e-3) / 1e9 / peak) if mm_ms > 0 else None, "bound": "hbm"},
        }
        return line
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=36)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--image", type=int, default=2048, help="image width (and height unless --image-h)")
    ap.add_argument("--image-h", type=int, default=0)
    ap.add_argument("--rate", type=float, default=2.0)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--classification", default="preint", choices=["preint", "tf"])
    ap.add_argument("--sampler", default="exact", choices=["exact", "tex"])
    ap.add_argument("--parallelism", default="auto", choices=["auto", "sort-first", "sort-last"])
    ap.add_argument("--transport", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N > 1: p2p = peer-memory loads/stores inside the kernels (symmetric memory), nccl = send/recv + gather")
    ap.add_argument("--treat-as-sharded-only", action="store_true", help="testing: take the C5 code path (volume never rendered unsharded)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sortlast-record", action="store_true", help="N > 1: skip the additional sort-last (sharded volume) measurement")
    ap.add_argument("--no-c5", action="store_true", help="N = 8: skip the 4096^3 @3840x2160 sort-last measurement")
    ap.add_argument("--no-c4", action="store_true", help="N = 1: skip the config-4 (multi-volume) record")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
