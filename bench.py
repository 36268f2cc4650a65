#!/usr/bin/env python
"""Benchmark of the Module2 netG hot path: generator frames/s at 256x256 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16|fp32_simt]
                    [--batch B] [--output-nc 1|3] [--workload batch|clip] [--frames T]

A step = one generator forward over one batch of B synthetic frames per GPU (BASELINE.json configs[1]:
batch=16 frames, 1xB200, fp32-accurate, line drawing).  N>1 is launched by torchrun, one rank per GPU; frames
shard data-parallel with no collective in the data path (weak scaling: B frames per GPU).
Prints ONE JSON line on rank 0.

`--workload clip` is BASELINE.json configs[2] (one photo, the T = 733 target landmark sets of a 12 s clip, bf16 convs):
a step = one pass over the whole clip through animateportrait_b200.clip.ClipRenderer -- landmark maps and the Delaunay
motion field made on the GPU from the 68x2 coordinates, generator, blend with the static drawing, uint8 frames.

`--impl reference` times the reference's CPU implementation of the same path on the host cores: the
oracle port of it (oracle/netg_oracle.py, bit-identical to the PyTorch reference, see tests/golden) --
the reference is pure Python/PyTorch and cannot travel to the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "generator frames/sec @256x256"
UNIT = "frames/s"
N_INPUT_SETS = 4  # 4 x 42 MB of distinct inputs > 126 MB L2; the per-step working set (GBs) flushes L2 anyway


_STDOUT_FD = None


def claim_stdout():
    """Rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL's version banner comes from NCCL_DEBUG in
    the environment or in /etc/nccl.conf), so file descriptor 1 is pointed at stderr for the whole run and the JSON line
    goes to the saved descriptor."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


def algorithmic_bytes_per_frame(precision: str, onc: int) -> float:
    return (257.0e6 if precision == "bf16" else 513.9e6) + (0.5e6 if onc == 3 else 0.0)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def reference_forward_fn(output_nc: int):
    """The reference arm's callable: the UNMODIFIED reference module built by its own define_G from baseline/_ref
    (vendored by tools/prep_ref.py) when it is there, else the oracle port (bit-identical, tests/golden).  Returns
    (fn(*six inputs) -> frames, kind)."""
    import torch
    from oracle import netg_oracle as O
    sd = O.make_state_dict(output_nc, seed=0)
    try:
        from baseline import reference_netg as R
        if R.available():
            net = R.make_reference_netG(output_nc, sd)

            def fn(*inputs):
                with torch.no_grad():
                    return net(*inputs)
            return fn, "reference"
    except Exception as e:  # the vendored files are optional; the port always exists
        print(f"[bench] reference module unavailable ({e!r}); timing the oracle port", file=sys.stderr)
    return (lambda *inputs: O.netg_forward(sd, *inputs)), "port"


def workload_name(onc: int, B: int, precision: str) -> str:
    return (f"configs[1]: netG resnet_9blocks_rcatland32_full_ifw, output_nc={onc}, batch={B} frames, 256x256, "
            f"precision={precision}")


def time_cpu_reference(B: int, output_nc: int, min_seconds: float, max_iters: int):
    """CPU reference frames/s on the host cores, bounded sample (the cpu_baseline leg of our own arm)."""
    import torch
    from oracle import netg_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    fn, kind = reference_forward_fn(output_nc)
    inputs = O.make_inputs(B, seed=1016, kind="smooth")
    fn(*inputs)  # warm-up
    times = []
    t_all = time.perf_counter()
    while len(times) < max_iters and (time.perf_counter() - t_all < min_seconds or len(times) < 2):
        t0 = time.perf_counter()
        fn(*inputs)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return B / med, len(times), torch.get_num_threads(), kind


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the path on the host cores (rank 0 only), on the
    batch our arm runs (configs[1]: B=16, fp32).  A step is the whole batch unless K steps of it would take more than
    about three minutes on this host; then it is the largest sample of the batch that fits, and the line says so."""
    if rank != 0:
        return
    import torch
    from oracle import netg_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    fn, kind = reference_forward_fn(args.output_nc)
    B = args.batch
    one = O.make_inputs(1, seed=1016, kind="smooth")
    fn(*one)
    t0 = time.perf_counter()
    fn(*one)
    t1 = time.perf_counter() - t0                      # one frame; batches are about 0.6x of that per frame
    budget = 180.0
    Bs = B
    while Bs > 1 and (args.steps + args.warmup) * Bs * t1 * 0.6 > budget:
        Bs //= 2
    inputs = O.make_inputs(B, seed=1016, kind="smooth")
    inputs = [t[:Bs].contiguous() for t in inputs]
    for _ in range(args.warmup):
        fn(*inputs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(*inputs)
    dt = time.perf_counter() - t0
    fps = Bs * args.steps / dt
    what = "the unmodified reference module (define_G, baseline/_ref)" if kind == "reference" else "oracle port of the reference"
    sample = (f"{args.steps} steps x {Bs} frames" + ("" if Bs == B else f" (a {Bs}-frame sample of the {B}-frame batch)")
              + f", {what}, torch {torch.__version__} CPU ops on all host threads")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.output_nc, B, "fp32") if Bs == B else
                       workload_name(args.output_nc, B, "fp32") + f"; step = {Bs}-frame sample of the batch"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


CLIP_METRIC_NOTE = ("configs[2]: one photo + {T} target landmark sets (12 s clip), landmark maps + Delaunay motion field on the "
                    "GPU, netG output_nc={onc} precision={prec} in batches of {B}{share}, blend with the static drawing, uint8 frames; "
                    "intrinsic flow / visibility mask: {flow}")


def _cpu_clip_frames(fwd, clip, frames):
    """The reference's per-frame loop on the CPU for the given frame indices (oracle port, batch size 1 as
    Module2/test.py:42): cal_motion256 through scipy's griddata -- the reference's own call -- when scipy is installed."""
    import numpy as np
    import torch
    from oracle import cond_oracle as OC
    from oracle import netg_oracle as O
    photo, matte, static, src, seq, flow, ifmask = clip
    try:
        from scipy.interpolate import griddata
    except Exception:  # pragma: no cover
        griddata = None
    real_A, mask = OC.matte_photo(photo.numpy(), matte.numpy())
    real_A, mask = torch.from_numpy(real_A), torch.from_numpy(mask)
    land1 = torch.from_numpy(OC.draw_landmarks(src.numpy()[None]))
    ys, xs = np.mgrid[0:256, 0:256]
    for t in frames:
        land2 = torch.from_numpy(OC.draw_landmarks(seq[t:t + 1].numpy()))
        if griddata is not None:
            sites, vals = OC.motion_sites(src.numpy(), seq[t].numpy())
            m = griddata(sites, vals, (xs.astype(np.float64), ys.astype(np.float64)), method="linear").astype(np.float32)
            motion = torch.from_numpy(m / np.float32(127.5) - np.float32(1))[None]
        else:
            motion = torch.from_numpy(OC.cal_motion(src.numpy(), seq[t].numpy()))[None]
        fake = fwd(real_A, land1, land2, motion, flow[t:t + 1], ifmask[t:t + 1])
        O.tensor2im_batch(O.blend_foreground(fake, mask, motion, static))
    return griddata is not None


def _time_cpu_clip(args, n_frames, steps, warmup):
    import torch
    from animateportrait_b200 import synth
    from oracle import netg_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    fwd, kind = reference_forward_fn(args.output_nc)
    clip = synth.make_clip(max(n_frames, 8), args.output_nc, seed=2000)
    frames = list(range(n_frames))
    for _ in range(warmup):
        _cpu_clip_frames(fwd, clip, frames[:1])
    t0 = time.perf_counter()
    for _ in range(steps):
        used_scipy = _cpu_clip_frames(fwd, clip, frames)
    dt = time.perf_counter() - t0
    return n_frames * steps / dt, dt / steps, torch.get_num_threads(), used_scipy, kind


def run_reference_clip(args, rank):
    """Reference arm of the clip workload: the per-frame CPU loop on a bounded sample of the clip's frames."""
    if rank != 0:
        return
    import torch
    n = 8
    fps, step_s, cores, used_scipy, kind = _time_cpu_clip(args, n, args.steps, max(args.warmup, 1))
    sample = (f"{args.steps} steps x {n} frames of the clip, one frame at a time (reference batch size 1), torch "
              f"{torch.__version__} CPU ops on all host threads, motion field by "
              f"{'scipy griddata (the reference call)' if used_scipy else 'the numpy restatement'}")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CLIP_METRIC_NOTE.format(T=args.frames, onc=args.output_nc, prec="fp32 (CPU)", B=1, share="", flow="device-resident stand-ins for netF's output")
                       + f"; step = {n}-frame sample of the clip"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def run_clip(args, rank, local_rank, world):
    """BASELINE.json configs[2] on N GPUs: frames of ONE clip shard over the ranks (strong scaling); only the target
    landmarks are scattered (68x2 floats per frame), uint8 frames are gathered on rank 0."""
    import torch
    import torch.distributed as dist
    import animateportrait_b200 as ap
    from animateportrait_b200 import synth
    from animateportrait_b200.clip import ClipRenderer
    from animateportrait_b200.frames import _gather, _scatter, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the generator")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    T, B, onc = args.frames, args.batch, args.output_nc
    net = ap.define_G(3, onc, 64, ap.NETG_NAME, "instance", False, "normal", 0.02, [local_rank], div=3, disp=3,
                      precision=args.precision).module
    net.load_state_dict(synth.make_state_dict(onc, seed=0))
    photo, matte, static, src, seq, flow, ifmask = synth.make_clip(T, onc, seed=2000)
    lo, hi = shard_range(T, world, rank)
    # netF stand-in: every rank holds the flow / visibility mask of its own frames, as a per-rank flow network would
    flow_d, ifm_d = flow[lo:hi].to(dev), ifmask[lo:hi].to(dev)
    seq_host = seq.pin_memory()
    seq_dev = seq.to(dev) if rank == 0 else None
    netF = None
    if args.flow_net:
        # row f3: netF makes iw_flow / if_mask per batch from the landmarks instead of the device-resident stand-ins.  The
        # released configuration is unknown (train_opt.json does not ship): "nf,start_scale,num_scale,norm", seeded weights
        from animateportrait_b200.flownet import FlowUnet
        nf_, ss_, ns_, norm_ = args.flow_net.split(",")
        netF = FlowUnet(136, nf=int(nf_), start_scale=int(ss_), num_scale=int(ns_), norm=norm_).to(dev).eval()
        g_ = torch.Generator().manual_seed(0)
        with torch.no_grad():
            for p_ in netF.parameters():
                p_.copy_(torch.randn(p_.shape, generator=g_) * 0.05)
    r = ClipRenderer(net, batch=B, share_photo=not args.no_share_photo, netF=netF)
    r.set_photo(photo.to(dev), src.to(dev), matte.to(dev), static.to(dev))
    mine = torch.empty((hi - lo, 256, 256, 3), dtype=torch.uint8, device=dev)
    frames_host = torch.empty((T, 256, 256, 3), dtype=torch.uint8, pin_memory=True) if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(lm_rank0):
        lm = _scatter(lm_rank0, (68, 2), T, dev, None) if world > 1 else lm_rank0
        if hi > lo:
            if netF is not None:
                r.render(lm, out=mine)
            else:
                r.render(lm, flow_d, ifm_d, out=mine)
        return _gather(mine, T, None) if world > 1 else mine

    # ---------------- device-resident: landmarks already in HBM, frames stay in HBM ----------------
    for _ in range(args.warmup):
        one_pass(seq_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_pass(seq_dev)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    value = T * args.steps / (ms_max * 1e-3)

    # ---------------- end to end: host landmarks in, host uint8 frames out ----------------
    e2e_steps = max(2, min(args.steps, 5))

    def one_e2e():
        lm0 = seq_host.to(dev, non_blocking=True) if rank == 0 else None
        fr = one_pass(lm0)
        if rank == 0:
            frames_host.copy_(fr, non_blocking=True)
        torch.cuda.synchronize()

    one_e2e()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        one_e2e()
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e = {"value": T * e2e_steps / (t2.item() * 1e-3), "unit": UNIT, "h2d_bytes_per_step": T * 68 * 2 * 4,
           "d2h_bytes_per_step": T * 256 * 256 * 3,
           "path": "pinned host target landmarks -> H2D -> (NCCL scatter) -> ClipRenderer.render per rank -> (NCCL gather) -> "
                   "D2H uint8 frames -> sync, every step"}

    # ---------------- per-kernel-class device time of the generator + the conditioning kernels ----------------
    peaks = measured_peaks()
    nb = min(B, hi - lo)
    lm_b = seq[lo:lo + nb].to(dev)
    from animateportrait_b200 import conditioning as cond
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(2):
        ev[0].record()
        cond.draw2(256, 256, lm_b, 3)
        ev[1].record()
        motion_b = cond.cal_motion256(src.to(dev), lm_b)
        ev[2].record()
    torch.cuda.synchronize()
    cond_ms = {"draw2": ev[0].elapsed_time(ev[1]), "cal_motion256": ev[1].elapsed_time(ev[2])}
    if netF is not None:
        from animateportrait_b200.flownet import flow_network_warp
        lm1_b = src.to(dev)[None].expand(nb, -1, -1)
        for _ in range(2):
            ev[0].record()
            flow_network_warp(netF, None, lm1_b, lm_b)
            ev[1].record()
        torch.cuda.synchronize()
        cond_ms["netF_flow_network_warp"] = ev[0].elapsed_time(ev[1])
    photo_b, land1_b = r._expanded(nb)[:2]
    land2_b = cond.draw2(256, 256, lm_b, 3)
    net.set_profiling(True)
    prof = None
    with torch.no_grad():
        for _ in range(3):
            if args.no_share_photo:
                net(photo_b, land1_b, land2_b, motion_b, flow_d[:nb], ifm_d[:nb])
            else:
                net.forward_shared_photo(photo_b[:1], land1_b[:1], land2_b, motion_b, flow_d[:nb], ifm_d[:nb])
            p = net.get_profile()
            if prof is None:
                prof = p
            else:
                for k in p:
                    for f in ("ms", "launches", "flops"):
                        prof[k][f] += p[k][f]
    net.set_profiling(False)
    trunk = prof["trunk_conv3x3"]
    trunk_tflops = trunk["flops"] / (trunk["ms"] * 1e-3) / 1e12 if trunk["ms"] > 0 else 0.0
    nprod = {"fp32": 3, "bf16": 1, "fp32_simt": 0}[args.precision]
    step_prof = sum(v["ms"] for v in prof.values()) / 3.0
    roofline = {"bound": "tensor", "kernel": f"conv_umma_kernel (3x3 s1 trunk convs @64x64, 22 launches per {nb}-frame batch)",
                "achieved": trunk_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": trunk_tflops / peaks["bf16_tflops_sustained"], "traffic": None,
                "mma_tflops": trunk_tflops * max(nprod, 1),
                "frac_of_mma_ceiling": trunk_tflops * max(nprod, 1) / peaks["bf16_tflops_sustained"],
                "peak_source": peaks["source"] + " bf16 sustained (kernel timed inside a long step)",
                "algorithmic_flops_per_launch": trunk["flops"] / max(trunk["launches"], 1),
                "avg_launch_ms": trunk["ms"] / max(trunk["launches"], 1), "mma_products_per_flop": nprod,
                "share_of_step": trunk["ms"] / 3.0 / step_prof if step_prof else None,
                "classes_ms_per_batch": {k: round(v["ms"] / 3.0, 4) for k, v in prof.items()},
                "conditioning_ms_per_batch": {k: round(v, 4) for k, v in cond_ms.items()}}

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            fps, step_s, cores, used_scipy, kind = _time_cpu_clip(args, 8, 2, 1)
            cpu_baseline = {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": "2 x 8 frames of the clip, frame by frame (reference batch size 1), the reference's "
                                      "loop (dataset item, generator, blend, tensor2im) on all host threads, motion field by "
                                      + ("scipy griddata (the reference call)" if used_scipy else "the numpy restatement")}
        n_batches = sum(-(-(shard_range(T, world, q)[1] - shard_range(T, world, q)[0]) // B) for q in range(world))
        launches_per_step = n_batches * (net.last_launch_count() + 4)   # + draw2, delaunay, raster, compose per batch
        flops_frame = synth.flops_per_frame(onc)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": {"fp32": "bf16x3 (hi/lo split, fp32 accumulate; fp32-accurate)", "bf16": "bf16",
                          "fp32_simt": "f32"}[args.precision],
                "data": "synthetic",
                "config": {"workload": CLIP_METRIC_NOTE.format(T=T, onc=onc, prec=args.precision, B=B,
                                                               share=" (photo-only encoder layers once per batch)"
                                                               if not args.no_share_photo else " (B copies of the photo)",
                                                               flow=(f"netF FlowUnet({args.flow_net}) per batch, seeded weights"
                                                                     if args.flow_net else
                                                                     "device-resident stand-ins for netF's output")),
                           "l2": f"every batch touches a {net.workspace_bytes(min(B, T)) / 1e9:.1f} GB working set (> 126 MB L2); "
                                 "per-frame inputs differ for every frame of the clip",
                           "parallelism": f"dp{world} (frames of one clip sharded; landmark scatter + frame gather only)"},
                "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "model_tflops": value * flops_frame / 1e12,
                "tensor_frac_of_sustained_bf16": value / world * flops_frame / 1e12 / peaks["bf16_tflops_sustained"],
                # algorithmic bytes per frame with everything fused (each activation written once, read once
                # per consumer): 513.9 MB with fp32 activations, 257.0 MB with bf16 activations, +0.5 MB for output_nc=3
                "hbm_frac_of_measured_copy": value / world * algorithmic_bytes_per_frame(args.precision, onc)
                                             / (peaks["hbm_gbs"] * 1e9)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--gpus", type=int, default=1)
    ap_.add_argument("--steps", type=int, default=None)
    ap_.add_argument("--warmup", type=int, default=5)
    ap_.add_argument("--workload", default="batch", choices=["batch", "clip"])
    ap_.add_argument("--frames", type=int, default=733, help="clip workload: frames per clip (12 s of audio = 733)")
    ap_.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap_.add_argument("--precision", default=None, choices=["fp32", "bf16", "fp32_simt"])
    ap_.add_argument("--batch", type=int, default=None)
    ap_.add_argument("--output-nc", type=int, default=1, dest="output_nc")
    ap_.add_argument("--no-cpu-baseline", action="store_true")
    ap_.add_argument("--frames-per-gpu", type=int, default=64, dest="frames_per_gpu",
                     help="N > 1: frames per GPU of the clip rank 0 owns (BASELINE.json configs[3]: 64)")
    ap_.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                     help="N > 1: how the frames return to rank 0 (frames.render_frames_sharded)")
    ap_.add_argument("--flow-net", default=None, dest="flow_net",
                     help="clip workload: run netF (row f3) per batch, e.g. 32,2,4,batch = nf,start_scale,num_scale,norm")
    ap_.add_argument("--no-share-photo", action="store_true",
                     help="clip workload: hand netG B copies of the photo instead of the shared-photo entry point")
    args = ap_.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    clip = args.workload == "clip"
    if args.steps is None:
        args.steps = (5 if clip else 100) if args.impl == "ours" else (2 if clip else 5)
    if args.precision is None:
        args.precision = "bf16" if clip else "fp32"      # configs[2] is quoted on bf16 convs, configs[1] on fp32
    if args.batch is None:
        args.batch = 64 if clip else 16

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    claim_stdout()

    if args.impl == "reference":
        (run_reference_clip if clip else run_reference)(args, rank)
        return
    if clip:
        run_clip(args, rank, local_rank, world)
        return

    run_batch(args, rank, local_rank, world)


def _max_over_ranks(ms: float, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def run_batch(args, rank, local_rank, world):
    """N = 1: BASELINE.json configs[1] (one batch of B=16 frames per step, fp32-accurate, line drawing).
    N > 1: configs[3] -- rank 0 owns N x F synthetic frames (F = 64 per GPU); a step scatters them chunk by chunk
    (NCCL), renders every shard in batches of B and collects the frames on rank 0 (SURVEY.md §8d "Config 4": the time
    includes scatter + compute + gather)."""
    import torch
    import torch.distributed as dist
    import animateportrait_b200 as ap
    from animateportrait_b200 import synth as O  # seeded stand-in checkpoint + synthetic batches (product side)
    from animateportrait_b200.frames import render_frames_sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the generator")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    B, onc = args.batch, args.output_nc
    F = args.frames_per_gpu

    net = ap.define_G(3, onc, 64, ap.NETG_NAME, "instance", False, "normal", 0.02, [local_rank], div=3, disp=3,
                      precision=args.precision).module
    net.load_state_dict(O.make_state_dict(onc, seed=0))
    host_sets = [[t.pin_memory() for t in O.make_inputs(B, seed=1016 + 97 * rank + i, kind="smooth")]
                 for i in range(N_INPUT_SETS)]
    dev_sets = [[t.to(dev) for t in s] for s in host_sets]
    h2d = sum(t_.numel() * 4 for t_ in host_sets[0])
    out_host = torch.empty((B, onc, 256, 256), dtype=torch.float32, pin_memory=True)
    d2h = out_host.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, warmup, finish=None):
        """fn(i) for `steps` steps between barriers (+ finish(), e.g. the drain of a pipelined API); device time, max
        over ranks, in ms."""
        for i in range(warmup):
            fn(i)
        if finish:
            finish()
        barrier()
        e0.record()
        for i in range(steps):
            fn(i)
        if finish:
            finish()
        e1.record()
        barrier()
        return _max_over_ranks(e0.elapsed_time(e1), dev, world)

    # ---------------- replicas: every rank renders its own resident batch, no communication ----------------
    def replica_step(i):
        with torch.no_grad():
            net(*dev_sets[i % N_INPUT_SETS])

    extra = {}
    sampler = ClockSampler(local_rank)
    if world == 1:
        for i in range(args.warmup):
            replica_step(i)
        barrier()
        if rank == 0:
            sampler.start()
        e0.record()
        for i in range(args.steps):
            replica_step(i)
        e1.record()
        barrier()
        ms_max = e0.elapsed_time(e1)
        clocks = sampler.stop()
        frames_per_step = B
        launches_per_step = net.last_launch_count()
        workload = workload_name(onc, B, args.precision)
        parallelism = "dp1 (one GPU)"
        # end to end: host buffers in, host frames out, through the host-buffer entry point; at least a second long so
        # that it runs at the sustained (power-capped) clocks like the leg above
        est = ms_max / args.steps
        e2e_steps = max(args.steps, 3, int(1200.0 / max(est, 0.1)))

        out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]

        def host_step(i):   # the clip loop of a user: batch after batch through the pipelined host-buffer entry point
            net.forward_host_async(*host_sets[i % N_INPUT_SETS], out=out_hosts[i % 2])

        def host_step_sync(i):
            net.forward_host(*host_sets[i % N_INPUT_SETS], out=out_host)

        ms2 = timed(host_step, e2e_steps, 2, finish=net.host_sync)
        ms2s = timed(host_step_sync, max(3, e2e_steps // 4), 2)
        e2e = {"value": B * e2e_steps / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps,
               "path": "ap_netg_forward_host_async, every step: pinned host inputs -> H2D -> forward -> D2H frames into pinned "
                       "host memory; two staging slots, so the copies of one step overlap the forward of its neighbours; "
                       "ap_netg_host_sync inside the timed region",
               "one_call_at_a_time": {"value": B * max(3, e2e_steps // 4) / (ms2s * 1e-3), "unit": UNIT,
                                      "path": "ap_netg_forward_host: H2D -> forward -> D2H -> sync per call, no overlap "
                                              "between calls"}}
        # batch size 1, the shape the reference's own loop calls the generator with (Module2/test.py:42): CUDA-graph replay
        one_dev = [t[:1].contiguous() for t in dev_sets[0]]
        one_host = [t[:1].contiguous().pin_memory() for t in host_sets[0]]
        one_out = torch.empty((1, onc, 256, 256), dtype=torch.float32, pin_memory=True)

        def b1_dev(i):
            with torch.no_grad():
                net(*one_dev)

        def b1_host(i):
            net.forward_host(*one_host, out=one_out)

        n1 = 300
        extra["batch1"] = {"device_frames_per_s": n1 / (timed(b1_dev, n1, 20) * 1e-3),
                           "e2e_frames_per_s": n1 / (timed(b1_host, n1, 20) * 1e-3),
                           "launches_per_frame": net.last_launch_count(), "steps": n1,
                           "note": "B=1 (the reference's call, Module2/test.py:42): CUDA-graph replay of the forward; e2e = "
                                   "ap_netg_forward_host with pinned host buffers"}
    else:
        T = world * F
        info = {}
        full_dev = full_host = frames_host = None
        if rank == 0:
            full_host = [torch.cat([O.make_inputs(F, seed=1016 + 97 * r, kind="smooth")[k] for r in range(world)]).pin_memory()
                         for k in range(6)]
            full_dev = [t.to(dev) for t in full_host]
            frames_host = torch.empty((T, onc, 256, 256), dtype=torch.float32, pin_memory=True)

        def sharded_step(i):
            render_frames_sharded(net, full_dev, T, onc, dev, batch=B, gather=args.gather, info=info)

        for i in range(max(args.warmup, 2)):
            sharded_step(i)
        barrier()
        if rank == 0:
            sampler.start()
        e0.record()
        for i in range(args.steps):
            sharded_step(i)
        e1.record()
        barrier()
        ms_max = _max_over_ranks(e0.elapsed_time(e1), dev, world)
        clocks = sampler.stop() if rank == 0 else None
        frames_per_step = T
        chunks = -(-F // B)
        launches_per_step = chunks * net.last_launch_count()
        workload = (f"configs[3]: {T} synthetic frames owned by rank 0 ({F} per GPU), netG resnet_9blocks_rcatland32_full_ifw "
                    f"output_nc={onc} precision={args.precision} in batches of {B}; every step = NCCL scatter of the six "
                    f"conditioning tensors + render + gather of the frames on rank 0")
        parallelism = (f"dp{world}: frames sharded; chunked NCCL send/recv scatter on a side stream under the render; gather = "
                       f"{info.get('gather')} ({'output kernels store into a CUDA-IPC mapping of rank 0 frame buffer over NVLink' if info.get('gather') == 'peer-store' else 'NCCL send/recv behind the render'})")

        # end to end: the clip starts in rank 0's pinned host memory and the frames end there
        def host_step(i):
            render_frames_sharded(net, full_host, T, onc, dev, batch=B, gather=args.gather, out_host=frames_host)
            torch.cuda.synchronize()

        e2e_steps = max(3, min(args.steps, 10))
        ms2 = timed(host_step, e2e_steps, 1)
        e2e = {"value": T * e2e_steps / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": T * h2d // B,
               "d2h_bytes_per_step": T * d2h // B, "steps": e2e_steps,
               "path": "rank 0 pinned host clip -> H2D round by round (ONE PCIe link feeds all GPUs) -> NCCL scatter -> render "
                       "per rank -> gather on rank 0 -> D2H -> sync, every step"}
        # the communication-free figure (every rank renders a resident batch of B): what the links cost is value vs this
        rsteps = max(10, min(args.steps, 30))
        ms3 = timed(replica_step, rsteps, 3)
        extra["replicas_no_communication"] = {"value": world * B * rsteps / (ms3 * 1e-3), "unit": UNIT, "steps": rsteps,
                                              "note": f"{world} independent replicas, B={B} resident per rank"}

    value = frames_per_step * args.steps / (ms_max * 1e-3)

    # ---------------- per-kernel-class device time (separate profiled pass, CUDA events per launch) ----------------
    peaks = measured_peaks()
    net.set_profiling(True)
    prof_acc = None
    with torch.no_grad():
        for i in range(3):
            net(*dev_sets[i % N_INPUT_SETS])
            p = net.get_profile()
            if prof_acc is None:
                prof_acc = p
            else:
                for k in p:
                    for f in ("ms", "launches", "flops"):
                        prof_acc[k][f] += p[k][f]
    net.set_profiling(False)
    step_ms_prof = sum(v["ms"] for v in prof_acc.values()) / 3.0
    trunk = prof_acc["trunk_conv3x3"]
    trunk_tflops = trunk["flops"] / (trunk["ms"] * 1e-3) / 1e12 if trunk["ms"] > 0 else 0.0
    nprod = {"fp32": 3, "bf16": 1, "fp32_simt": 0}[args.precision]
    traffic = None  # dram bytes per launch of the dominant kernel, from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "trunk_traffic.json")
    if nprod and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        for ent in (tj if isinstance(tj, list) else [tj]):
            if ent.get("batch") == B and ent.get("precision") == args.precision:
                traffic = ent.get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "conv_umma_kernel (3x3 s1 trunk convs @64x64, 22 launches per batch)"
                if nprod else "conv_simt_kernel (CUDA-core validation path)",
                "achieved": trunk_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": trunk_tflops / peaks["bf16_tflops_sustained"], "traffic": traffic,
                "mma_tflops": trunk_tflops * max(nprod, 1),
                "frac_of_mma_ceiling": trunk_tflops * max(nprod, 1) / peaks["bf16_tflops_sustained"],
                "peak_source": peaks["source"] + " bf16 sustained (kernel timed inside a long step)",
                "algorithmic_flops_per_launch": trunk["flops"] / max(trunk["launches"], 1),
                "avg_launch_ms": trunk["ms"] / max(trunk["launches"], 1),
                "mma_products_per_flop": nprod,
                "share_of_step": trunk["ms"] / 3.0 / step_ms_prof if step_ms_prof else None,
                "classes_ms_per_batch": {k: round(v["ms"] / 3.0, 4) for k, v in prof_acc.items()}}

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            fps, iters, cores, kind = time_cpu_reference(4, onc, 12.0, 40)  # ~12 s of CPU work on all host cores
            import torch as _t
            cpu_baseline = {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": f"{iters} forwards of a 4-frame sample of the batch, "
                                      + ("the unmodified reference module (baseline/_ref)" if kind == "reference"
                                         else "oracle port of the reference")
                                      + f" (torch {_t.__version__} CPU ops), median"}
        flops_frame = O.flops_per_frame(onc)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "bf16x3 (hi/lo split, fp32 accumulate; fp32-accurate)", "bf16": "bf16",
                          "fp32_simt": "f32"}[args.precision],
                "data": "synthetic",
                "config": {"workload": workload,
                           "l2": f"{N_INPUT_SETS} rotating input sets ({N_INPUT_SETS * h2d / 1e6:.0f} MB) and a "
                                 f"{net.workspace_bytes(B) / 1e9:.1f} GB per-batch working set, both larger than the 126 MB L2"
                                 if world == 1 else f"{F} frames per GPU per step ({F * h2d / B / 1e6:.0f} MB of inputs) and a "
                                 f"{net.workspace_bytes(B) / 1e9:.1f} GB per-batch working set, both larger than the 126 MB L2",
                           "parallelism": parallelism},
                "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "model_tflops": value * flops_frame / 1e12,
                "tensor_frac_of_sustained_bf16": value / world * flops_frame / 1e12 / peaks["bf16_tflops_sustained"],
                # algorithmic bytes per frame with everything fused (each activation written once, read once
                # per consumer): 513.9 MB with fp32 activations, 257.0 MB with bf16 activations, +0.5 MB for output_nc=3
                "hbm_frac_of_measured_copy": value / world * algorithmic_bytes_per_frame(args.precision, onc)
                                             / (peaks["hbm_gbs"] * 1e9)}
        line.update(extra)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
