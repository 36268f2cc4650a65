"""Time netF (`FlowUnet`, csrc/flownet.cu) alone on one GPU: `flow_network_warp` of a batch of landmark pairs, CUDA events.

    python tools/flow_probe.py [--batch 64] [--cfg 32,2,4,batch] [--reps 10] [--once]
    AP_FLOW_SPARSE=0 AP_FLOW_TILED=0 python tools/flow_probe.py     # the generic kernel everywhere (A/B)

Prints one JSON line.  `--once` runs a single warm forward (for an ncu launch list of the same command)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animateportrait_b200 import synth  # noqa: E402
from animateportrait_b200.flownet import FlowUnet, flow_network_warp  # noqa: E402


def dense_gflop(nf, ss, ns, max_nf=512, size=224, input_nc=136):
    """Multiply-add FLOPs (2 * MACs) of one frame, layer by layer (networks.py:601-627), in GFLOP."""
    f = {"first7x7": 2.0 * size * size * 49 * input_nc * nf}
    s, nc, rest = size, nf, 0.0
    k = ss
    while k > 1:
        s //= 2
        rest += 2.0 * s * s * 9 * nc * 2 * nc
        nc *= 2
        k //= 2
    for l in range(ns):
        outer, inner = min(max_nf, nc * 2 ** l), min(max_nf, nc * 2 ** (l + 1))
        sl = s // 2 ** l
        rest += 2.0 * (sl // 2) ** 2 * 16 * outer * inner                                  # down conv 4x4 s2
        rest += 2.0 * sl * sl * 4 * (inner if l == ns - 1 else 2 * inner) * outer           # up: 4 taps per output pixel
    rest += 2.0 * s * s * 9 * min(max_nf, nc) * 5                                          # the two heads
    f["rest"] = rest
    return {k_: v / 1e9 for k_, v in f.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--cfg", default="32,2,4,batch")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    nf, ss, ns, norm = a.cfg.split(",")
    nf, ss, ns = int(nf), int(ss), int(ns)
    dev = torch.device("cuda", 0)
    net = FlowUnet(136, nf=nf, start_scale=ss, num_scale=ns, norm=norm).to(dev).eval()
    g = torch.Generator().manual_seed(1)
    sd = net.state_dict()
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            sd[k] = (torch.rand(v.shape, generator=g) + 0.5) if k.endswith("running_var") else 0.05 * torch.randn(v.shape, generator=g)
    net.load_state_dict(sd)
    _, _, _, src, seq, _, _ = synth.make_clip(a.batch, output_nc=1, seed=5)
    lm1 = src.to(dev)[None].expand(a.batch, -1, -1).contiguous()
    lm2 = seq.to(dev)
    for _ in range(1 if a.once else 3):
        iw, ifm = flow_network_warp(net, None, lm1, lm2)
    torch.cuda.synchronize()
    if a.once:
        print(json.dumps({"launches": net.last_launch_count()}))
        return
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(a.reps):
        iw, ifm = flow_network_warp(net, None, lm1, lm2)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / a.reps
    gf = dense_gflop(nf, ss, ns)
    print(json.dumps({"what": "flow_network_warp from landmarks (ap_flow_warp_landmarks: key-point discs + netF + arg-max/mask/resize)", "cfg": a.cfg, "batch": a.batch,
                      "ms_per_batch": ms, "ms_per_frame": ms / a.batch, "frames_per_s": 1e3 * a.batch / ms,
                      "launches": net.last_launch_count(), "dense_gflop_per_frame": gf,
                      "fp32_tflops_on_the_dense_count": sum(gf.values()) * a.batch / ms,
                      "fp32_tflops_without_the_first_conv": gf["rest"] * a.batch / ms,
                      "sparse": os.environ.get("AP_FLOW_SPARSE", "1"), "tiled": os.environ.get("AP_FLOW_TILED", "1"),
                      "mask_mean": float(ifm.mean()), "iw_abs_max": float(iw.abs().max())}))


if __name__ == "__main__":
    main()
