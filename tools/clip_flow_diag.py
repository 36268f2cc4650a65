"""Where does a clip rendered with netF in the loop spend its time?  (one GPU, CUDA events + allocator statistics)"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import animateportrait_b200 as ap
from animateportrait_b200 import synth
from animateportrait_b200.clip import ClipRenderer
from animateportrait_b200.flownet import FlowUnet, flow_network_warp

dev = torch.device("cuda", 0)
T, B = 733, 64
netF = FlowUnet(136, nf=32, start_scale=2, num_scale=4, norm="batch").to(dev).eval()
net = ap.define_G(3, 1, 64, ap.NETG_NAME, "instance", False, "normal", 0.02, [0], div=3, disp=3, precision="bf16")
net.module.load_state_dict(synth.make_state_dict(1, seed=3, bias_std=0.3))
photo, matte, static, src, seq, flow, ifm = synth.make_clip(T, output_nc=1, seed=2000)
seq_d = seq.to(dev)

def timed(fn, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    st0 = torch.cuda.memory_stats(dev)
    t0 = time.perf_counter()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps): fn()
    ev[1].record(); torch.cuda.synchronize()
    st1 = torch.cuda.memory_stats(dev)
    return {"ms": ev[0].elapsed_time(ev[1]) / reps, "wall_ms": (time.perf_counter() - t0) * 1e3 / reps,
            "cudaMalloc_calls": st1["num_device_alloc"] - st0["num_device_alloc"],
            "cudaFree_calls": st1["num_device_free"] - st0["num_device_free"]}

out = {}
r = ClipRenderer(net, batch=B, netF=netF)
r.set_photo(photo.to(dev), src.to(dev), matte.to(dev), static.to(dev))
out["clip_with_netF"] = timed(lambda: r.render(seq_d))
r0 = ClipRenderer(net, batch=B)
r0.set_photo(photo.to(dev), src.to(dev), matte.to(dev), static.to(dev))
fl, im = flow.to(dev), ifm.to(dev)
out["clip_given_flow"] = timed(lambda: r0.render(seq_d, fl, im))
lm1 = src.to(dev)[None].expand(B, -1, -1).contiguous()
def only_flow():
    for s in range(0, T, B):
        e = min(s + B, T)
        flow_network_warp(netF, None, lm1[:e - s], seq_d[s:e])
out["netF_only_12_batches"] = timed(only_flow)
print(json.dumps(out))
