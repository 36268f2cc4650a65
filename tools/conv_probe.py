"""Timing probe of ONE tcgen05 conv layer through ap_conv2d_debug (run under ncu; see tools/conv_probe.sh).
    python tools/conv_probe.py Cin Cout S stride pad_mode transposed B impl [reps]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import animateportrait_b200 as ap

Cin, Cout, S, stride = (int(a) for a in sys.argv[1:5])
pad_mode, transposed, B, impl = sys.argv[5], sys.argv[6] == "1", int(sys.argv[7]), sys.argv[8]
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 3
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, Cin, S, S, generator=g).to(dev)
w = (torch.randn((Cin, Cout, 3, 3) if transposed else (Cout, Cin, 3, 3), generator=g) * 0.05).to(dev)
for _ in range(reps):
    y, st = ap.conv2d_debug(x, w, stride=stride, pad=1, pad_mode=pad_mode, transposed=transposed, impl=impl)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
