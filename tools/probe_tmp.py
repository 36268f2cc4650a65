import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import animateportrait_b200 as ap
from oracle import netg_oracle as O
dev = torch.device("cuda", 0)
sd = O.make_state_dict(1, seed=0)
inputs = [t.to(dev) for t in O.make_inputs(1, seed=1001, kind="smooth")]
names = ["tri00","warp0","tri01","tri02","tri11","warp1","tri12","tri21","tri22","warp2","land1","land2","merge"] + [f"block{i}" for i in range(9)] + ["up0","up1"]
res = {}
for keep in (1, 0):
    for overlap in ((1,) if keep else (1, 0)):
        net = ap.define_G(3, 1, 64, ap.NETG_NAME, "instance", False, "normal", 0.02, [0], div=3, disp=3, precision="fp32_simt").module
        net.load_state_dict(sd)
        net.set_option("keep_intermediates", keep)
        net.set_option("overlap", overlap)
        with torch.no_grad():
            y = net(*inputs)
        torch.cuda.synchronize()
        taps = {k: net.debug_read(k).cpu() for k in names}
        taps["out"] = y.cpu()
        res[(keep, overlap)] = taps
ref = res[(1, 1)]
for key, taps in res.items():
    if key == (1, 1): continue
    print("keep,overlap =", key, {k: round((taps[k] - ref[k]).abs().max().item(), 5) for k in taps})
