"""Pin oracle/flow_oracle.py against the UNMODIFIED reference FlowUnet and write tests/golden/flow_*.npz.

Run in the build container only (needs /root/reference):    python tests/golden/make_flow_golden.py

The reference class is imported from /root/reference/Module2 with two shims for the environment -- `skimage.measure`
(absent) and `np.int` (removed from numpy 2; the class calls `.astype(np.int)`, intrinsic_flow_models/networks.py:605) --
built for every configuration below, loaded with the seeded stand-in checkpoint through `load_state_dict(strict=True)`
(so the key names / shapes of oracle.flow_oracle.state_dict_spec are checked against the real module), put in eval mode
like the caller does (geomcgt_ifw_test_model.py:216) and run on binary key-point maps made by the reference's own
`kp_to_map_some` recipe.  The oracle must reproduce flow_out and vis to 0.0 max-abs.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import cond_oracle as OC  # noqa: E402
from oracle import flow_oracle as FO  # noqa: E402

# name: (nf, start_scale, num_scale, norm, batch)
CASES = {
    "flow_bn_nf16_s2_n4": (16, 2, 4, "batch", 1),
    "flow_bn_nf32_s1_n5": (32, 1, 5, "batch", 1),
    "flow_in_nf16_s4_n3": (16, 4, 3, "instance", 2),
    "flow_bn_nf8_s2_n2": (8, 2, 2, "batch", 2),
}


def import_reference():
    if not hasattr(np, "int"):
        np.int = int  # noqa: NPY001  (what numpy < 1.24 exported)
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.measure")
    skm.compare_ssim = skm.compare_psnr = None
    sk.measure = skm
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.measure", skm)
    sys.path.insert(0, "/root/reference/Module2")
    from intrinsic_flow_models import networks  # type: ignore
    return networks


def kp_maps(B, seed):
    src, seq = OC.landmark_sequence(B, seed=seed)
    lm1 = np.repeat(src[None], B, 0) * 7 / 8
    lm2 = seq * 7 / 8
    return torch.from_numpy(np.concatenate([OC.kp_to_map(lm1), OC.kp_to_map(lm2)], 1))


def sample(t):
    return t[:, :, ::4, ::4].contiguous().numpy()


def main():
    nets = import_reference()
    for name, (nf, ss, ns, norm, B) in CASES.items():
        assert FO.consistent(224, ss, ns), name
        sd = FO.make_state_dict(136, nf, ss, ns, norm, seed=sum(map(ord, name)) % 1000)
        net = nets.FlowUnet(136, nf=nf, start_scale=ss, num_scale=ns, norm=norm)
        float_keys = [k for k in net.state_dict().keys() if not k.endswith("num_batches_tracked")]
        assert float_keys == list(sd.keys()), (name, [k for k in float_keys if k not in sd], [k for k in sd if k not in float_keys])
        net.load_state_dict(sd, strict=False)
        net.eval()
        x = kp_maps(B, seed=7 + B)
        with torch.no_grad():
            flow_ref, vis_ref, pyr, feat = net(x)
        flow, vis, flow0, feat_o = FO.flow_unet_forward(sd, x, nf, ss, ns, norm)
        d = max((flow - flow_ref).abs().max().item(), (vis - vis_ref).abs().max().item(), (feat_o - feat).abs().max().item())
        assert d == 0.0, (name, d)
        wf, wm = FO.warp_outputs(flow, vis)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), flow=sample(flow_ref), vis=sample(vis_ref), iw_flow=sample(wf),
                            if_mask=sample(wm), flow_abs_sum=np.float64(flow_ref.double().abs().sum().item()),
                            vis_abs_sum=np.float64(vis_ref.double().abs().sum().item()), seed=np.int64(sum(map(ord, name)) % 1000))
        print(f"{name}: oracle == reference (0.0), flow |max| {flow_ref.abs().max():.3f}, mask mean {wm.mean():.3f}")


if __name__ == "__main__":
    main()
