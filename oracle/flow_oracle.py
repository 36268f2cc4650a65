"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's flow network path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

What it restates, with plain torch functional ops on CPU fp32 (no module tree):
  * `FlowUnet.forward`                  Module2/intrinsic_flow_models/networks.py:577-644
  * `FlowUnetSkipConnectionBlock`       Module2/intrinsic_flow_models/networks.py:510-575, including its in-place
    activations: `nn.LeakyReLU(0.2, True)` at the head of `down` rewrites the block's input, so the skip half of
    `torch.cat((x, x_), 1)` is LeakyReLU_0.2(x), and the parent's in-place `nn.ReLU(True)` then acts on that
  * `flow_network_warp`                 Module2/models/geomcgt_ifw_test_model.py:62-76 (key-point maps at 7/8 scale,
    visibility arg-max -> mask, flow * 20 * mask * 8/7, bilinear 224 -> 256 with align_corners=True)
  * `kp_to_map_some`                    lives in oracle/cond_oracle.py (row f2)

Pinning: tests/golden/make_flow_golden.py imports the reference class itself (with an `np.int` alias, removed from
numpy 2) in the build container, checks this restatement against it to 0.0 max-abs on seeded weights for every
configuration that is shape-consistent on the 224x224 input the caller feeds, and commits sampled outputs under
tests/golden/flow_*.npz.  The configuration the released checkpoint uses lives in `checkpoints/FlowReg_id_flow_faces/
train_opt.json`, which does not ship: the network is therefore parametric in exactly the arguments
`FlowRegressionModel.initialize` passes (flow_regression_model.py:19-38).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

EPS = 1e-5
UPSAMPLE = 2  # networks.py:583


def level_channels(nf: int, start_scale: int, num_scale: int, max_nf: int = 512) -> Tuple[int, List[Tuple[int, int]]]:
    """(channels entering the U-Net, [(outer_nc, inner_nc) per level, outermost first])  -- networks.py:605-620."""
    nc = nf * start_scale  # doubled log2(start_scale) times
    return nc, [(min(max_nf, nc * 2 ** l), min(max_nf, nc * 2 ** (l + 1))) for l in range(num_scale)]


def consistent(size: int, start_scale: int, num_scale: int) -> bool:
    """Every level halves exactly (4x4 stride-2 convs and their transposes must round-trip for the skip concat)."""
    s = size
    for _ in range(int(math.log2(start_scale))):
        s = (s + 2 - 3) // 2 + 1
    for _ in range(num_scale):
        if s % 2:
            return False
        s //= 2
    return s >= 1


def state_dict_spec(input_nc: int, nf: int, start_scale: int, num_scale: int, norm: str, max_nf: int = 512):
    """[(key, shape)] of the reference FlowUnet's float tensors, in registration order."""
    bias = norm == "instance"
    spec = []

    def conv(key, cout, cin, k, with_bias, transposed=False):
        spec.append((key + ".weight", (cin, cout, k, k) if transposed else (cout, cin, k, k)))
        if with_bias:
            spec.append((key + ".bias", (cout,)))

    def normp(key, c):
        if norm == "batch":
            spec.extend([(key + ".weight", (c,)), (key + ".bias", (c,)), (key + ".running_mean", (c,)),
                         (key + ".running_var", (c,))])

    conv("conv_downsample.0", nf, input_nc, 7, bias)
    normp("conv_downsample.1", nf)
    nc = nf
    for i in range(int(math.log2(start_scale))):
        conv(f"conv_downsample.{3 * (i + 1)}", 2 * nc, nc, 3, bias)
        normp(f"conv_downsample.{3 * (i + 1) + 1}", 2 * nc)
        nc *= 2
    _, levels = level_channels(nf, start_scale, num_scale, max_nf)

    def block(l, prefix):
        outer, inner = levels[l]
        outermost, innermost = l == 0, l == num_scale - 1
        d = 0 if outermost else 1
        conv(f"{prefix}down.{d}", inner, outer, 4, bias)
        if not innermost:
            normp(f"{prefix}down.{d + 1}", inner)
        conv(f"{prefix}up.1", outer, inner if innermost else inner * 2, 4, True if outermost else bias, transposed=True)
        normp(f"{prefix}up.2", outer)
        if not innermost:
            block(l + 1, prefix + "submodule.")
        conv(f"{prefix}predict_flow.1", 2, outer, 3, True)

    block(0, "unet_block.")
    conv("predict_vis.1", 3, nc, 3, True)
    return spec


def make_state_dict(input_nc=136, nf=16, start_scale=2, num_scale=4, norm="batch", max_nf=512, seed=0, gain=0.05):
    """Seeded stand-in checkpoint: N(0, gain) weights, N(0, 0.1) biases, non-trivial BatchNorm statistics."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape in state_dict_spec(input_nc, nf, start_scale, num_scale, norm, max_nf):
        if key.endswith("running_var"):
            sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif key.endswith("running_mean"):
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 1 and key.endswith(".weight"):
            sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.75      # BatchNorm gamma
        elif len(shape) == 1:
            sd[key] = torch.randn(shape, generator=g) * 0.1
        else:
            sd[key] = torch.randn(shape, generator=g) * gain
    return sd


def _norm(x, sd, key, norm):
    if norm == "batch":  # eval mode: running statistics (geomcgt_ifw_test_model.py:216 calls netF.eval())
        return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"],
                            False, 0.0, EPS)
    return F.instance_norm(x, eps=EPS)


def flow_unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, nf: int, start_scale: int, num_scale: int, norm: str,
                      max_nf: int = 512, tap=None):
    """networks.py:629-644: returns (flow_out, vis, flow_pyr[0], feat_out)."""
    b = lambda k: sd.get(k + ".bias")  # noqa: E731
    x = F.leaky_relu(_norm(F.conv2d(x, sd["conv_downsample.0.weight"], b("conv_downsample.0"), padding=3), sd,
                           "conv_downsample.1", norm), 0.1)
    for i in range(int(math.log2(start_scale))):
        k = f"conv_downsample.{3 * (i + 1)}"
        x = F.leaky_relu(_norm(F.conv2d(x, sd[k + ".weight"], b(k), stride=2, padding=1), sd,
                               f"conv_downsample.{3 * (i + 1) + 1}", norm), 0.1)
    if tap:
        tap("x0", x)

    def block(l, prefix, x):
        outermost, innermost = l == 0, l == num_scale - 1
        d = 0 if outermost else 1
        if not outermost:
            x = F.leaky_relu(x, 0.2)  # in place in the reference: the skip connection below sees it
        y = F.conv2d(x, sd[f"{prefix}down.{d}.weight"], b(f"{prefix}down.{d}"), stride=2, padding=1)
        if not innermost:
            y = _norm(y, sd, f"{prefix}down.{d + 1}", norm)
            y = block(l + 1, prefix + "submodule.", y)
        y = F.relu(y)
        y = F.conv_transpose2d(y, sd[f"{prefix}up.1.weight"], b(f"{prefix}up.1"), stride=2, padding=1)
        y = _norm(y, sd, f"{prefix}up.2", norm)
        if tap:
            tap(f"u{l}", y)
        if outermost:
            return y
        return torch.cat((x, y), 1)

    feat = block(0, "unet_block.", x)
    flow0 = F.conv2d(F.leaky_relu(feat, 0.1), sd["unet_block.predict_flow.1.weight"], sd["unet_block.predict_flow.1.bias"], padding=1)
    vis0 = F.conv2d(F.leaky_relu(feat, 0.1), sd["predict_vis.1.weight"], sd["predict_vis.1.bias"], padding=1)
    # the reference up-samples by `self.start_scale`, which its constructor hard-codes to 2 whatever the argument was
    # (networks.py:583 `self.start_scale = 2`, used at :641-642): outputs are 2x the U-Net resolution
    flow_out = F.interpolate(flow0, scale_factor=UPSAMPLE, mode="bilinear", align_corners=False)
    vis = F.interpolate(vis0, scale_factor=UPSAMPLE, mode="bilinear", align_corners=False)
    return flow_out, vis, flow0, feat


def warp_outputs(flow_out: torch.Tensor, vis: torch.Tensor):
    """geomcgt_ifw_test_model.py:68-75: (iw_flow [B,2,256,256], real_A_if_mask [B,1,256,256])."""
    vis_out = vis.argmax(dim=1, keepdim=True).float()
    mask_out = (vis_out < 2).float()
    flow = flow_out * 20.0 * mask_out
    warp_flow = F.interpolate(flow / 7 * 8, size=(256, 256), mode="bilinear", align_corners=True)
    res_mask = F.interpolate(mask_out, size=(256, 256), mode="bilinear", align_corners=True)
    return warp_flow, res_mask
