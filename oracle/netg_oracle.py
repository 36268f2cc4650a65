"""CPU oracle for the Module2 netG hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch CPU fp32 ops, no autograd, no nn.Module tree) of the
reference generator `ResnetConditionTriGenerator32_full_ifw`
(/root/reference/Module2/models/networks.py:1190-1340) and of the helpers it reaches
(`ResnetBlock` networks.py:2303-2361, `ResnetBlock2` networks.py:2363-2421,
`warp_acc_flow` intrinsic_flow_models/modules.py:596-625,
`get_norm_layer('instance')` networks.py:33-34).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it, and only as the checker / the CPU baseline.  The product package
(`animateportrait_b200`) never imports anything from `oracle/`.

Parity pinning: the reference has NO tests or golden vectors of its own.  The
oracle is pinned against the reference itself, imported in the build container from
/root/reference by `tests/golden/make_golden.py`; that script asserts oracle == reference to
<= 1e-6 max-abs on every fixture case and commits the reference outputs under `tests/golden/`.
`tests/test_oracle_golden.py` re-checks the oracle against those committed vectors on CPU.

All functions take the reference's 74-tensor state_dict as a plain dict.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

NETG_NAME = "resnet_9blocks_rcatland32_full_ifw"
N_BLOCKS = 9
DIV = 3
DISP = 3
NGF = 64
CON_DIM = 16
IN_EPS = 1e-5  # nn.InstanceNorm2d default, networks.py:34


# --------------------------------------------------------------------------------------
# state_dict layout (reference registration order: networks.py:1251 registers model_tri_merge
# first, the Sequentials follow at networks.py:1284-1295)
# --------------------------------------------------------------------------------------
def state_dict_spec(output_nc: int = 1, input_nc: int = 3, ngf: int = NGF):
    """Ordered {key: shape} of the generator's checkpoint."""
    spec = OrderedDict()

    def conv(prefix, cout, cin, k):
        spec[prefix + ".weight"] = (cout, cin, k, k)
        spec[prefix + ".bias"] = (cout,)

    def convT(prefix, cin, cout, k):
        spec[prefix + ".weight"] = (cin, cout, k, k)
        spec[prefix + ".bias"] = (cout,)

    conv("model_tri_merge", ngf * 4, ngf * 12, 3)
    conv("model_tri00.1", ngf // 2, input_nc, 7)
    conv("model_tri01.0", ngf * 2, ngf, 3)
    conv("model_tri02.0", ngf * 4, ngf * 2, 3)
    conv("model_tri10.1", ngf, input_nc, 7)
    conv("model_tri11.0", ngf, ngf, 3)
    conv("model_tri12.0", ngf * 4, ngf * 2, 3)
    conv("model_tri20.1", ngf, input_nc, 7)
    conv("model_tri21.0", ngf * 2, ngf, 3)
    conv("model_tri22.0", ngf * 2, ngf * 2, 3)
    dim = ngf * 4
    for i in range(N_BLOCKS):
        if (i + DISP) % DIV == 0:  # ResnetBlock2, networks.py:1259-1263
            conv(f"model2.{i}.conv_block.1", dim, dim + 2 * CON_DIM, 3)
            conv(f"model2.{i}.conv_block.5", dim, dim, 3)
            conv(f"model2.{i}.shortcut.0", dim, dim + 2 * CON_DIM, 3)
        else:
            conv(f"model2.{i}.conv_block.1", dim, dim, 3)
            conv(f"model2.{i}.conv_block.5", dim, dim, 3)
    convT("model3.0", ngf * 4, ngf * 2, 3)
    convT("model3.3", ngf * 2, ngf, 3)
    conv("model3.7", output_nc, ngf, 7)
    conv("model_landmark_trans.0", 8, 1, 3)
    conv("model_landmark_trans.3", CON_DIM, 8, 3)
    conv("model_landmark_trans.6", CON_DIM, CON_DIM, 3)
    return spec


def _key_seed(seed: int, key: str) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def make_state_dict(output_nc: int = 1, seed: int = 0, weight_std: float = 0.02,
                    bias_std: float = 0.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded stand-in checkpoint (no pretrained weights ship with the reference, readme.md:42-44).

    weights ~ N(0, weight_std) like `init_weights(..., 'normal', 0.02)` (networks.py:82-102);
    biases 0 (reference init) or N(0, bias_std) for the stress variant that catches a dropped
    `model_tri_merge.bias` / `model3.7.bias`.  Each tensor has its own
    generator so the recipe is independent of construction order.
    """
    sd = OrderedDict()
    for key, shape in state_dict_spec(output_nc).items():
        g = torch.Generator().manual_seed(_key_seed(seed, key))
        if key.endswith(".weight"):
            sd[key] = torch.randn(shape, generator=g, dtype=torch.float32) * weight_std
        elif bias_std > 0:
            sd[key] = torch.randn(shape, generator=g, dtype=torch.float32) * bias_std
        else:
            sd[key] = torch.zeros(shape, dtype=torch.float32)
    return sd


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def _smooth_field(B, C, S, coarse, std, g):
    z = torch.randn(B, C, coarse, coarse, generator=g) * std
    return F.interpolate(z, size=(S, S), mode="bilinear", align_corners=True)


def make_inputs(B: int = 1, seed: int = 1000, kind: str = "smooth", S: int = 256):
    """Seeded synthetic inputs of the shapes `netG.forward` takes (networks.py:1315).

    kind='smooth': identity motion grid + smooth displacement, smooth pixel flow times a binary
                   region, elliptical ifmask softened to [0,1], landmark disc maps in {-1,+1}.
    kind='noise' : the adversarial set (white-noise motion/flow, uniform ifmask, Bernoulli landmarks).
    Returns (input, land1, land2, motion, flow, ifmask).
    """
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    lin = torch.linspace(-1, 1, S)
    ys, xs = torch.meshgrid(lin, lin, indexing="ij")
    base = torch.stack([xs, ys], -1)[None].repeat(B, 1, 1, 1)  # (...,0)=x, (...,1)=y
    if kind == "noise":
        l1 = (torch.rand(B, 1, S, S, generator=g) > 0.98).float() * 2 - 1
        l2 = (torch.rand(B, 1, S, S, generator=g) > 0.98).float() * 2 - 1
        motion = base + 0.05 * torch.randn(B, S, S, 2, generator=g)
        flow = 4 * torch.randn(B, 2, S, S, generator=g)
        ifmask = torch.rand(B, 1, S, S, generator=g)
        return x, l1, l2, motion, flow, ifmask
    if kind != "smooth":
        raise ValueError(kind)

    def discs():
        # 68 filled discs of radius 3 (draw2 op 0, data/umlvdfw_test_dataset.py:35-41)
        pts = torch.rand(B, 68, 2, generator=g) * (S * 0.6) + S * 0.2
        yy, xx = torch.meshgrid(torch.arange(S).float(), torch.arange(S).float(), indexing="ij")
        d2 = (xx[None, None] - pts[..., 0, None, None]) ** 2 + (yy[None, None] - pts[..., 1, None, None]) ** 2
        return ((d2 <= 9.0).any(1, keepdim=True).float() * 2 - 1), pts

    l1, _ = discs()
    l2, _ = discs()
    motion = base + _smooth_field(B, 2, S, 8, 0.02, g).permute(0, 2, 3, 1)
    yy, xx = torch.meshgrid(lin, lin, indexing="ij")
    cx = torch.rand(B, 1, 1, 1, generator=g) * 0.2 - 0.1
    cy = torch.rand(B, 1, 1, 1, generator=g) * 0.2 - 0.1
    ell = (((xx[None, None] - cx) / 0.55) ** 2 + ((yy[None, None] - cy) / 0.75) ** 2 <= 1.0).float()
    ifmask = F.avg_pool2d(ell, 5, stride=1, padding=2).clamp(0, 1)
    flow = _smooth_field(B, 2, S, 8, 4.0, g) * ell
    return x, l1, l2, motion.contiguous(), flow.contiguous(), ifmask.contiguous()


# --------------------------------------------------------------------------------------
# the generator, functionally
# --------------------------------------------------------------------------------------
def instance_norm(x: torch.Tensor) -> torch.Tensor:
    """nn.InstanceNorm2d(affine=False, track_running_stats=False) (networks.py:34): per (n,c)
    biased variance over HxW, eps 1e-5; identical in train and eval mode."""
    return F.instance_norm(x, eps=IN_EPS)


def _conv(sd, key, x, stride=1, padding=0):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=padding)


def _refpad(x, p):
    return F.pad(x, (p, p, p, p), mode="reflect")


def warp_acc_flow(x, flow, mask=None, mask_value=-1.0):
    """intrinsic_flow_models/modules.py:596-625: pixel-unit flow -> normalised grid
    (`2*g/(W-1)-1`, :615-616) -> grid_sample bilinear/zeros with the DEFAULT align_corners=False
    (:619) -> where(mask>0.5, out, -1) (:623-624)."""
    B, C, H, W = x.shape
    xx = torch.arange(W, dtype=x.dtype).view(1, -1).repeat(H, 1).view(1, 1, H, W).repeat(B, 1, 1, 1)
    yy = torch.arange(H, dtype=x.dtype).view(-1, 1).repeat(1, W).view(1, 1, H, W).repeat(B, 1, 1, 1)
    grid = torch.cat((xx, yy), 1).float() + flow
    grid[:, 0] = 2.0 * grid[:, 0] / max(W - 1, 1) - 1.0
    grid[:, 1] = 2.0 * grid[:, 1] / max(H - 1, 1) - 1.0
    out = F.grid_sample(x, grid.permute(0, 2, 3, 1), mode="bilinear", padding_mode="zeros",
                        align_corners=False)
    if mask is not None:
        out = torch.where(mask > 0.5, out, out.new_ones(1).mul_(mask_value))
    return out


def double_feature_warping(x, motion, flow, ifmask, level):
    """networks.py:1298-1313."""
    if level in (1, 2):
        S = 128 if level == 1 else 64
        motion = F.interpolate(motion.permute(0, 3, 1, 2), size=(S, S), mode="bilinear",
                               align_corners=True).permute(0, 2, 3, 1)
        flow = F.interpolate(flow / (2 ** level), size=(S, S), mode="bilinear", align_corners=True)
        ifmask = F.interpolate(ifmask, size=(S, S), mode="bilinear", align_corners=True)
    x1 = F.grid_sample(x, motion, mode="bilinear", padding_mode="zeros", align_corners=False)
    x2 = warp_acc_flow(x, flow, mask=ifmask)
    return torch.cat([x1, x2], 1)


def _stem(sd, name, x):
    """RefPad3 + Conv7 + IN + ReLU (networks.py:1218-1221 / 1229-1232 / 1240-1243)."""
    return F.relu(instance_norm(_conv(sd, name + ".1", _refpad(x, 3))))


def _down(sd, name, x):
    """Conv3 s2 p1 + IN + ReLU (networks.py:1222-1227 etc.)."""
    return F.relu(instance_norm(_conv(sd, name + ".0", x, stride=2, padding=1)))


def landmark_trans(sd, land):
    """networks.py:1280-1282."""
    p = "model_landmark_trans"
    y = F.relu(instance_norm(_conv(sd, p + ".0", land, 1, 1)))
    y = F.relu(instance_norm(_conv(sd, p + ".3", y, 2, 1)))
    return instance_norm(_conv(sd, p + ".6", y, 2, 1))


def resnet_block(sd, p, x):
    """networks.py:2303-2361 (reflect padding, no dropout)."""
    y = F.relu(instance_norm(_conv(sd, p + ".conv_block.1", _refpad(x, 1))))
    y = instance_norm(_conv(sd, p + ".conv_block.5", _refpad(y, 1)))
    return x + y


def resnet_block2(sd, p, x):
    """networks.py:2363-2421: zero-padded conv shortcut + IN, reflect-padded main branch."""
    s = instance_norm(_conv(sd, p + ".shortcut.0", x, 1, 1))
    y = F.relu(instance_norm(_conv(sd, p + ".conv_block.1", _refpad(x, 1))))
    y = instance_norm(_conv(sd, p + ".conv_block.5", _refpad(y, 1)))
    return s + y


def decoder(sd, x):
    """model3, networks.py:1268-1279."""
    y = F.conv_transpose2d(x, sd["model3.0.weight"], sd["model3.0.bias"], stride=2, padding=1,
                           output_padding=1)
    y = F.relu(instance_norm(y))
    y = F.conv_transpose2d(y, sd["model3.3.weight"], sd["model3.3.bias"], stride=2, padding=1,
                           output_padding=1)
    y = F.relu(instance_norm(y))
    pre = _conv(sd, "model3.7", _refpad(y, 3))
    return pre


@torch.no_grad()
def netg_forward(sd: Dict[str, torch.Tensor], input, land1, land2, motion, flow, ifmask,
                 tap: Optional[Callable[[str, torch.Tensor], None]] = None) -> torch.Tensor:
    """`ResnetConditionTriGenerator32_full_ifw.forward` (networks.py:1315-1340).

    `tap(name, tensor)` (optional) receives named NCHW intermediates; the names are the ones the
    CUDA library's debug tap API (`ap_netg_debug_read`, include/ap_netg.h) uses.
    """
    t = tap if tap is not None else (lambda n, v: None)
    x1 = _stem(sd, "model_tri00", input); t("tri00", x1)
    x1 = double_feature_warping(x1, motion, flow, ifmask, 0); t("warp0", x1)
    x1 = _down(sd, "model_tri01", x1); t("tri01", x1)
    x1 = _down(sd, "model_tri02", x1); t("tri02", x1)
    x2 = _stem(sd, "model_tri10", input); t("tri10", x2)
    x2 = _down(sd, "model_tri11", x2); t("tri11", x2)
    x2 = double_feature_warping(x2, motion, flow, ifmask, 1); t("warp1", x2)
    x2 = _down(sd, "model_tri12", x2); t("tri12", x2)
    x3 = _stem(sd, "model_tri20", input); t("tri20", x3)
    x3 = _down(sd, "model_tri21", x3); t("tri21", x3)
    x3 = _down(sd, "model_tri22", x3); t("tri22", x3)
    x3 = double_feature_warping(x3, motion, flow, ifmask, 2); t("warp2", x3)
    x = _conv(sd, "model_tri_merge", torch.cat([x1, x2, x3], 1), 1, 1); t("merge", x)
    l1 = landmark_trans(sd, land1); t("land1", l1)
    l2 = landmark_trans(sd, land2); t("land2", l2)
    for i in range(N_BLOCKS):
        if (i + DISP) % DIV == 0:
            x = resnet_block2(sd, f"model2.{i}", torch.cat([x, l1, l2], 1))
        else:
            x = resnet_block(sd, f"model2.{i}", x)
        t(f"block{i}", x)
    pre = decoder(sd, x); t("pre_tanh", pre)
    return torch.tanh(pre)


# --------------------------------------------------------------------------------------
# closed-form restatement of the warps.  This is the arithmetic the
# CUDA gather kernel follows; tests check it against the grid_sample formulation above.
# --------------------------------------------------------------------------------------
def _resize_ac_true(v: torch.Tensor, S: int) -> torch.Tensor:
    """F.interpolate(bilinear, align_corners=True) written out: src = dst * fp32((in-1)/(out-1)),
    i0 = floor(src), l1 = src - i0, l0 = 1 - l1; out = l0h*(l0w*v00 + l1w*v01) + l1h*(l0w*v10 + l1w*v11)."""
    B, C, H, W = v.shape
    if H == S and W == S:
        return v
    sc_h = torch.tensor((H - 1) / (S - 1), dtype=torch.float32)
    sc_w = torch.tensor((W - 1) / (S - 1), dtype=torch.float32)
    sy = torch.arange(S, dtype=torch.float32) * sc_h
    sx = torch.arange(S, dtype=torch.float32) * sc_w
    y0 = sy.floor().long().clamp(max=H - 1); x0 = sx.floor().long().clamp(max=W - 1)
    y1 = (y0 + 1).clamp(max=H - 1); x1 = (x0 + 1).clamp(max=W - 1)
    ly1 = (sy - y0.float()).view(1, 1, S, 1); lx1 = (sx - x0.float()).view(1, 1, 1, S)
    ly0 = 1 - ly1; lx0 = 1 - lx1
    v00 = v[:, :, y0][:, :, :, x0]; v01 = v[:, :, y0][:, :, :, x1]
    v10 = v[:, :, y1][:, :, :, x0]; v11 = v[:, :, y1][:, :, :, x1]
    return ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11)


def _bilinear_zeros(x: torch.Tensor, ix: torch.Tensor, iy: torch.Tensor) -> torch.Tensor:
    """4-tap bilinear gather at pixel coordinates (ix, iy) [B,S,S]; taps outside contribute 0."""
    B, C, H, W = x.shape
    x0 = ix.floor(); y0 = iy.floor()
    wx1 = ix - x0; wy1 = iy - y0
    wx0 = 1 - wx1; wy0 = 1 - wy1
    out = torch.zeros(B, C, ix.shape[1], ix.shape[2], dtype=x.dtype)
    flat = x.reshape(B, C, H * W)
    for (yy, xx, ww) in ((y0, x0, wy0 * wx0), (y0, x0 + 1, wy0 * wx1),
                         (y0 + 1, x0, wy1 * wx0), (y0 + 1, x0 + 1, wy1 * wx1)):
        ok = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long().view(B, 1, -1).expand(B, C, -1)
        val = torch.gather(flat, 2, idx).view(B, C, ix.shape[1], ix.shape[2])
        out = out + val * (ww * ok.float()).unsqueeze(1)
    return out


def double_feature_warping_closed_form(x, motion, flow, ifmask, level):
    B, C, S, _ = x.shape
    m = _resize_ac_true(motion.permute(0, 3, 1, 2), S)
    f = _resize_ac_true(flow / (2 ** level), S)
    k = _resize_ac_true(ifmask, S)
    # motion warp: align_corners=False un-normalisation  ix = ((g + 1) * S - 1) / 2
    ix = ((m[:, 0] + 1) * S - 1) / 2
    iy = ((m[:, 1] + 1) * S - 1) / 2
    x1 = _bilinear_zeros(x, ix, iy)
    # flow warp: gx = 2*(j + fx)/(S-1) - 1, then the same un-normalisation (reference op order)
    jj = torch.arange(S, dtype=torch.float32).view(1, 1, S)
    ii = torch.arange(S, dtype=torch.float32).view(1, S, 1)
    gx = 2.0 * (jj + f[:, 0]) / (S - 1) - 1.0
    gy = 2.0 * (ii + f[:, 1]) / (S - 1) - 1.0
    x2 = _bilinear_zeros(x, ((gx + 1) * S - 1) / 2, ((gy + 1) * S - 1) / 2)
    x2 = torch.where(k > 0.5, x2, torch.full_like(x2, -1.0))
    return torch.cat([x1, x2], 1)


# --------------------------------------------------------------------------------------
# output stage after netG ("next" rows f1 / f4): blend with the static drawing and uint8 conversion
# --------------------------------------------------------------------------------------
def blend_foreground(fake_B, mask, warp_motion, fakeB_static):
    """GeomCGTIFWTestModel.forward, Module2/models/geomcgt_ifw_test_model.py:297-300."""
    mask1 = F.grid_sample(mask, warp_motion, align_corners=True)
    return ((fake_B / 2 + 0.5) * mask1 + (fakeB_static / 2 + 0.5) * (1 - mask1)) * 2 - 1


def tensor2im_batch(x: torch.Tensor):
    """util.tensor2im (Module2/util/util.py:9-29) applied to every frame of a batch: uint8 [B,H,W,3]."""
    import numpy as np
    arr = x.detach().cpu().float().numpy()
    if arr.shape[1] == 1:  # grayscale to RGB
        arr = np.tile(arr, (1, 3, 1, 1))
    img = (np.transpose(arr, (0, 2, 3, 1)) + 1) / 2.0 * 255.0
    return img.astype(np.uint8)


def make_compose_inputs(B: int, onc: int, seed: int):
    """Seeded stand-ins for (fake_B, mask, warp_motion, fakeB_static): a tanh-range frame, a binary matte
    ((matte > 0.5).float(), geomcgt_ifw_test_model.py:280), the smooth motion grid of make_inputs, a static drawing."""
    g = torch.Generator().manual_seed(seed)
    fake = torch.tanh(_smooth_field(B, onc, 256, 16, 1.5, g))
    stat = torch.tanh(_smooth_field(B, onc, 256, 32, 1.5, g))
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, 256), torch.linspace(-1, 1, 256), indexing="ij")
    cx, cy = 0.1 * torch.randn(B, 1, 1, generator=g), 0.1 * torch.randn(B, 1, 1, generator=g)
    mask = ((((xs[None] - cx) / 0.6) ** 2 + ((ys[None] - cy) / 0.8) ** 2) < 1.0).float()[:, None]
    motion = make_inputs(B, seed=seed + 1, kind="smooth")[3]
    return fake, mask, motion, stat


def flops_per_frame(output_nc: int = 1) -> float:
    """2*MACs of the 38 Conv2d + 2 ConvTranspose2d calls per frame: 140.125e9 / 140.947e9."""
    total = 0.0
    res = {"model_tri00.1": 256, "model_tri01.0": 128, "model_tri02.0": 64, "model_tri10.1": 256,
           "model_tri11.0": 128, "model_tri12.0": 64, "model_tri20.1": 256, "model_tri21.0": 128,
           "model_tri22.0": 64, "model_tri_merge": 64, "model3.0": 64, "model3.3": 128, "model3.7": 256,
           "model_landmark_trans.0": 256, "model_landmark_trans.3": 128, "model_landmark_trans.6": 64}
    for key, shape in state_dict_spec(output_nc).items():
        if not key.endswith(".weight"):
            continue
        name = key[:-7]
        macs_per_px = shape[0] * shape[1] * shape[2] * shape[3]
        if name.startswith("model2."):
            r = 64
        else:
            r = res[name]
        n = 2 if name.startswith("model_landmark_trans") else 1
        # model3.0/.3 are transposed convs: MACs counted on the INPUT resolution (Appendix A.4)
        total += 2.0 * macs_per_px * r * r * n
    return total
