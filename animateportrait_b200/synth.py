"""Synthetic workload of the Module2 netG path: seeded stand-in checkpoints and input batches.

The reference ships no pretrained checkpoints (readme.md:42-44, Google-Drive links only) and the GPU box has no
network, so benchmarks and smoke tests run on seeded weights of the reference architecture and on synthetic
(photo, landmark maps, motion grid, intrinsic flow, visibility mask) batches shaped like the tensors
`GeomCGTIFWTestModel.forward` hands to netG (Module2/models/geomcgt_ifw_test_model.py:295).  SURVEY.md §8d
describes the recipes.  `oracle/netg_oracle.py` carries its own identical copy for the tests (the product package
never imports the oracle); tests/test_host_module.py checks that the two stay bit-identical.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

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


# --------------------------------------------------------------------------------------
# synthetic clip: one photo + a landmark trajectory standing in for Module1's output
# --------------------------------------------------------------------------------------
def face_template():
    """A plausible 68-point layout in the 256x256 window (jaw arc, brows, nose, eyes, two mouth rings), float32 [68,2]."""
    import numpy as np
    p = []
    for i in range(17):                      # jaw
        a = np.pi * (0.08 + 0.84 * i / 16)
        p.append([128 - 78 * np.cos(a), 118 + 92 * np.sin(a)])
    for i in range(5):
        p.append([62 + 11 * i, 92 - 6 * np.sin(np.pi * i / 4)])    # right brow
    for i in range(5):
        p.append([150 + 11 * i, 92 - 6 * np.sin(np.pi * i / 4)])   # left brow
    for i in range(4):
        p.append([128, 104 + 11 * i])                              # nose bridge
    for i in range(5):
        p.append([110 + 9 * i, 150 + 3 * np.sin(np.pi * i / 4)])   # nostrils
    for cx in (84, 172):                                           # eyes
        for i in range(6):
            a = 2 * np.pi * i / 6
            p.append([cx - 13 * np.cos(a), 112 - 6 * np.sin(a)])
    for i in range(12):                                            # outer lips
        a = 2 * np.pi * i / 12
        p.append([128 - 28 * np.cos(a), 182 - 11 * np.sin(a)])
    for i in range(8):                                             # inner lips
        a = 2 * np.pi * i / 8
        p.append([128 - 17 * np.cos(a), 182 - 5 * np.sin(a)])
    return np.asarray(p, dtype=np.float32)


def landmark_sequence(T: int, seed: int = 0, amp: float = 4.0):
    """(source landmarks [68,2], target sequence [T,68,2]) float32 tensors: template + jitter, then a smooth head sway
    plus a mouth opening cycle (Module1 cannot run here).  Same recipe as oracle/cond_oracle.py."""
    import numpy as np
    rng = np.random.RandomState(seed)
    src = face_template() + rng.normal(0, 0.7, (68, 2)).astype(np.float32)
    t = np.arange(T, dtype=np.float32)[:, None, None]
    sway = np.concatenate([amp * np.sin(2 * np.pi * t / 90.0), 0.5 * amp * np.sin(2 * np.pi * t / 57.0 + 1.0)], 2)
    seq = src[None] + sway
    mouth = np.zeros((1, 68, 2), dtype=np.float32)
    mouth[0, 48:68, 1] = (src[48:68, 1] - 182.0) * 0.6
    seq = seq + mouth * (0.5 + 0.5 * np.sin(2 * np.pi * t / 11.0))
    seq = seq + rng.normal(0, 0.15, seq.shape)
    return torch.from_numpy(src.astype(np.float32)), torch.from_numpy(seq.astype(np.float32))


def make_clip(T: int, output_nc: int = 1, seed: int = 2000):
    """Synthetic clip: photo [1,3,256,256], matte [1,1,256,256], static drawing [1,onc,256,256], source landmarks [68,2],
    target landmarks [T,68,2], and per-frame intrinsic flow [T,2,256,256] / visibility mask [T,1,256,256] standing in
    for the flow network's output (netF, SURVEY.md §8 row f3, is not built)."""
    g = torch.Generator().manual_seed(seed)
    photo = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    lin = torch.linspace(-1, 1, 256)
    yy, xx = torch.meshgrid(lin, lin, indexing="ij")
    ell = (((xx / 0.7) ** 2 + ((yy - 0.05) / 0.85) ** 2) <= 1.0).float()[None, None]
    matte = F.avg_pool2d(ell, 9, stride=1, padding=4).clamp(0, 1)
    static = torch.rand(1, output_nc, 256, 256, generator=g) * 2 - 1
    src, seq = landmark_sequence(T, seed=seed)
    coarse = torch.randn(8, 2, 8, 8, generator=g) * 4.0      # 8 key flows, cycled and blended along the clip
    key = F.interpolate(coarse, size=(256, 256), mode="bilinear", align_corners=True) * ell
    w = (torch.arange(T, dtype=torch.float32) / 12.0)
    i0 = w.floor().long() % 8
    a = (w - w.floor())[:, None, None, None]
    flow = (1 - a) * key[i0] + a * key[(i0 + 1) % 8]
    ifmask = matte.expand(T, 1, 256, 256).contiguous()
    return photo, matte, static, src, seq, flow.contiguous(), ifmask
