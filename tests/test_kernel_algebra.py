"""CPU checks of the two algebraic reformulations the CUDA kernels rely on (pure torch, no GPU): they document the
kernels' math and guard the index conventions used by the weight packers in animateportrait_b200/csrc.

1. ConvTranspose2d(k3,s2,p1,op1) == ONE stride-1 conv over the 2x2 input taps (dy,dx) in {0,1}^2 whose output channels
   pack the four output phases (csrc/common.cuh: PhasePack; csrc/elementwise.cu: pack_convT_phases_kernel).
   Reference layer: Module2/models/networks.py:1271-1274.
2. ReflectionPad2d(3) + Conv2d(64 -> onc, 7x7) == a per-pixel [64 x 49] channel contraction followed by a 49-term
   shifted sum (csrc/conv_out.cu).  Reference layer: Module2/models/networks.py:1277-1278.
"""
import torch
import torch.nn.functional as F


def convT_k(p, d):
    """kernel row used by output phase p (0/1) at input offset d (0/1); -1 = none (same rule as the CUDA packer)"""
    return (1 if d == 0 else -1) if p == 0 else (0 if d == 1 else 2)


def test_transposed_conv_equals_phase_packed_stride1_conv():
    g = torch.Generator().manual_seed(0)
    B, Cin, Cout, H = 2, 16, 8, 12
    x = torch.randn(B, Cin, H, H, generator=g, dtype=torch.float64)
    w = torch.randn(Cin, Cout, 3, 3, generator=g, dtype=torch.float64)  # ConvTranspose2d layout [Cin, Cout, kh, kw]
    ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    # packed weights [tap t = dy*2+dx][phase ph = py*2+px][Cout][Cin]
    wp = torch.zeros(4, 4, Cout, Cin, dtype=torch.float64)
    for t in range(4):
        dy, dx = t >> 1, t & 1
        for ph in range(4):
            ky, kx = convT_k(ph >> 1, dy), convT_k(ph & 1, dx)
            if ky >= 0 and kx >= 0:
                wp[t, ph] = w[:, :, ky, kx].t()
    # stride-1 conv over taps (dy,dx) with zero padding on the bottom/right edge (what TMA out-of-bounds fill provides)
    xp = F.pad(x, (0, 1, 0, 1))
    out = torch.zeros(B, 4, Cout, H, H, dtype=torch.float64)
    for t in range(4):
        dy, dx = t >> 1, t & 1
        patch = xp[:, :, dy:dy + H, dx:dx + H]
        out += torch.einsum("bchw,poc->bpohw", patch, wp[t])
    # the epilogue routes phase (py,px) to output pixels (2i+py, 2j+px)
    got = torch.zeros_like(ref)
    for ph in range(4):
        got[:, :, (ph >> 1)::2, (ph & 1)::2] = out[:, ph]
    assert (got - ref).abs().max().item() < 1e-12
    # the packing wastes 7 of 16 (tap, phase) blocks: they must be exactly the all-zero ones
    zero_blocks = sum(int(wp[t, ph].abs().sum() == 0) for t in range(4) for ph in range(4))
    assert zero_blocks == 7


def test_output_conv_equals_channel_contraction_plus_shifted_sums():
    g = torch.Generator().manual_seed(1)
    B, C, onc, S = 1, 64, 3, 20
    a = torch.randn(B, C, S, S, generator=g, dtype=torch.float64)
    w = torch.randn(onc, C, 7, 7, generator=g, dtype=torch.float64)
    bias = torch.randn(onc, generator=g, dtype=torch.float64)
    ref = torch.tanh(F.conv2d(F.pad(a, (3, 3, 3, 3), mode="reflect"), w, bias))
    ap = F.pad(a, (3, 3, 3, 3), mode="reflect")                     # patch positions incl. the reflected halo
    # t[o][j = ky*7+kx][position] = sum_c a[c][position] * w[o][c][ky][kx]   (the tcgen05 GEMM, N = 49 padded to 64)
    t = torch.einsum("bchw,ocj->bojhw", ap, w.reshape(onc, C, 49))
    out = torch.zeros(B, onc, S, S, dtype=torch.float64)
    for ky in range(7):
        for kx in range(7):
            out += t[:, :, ky * 7 + kx, ky:ky + S, kx:kx + S]      # the 49-term shifted sum out of shared memory
    got = torch.tanh(out + bias.view(1, onc, 1, 1))
    assert (got - ref).abs().max().item() < 1e-12
