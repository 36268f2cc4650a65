"""CPU checks of the two algebraic reformulations the CUDA kernels rely on (pure torch, no GPU): they document the
kernels' math and guard the index conventions used by the weight packers in animateportrait_b200/csrc.

1. ConvTranspose2d(k3,s2,p1,op1) == ONE stride-1 conv over the 2x2 input taps (dy,dx) in {0,1}^2 whose output channels
   pack the four output phases (csrc/common.cuh: PhasePack; csrc/elementwise.cu: pack_convT_phases_kernel).
   Reference layer: Module2/models/networks.py:1271-1274.
2. ReflectionPad2d(3) + Conv2d(64 -> onc, 7x7) == a per-pixel [64 x 49] channel contraction followed by a 49-term
   shifted sum (csrc/conv_out.cu).  Reference layer: Module2/models/networks.py:1277-1278.
3. netF (csrc/flownet.cu): the zero-skipping first conv == the dense conv; the shared-memory offsets of the tensor-core
   conv's builder warps and the host weight packer are the K-major SWIZZLE_128B layout, and the three-product bf16 split
   is fp32-accurate; the key-point boxes derived from coordinates contain their discs (csrc/conditioning.cu).
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


def test_sparse_first_conv_of_the_flow_network_equals_the_dense_conv():
    """csrc/flownet.cu: fbbox_kernel + fconv_sparse_kernel.  The first conv of netF (7x7, pad 3; reference layer
    Module2/intrinsic_flow_models/networks.py:601) reads key-point discs: per output pixel only the (channel, tap) pairs
    whose input lies inside the channel's non-zero bounding box are visited.  This restates the kernel's loops (tile-level
    box test, per-pixel window, slab index into the packed weights [ky*k+kx][Cin][Cout]) in numpy and compares with
    F.conv2d, including an empty plane, a disc cut by the border and a plane that is dense."""
    import numpy as np
    g = torch.Generator().manual_seed(3)
    H = W = 40
    Cin, Cout, k, pad = 6, 5, 7, 3
    x = torch.zeros(1, Cin, H, W, dtype=torch.float64)
    yy, xx = np.mgrid[0:H, 0:W]
    for c, (cy, cx) in enumerate([(10.3, 12.7), (1.2, 38.9), (39.0, 0.0), (20.5, 20.5)]):
        x[0, c] = torch.from_numpy(((xx - cx) ** 2 + (yy - cy) ** 2 <= 16).astype(np.float64))
    x[0, 5] = torch.randn(H, W, generator=g, dtype=torch.float64)   # dense plane; plane 4 stays empty
    w = torch.randn(Cout, Cin, k, k, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, padding=pad)[0].numpy()
    packed = w.permute(2, 3, 1, 0).reshape(k * k, Cin, Cout).numpy()   # [slab][Cin][Cout], slab = ky*k + kx
    xn = x[0].numpy()
    bbox = []
    for c in range(Cin):
        nz = np.argwhere(xn[c] != 0)
        bbox.append((H, -1, W, -1) if len(nz) == 0 else (nz[:, 0].min(), nz[:, 0].max(), nz[:, 1].min(), nz[:, 1].max()))
    out = np.zeros((Cout, H, W))
    visited = 0
    for ty0 in range(0, H, 16):
        for tx0 in range(0, W, 16):
            ry0, ry1, rx0, rx1 = ty0 - pad, ty0 + 15 - pad + k - 1, tx0 - pad, tx0 + 15 - pad + k - 1
            for c in range(Cin):
                y0, y1, x0, x1 = bbox[c]
                if y1 < ry0 or y0 > ry1 or x1 < rx0 or x0 > rx1:
                    continue
                for y in range(ty0, min(ty0 + 16, H)):
                    for xq in range(tx0, min(tx0 + 16, W)):
                        wy0, wx0 = y - pad, xq - pad
                        for iy in range(max(wy0, y0), min(wy0 + k - 1, y1) + 1):
                            for ix in range(max(wx0, x0), min(wx0 + k - 1, x1) + 1):
                                v = xn[c, iy, ix]
                                if v != 0:
                                    visited += 1
                                    out[:, y, xq] += v * packed[(iy - wy0) * k + (ix - wx0), c]
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    assert visited < 0.5 * H * W * Cin * k * k   # and most of the dense conv's terms were never touched


def _bf16_rn(x):
    """float32 -> bfloat16 (round to nearest even) -> float32, as __float2bfloat16_rn / the host packer of flownet.cu"""
    import numpy as np
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def test_flow_tensor_core_operand_layout_and_three_product_precision():
    """csrc/flownet.cu: fconv_umma_kernel.  (1) The builder warps' shared-memory offsets and the host packer's weight image
    are both the K-major SWIZZLE_128B layout tcgen05.mma reads: element (row r, k) of a [rows x 64] bf16 tile lives at
    (r>>3)*1024 + (r&7)*128 + (((k>>3) ^ (r&7)) << 4) + (k&7)*2.  (2) x = hi + lo with bf16 halves and the three products
    A_hi*W_hi + A_hi*W_lo + A_lo*W_hi reproduce an fp32 dot product far inside the 1e-3 gate (the dropped term is 2^-16)."""
    import numpy as np
    rng = np.random.default_rng(0)

    def canonical(r, k):
        return (r >> 3) * 1024 + (r & 7) * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2

    # (1a) builder: thread (row, col) writes the 8 bytes of channels 4*col .. 4*col+3 at soff
    tile = np.full(128 * 128, -1, dtype=np.int64)     # byte -> (row * 64 + k) that owns it
    for row in range(128):
        for col in range(16):
            soff = (row >> 3) * 1024 + (row & 7) * 128 + (((col >> 1) ^ (row & 7)) << 4) + (col & 1) * 8
            for e in range(4):
                k = 4 * col + e
                assert tile[soff + 2 * e] == -1
                tile[soff + 2 * e] = tile[soff + 2 * e + 1] = row * 64 + k
    for r in range(128):
        for k in range(64):
            assert tile[canonical(r, k)] == r * 64 + k
    # (1b) host packer: [slab][cin/kc][plane][wrows][64] with the same row swizzle; an N tile of BN rows from n0 (multiple
    # of 8) is one contiguous block of BN * 128 bytes starting (n0 >> 3) * 1024 into the plane
    cin, cout, kc = 128, 256, 64
    ncb = cin // kc
    owner = {}
    for ci in range(cin):
        for co in range(cout):
            cb, kk = divmod(ci, kc)
            row = (co >> 3) * 1024 + (co & 7) * 128 + (((kk >> 3) ^ (co & 7)) << 4) + (kk & 7) * 2
            base = ((0 * ncb + cb) * 2) * cout * 128
            owner[base + row] = (ci, co)
    n0, bn, cb = 128, 128, 1
    block0 = ((0 * ncb + cb) * 2) * cout * 128 + (n0 >> 3) * 1024
    for r in range(bn):
        for k in range(64):
            assert owner[block0 + canonical(r, k)] == (cb * kc + k, n0 + r)
    # (2) precision of the split
    K = 2048
    a = rng.standard_normal(K).astype(np.float32) * np.float32(3.0)
    w = rng.standard_normal(K).astype(np.float32) * np.float32(0.05)
    a_hi, w_hi = _bf16_rn(a), _bf16_rn(w)
    a_lo, w_lo = _bf16_rn(a - a_hi), _bf16_rn(w - w_hi)
    assert np.abs((a_hi.astype(np.float64) + a_lo) - a).max() <= 2.0 ** -16 * np.abs(a).max()
    exact = float(np.dot(a.astype(np.float64), w.astype(np.float64)))
    three = float(np.dot(a_hi.astype(np.float64), w_hi) + np.dot(a_hi.astype(np.float64), w_lo) + np.dot(a_lo.astype(np.float64), w_hi))
    one = float(np.dot(a_hi.astype(np.float64), w_hi))
    scale = float(np.sqrt(np.sum((a.astype(np.float64) * w) ** 2)))
    assert abs(three - exact) <= 1e-4 * scale and abs(one - exact) >= 10 * abs(three - exact)


def test_key_point_boxes_from_coordinates_contain_their_discs():
    """csrc/conditioning.cu: kp_box_kernel (used by ap_flow_warp_landmarks).  The box floor(c - r) - 1 .. ceil(c + r) + 1,
    clipped to the map, contains every pixel of the disc (x - cx)^2 + (y - cy)^2 <= r^2 the fp64 test of kp_kernel can set
    (Module2/models/geomcgt_ifw_test_model.py:12-37), also for centres outside the map and for the missing-point marker."""
    import numpy as np
    rng = np.random.default_rng(1)
    size, r = 224, 4.0
    yy, xx = np.mgrid[0:size, 0:size]
    pts = np.concatenate([rng.uniform(-8, size + 8, (200, 2)), [[0.0, 0.0], [223.0, 223.0], [-4.0, 100.0], [227.0, 3.5],
                                                                 [-1.0, 50.0], [100.25, -1.0], [-4.5, -4.5]]]).astype(np.float32)
    for cx, cy in pts:
        disc = ((xx - np.float64(cx)) ** 2 + (yy - np.float64(cy)) ** 2 <= r * r)
        if cx == -1 or cy == -1:
            disc[:] = False                                   # the reference's missing point: an empty map
            box = None
        else:
            y0, y1 = int(max(np.floor(np.float32(cy - r)) - 1, 0)), int(min(np.ceil(np.float32(cy + r)) + 1, size - 1))
            x0, x1 = int(max(np.floor(np.float32(cx - r)) - 1, 0)), int(min(np.ceil(np.float32(cx + r)) + 1, size - 1))
            box = (y0, y1, x0, x1) if (y1 >= y0 and x1 >= x0) else None
        if box is None:
            assert not disc.any(), (cx, cy)
        else:
            inside = np.zeros_like(disc)
            inside[box[0]:box[1] + 1, box[2]:box[3] + 1] = True
            assert not (disc & ~inside).any(), (cx, cy, box)
            assert (box[1] - box[0] + 1) * (box[3] - box[2] + 1) <= 13 * 13
