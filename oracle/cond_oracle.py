"""CPU oracle of the per-frame CONDITIONING producers that feed netG (plus the photo matting
line of row f1).  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg; the product package never imports it.

What the reference does on the CPU for every frame before the generator runs:

  * `draw2(..., op=0)`       Module2/data/umlvdfw_test_dataset.py:34-41   68 filled discs -> land map in {-1,+1}
  * `cal_motion256`          Module2/data/umlvdfw_test_dataset.py:67-81   piecewise-linear (Delaunay) motion field
  * `kp_to_map(_some)`       Module2/models/geomcgt_ifw_test_model.py:12-44   68 binary key-point maps for netF
  * photo matting            Module2/models/geomcgt_ifw_test_model.py:292     real_A = ((real_A/2+.5)*mask + 1-mask)*2-1

Third-party arithmetic these call and that is NOT in /root/reference:
  * `cv2.circle(img, c, r, color, -1)` (OpenCV 4.x, imgproc/drawing.cpp): integer midpoint circle, filled with
    horizontal spans.  Restated in `cv_circle_halfwidths` and checked against cv2 itself for r = 0..12
    (tests/test_conditioning.py).
  * `scipy.interpolate.griddata(method='linear')` (scipy 1.x: Qhull Delaunay "Qbb Qc Qz Q12 Qt" + barycentric
    interpolation in `LinearNDInterpolator`).  Restated as an exhaustive empty-circumcircle Delaunay over the 72
    distinct sites + barycentric interpolation in float64.  For sites in general position the Delaunay triangulation
    is unique, so this equals scipy up to float64 rounding; for co-circular sites every triangulation of the
    co-circular cell is a Delaunay triangulation and Qhull picks one by its own facet order -- there the restatement
    picks the lexicographically first containing triangle (documented difference, still a linear interpolant of the
    same sites; tests cover it through affine reproduction).
Pinned by tests/golden/cond_*.npz, produced by tests/golden/make_cond_golden.py from the reference's own functions.
"""
from __future__ import annotations

import itertools

import numpy as np

N_LM = 68
# cal_motion256's `edges` (umlvdfw_test_dataset.py:69) lists 8 rows but only these 4 distinct sites
CORNERS = np.array([[0.0, 0.0], [255.0, 255.0], [0.0, 255.0], [255.0, 0.0]])
INCIRCLE_TOL = 1e-9   # relative to the magnitude of the determinant's terms
INSIDE_TOL = 1e-9     # barycentric tolerance of the point-in-triangle test


# --------------------------------------------------------------------------------------
# draw2 op 0
# --------------------------------------------------------------------------------------
def cv_circle_halfwidths(radius: int):
    """Half-width of the filled span at row offset |dy| = 0..radius of OpenCV's filled circle.

    Midpoint circle walking the first octant from (dx, dy) = (radius, 0): each step emits the spans
    rows +-dy with half-width dx and rows +-dx with half-width dy; the error term adds the odd numbers
    1,3,5,.. and, once positive, takes back 2*dx-1 while dx steps inwards."""
    hw = [0] * (radius + 1)
    dx, dy, err = radius, 0, 0
    inc, dec = 1, 2 * radius - 1
    while dx >= dy:
        hw[dy] = max(hw[dy], dx)
        hw[dx] = max(hw[dx], dy)
        dy += 1
        err += inc
        inc += 2
        if err > 0:
            err -= dec
            dx -= 1
            dec -= 2
    return hw


def draw_landmarks(lands: np.ndarray, size: int = 256, radius: int = 3) -> np.ndarray:
    """draw2(size, size, lands, radius, _, op=0) for a batch: lands [T,68,2] (x,y) -> [T,1,size,size] float32.

    umlvdfw_test_dataset.py:36-41: np.round (half to even) -> int, cv2.circle filled with 255 on a uint8 canvas
    (clipped at the border), then /255*2-1."""
    lands = np.asarray(lands, dtype=np.float32)
    T = lands.shape[0]
    hw = cv_circle_halfwidths(radius)
    out = np.zeros((T, size, size), dtype=np.uint8)
    c = np.round(lands).astype(np.int64)
    for t in range(T):
        for x, y in c[t]:
            for dy in range(-radius, radius + 1):
                yy = y + dy
                if 0 <= yy < size:
                    a, b = max(x - hw[abs(dy)], 0), min(x + hw[abs(dy)], size - 1)
                    if a <= b:
                        out[t, yy, a:b + 1] = 255
    return (out[:, None].astype(np.float32) / 255.0 * 2 - 1).astype(np.float32)


# --------------------------------------------------------------------------------------
# kp_to_map / kp_to_map_some (binary mode)
# --------------------------------------------------------------------------------------
def kp_to_map(kps: np.ndarray, size: int = 224, radius: float = 4) -> np.ndarray:
    """kp_to_map_some((size,size), kps): kps [T,68,2] float32 (x,y) -> [T,68,size,size] float32 in {0,1}.

    geomcgt_ifw_test_model.py:25-36: a point with x == -1 or y == -1 gives an empty map; otherwise
    (x_grid-x)**2 + (y_grid-y)**2 <= radius**2 with integer grids minus a float32 scalar, i.e. float64 arithmetic."""
    kps = np.asarray(kps, dtype=np.float32)
    T, K = kps.shape[:2]
    g = np.arange(size, dtype=np.float64)
    x = kps[..., 0].astype(np.float64)[:, :, None, None]
    y = kps[..., 1].astype(np.float64)[:, :, None, None]
    m = ((g[None, None, None, :] - x) ** 2 + (g[None, None, :, None] - y) ** 2 <= float(radius) ** 2)
    m &= ~((kps[..., 0] == -1) | (kps[..., 1] == -1))[:, :, None, None]
    return m.astype(np.float32).reshape(T, K, size, size)


# --------------------------------------------------------------------------------------
# cal_motion256
# --------------------------------------------------------------------------------------
def delaunay_triangles(pts: np.ndarray):
    """All triangles (i<j<k) of `pts` [n,2] float64 whose circumcircle holds no other site strictly inside.

    In general position this is THE Delaunay triangulation; with co-circular sites it is the union of all of them."""
    n = len(pts)
    trip = np.array(list(itertools.combinations(range(n), 3)), dtype=np.int64)
    a, b, c = pts[trip[:, 0]], pts[trip[:, 1]], pts[trip[:, 2]]
    orient = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    span = np.maximum(np.abs(b - a).max(1), np.abs(c - a).max(1))
    ok = np.abs(orient) > 1e-12 * np.maximum(span * span, 1e-300)   # drop degenerate (collinear / repeated) triples
    trip, a, b, c, orient = trip[ok], a[ok], b[ok], c[ok], orient[ok]
    keep = np.ones(len(trip), dtype=bool)
    for d in range(n):
        ad, bd, cd = a - pts[d], b - pts[d], c - pts[d]
        a2, b2, c2 = (ad * ad).sum(1), (bd * bd).sum(1), (cd * cd).sum(1)
        t1 = ad[:, 0] * (bd[:, 1] * c2 - b2 * cd[:, 1])
        t2 = ad[:, 1] * (bd[:, 0] * c2 - b2 * cd[:, 0])
        t3 = a2 * (bd[:, 0] * cd[:, 1] - bd[:, 1] * cd[:, 0])
        det = (t1 - t2 + t3) * np.sign(orient)      # > 0: d strictly inside the circumcircle
        mag = np.abs(t1) + np.abs(t2) + np.abs(t3)
        inside = det > INCIRCLE_TOL * mag
        inside &= ~((trip == d).any(1))
        keep &= ~inside
    return trip[keep]


def interpolate_linear(sites: np.ndarray, values: np.ndarray, tri: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Barycentric interpolation of `values` [n,m] given at `sites` [n,2] at the query points q [Q,2] (float64);
    the first triangle of `tri` (lexicographic order) that contains a query point wins; NaN outside the hull."""
    out = np.full((len(q), values.shape[1]), np.nan)
    done = np.zeros(len(q), dtype=bool)
    for i, j, k in tri:
        r0, r1, r2 = sites[i], sites[j], sites[k]
        m00, m01 = r0[0] - r2[0], r1[0] - r2[0]
        m10, m11 = r0[1] - r2[1], r1[1] - r2[1]
        det = m00 * m11 - m01 * m10
        dx, dy = q[:, 0] - r2[0], q[:, 1] - r2[1]
        c0 = (m11 * dx - m01 * dy) / det
        c1 = (-m10 * dx + m00 * dy) / det
        c2 = 1.0 - c0 - c1
        hit = (~done) & (c0 >= -INSIDE_TOL) & (c1 >= -INSIDE_TOL) & (c2 >= -INSIDE_TOL)
        if hit.any():
            out[hit] = c0[hit, None] * values[i] + c1[hit, None] * values[j] + c2[hit, None] * values[k]
            done |= hit
    return out


def motion_sites(lm_src: np.ndarray, lm_dst: np.ndarray):
    """Sites (destination space) and values (source space) of cal_motion256, both as (x, y) float64: the 68
    landmarks (float32 values widened, as np.concatenate with the int64 `edges` does) + the 4 corners."""
    dst = np.concatenate([np.asarray(lm_dst, dtype=np.float32).astype(np.float64), CORNERS])
    src = np.concatenate([np.asarray(lm_src, dtype=np.float32).astype(np.float64), CORNERS])
    return dst, src


def cal_motion(lm_src: np.ndarray, lm_dst: np.ndarray, size: int = 256) -> np.ndarray:
    """cal_motion256(lm2d0=lm_src, lm2d=lm_dst) -> [size,size,2] float32 sampling grid for F.grid_sample.

    umlvdfw_test_dataset.py:67-81: griddata maps each output pixel (given in destination = target-landmark space)
    to its source position by linear interpolation over the Delaunay triangulation of the target landmarks + image
    corners; channel 0 = x, channel 1 = y; float32(value) / 127.5 - 1 in float32."""
    dst, src = motion_sites(lm_src, lm_dst)
    tri = delaunay_triangles(dst)
    ys, xs = np.mgrid[0:size, 0:size]
    q = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float64) * (255.0 / (size - 1))
    val = interpolate_linear(dst, src, tri, q).astype(np.float32).reshape(size, size, 2)
    return (val / np.float32(127.5) - np.float32(1.0)).astype(np.float32)


def cal_motion_batch(lm_src: np.ndarray, lm_dst_seq: np.ndarray, size: int = 256) -> np.ndarray:
    """One source landmark set [68,2] (or [T,68,2]) against T target sets [T,68,2] -> [T,size,size,2]."""
    lm_dst_seq = np.asarray(lm_dst_seq)
    lm_src = np.asarray(lm_src)
    return np.stack([cal_motion(lm_src if lm_src.ndim == 2 else lm_src[t], lm_dst_seq[t], size)
                     for t in range(len(lm_dst_seq))])


# --------------------------------------------------------------------------------------
# photo matting (frame-invariant, row f1)
# --------------------------------------------------------------------------------------
def matte_photo(real_A: np.ndarray, matte: np.ndarray):
    """geomcgt_ifw_test_model.py:280,292: mask = (matte > 0.5).float(); real_A = ((real_A/2+0.5)*mask + 1-mask)*2-1.
    real_A [B,3,H,W], matte [B,1,H,W] float32 -> (matted photo, mask), float32, in the reference's op order."""
    real_A = np.asarray(real_A, dtype=np.float32)
    mask = (np.asarray(matte, dtype=np.float32) > np.float32(0.5)).astype(np.float32)
    one, half, two = np.float32(1), np.float32(0.5), np.float32(2)
    out = ((real_A / two + half) * mask + one - mask) * two - one
    return out.astype(np.float32), mask


# --------------------------------------------------------------------------------------
# synthetic landmark sequences (fixed face template + smooth per-frame trajectory)
# --------------------------------------------------------------------------------------
def face_template() -> np.ndarray:
    """A plausible 68-point layout in the 256x256 window (jaw arc, brows, nose, eyes, two mouth rings)."""
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
    """(source landmarks [68,2], target sequence [T,68,2]) float32: template + jitter, then a smooth head sway plus a
    mouth opening cycle -- a stand-in for Module1's predicted landmarks (Module1 cannot run here)."""
    rng = np.random.RandomState(seed)
    src = face_template() + rng.normal(0, 0.7, (N_LM, 2)).astype(np.float32)
    t = np.arange(T, dtype=np.float32)[:, None, None]
    sway = np.concatenate([amp * np.sin(2 * np.pi * t / 90.0), 0.5 * amp * np.sin(2 * np.pi * t / 57.0 + 1.0)], 2)
    seq = src[None] + sway
    mouth = np.zeros((1, N_LM, 2), dtype=np.float32)
    mouth[0, 48:68, 1] = (src[48:68, 1] - 182.0) * 0.6
    seq = seq + mouth * (0.5 + 0.5 * np.sin(2 * np.pi * t / 11.0))
    seq = seq + rng.normal(0, 0.15, seq.shape)
    return src.astype(np.float32), seq.astype(np.float32)


# --------------------------------------------------------------------------------------
# one clip frame end to end, the way the reference's per-frame loop produces it
# --------------------------------------------------------------------------------------
def render_clip_frames(sd, photo, matte, static, lm_src, lm_seq, flow=None, ifmask=None):
    """Restates, frame by frame, UMLVDFWTestDataset.__getitem__ (umlvdfw_test_dataset.py:145-166) ->
    GeomCGTIFWTestModel.forward (geomcgt_ifw_test_model.py:279-300) -> tensor2im (util/util.py:9-29) with the flow
    network's outputs given.  torch CPU tensors in; returns (fp32 blended frames [T,onc,256,256], uint8 [T,256,256,3])."""
    import torch
    from oracle import netg_oracle as N
    T = len(lm_seq)
    real_A, mask = matte_photo(photo.numpy(), matte.numpy())
    real_A, mask = torch.from_numpy(real_A), torch.from_numpy(mask)
    land1 = torch.from_numpy(draw_landmarks(np.asarray(lm_src)[None]))
    outs = []
    for t in range(T):
        land2 = torch.from_numpy(draw_landmarks(np.asarray(lm_seq[t])[None]))
        motion = torch.from_numpy(cal_motion(np.asarray(lm_src), np.asarray(lm_seq[t])))[None]
        f = flow[t:t + 1] if flow is not None else torch.zeros(1, 2, 256, 256)
        m = ifmask[t:t + 1] if ifmask is not None else torch.ones(1, 1, 256, 256)
        fake = N.netg_forward(sd, real_A, land1, land2, motion, f, m)
        outs.append(N.blend_foreground(fake, mask, motion, static))
    blended = torch.cat(outs)
    return blended, N.tensor2im_batch(blended)
