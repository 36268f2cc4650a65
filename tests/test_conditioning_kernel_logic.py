"""CPU checks of the arguments the conditioning kernels (animateportrait_b200/csrc/conditioning.cu) rest on, emulated in
numpy with the kernels' own arithmetic (pure CPU, no GPU):

1. delaunay_kernel unranks the flat thread index into the site triple i < j < k with a binary search + a float32 square
   root: the decoding must be the lexicographic order of itertools.combinations for all C(72,3) triples.
2. Its fp32 screen (`incircle_screen`) may only decide what the float64 in-circle test of the oracle decides: never
   "inside" when the oracle says outside, never "outside" when the oracle says inside -- also for co-circular sites.
3. Its phase A tests each triple only against the index-neighbours of its vertices; that is a REJECTION shortcut, so the
   survivors must contain every Delaunay triangle, and few others (the reason the shortcut pays).
4. motion_raster_kernel's division-free screen (barycentric numerators times det >= -1e-6 det^2) must be a superset of
   the exact containment test (tolerance 1e-9), or pixels would lose their triangle.
5. draw_kernel receives OpenCV's circle span table as 16 nibbles.
"""
import itertools

import numpy as np

from oracle import cond_oracle as O

NS = 72
NT = NS * (NS - 1) * (NS - 2) // 6


def _sites(kind, seed=0):
    rng = np.random.RandomState(seed)
    if kind == "clip":
        src, seq = O.landmark_sequence(3, seed)
        return O.motion_sites(src, seq[2])[0]
    ys, xs = np.mgrid[0:9, 0:8]
    lat = np.stack([20 + 25 * xs.ravel(), 15 + 25 * ys.ravel()], 1)[:68].astype(np.float32)
    if kind == "lattice":
        return O.motion_sites(lat, lat)[0]
    if kind == "near-lattice":
        near = (lat + rng.normal(0, 1e-4, lat.shape)).astype(np.float32)
        return O.motion_sites(near, near)[0]
    r = rng.uniform(-5, 260, (68, 2)).astype(np.float32)
    return O.motion_sites(r, r)[0]


def test_triple_unranking_is_lexicographic():
    before = [NT - (NS - i) * (NS - i - 1) * (NS - i - 2) // 6 for i in range(NS)]
    t = np.arange(NT)
    i = np.searchsorted(np.array(before[:NS - 2]), t, side="right") - 1        # the kernel's binary search
    m = NS - 1 - i
    r = t - np.array(before)[i]
    b2m = (2 * m - 1).astype(np.float32)
    disc = np.maximum(b2m * b2m - np.float32(8) * r.astype(np.float32), np.float32(0))
    jj = ((b2m - np.sqrt(disc)) * np.float32(0.5)).astype(np.int64)
    jj = np.clip(jj, 0, m - 2)
    S = lambda q: q * m - q * (q + 1) // 2  # noqa: E731
    for _ in range(3):                                                           # the kernel's two fix-up loops
        jj = np.where(S(jj + 1) <= r, jj + 1, jj)
    for _ in range(3):
        jj = np.where(S(jj) > r, jj - 1, jj)
    kk = r - S(jj) + jj + 1
    want = np.array(list(itertools.combinations(range(NS), 3)))
    assert np.array_equal(np.stack([i, i + 1 + jj, i + 1 + kk], 1), want)


def _incircle64(a, b, c, d, sgn):
    ad, bd, cd = a - d, b - d, c - d
    a2, b2, c2 = (ad * ad).sum(1), (bd * bd).sum(1), (cd * cd).sum(1)
    t1 = ad[:, 0] * (bd[:, 1] * c2 - b2 * cd[:, 1])
    t2 = ad[:, 1] * (bd[:, 0] * c2 - b2 * cd[:, 0])
    t3 = a2 * (bd[:, 0] * cd[:, 1] - bd[:, 1] * cd[:, 0])
    return (t1 - t2 + t3) * sgn > O.INCIRCLE_TOL * (np.abs(t1) + np.abs(t2) + np.abs(t3))


def _screen32(a, b, c, d, sgn):
    f = np.float32
    A, B, C, D, s = a.astype(f), b.astype(f), c.astype(f), d.astype(f), sgn.astype(f)
    ad, bd, cd = A - D, B - D, C - D
    a2 = ad[:, 0] * ad[:, 0] + ad[:, 1] * ad[:, 1]
    b2 = bd[:, 0] * bd[:, 0] + bd[:, 1] * bd[:, 1]
    c2 = cd[:, 0] * cd[:, 0] + cd[:, 1] * cd[:, 1]
    u1, u2, u3, u4, u5, u6 = bd[:, 1] * c2, b2 * cd[:, 1], bd[:, 0] * c2, b2 * cd[:, 0], bd[:, 0] * cd[:, 1], bd[:, 1] * cd[:, 0]
    det = (ad[:, 0] * (u1 - u2) - ad[:, 1] * (u3 - u4) + a2 * (u5 - u6)) * s
    M = (np.abs(ad[:, 0]) * (np.abs(u1) + np.abs(u2)) + np.abs(ad[:, 1]) * (np.abs(u3) + np.abs(u4))
         + a2 * (np.abs(u5) + np.abs(u6)))
    thr = f(1e-5) * M
    return np.where(det > thr, 1, np.where(det < -thr, -1, 0))


def test_fp32_screen_never_contradicts_the_float64_decision():
    rng = np.random.RandomState(3)
    for kind in ("clip", "lattice", "near-lattice", "random"):
        pts = _sites(kind)
        idx = rng.randint(0, NS, (300000, 4))
        ok = np.array([len(set(q)) == 4 for q in idx[:, :4].tolist()])
        idx = idx[ok]
        a, b, c, d = (pts[idx[:, q]] for q in range(4))
        orient = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
        keep = orient != 0
        a, b, c, d, sgn = a[keep], b[keep], c[keep], d[keep], np.sign(orient[keep])
        inside = _incircle64(a, b, c, d, sgn)
        scr = _screen32(a, b, c, d, sgn)
        assert not ((scr > 0) & ~inside).any() and not ((scr < 0) & inside).any(), kind
        # and it decides nearly everything, so the fp64 pipe only sees nearly co-circular quadruples
        assert (scr == 0).mean() < (0.05 if "lattice" in kind else 1e-3), kind


def test_neighbour_screen_keeps_every_delaunay_triangle_and_little_else():
    pts = _sites("clip", seed=2)
    trip = np.array(list(itertools.combinations(range(NS), 3)))
    a, b, c = pts[trip[:, 0]], pts[trip[:, 1]], pts[trip[:, 2]]
    orient = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    sgn = np.sign(orient)
    alive = np.abs(orient) > 0
    for q in range(12):                                   # the kernel's candidate order: +-1, +-2 around i, j, k
        off, r6 = 1 + q // 6, q % 6
        base = trip[:, 0] if r6 < 2 else (trip[:, 1] if r6 < 4 else trip[:, 2])
        d = (base + (-off if r6 & 1 else off)) % NS
        skip = (d == trip[:, 0]) | (d == trip[:, 1]) | (d == trip[:, 2])
        alive &= skip | ~_incircle64(a, b, c, pts[d], sgn)
    survivors = {tuple(t) for t in trip[alive].tolist()}
    delaunay = {tuple(t) for t in O.delaunay_triangles(pts).tolist()}
    assert delaunay <= survivors
    assert len(survivors) < 0.03 * len(trip)              # ~1 % of the triples reach the warp-parallel phase


def test_division_free_raster_screen_is_a_superset_of_the_exact_test():
    pts = _sites("clip", seed=4)
    tri = O.delaunay_triangles(pts)
    ys, xs = np.mgrid[0:256, 0:256]
    q = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float64)[::7]
    for i, j, k in tri[::3]:
        r0, r1, r2 = pts[i], pts[j], pts[k]
        m00, m01, m10, m11 = r0[0] - r2[0], r1[0] - r2[0], r0[1] - r2[1], r1[1] - r2[1]
        det = m00 * m11 - m01 * m10
        dx, dy = q[:, 0] - r2[0], q[:, 1] - r2[1]
        c0 = (m11 * dx - m01 * dy) / det
        c1 = (-m10 * dx + m00 * dy) / det
        c2 = 1.0 - c0 - c1
        exact = (c0 >= -O.INSIDE_TOL) & (c1 >= -O.INSIDE_TOL) & (c2 >= -O.INSIDE_TOL)
        a0, a1 = m11 * dx - m01 * dy, m00 * dy - m10 * dx
        lim = -1e-6 * det * det
        loose = (a0 * det >= lim) & (a1 * det >= lim) & ((det - a0 - a1) * det >= lim)
        assert not (exact & ~loose).any()
        assert loose.sum() <= exact.sum() + 8             # and it is tight: only pixels within 1e-6 of an edge are extra


def test_circle_span_table_fits_sixteen_nibbles():
    for radius in range(16):
        hw = O.cv_circle_halfwidths(radius) + [0] * (15 - radius)
        packed = sum(h << (4 * i) for i, h in enumerate(hw))
        assert max(hw) <= 15 and packed < (1 << 64)
        assert [(packed >> (4 * i)) & 15 for i in range(16)] == hw
