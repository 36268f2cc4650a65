"""Conditioning producers (landmark maps, Delaunay motion field, key-point maps, photo matting).

CPU: oracle/cond_oracle.py against the golden vectors made from the reference's own functions
(tests/golden/make_cond_golden.py), against cv2 / scipy where they are installed, and through properties.
GPU (-m gpu): the CUDA kernels through the C ABI (include/ap_cond.h) against the golden vectors and the oracle.
Bars: landmark maps, key-point maps and matting are bit-exact; the motion field is float32(float64 barycentric
interpolation) / 127.5 - 1, tolerance 1e-6 on the [-1,1] grid (measured 0.0 against scipy on every golden frame)."""
import os

import numpy as np
import pytest
import torch

from oracle import cond_oracle as O

MOTION_TOL = 1e-6


def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _unpack(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(bool)


# ------------------------------------------------------------------------------------------------------
# CPU: the oracle is pinned to the reference
# ------------------------------------------------------------------------------------------------------
def test_oracle_landmark_maps_match_reference_golden(golden_dir):
    g = _golden(golden_dir, "cond_draw")
    want = _unpack(g["bits"], (len(g["lands"]), 1, 256, 256))
    got = O.draw_landmarks(g["lands"])
    assert got.dtype == np.float32 and set(np.unique(got)) == {-1.0, 1.0}
    assert np.array_equal(got > 0, want)


def test_oracle_circle_is_opencv_circle():
    cv2 = pytest.importorskip("cv2")
    for r in range(0, 13):
        img = np.zeros((48, 48), np.uint8)
        cv2.circle(img, (24, 24), r, 255, -1)
        hw = O.cv_circle_halfwidths(r)
        for dy in range(-r, r + 1):
            row = np.flatnonzero(img[24 + dy])
            assert row.min() == 24 - hw[abs(dy)] and row.max() == 24 + hw[abs(dy)]
        assert img.sum() // 255 == sum(2 * hw[abs(dy)] + 1 for dy in range(-r, r + 1))


def test_oracle_motion_matches_reference_golden(golden_dir):
    g = _golden(golden_dir, "cond_motion")
    for slot, t in enumerate(g["full_index"]):
        got = O.cal_motion(g["src"][t], g["dst"][t])
        assert got.shape == (256, 256, 2) and got.dtype == np.float32
        assert np.abs(got - g["full"][slot]).max() <= MOTION_TOL
    t = 3   # one more frame on the strided lattice
    assert np.abs(O.cal_motion(g["src"][t], g["dst"][t])[::4, ::4] - g["strided"][t]).max() <= MOTION_TOL


def test_oracle_motion_matches_scipy_griddata_on_fresh_landmarks():
    interp = pytest.importorskip("scipy.interpolate")
    rng = np.random.RandomState(5)
    src = rng.uniform(10, 245, (68, 2)).astype(np.float32)
    dst = (src + rng.normal(0, 5, (68, 2))).astype(np.float32)
    sites, vals = O.motion_sites(src, dst)
    ys, xs = np.mgrid[0:256, 0:256]
    ref = interp.griddata(sites, vals, (xs.astype(np.float64), ys.astype(np.float64)), method="linear")
    ref = (ref.astype(np.float32) / np.float32(127.5) - np.float32(1)).astype(np.float32)
    assert np.abs(O.cal_motion(src, dst) - ref).max() <= MOTION_TOL


def test_oracle_triangulation_properties():
    src, seq = O.landmark_sequence(2, seed=4)
    sites, _ = O.motion_sites(src, seq[1])
    tri = O.delaunay_triangles(sites)
    # Euler: a triangulation of n sites with h on the hull has 2n - 2 - h triangles (hull = the 4 corners here)
    assert len(tri) == 2 * len(sites) - 2 - 4
    a, b, c = sites[tri[:, 0]], sites[tri[:, 1]], sites[tri[:, 2]]
    area = 0.5 * np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
    assert abs(area.sum() - 255.0 * 255.0) < 1e-6     # the triangles tile the image square exactly once


def test_oracle_triangulation_is_qhulls_on_sites_in_general_position():
    spatial = pytest.importorskip("scipy.spatial")
    rng = np.random.RandomState(17)
    for trial in range(6):
        lm = rng.uniform(-10 if trial % 2 else 5, 265 if trial % 2 else 250, (68, 2)).astype(np.float32)
        sites, _ = O.motion_sites(lm, lm)
        mine = {tuple(sorted(t)) for t in O.delaunay_triangles(sites).tolist()}
        qhull = {tuple(sorted(t)) for t in spatial.Delaunay(sites).simplices.tolist()}
        assert mine == qhull


def test_oracle_motion_identity_and_affine_reproduction_under_cocircular_sites():
    # landmarks on an integer lattice: many co-circular quadruples, the Delaunay triangulation is not unique; every
    # valid choice still reproduces an affine map exactly (source = A * destination + b at all sites incl. the corners)
    ys, xs = np.mgrid[0:9, 0:8]
    dst = np.stack([20 + 25 * xs.ravel(), 15 + 25 * ys.ravel()], 1)[:68].astype(np.float32)
    ident = O.cal_motion(dst, dst)
    lat = (np.arange(256, dtype=np.float32) / np.float32(127.5) - np.float32(1)).astype(np.float32)
    assert np.array_equal(ident[..., 0], np.broadcast_to(lat[None, :], (256, 256)))
    assert np.array_equal(ident[..., 1], np.broadcast_to(lat[:, None], (256, 256)))
    assert not np.isnan(ident).any()


def test_oracle_key_point_maps_match_reference_golden(golden_dir):
    g = _golden(golden_dir, "cond_kp")
    want = _unpack(g["bits"], (2, 68, 224, 224))
    got = O.kp_to_map(g["kps"])
    assert got.dtype == np.float32 and np.array_equal(got > 0, want)
    assert got[1, 10].sum() == 0          # x == -1: the "missing point" branch


def test_oracle_matting_matches_reference_golden(golden_dir):
    g = _golden(golden_dir, "cond_matte")
    out, mask = O.matte_photo(g["real_A"], g["matte"])
    assert np.array_equal(out, g["out"]) and np.array_equal(mask, g["mask"])
    assert mask[0, 0, 0, 0] == 0 and mask[0, 0, 0, 1] == 1   # strict > 0.5


def test_wrappers_reject_cpu_tensors_and_unbuilt_modes():
    from animateportrait_b200 import conditioning as cond
    with pytest.raises(RuntimeError, match="CUDA"):
        cond.draw2(256, 256, torch.zeros(1, 68, 2), 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        cond.cal_motion256(torch.zeros(68, 2), torch.zeros(2, 68, 2))
    with pytest.raises(NotImplementedError):
        cond.draw2(256, 256, torch.zeros(1, 68, 2), 3, op=1)
    with pytest.raises(NotImplementedError):
        cond.kp_to_map_some((224, 224), torch.zeros(1, 68, 2), mode="gaussian")


# ------------------------------------------------------------------------------------------------------
# GPU: the CUDA kernels through the C ABI
# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.mark.gpu
def test_gpu_landmark_maps_bit_exact(golden_dir, dev):
    from animateportrait_b200 import conditioning as cond
    g = _golden(golden_dir, "cond_draw")
    want = _unpack(g["bits"], (len(g["lands"]), 1, 256, 256))
    got = cond.draw2(256, 256, torch.from_numpy(g["lands"]).to(dev), 3).cpu().numpy()
    assert got.shape == want.shape and set(np.unique(got)) == {-1.0, 1.0}
    assert np.array_equal(got > 0, want)
    # crop_size 512 variant (radius 5), one frame in the reference's single-frame shape
    l5 = 2 * g["lands"][:2]
    got5 = cond.draw2(512, 512, torch.from_numpy(l5[0]).to(dev), 5).cpu().numpy()
    assert got5.shape == (1, 512, 512) and np.array_equal(got5, O.draw_landmarks(l5[:1], 512, 5)[0])
    # NaN / far-away points never reach the canvas
    bad = torch.full((1, 68, 2), float("nan"), device=dev)
    bad[0, 1] = 1e30
    assert (cond.draw2(256, 256, bad, 3) == -1).all()


@pytest.mark.gpu
def test_gpu_motion_matches_reference_golden_and_oracle(golden_dir, dev):
    from animateportrait_b200 import conditioning as cond
    g = _golden(golden_dir, "cond_motion")
    src, dst = torch.from_numpy(g["src"]).to(dev), torch.from_numpy(g["dst"]).to(dev)
    motion, count = cond.cal_motion256(src, dst, return_triangle_count=True)
    motion, count = motion.cpu().numpy(), count.cpu().numpy()
    assert motion.shape == (8, 256, 256, 2) and not np.isnan(motion).any()
    # general position: THE Delaunay triangulation, 2n-2-h triangles (h = 4 hull corners when all landmarks are inside
    # the window; the two scattered-landmark frames have points outside it)
    assert (count[:6] == 2 * 72 - 2 - 4).all()
    for t in (6, 7):
        assert count[t] == len(O.delaunay_triangles(O.motion_sites(g["src"][t], g["dst"][t])[0]))
    for slot, t in enumerate(g["full_index"]):
        assert np.abs(motion[t] - g["full"][slot]).max() <= MOTION_TOL
    assert np.abs(motion[:, ::4, ::4] - g["strided"]).max() <= MOTION_TOL
    # same op order as the oracle: bit-identical
    assert np.array_equal(motion[3], O.cal_motion(g["src"][3], g["dst"][3]))
    # one photo, many frames (shared source landmarks) and the single-frame shape of the reference
    shared = cond.cal_motion256(src[0], dst[:6]).cpu().numpy()
    assert np.array_equal(shared, motion[:6])
    one = cond.cal_motion256(src[2], dst[2]).cpu().numpy()
    assert one.shape == (256, 256, 2) and np.array_equal(one, motion[2])


@pytest.mark.gpu
def test_gpu_motion_degenerate_sites(dev):
    from animateportrait_b200 import conditioning as cond
    # co-circular lattice sites: same lexicographic tie-break as the oracle -> identical, and the identity map is exact
    ys, xs = np.mgrid[0:9, 0:8]
    lat = np.stack([20 + 25 * xs.ravel(), 15 + 25 * ys.ravel()], 1)[:68].astype(np.float32)
    rng = np.random.RandomState(9)
    src = (lat + rng.normal(0, 2, lat.shape)).astype(np.float32)
    got, count = cond.cal_motion256(torch.from_numpy(src).to(dev), torch.from_numpy(lat).to(dev), return_triangle_count=True)
    assert int(count) <= 512 and not torch.isnan(got).any()
    assert np.abs(got.cpu().numpy() - O.cal_motion(src, lat)).max() <= MOTION_TOL
    ident = cond.cal_motion256(torch.from_numpy(lat).to(dev), torch.from_numpy(lat).to(dev)).cpu().numpy()
    grid = (np.arange(256, dtype=np.float32) / np.float32(127.5) - np.float32(1)).astype(np.float32)
    assert np.array_equal(ident[..., 0], np.broadcast_to(grid[None, :], (256, 256)))
    assert np.array_equal(ident[..., 1], np.broadcast_to(grid[:, None], (256, 256)))
    # repeated landmarks (two points coincide) and landmarks outside the window
    dup = src.copy()
    dup[5] = dup[4]
    dup[9] = [-20.0, 300.0]
    got = cond.cal_motion256(torch.from_numpy(src).to(dev), torch.from_numpy(dup).to(dev)).cpu().numpy()
    assert np.abs(got - O.cal_motion(src, dup)).max() <= MOTION_TOL


@pytest.mark.gpu
def test_gpu_motion_full_clip_properties(dev):
    """733 frames (a 12 s clip) in one call: no holes, 138 triangles everywhere, the field
    interpolates the landmarks (sampling it at a target landmark returns the source landmark), batch == per-frame."""
    from animateportrait_b200 import conditioning as cond
    src, seq = O.landmark_sequence(733, seed=2)
    s, d = torch.from_numpy(src).to(dev), torch.from_numpy(seq).to(dev)
    motion, count = cond.cal_motion256(s, d, return_triangle_count=True)
    assert motion.shape == (733, 256, 256, 2) and not torch.isnan(motion).any()
    assert (count == 138).all()
    assert torch.equal(cond.cal_motion256(s, d[700:701])[0], motion[700])
    # corners map to themselves
    assert motion[:, 0, 0].abs().sub(1).abs().max() == 0 and motion[:, 255, 255].sub(1).abs().max() == 0
    # bilinear sample of the field at a target landmark ~ the source landmark (exact at vertices of a linear interpolant,
    # up to the kink inside the sampled pixel: the field's gradient changes by < 1 px/px across an edge)
    t = 321
    pix = (motion[t].cpu().numpy() + 1) * 127.5
    for k in (30, 36, 48, 8):
        x, y = seq[t, k]
        x0, y0 = int(np.floor(x)), int(np.floor(y))
        wx, wy = x - x0, y - y0
        val = ((1 - wx) * (1 - wy) * pix[y0, x0] + wx * (1 - wy) * pix[y0, x0 + 1] + (1 - wx) * wy * pix[y0 + 1, x0]
               + wx * wy * pix[y0 + 1, x0 + 1])
        assert np.abs(val - src[k]).max() < 0.5


@pytest.mark.gpu
def test_gpu_key_point_maps_bit_exact(golden_dir, dev):
    from animateportrait_b200 import conditioning as cond
    g = _golden(golden_dir, "cond_kp")
    want = _unpack(g["bits"], (2, 68, 224, 224))
    got = cond.kp_to_map_some((224, 224), torch.from_numpy(g["kps"]).to(dev)).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got > 0, want) and set(np.unique(got)) <= {0.0, 1.0}
    assert got[1, 10].sum() == 0


@pytest.mark.gpu
def test_gpu_matting_bit_exact(golden_dir, dev):
    from animateportrait_b200 import conditioning as cond
    g = _golden(golden_dir, "cond_matte")
    out, mask = cond.matte_photo(torch.from_numpy(g["real_A"]).to(dev), torch.from_numpy(g["matte"]).to(dev))
    assert np.array_equal(out.cpu().numpy(), g["out"]) and np.array_equal(mask.cpu().numpy(), g["mask"])
