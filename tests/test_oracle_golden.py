"""CPU: pin the oracle (oracle/netg_oracle.py) to the golden vectors produced by the reference itself
(tests/golden/make_golden.py ran the unmodified /root/reference generator)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import netg_oracle as O
from tests.golden.make_golden import CASES, sample_tap

FAST_CASES = ["c1_line_smooth", "c1_line_noise", "c1_line_bias", "c5_cartoon_noise", "b2_line_mixed"]


def _case(name):
    onc, B, wseed, bstd, iseed, kind = CASES[name]
    sd = O.make_state_dict(onc, seed=wseed, bias_std=bstd)
    inputs = O.make_inputs(B, seed=iseed, kind=kind)
    return onc, B, sd, inputs


def test_state_dict_spec_matches_survey_appendix_b():
    spec = O.state_dict_spec(1)
    assert len(spec) == 74
    assert list(spec)[0] == "model_tri_merge.weight"  # registered first (networks.py:1251)
    assert spec["model_tri_merge.weight"] == (256, 768, 3, 3)
    assert spec["model3.0.weight"] == (256, 128, 3, 3)  # ConvTranspose2d: [Cin, Cout, kh, kw]
    assert spec["model2.0.shortcut.0.weight"] == (256, 288, 3, 3)
    assert "model2.1.shortcut.0.weight" not in spec
    n1 = sum(int(np.prod(s)) for s in spec.values())
    n3 = sum(int(np.prod(s)) for s in O.state_dict_spec(3).values())
    assert (n1, n3) == (15925553, 15931827)  # SURVEY.md Appendix D anchors


def test_flops_per_frame_matches_survey():
    assert O.flops_per_frame(1) == pytest.approx(140.125e9, rel=1e-4)
    assert O.flops_per_frame(3) == pytest.approx(140.947e9, rel=1e-4)


def test_reference_anchors_recorded(manifest):
    a = manifest["_anchors_define_G_seed0"]
    # values measured by the survey on the reference
    assert a["1"]["params"] == 15925553 and a["3"]["params"] == 15931827
    assert a["1"]["weight_sum"] == pytest.approx(-17.337519, abs=1e-4)
    assert a["1"]["y000"] == pytest.approx(-0.765963, abs=1e-5)
    assert a["1"]["y_center"] == pytest.approx(-0.958812, abs=1e-5)
    assert a["3"]["y000"] == pytest.approx(-0.827040, abs=1e-5)
    assert a["1"]["oracle_vs_reference"] <= 1e-6 and a["3"]["oracle_vs_reference"] <= 1e-6


@pytest.mark.parametrize("name", FAST_CASES)
def test_oracle_reproduces_reference_golden(name, golden_dir, manifest):
    onc, B, sd, inputs = _case(name)
    m = manifest[name]
    # the seeded recipes must regenerate the very tensors the reference saw
    wsum = sum(float(t.double().sum()) for t in sd.values())
    assert wsum == pytest.approx(m["weight_sum"], abs=1e-6), "weight recipe drifted from the fixture"
    for t, s in zip(inputs, m["input_sums"]):
        assert float(t.double().sum()) == pytest.approx(s, rel=1e-9, abs=1e-6), "input recipe drifted from the fixture"
    taps = {}
    y = O.netg_forward(sd, *inputs, tap=lambda n, v: taps.__setitem__(n, v))
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    assert y.shape == (B, onc, 256, 256)
    assert np.abs(y.numpy() - g["y"]).max() <= 1e-5
    for k, st in m["taps"].items():
        v = taps[k]
        assert list(v.shape) == st["shape"]
        assert np.abs(sample_tap(v) - g["tap_" + k]).max() <= 1e-4, k
        assert v.mean().item() == pytest.approx(st["mean"], abs=1e-4)
        assert v.abs().max().item() == pytest.approx(st["absmax"], rel=1e-4)


def test_closed_form_warp_matches_grid_sample_formulation():
    g = torch.Generator().manual_seed(5)
    x, l1, l2, motion, flow, ifmask = O.make_inputs(1, seed=77, kind="noise")
    for level, (C, S) in enumerate([(32, 256), (64, 128), (128, 64)]):
        feat = torch.randn(1, C, S, S, generator=g)
        a = O.double_feature_warping(feat, motion, flow, ifmask, level)
        b = O.double_feature_warping_closed_form(feat, motion, flow, ifmask, level)
        assert a.shape == (1, 2 * C, S, S)
        d = (a - b).abs()
        # coordinate rounding only; a flipped mask pixel would show up as an O(1) error
        assert d.max().item() < 5e-4, (level, d.max().item())


def test_batched_equals_per_sample():
    onc, B, sd, inputs = _case("b2_line_mixed")
    y = O.netg_forward(sd, *inputs)
    for i in range(B):
        yi = O.netg_forward(sd, *[t[i:i + 1] for t in inputs])
        assert (y[i:i + 1] - yi).abs().max().item() <= 1e-6


def test_only_two_biases_reach_the_output():
    # biases in front of an affine-less InstanceNorm cancel
    sd = O.make_state_dict(1, seed=7, bias_std=0.5)
    inputs = O.make_inputs(1, seed=1003, kind="smooth")
    y = O.netg_forward(sd, *inputs)
    sd2 = {k: (torch.zeros_like(v) if k.endswith(".bias") and not k.startswith(("model_tri_merge", "model3.7")) else v)
           for k, v in sd.items()}
    assert (O.netg_forward(sd2, *inputs) - y).abs().max().item() < 1e-4
    sd3 = dict(sd2)
    sd3["model_tri_merge.bias"] = torch.zeros(256)
    assert (O.netg_forward(sd3, *inputs) - y).abs().max().item() > 1e-3
