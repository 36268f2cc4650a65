"""netF, the intrinsic-flow network (include/ap_flow.h, animateportrait_b200/flownet.py).

CPU: the oracle restatement against the golden vectors generated from the reference class itself
(tests/golden/make_flow_golden.py), the checkpoint layout of the host module against the oracle's spec (which that script
checked against the real module's state_dict), configuration errors.
GPU (-m gpu): the CUDA path through the C ABI against the oracle for every golden configuration: flow / visibility within
1e-3 max-abs (the fp32 gate of the path), the 256x256 tensors the generator consumes, `flow_network_warp` from landmarks."""
import os

import numpy as np
import pytest
import torch

from animateportrait_b200.flownet import FlowUnet, flow_network_warp
from oracle import cond_oracle as OC
from oracle import flow_oracle as FO
from tests.golden.make_flow_golden import CASES, kp_maps

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _seed(name):
    return sum(map(ord, name)) % 1000


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_the_reference_vectors(name):
    nf, ss, ns, norm, B = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    sd = FO.make_state_dict(136, nf, ss, ns, norm, seed=int(g["seed"]))
    flow, vis, _, _ = FO.flow_unet_forward(sd, kp_maps(B, seed=7 + B), nf, ss, ns, norm)
    iw, ifm = FO.warp_outputs(flow, vis)
    assert np.array_equal(flow[:, :, ::4, ::4].numpy(), g["flow"]) and np.array_equal(vis[:, :, ::4, ::4].numpy(), g["vis"])
    assert np.array_equal(iw[:, :, ::4, ::4].numpy(), g["iw_flow"]) and np.array_equal(ifm[:, :, ::4, ::4].numpy(), g["if_mask"])
    assert abs(float(flow.double().abs().sum()) - float(g["flow_abs_sum"])) <= 1e-9 * float(g["flow_abs_sum"])
    assert iw.shape == (B, 2, 256, 256) and ifm.shape == (B, 1, 256, 256) and set(np.unique(g["if_mask"] >= 0)) == {True}


@pytest.mark.parametrize("cfg", [(16, 2, 4, "batch"), (32, 1, 5, "batch"), (16, 4, 3, "instance"), (64, 2, 4, "instance")])
def test_host_module_has_the_reference_checkpoint_layout(cfg):
    nf, ss, ns, norm = cfg
    net = FlowUnet(136, nf=nf, start_scale=ss, num_scale=ns, norm=norm)
    keys = [(k, tuple(v.shape)) for k, v in net.state_dict().items() if not k.endswith("num_batches_tracked")]
    assert keys == FO.state_dict_spec(136, nf, ss, ns, norm)
    net.load_state_dict(FO.make_state_dict(136, nf, ss, ns, norm), strict=False)


def test_refuses_cpu_tensors_training_mode_and_inconsistent_pyramids():
    net = FlowUnet(136, nf=8, start_scale=2, num_scale=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 136, 224, 224))
    with pytest.raises(RuntimeError, match="CUDA"):       # the landmark-level entry point has no CPU path either
        net.warp_landmarks(torch.zeros(68, 2), torch.zeros(1, 68, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        flow_network_warp(net, None, torch.zeros(1, 68, 2), torch.zeros(1, 68, 2))
    assert not FO.consistent(224, 2, 5) and FO.consistent(224, 2, 4) and FO.consistent(224, 1, 5)


# ------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_flow_network_matches_the_oracle(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda", 0)
    nf, ss, ns, norm, B = CASES[name]
    sd = FO.make_state_dict(136, nf, ss, ns, norm, seed=_seed(name))
    net = FlowUnet(136, nf=nf, start_scale=ss, num_scale=ns, norm=norm).to(dev).eval()
    net.load_state_dict(sd, strict=False)
    x = kp_maps(B, seed=7 + B)
    flow, vis, _, _ = net(x.to(dev))
    assert net.last_launch_count() >= 10   # the CUDA path ran (no fallback exists)
    want_flow, want_vis, _, _ = FO.flow_unet_forward(sd, x, nf, ss, ns, norm)
    ef, ev = (flow.cpu() - want_flow).abs().max().item(), (vis.cpu() - want_vis).abs().max().item()
    scale = max(1.0, want_flow.abs().max().item(), want_vis.abs().max().item())
    assert ef <= 1e-3 * scale and ev <= 1e-3 * scale, (ef, ev, scale)
    # the tensors the generator consumes: equal wherever the visibility arg-max is not a near-tie
    iw, ifm = net.warp_tensors(x.to(dev))
    want_iw, want_ifm = FO.warp_outputs(want_flow, want_vis)
    top2 = want_vis.topk(2, dim=1).values
    near_tie = ((top2[:, :1] - top2[:, 1:]) < 1e-3 * scale).float()
    frac_tie = near_tie.mean().item()
    bad = ((ifm.cpu() - want_ifm).abs() > 1e-6).float().mean().item()
    assert bad <= 4 * frac_tie + 1e-4, (bad, frac_tie)
    same = (ifm.cpu() - want_ifm).abs() <= 1e-6
    d = ((iw.cpu() - want_iw).abs() * same).max().item()
    assert d <= 20 * 8 / 7 * 1e-3 * scale + 1e-4, d
    # run-to-run and batch-vs-single determinism
    flow2, vis2, _, _ = net(x.to(dev))
    assert torch.equal(flow, flow2) and torch.equal(vis, vis2)
    one, _, _, _ = net(x[:1].to(dev))
    assert torch.equal(one, flow[:1])


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_flow_network_warp_from_landmarks():
    """geomcgt_ifw_test_model.py:62-76 end to end: landmarks -> key-point maps (GPU) -> netF -> iw_flow / mask."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda", 0)
    nf, ss, ns, norm = 16, 2, 4, "batch"
    sd = FO.make_state_dict(136, nf, ss, ns, norm, seed=5)
    net = FlowUnet(136, nf=nf, start_scale=ss, num_scale=ns, norm=norm).to(dev).eval()
    net.load_state_dict(sd, strict=False)
    src, seq = OC.landmark_sequence(3, seed=11)
    lm1 = torch.from_numpy(np.repeat(src[None], 3, 0))
    lm2 = torch.from_numpy(seq)
    iw, ifm = flow_network_warp(net, None, lm1.to(dev), lm2.to(dev))
    x = torch.from_numpy(np.concatenate([OC.kp_to_map(lm1.numpy() * 7 / 8), OC.kp_to_map(lm2.numpy() * 7 / 8)], 1))
    want_iw, want_ifm = FO.warp_outputs(*FO.flow_unet_forward(sd, x, nf, ss, ns, norm)[:2])
    same = (ifm.cpu() - want_ifm).abs() <= 1e-6
    assert same.float().mean().item() >= 0.999
    assert ((iw.cpu() - want_iw).abs() * same).max().item() <= 0.05
    assert iw.shape == (3, 2, 256, 256) and ifm.shape == (3, 1, 256, 256)
    # the landmark-level entry point draws the same maps into the operand itself and takes the boxes of the zero-skipping
    # conv from the coordinates (supersets of the exact ones): bit-identical to the tensor-level call, and a shared source
    # landmark set [68,2] equals B copies of it
    iw2, ifm2 = net.warp_tensors(x.to(dev))
    assert torch.equal(iw, iw2) and torch.equal(ifm, ifm2)
    iw3, ifm3 = net.warp_landmarks(torch.from_numpy(src).to(dev), lm2.to(dev))
    assert torch.equal(iw, iw3) and torch.equal(ifm, ifm3)
    lm_missing = lm2.clone()
    lm_missing[1, 5] = -8.0 / 7.0        # scales to exactly -1: the reference's "missing point" marker -> an empty map
    xm = torch.from_numpy(np.concatenate([OC.kp_to_map(lm1.numpy() * 7 / 8), OC.kp_to_map(lm_missing.numpy() * 7 / 8)], 1))
    assert float(xm[1, 68 + 5].sum()) == 0.0
    iw4, ifm4 = net.warp_landmarks(lm1.to(dev), lm_missing.to(dev))
    iw5, ifm5 = net.warp_tensors(xm.to(dev))
    assert torch.equal(iw4, iw5) and torch.equal(ifm4, ifm5)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_clip_renderer_makes_its_own_flow_with_netF():
    """ClipRenderer(netF=...) == ClipRenderer fed with flow_network_warp's tensors (the reference's set_input + forward)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import animateportrait_b200 as ap
    from animateportrait_b200 import synth
    from animateportrait_b200.clip import ClipRenderer
    dev = torch.device("cuda", 0)
    T = 5
    netF = FlowUnet(136, nf=16, start_scale=2, num_scale=4, norm="batch").to(dev).eval()
    netF.load_state_dict(FO.make_state_dict(136, 16, 2, 4, "batch", seed=2, gain=0.08), strict=False)
    net = ap.define_G(3, 1, 64, ap.NETG_NAME, "instance", False, "normal", 0.02, [0], div=3, disp=3)
    net.module.load_state_dict(synth.make_state_dict(1, seed=3, bias_std=0.3))
    photo, matte, static, src, seq, _, _ = synth.make_clip(T, output_nc=1, seed=33)
    r = ClipRenderer(net, batch=2, netF=netF)
    r.set_photo(photo.to(dev), src.to(dev), matte.to(dev), static.to(dev))
    got = r.render(seq.to(dev))
    iw, ifm = flow_network_warp(netF, None, src.to(dev)[None].expand(T, -1, -1), seq.to(dev))
    assert 0.0 < ifm.mean().item() and iw.abs().max().item() > 0.0
    want = ClipRenderer(net, batch=2)
    want.set_photo(photo.to(dev), src.to(dev), matte.to(dev), static.to(dev))
    assert torch.equal(got, want.render(seq.to(dev), iw, ifm))


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("cfg", [(16, 2, 4, "batch"), (8, 1, 5, "instance"), (32, 2, 4, "instance"), (64, 2, 3, "batch")])
def test_fast_kernels_agree_with_the_generic_kernel(cfg, monkeypatch):
    """The default path (box-restricted first conv, tcgen05 convs where Cin % 64 == 0, register-tiled fp32 convs
    elsewhere; csrc/flownet.cu) against the fp32 kernels alone (AP_FLOW_UMMA=0) and against the one generic kernel it all
    replaced (AP_FLOW_SPARSE=0 AP_FLOW_TILED=0 AP_FLOW_UMMA=0): same sums, other association / a bf16 hi-lo split."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda", 0)
    nf, ss, ns, norm = cfg
    sd = FO.make_state_dict(136, nf, ss, ns, norm, seed=21)
    x = kp_maps(3, seed=5).to(dev)

    def run():
        net = FlowUnet(136, nf=nf, start_scale=ss, num_scale=ns, norm=norm).to(dev).eval()
        net.load_state_dict(sd, strict=False)
        flow, vis, _, _ = net(x)
        one, _, _, _ = net(x[1:2])
        assert torch.equal(one, flow[1:2])      # a frame does not depend on its batch, whatever the kernels
        return flow.cpu(), vis.cpu(), net.last_launch_count()

    flow, vis, n_fast = run()
    monkeypatch.setenv("AP_FLOW_UMMA", "0")
    flow1, vis1, n_ffma = run()
    monkeypatch.setenv("AP_FLOW_SPARSE", "0")
    monkeypatch.setenv("AP_FLOW_TILED", "0")
    flow0, vis0, n_plain = run()
    assert n_fast == n_ffma == n_plain + 2   # the box pass and the sparse kernel; the other kernels replace one for one
    scale = max(1.0, flow0.abs().max().item(), vis0.abs().max().item())
    assert (flow1 - flow0).abs().max().item() <= 2e-4 * scale and (vis1 - vis0).abs().max().item() <= 2e-4 * scale
    assert (flow - flow0).abs().max().item() <= 5e-4 * scale and (vis - vis0).abs().max().item() <= 5e-4 * scale


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_dense_operand_stays_with_the_dense_first_conv():
    """ap_flow_forward assumes nothing about its operand: in one batch, an image of key-point discs takes the sparse
    first conv and an image of noise the dense one (per-image decision on the device); both match the oracle, and the
    disc image equals its own single-image result bit for bit."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda", 0)
    nf, ss, ns, norm = 16, 2, 4, "instance"
    sd = FO.make_state_dict(136, nf, ss, ns, norm, seed=9)
    net = FlowUnet(136, nf=nf, start_scale=ss, num_scale=ns, norm=norm).to(dev).eval()
    net.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(4)
    x = kp_maps(3, seed=6)
    x[1] = 0.05 * torch.randn(136, 224, 224, generator=g)
    x[2, 7] = 0.0          # an empty plane
    flow, vis, _, _ = net(x.to(dev))
    want_flow, want_vis, _, _ = FO.flow_unet_forward(sd, x, nf, ss, ns, norm)
    scale = max(1.0, want_flow.abs().max().item(), want_vis.abs().max().item())
    assert (flow.cpu() - want_flow).abs().max().item() <= 1e-3 * scale
    assert (vis.cpu() - want_vis).abs().max().item() <= 1e-3 * scale
    one, _, _, _ = net(x[:1].to(dev))
    two, _, _, _ = net(x[1:2].to(dev))
    assert torch.equal(one, flow[:1]) and torch.equal(two, flow[1:2])
