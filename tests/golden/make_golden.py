"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference generator.

Run in the build container only (needs /root/reference; the GPU box has no copy of it):

    python tests/golden/make_golden.py

What it does, per fixture case:
  1. imports the reference `models.networks` from /root/reference/Module2 (with a 2-symbol stub for
     the absent `skimage.measure`, the only import blocker -- SURVEY.md §8c),
  2. builds netG with the reference's own `define_G(3, onc, 64, 'resnet_9blocks_rcatland32_full_ifw',
     'instance', False, 'normal', 0.02, [], div=3, disp=3)` (call site
     Module2/models/geomcgt_ifw_test_model.py:207-209),
  3. loads the seeded stand-in checkpoint through `load_state_dict` (the checkpoint layout the
     reference's `BaseModel.load_networks`, base_model.py:179-202, uses),
  4. runs the reference forward on the seeded inputs, with forward hooks capturing intermediates,
  5. asserts the oracle restatement (oracle/netg_oracle.py) reproduces output and every tap,
  6. stores output + per-tap statistics + small strided samples in tests/golden/<case>.npz.

It also records the survey's define_G(seed 0) anchors in anchors.json so the
reference import itself is sanity-checked.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/Module2"

from oracle import netg_oracle as O  # noqa: E402

CASES = {
    # name: (output_nc, B, weight seed, bias_std, input seed, input kind)
    "c1_line_smooth": (1, 1, 0, 0.0, 1001, "smooth"),
    "c1_line_noise": (1, 1, 0, 0.0, 1002, "noise"),
    "c1_line_bias": (1, 1, 7, 0.5, 1003, "smooth"),
    "c5_cartoon_noise": (3, 1, 3, 0.5, 1005, "noise"),
    "b2_line_mixed": (1, 2, 11, 0.1, 1004, "smooth"),
}


def import_reference():
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.measure")
    skm.compare_ssim = skm.compare_psnr = None
    sk.measure = skm
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.measure", skm)
    sys.path.insert(0, REF)
    from models import networks  # type: ignore
    return networks


def sample_tap(v: torch.Tensor) -> np.ndarray:
    """Small deterministic sample: 3 channels x border ring rows/cols + strided interior."""
    B, C, H, W = v.shape
    ch = sorted(set([0, C // 2, C - 1]))
    s = max(H // 16, 1)
    rows = v[:, ch][:, :, [0, 1, H // 2, H - 2, H - 1], :]
    grid = v[:, ch][:, :, ::s, ::s]
    return np.concatenate([rows.reshape(-1).numpy(), grid.reshape(-1).numpy()]).astype(np.float32)


def run_reference(networks, onc, sd, inputs):
    net = networks.define_G(3, onc, 64, O.NETG_NAME, "instance", False, "normal", 0.02, [], div=3, disp=3)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(net.state_dict().keys()) == list(sd.keys()), "state_dict key order differs from Appendix B"
    taps = {}

    def hook(name):
        def f(mod, inp, out):
            taps.setdefault(name, out.detach().clone())
        return f

    names = {"model_tri00": "tri00", "model_tri01": "tri01", "model_tri02": "tri02", "model_tri10": "tri10",
             "model_tri11": "tri11", "model_tri12": "tri12", "model_tri20": "tri20", "model_tri21": "tri21",
             "model_tri22": "tri22", "model_tri_merge": "merge"}
    for mname, tname in names.items():
        getattr(net, mname).register_forward_hook(hook(tname))
    for i in range(9):
        net.model2[i].register_forward_hook(hook(f"block{i}"))
    net.model3[7].register_forward_hook(hook("pre_tanh"))
    lcount = [0]

    def lhook(mod, inp, out):
        lcount[0] += 1
        taps[f"land{lcount[0]}"] = out.detach().clone()

    net.model_landmark_trans.register_forward_hook(lhook)
    orig = net.double_feature_warping

    def wrapped(x, motion, flow, ifmask, level):
        out = orig(x, motion, flow, ifmask, level)
        taps[f"warp{level}"] = out.detach().clone()
        return out

    net.double_feature_warping = wrapped
    with torch.no_grad():
        y = net(*inputs)
    return y, taps


def main():
    torch.set_num_threads(os.cpu_count())
    networks = import_reference()
    manifest = {}
    for name, (onc, B, wseed, bstd, iseed, kind) in CASES.items():
        sd = O.make_state_dict(onc, seed=wseed, bias_std=bstd)
        inputs = O.make_inputs(B, seed=iseed, kind=kind)
        y_ref, taps_ref = run_reference(networks, onc, sd, inputs)
        taps_or = {}
        y_or = O.netg_forward(sd, *inputs, tap=lambda n, v: taps_or.__setitem__(n, v.clone()))
        d = (y_or - y_ref).abs().max().item()
        assert d <= 1e-6, f"{name}: oracle vs reference output differs by {d}"
        assert set(taps_or) == set(taps_ref), (sorted(taps_or), sorted(taps_ref))
        worst = 0.0
        for k in taps_ref:
            dk = (taps_or[k] - taps_ref[k]).abs().max().item()
            worst = max(worst, dk)
            assert dk <= 1e-5, f"{name}: tap {k} differs by {dk}"
        # closed-form warp restatement vs reference taps
        feats = {0: taps_ref["tri00"], 1: taps_ref["tri11"], 2: taps_ref["tri22"]}
        cf = {}
        for lvl in (0, 1, 2):
            w = O.double_feature_warping_closed_form(feats[lvl], inputs[3], inputs[4], inputs[5], lvl)
            diff = (w - taps_ref[f"warp{lvl}"]).abs()
            cf[lvl] = (diff.max().item(), (diff > 1e-3).float().mean().item())
        out = {"y": y_ref.numpy().astype(np.float32)}
        stats = {}
        for k, v in taps_ref.items():
            out["tap_" + k] = sample_tap(v)
            stats[k] = {"shape": list(v.shape), "mean": v.mean().item(), "std": v.std().item(),
                        "absmax": v.abs().max().item(), "first": v.flatten()[0].item()}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        wsum = sum(float(t.double().sum()) for t in sd.values())
        wabs = sum(float(t.double().abs().sum()) for t in sd.values())
        isum = [float(t.double().sum()) for t in inputs]
        manifest[name] = {"output_nc": onc, "B": B, "weight_seed": wseed, "bias_std": bstd,
                          "input_seed": iseed, "input_kind": kind, "weight_sum": wsum, "weight_abs_sum": wabs,
                          "input_sums": isum, "oracle_vs_reference_out": d, "oracle_vs_reference_worst_tap": worst,
                          "closed_form_warp_maxabs_fracbad": cf, "taps": stats,
                          "y_mean": y_ref.mean().item(), "y_absmean": y_ref.abs().mean().item()}
        print(name, "out diff", d, "worst tap", worst, "closed-form warp", cf)

    # survey anchors: the reference's own define_G init with torch.manual_seed(0)
    anchors = {}
    for onc in (1, 3):
        torch.manual_seed(0)
        net = networks.define_G(3, onc, 64, O.NETG_NAME, "instance", False, "normal", 0.02, [], div=3, disp=3).eval()
        nparam = sum(p.numel() for p in net.parameters())
        wsum = sum(float(p.double().sum()) for p in net.parameters())
        wabs = sum(float(p.double().abs().sum()) for p in net.parameters())
        keys = list(net.state_dict().keys())
        shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
        spec = O.state_dict_spec(onc)
        assert keys == list(spec.keys())
        assert all(tuple(shapes[k]) == tuple(spec[k]) for k in keys)
        g = torch.Generator().manual_seed(1)
        x = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
        l1 = (torch.rand(1, 1, 256, 256, generator=g) > 0.98).float() * 2 - 1
        l2 = (torch.rand(1, 1, 256, 256, generator=g) > 0.98).float() * 2 - 1
        ys, xs = torch.meshgrid(torch.linspace(-1, 1, 256), torch.linspace(-1, 1, 256), indexing="ij")
        base = torch.stack([xs, ys], -1)[None]
        motion = base + 0.05 * torch.randn(1, 256, 256, 2, generator=g)
        flow = 4 * torch.randn(1, 2, 256, 256, generator=g)
        ifmask = torch.rand(1, 1, 256, 256, generator=g)
        with torch.no_grad():
            y = net(x, l1, l2, motion, flow, ifmask)
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        yo = O.netg_forward(sd, x, l1, l2, motion, flow, ifmask)
        anchors[str(onc)] = {"params": nparam, "weight_sum": wsum, "weight_abs_sum": wabs, "n_keys": len(keys),
                             "y_mean": y.mean().item(), "y_absmean": y.abs().mean().item(),
                             "y000": y[0, 0, 0, 0].item(), "y_center": y[0, 0, 128, 128].item(),
                             "oracle_vs_reference": (yo - y).abs().max().item(),
                             "flops_per_frame": O.flops_per_frame(onc)}
        print("anchor", onc, anchors[str(onc)])
    manifest["_anchors_define_G_seed0"] = anchors
    manifest["_torch"] = torch.__version__
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
