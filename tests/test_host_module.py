"""CPU: the host-side mirror keeps the reference generator's interface (constructor, state_dict layout,
error behaviour) -- Module2/models/networks.py:123-201,1190-1340."""
import functools
import inspect

import pytest
import torch
import torch.nn as nn

import animateportrait_b200 as ap
from oracle import netg_oracle as O


def _make(onc=1, **kw):
    return ap.define_G(3, onc, 64, ap.NETG_NAME, "instance", False, "normal", 0.02, [], div=3, disp=3, **kw)


@pytest.mark.parametrize("onc", [1, 3])
def test_state_dict_layout_is_the_reference_checkpoint_layout(onc):
    net = _make(onc)
    sd = net.state_dict()
    spec = O.state_dict_spec(onc)
    assert list(sd.keys()) == list(spec.keys())  # same keys, same ORDER, no 'module.' prefix
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    # a checkpoint written by the reference (torch.save(net.state_dict())) loads strictly
    res = net.load_state_dict(O.make_state_dict(onc, seed=1))
    assert not res.missing_keys and not res.unexpected_keys


def test_load_networks_style_key_walk():
    # BaseModel.__patch_instance_norm_state_dict walks getattr(module, part) along every key (base_model.py:165-177)
    net = _make(1)
    for key in net.state_dict():
        obj = net
        for part in key.split(".")[:-1]:
            obj = getattr(obj, part)
        assert hasattr(obj, key.split(".")[-1])


def test_init_weights_is_normal_0_02_and_zero_bias():
    torch.manual_seed(0)
    net = _make(1)
    w = net.model2[1].conv_block[1].weight
    assert abs(w.std().item() - 0.02) < 1e-3 and abs(w.mean().item()) < 1e-3
    assert all(float(p.abs().sum()) == 0.0 for n, p in net.named_parameters() if n.endswith(".bias"))


def test_define_g_signature_matches_reference():
    params = list(inspect.signature(ap.define_G).parameters)
    assert params[:15] == ["input_nc", "output_nc", "ngf", "netG", "norm", "use_dropout", "init_type", "init_gain",
                           "gpu_ids", "model0_res", "model1_res", "extra_channel", "div", "disp", "regarch"]
    fwd = list(inspect.signature(ap.ResnetConditionTriGenerator32_full_ifw.forward).parameters)
    # the reference's six positional tensors (networks.py:1315); `out=` is a keyword-only-in-practice extension with a default
    assert fwd[:7] == ["self", "input", "land1", "land2", "motion", "flow", "ifmask"] and fwd[7:] == ["out"]
    assert inspect.signature(ap.ResnetConditionTriGenerator32_full_ifw.forward).parameters["out"].default is None


def test_unsupported_configurations_raise_like_the_reference():
    with pytest.raises(NotImplementedError):
        ap.define_G(3, 1, 64, "resnet_9blocks", "instance")
    with pytest.raises(NotImplementedError):
        ap.define_G(3, 1, 64, ap.NETG_NAME, "batch", div=3, disp=3)
    with pytest.raises(NotImplementedError):
        ap.define_G(3, 2, 64, ap.NETG_NAME, "instance", div=3, disp=3)
    with pytest.raises(NotImplementedError):
        ap.define_G(3, 1, 32, ap.NETG_NAME, "instance", div=3, disp=3)
    with pytest.raises(NotImplementedError):
        ap.define_G(3, 1, 64, ap.NETG_NAME, "instance", div=3, disp=1)  # reference default disp=1: other block layout
    with pytest.raises(NotImplementedError):
        ap.define_G(3, 1, 64, ap.NETG_NAME, "instance", div=3, disp=3, init_type="xavier")


def test_no_cpu_fallback():
    net = _make(1)
    x, l1, l2, motion, flow, ifmask = O.make_inputs(1, seed=1)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        net(x, l1, l2, motion, flow, ifmask)
    with pytest.raises(RuntimeError):
        net.model2[0](x)  # parameter holders are not executable


def test_product_package_never_imports_the_oracle():
    import os
    root = os.path.dirname(os.path.abspath(ap.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_install_patches_a_networks_namespace():
    import types
    fake = types.SimpleNamespace(ResnetConditionTriGenerator32_full_ifw=None)
    ap.install(fake, precision="bf16")
    norm = functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    net = fake.ResnetConditionTriGenerator32_full_ifw(3, 1, 64, norm_layer=norm, use_dropout=False, n_blocks=9, div=3, disp=3)
    assert isinstance(net, ap.ResnetConditionTriGenerator32_full_ifw) and net.precision == "bf16"


def test_package_synthetic_workload_is_the_oracles_recipe():
    """bench.py / smoke take weights and inputs from animateportrait_b200.synth (the product never imports the
    oracle); the oracle's own copy must stay bit-identical."""
    from animateportrait_b200 import synth as S
    for onc in (1, 3):
        a, b = S.make_state_dict(onc, seed=3, bias_std=0.1), O.make_state_dict(onc, seed=3, bias_std=0.1)
        assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    for kind in ("smooth", "noise"):
        x, y = S.make_inputs(2, seed=5, kind=kind), O.make_inputs(2, seed=5, kind=kind)
        assert all(torch.equal(p, q) for p, q in zip(x, y))
    assert S.flops_per_frame(1) == O.flops_per_frame(1) == 140125405184.0
    assert S.flops_per_frame(3) == O.flops_per_frame(3)


REF_MODULE2 = "/root/reference/Module2"


@pytest.mark.skipif(not __import__("os").path.isdir(REF_MODULE2), reason="reference tree not present (GPU box)")
def test_sitecustomize_shim_swaps_the_class_inside_the_reference_define_G():
    """The drop-in boundary, end to end on the host side: a fresh interpreter with the shim on PYTHONPATH imports
    the UNMODIFIED reference `models.networks`; its own define_G (networks.py:175-176) then builds the B200 module,
    and a checkpoint in the reference layout loads through load_state_dict."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from models import networks\n"
        "import animateportrait_b200 as ap\n"
        "net = networks.define_G(3, 1, 64, 'resnet_9blocks_rcatland32_full_ifw', 'instance', False, 'normal', 0.02, [], div=3, disp=3)\n"
        "assert isinstance(net, ap.ResnetConditionTriGenerator32_full_ifw), type(net)\n"
        "from animateportrait_b200 import synth\n"
        "r = net.load_state_dict(synth.make_state_dict(1, seed=2))\n"
        "assert not r.missing_keys and not r.unexpected_keys\n"
        "other = networks.define_G(3, 1, 64, 'resnet_9blocks', 'instance')\n"
        "assert type(other).__name__ == 'ResnetGenerator'\n"
        "print('SHIM_OK', len(net.state_dict()))\n" % REF_MODULE2)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "animateportrait_b200", "shim"), root]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=REF_MODULE2, timeout=300)
    assert "SHIM_OK 74" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]


def test_native_handle_is_never_shared_with_replicas_or_copies():
    """nn.DataParallel.replicate copies the module's __dict__ and copy.deepcopy cannot carry a ctypes pointer: every such
    copy must start without a native handle (and create its own on first use) instead of sharing -- and later freeing --
    the original's."""
    import copy
    net = ap.ResnetConditionTriGenerator32_full_ifw(3, 1, 64, norm_layer=ap.get_norm_layer("instance"), n_blocks=9, div=3, disp=3)
    net._handle = 1234            # stands in for a live ap_netg*
    net._handle_device = 0
    rep = net._replicate_for_data_parallel()
    assert rep._handle is None and rep._handle_owner == id(rep) and net._handle == 1234
    for c in (copy.copy(net), copy.deepcopy(net)):
        assert c._handle is None and c._handle_owner == id(c) and c._dirty
    assert len(dict(rep._weight_tensors())) == 0 or len(dict(rep._weight_tensors())) == 74
    assert sorted(dict(net._weight_tensors())) == sorted(net.state_dict())
    net._handle = None


def test_in_place_weight_changes_are_seen_without_a_hint():
    net = ap.ResnetConditionTriGenerator32_full_ifw(3, 1, 64, norm_layer=ap.get_norm_layer("instance"), n_blocks=9, div=3, disp=3)
    a = net._weights_signature()
    with torch.no_grad():
        net.model3[7].bias.add_(1.0)
    b = net._weights_signature()
    net.load_state_dict(net.state_dict())
    assert a != b and net._weights_signature() != b
