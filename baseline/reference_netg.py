"""Loader of the UNMODIFIED reference generator from baseline/_ref (vendored by tools/prep_ref.py, git-ignored).

Test / bench infrastructure only (the `--impl reference` arm of bench.py): the product package never imports this.
`skimage.measure.compare_ssim/compare_psnr` -- imported by the reference's flow modules, removed upstream and not
installed here -- is the only import blocker; a two-symbol stub stands in for it.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODULE2 = os.path.join(ROOT, "baseline", "_ref", "Module2")
NETG_NAME = "resnet_9blocks_rcatland32_full_ifw"


def available() -> bool:
    return os.path.exists(os.path.join(REF_MODULE2, "models", "networks.py"))


def import_networks():
    """The reference's `models.networks` module, imported from the vendored files."""
    if not available():
        raise ImportError(f"{REF_MODULE2} not found: run tools/prep_ref.py where /root/reference exists")
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.measure")
    skm.compare_ssim = skm.compare_psnr = None
    sk.measure = skm
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.measure", skm)
    if REF_MODULE2 not in sys.path:
        sys.path.insert(0, REF_MODULE2)
    from models import networks  # type: ignore
    return networks


def make_reference_netG(output_nc: int, state_dict=None):
    """networks.define_G(...) exactly as Module2/models/geomcgt_ifw_test_model.py:207-209 calls it (CPU, gpu_ids=[])."""
    networks = import_networks()
    net = networks.define_G(3, output_nc, 64, NETG_NAME, "instance", False, "normal", 0.02, [], div=3, disp=3)
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    return net.eval()
