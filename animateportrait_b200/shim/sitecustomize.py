"""Drop-in injection for the reference's Module2 subprocess.

`main_end2end_module2.py:108-110` starts `cd Module2/; python test.py ...` in a NEW interpreter, so the
B200 generator cannot be handed over in-process.  Putting this directory on PYTHONPATH makes that
interpreter import this `sitecustomize` at start-up; it registers a post-import hook that, the moment
the reference imports `models.networks` (Module2/models/networks.py), swaps its class
`ResnetConditionTriGenerator32_full_ifw` (networks.py:1190) for the B200-native one.  The reference's own
`define_G` (networks.py:175-176), `init_net`, `BaseModel.load_networks` and `GeomCGTIFWTestModel.forward`
(geomcgt_ifw_test_model.py:207-209,295) then run unchanged on top of libapnetg.so.

    PYTHONPATH=/path/to/repo/animateportrait_b200/shim:/path/to/repo python main_end2end_module2.py ...

Environment: AP_B200_PRECISION = fp32 (default) | bf16 | fp32_simt;  AP_B200_DISABLE=1 turns the hook off.
It also installs the two-symbol `skimage.measure` stub the reference needs on current scikit-image
(`compare_ssim`/`compare_psnr` were removed upstream; intrinsic_flow_models/modules.py:11) -- only when
the real names are missing.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

_TARGET = "models.networks"


def _stub_skimage():
    try:
        import skimage.measure as m  # noqa: F401
        if hasattr(m, "compare_ssim") and hasattr(m, "compare_psnr"):
            return
        m.compare_ssim = getattr(m, "compare_ssim", None)
        m.compare_psnr = getattr(m, "compare_psnr", None)
    except Exception:
        sk = sys.modules.get("skimage") or types.ModuleType("skimage")
        skm = types.ModuleType("skimage.measure")
        skm.compare_ssim = skm.compare_psnr = None
        sk.measure = skm
        sys.modules.setdefault("skimage", sk)
        sys.modules.setdefault("skimage.measure", skm)


def _patch(module):
    import animateportrait_b200 as ap
    ap.install(module, precision=os.environ.get("AP_B200_PRECISION", "fp32"))


class _Loader(importlib.abc.Loader):
    def __init__(self, inner):
        self.inner = inner

    def create_module(self, spec):
        return self.inner.create_module(spec)

    def exec_module(self, module):
        self.inner.exec_module(module)
        _patch(module)


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name != _TARGET:
            return None
        for f in sys.meta_path:
            if f is self or not hasattr(f, "find_spec"):
                continue
            spec = f.find_spec(name, path, target)
            if spec is not None and spec.loader is not None:
                spec.loader = _Loader(spec.loader)
                return spec
        return None


def activate():
    if os.environ.get("AP_B200_DISABLE") == "1":
        return
    _stub_skimage()
    if _TARGET in sys.modules:
        _patch(sys.modules[_TARGET])
    elif not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())


activate()
