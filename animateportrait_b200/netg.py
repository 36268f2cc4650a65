"""Host-side mirror of the reference generator interface, backed by libapnetg.so.

`ResnetConditionTriGenerator32_full_ifw` here has the constructor, the submodule tree (hence the
74-key state_dict) and the `forward(input, land1, land2, motion, flow, ifmask)`
signature of the reference class (Module2/models/networks.py:1190-1340), so
`networks.define_G(..., 'resnet_9blocks_rcatland32_full_ifw', ...)`, `init_net`/`init_weights`
(networks.py:71-120) and `BaseModel.load_networks` (base_model.py:179-202) work on it unchanged.
The nn.Conv2d / nn.ConvTranspose2d children only HOLD parameters; `forward` hands the six tensors to
the C ABI (`ap_netg_forward`, include/ap_netg.h) on the current CUDA stream.  There is no PyTorch
fallback: without the built library or without a B200 the module raises.
"""
from __future__ import annotations

import ctypes as C
import functools
import weakref
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _capi

NETG_NAME = "resnet_9blocks_rcatland32_full_ifw"


def _norm_is_instance(norm_layer) -> bool:
    f = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    return f is nn.InstanceNorm2d


class _Holder(nn.Module):
    """Parameter holder: never executed."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder of the B200 generator; call the generator, not its children")


class ResnetBlock(_Holder):
    """Holder with the key layout of networks.py:2303-2361: conv_block.{1,5}."""

    def __init__(self, dim, norm_layer):
        super().__init__()
        self.conv_block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3), norm_layer(dim), nn.ReLU(True),
                                        nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3), norm_layer(dim))


class ResnetBlock2(_Holder):
    """Holder with the key layout of networks.py:2363-2421: conv_block.{1,5} and shortcut.0."""

    def __init__(self, dim_in, dim_out, norm_layer):
        super().__init__()
        self.conv_block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim_in, dim_out, 3), norm_layer(dim_out),
                                        nn.ReLU(True), nn.ReflectionPad2d(1), nn.Conv2d(dim_out, dim_out, 3),
                                        norm_layer(dim_out))
        self.shortcut = nn.Sequential(nn.Conv2d(dim_in, dim_out, 3, padding=1), norm_layer(dim_out))


class ResnetConditionTriGenerator32_full_ifw(nn.Module):
    """Drop-in for the reference class of the same name (networks.py:1190).

    Supported configuration (anything else raises NotImplementedError, the reference's own error
    style at networks.py:38,200): input_nc=3, ngf=64, InstanceNorm, no dropout, 9 blocks, reflect
    padding, div=3, disp=3, output_nc in {1, 3}.

    Extra keyword (not in the reference): precision = 'fp32' (default; bf16 hi/lo 3-product tcgen05,
    <=1e-3 of the fp32 reference), 'bf16', or 'fp32_simt' (CUDA-core validation path).
    """

    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False, n_blocks=6,
                 padding_type="reflect", div=3, disp=1, precision: str = "fp32"):
        super().__init__()
        if not (input_nc == 3 and ngf == 64 and output_nc in (1, 3) and n_blocks == 9 and div == 3 and disp == 3
                and padding_type == "reflect" and not use_dropout and _norm_is_instance(norm_layer)):
            raise NotImplementedError(
                "B200 generator supports input_nc=3, output_nc in {1,3}, ngf=64, norm=instance, no dropout, "
                f"n_blocks=9, padding=reflect, div=3, disp=3; got input_nc={input_nc}, output_nc={output_nc}, ngf={ngf}, "
                f"norm={norm_layer}, use_dropout={use_dropout}, n_blocks={n_blocks}, padding={padding_type}, div={div}, disp={disp}")
        if precision not in _capi.PRECISIONS:
            raise NotImplementedError(f"precision [{precision}] is not recognized")
        self.n_blocks, self.div, self.disp = n_blocks, div, disp
        self.output_nc = output_nc
        self.precision = precision
        nl = norm_layer

        def stem(cout):
            return nn.Sequential(nn.ReflectionPad2d(3), nn.Conv2d(input_nc, cout, 7), nl(cout), nn.ReLU(True))

        def down(cin, cout):
            return nn.Sequential(nn.Conv2d(cin, cout, 3, stride=2, padding=1), nl(cout), nn.ReLU(True))

        # registration order follows the reference so state_dict() key ORDER matches too (Appendix B)
        self.model_tri_merge = nn.Conv2d(ngf * 12, ngf * 4, 3, padding=1)
        self.model_tri00, self.model_tri01, self.model_tri02 = stem(ngf // 2), down(ngf, ngf * 2), down(ngf * 2, ngf * 4)
        self.model_tri10, self.model_tri11, self.model_tri12 = stem(ngf), down(ngf, ngf), down(ngf * 2, ngf * 4)
        self.model_tri20, self.model_tri21, self.model_tri22 = stem(ngf), down(ngf, ngf * 2), down(ngf * 2, ngf * 2)
        dim, con = ngf * 4, 16
        self.model2 = nn.Sequential(*[ResnetBlock2(dim + 2 * con, dim, nl) if (i + disp) % div == 0 else ResnetBlock(dim, nl)
                                      for i in range(n_blocks)])
        self.model3 = nn.Sequential(
            nn.ConvTranspose2d(ngf * 4, ngf * 2, 3, stride=2, padding=1, output_padding=1), nl(ngf * 2), nn.ReLU(True),
            nn.ConvTranspose2d(ngf * 2, ngf, 3, stride=2, padding=1, output_padding=1), nl(ngf), nn.ReLU(True),
            nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, 7), nn.Tanh())
        self.model_landmark_trans = nn.Sequential(
            nn.Conv2d(1, 8, 3, padding=1), nl(8), nn.ReLU(True),
            nn.Conv2d(8, con, 3, stride=2, padding=1), nl(con), nn.ReLU(True),
            nn.Conv2d(con, con, 3, stride=2, padding=1), nl(con))
        self._options: Dict[str, int] = {}
        self._reset_native()

    # ---- the native handle --------------------------------------------------------------------------
    # The handle belongs to ONE module object.  Copies made behind the module's back share its __dict__ entries
    # (nn.DataParallel.replicate: replica.__dict__ = self.__dict__.copy(); copy.copy) or cannot carry a ctypes pointer
    # (copy.deepcopy, pickle): every such copy starts without a handle and creates its own on first use, and the handle
    # is destroyed by a finalizer tied to the object that created it, never by a replica.
    def _reset_native(self):
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[int] = None
        self._handle_owner: int = id(self)
        self._finalizer = None
        self._dirty = True
        self._weights_seen = None

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica._reset_native()
        return replica

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_handle", "_handle_device", "_finalizer", "_weights_seen"):
            state[k] = None
        state["_dirty"] = True
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._handle_owner = id(self)

    @staticmethod
    def _destroy(handle):
        try:
            _capi.lib().ap_netg_destroy(handle)
        except Exception:
            pass

    def _release(self):
        if self.__dict__.get("_handle") is not None and self.__dict__.get("_handle_owner") == id(self):
            fin = self.__dict__.get("_finalizer")
            if fin is not None:
                fin()  # runs _destroy once
        self._handle, self._finalizer = None, None

    # ---- weight synchronisation with the library -------------------------------------------------
    def mark_weights_dirty(self):
        """Forces a re-upload at the next call.  Not needed after load_state_dict / .to() / in-place parameter changes
        (optimizer steps, net.apply(init_func), p.data.copy_): those are detected through the tensors' version counters."""
        self._dirty = True

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._dirty = True
        return r

    def _weight_tensors(self):
        """(checkpoint key, tensor) of the 74 conv weights / biases.  Walks the children instead of state_dict():
        DataParallel replicas carry their broadcast copies as plain attributes, not as registered parameters."""
        for name, m in self.named_modules():
            for pn in ("weight", "bias"):
                t = getattr(m, pn, None)
                if isinstance(t, torch.Tensor):
                    yield f"{name}.{pn}", t

    def _weights_signature(self):
        # (storage address, in-place version) of every parameter: changes whenever a parameter is rebound or written
        return tuple((t.data_ptr(), t._version) for _, t in self._weight_tensors())

    def set_option(self, name: str, value: int) -> None:
        """Execution options of the native handle (`ap_netg_set_option`): graphs, overlap, keep_intermediates."""
        self._options[name] = int(value)
        if self._handle is not None:
            _capi.check(_capi.lib().ap_netg_set_option(self._handle, name.encode(), int(value)), "ap_netg_set_option")

    def _sync(self, device: torch.device):
        lib = _capi.lib()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self.__dict__.get("_handle_owner") != id(self):  # a shallow copy of another module: never touch its handle
            self._reset_native()
        if self._handle is None or self._handle_device != idx:
            self._release()
            h = C.c_void_p()
            _capi.check(lib.ap_netg_create(C.byref(h), self.output_nc, _capi.PRECISIONS[self.precision], idx),
                        "ap_netg_create")
            self._handle, self._handle_device = h, idx
            self._finalizer = weakref.finalize(self, ResnetConditionTriGenerator32_full_ifw._destroy, h)
            for k, v in self._options.items():
                _capi.check(lib.ap_netg_set_option(h, k.encode(), v), "ap_netg_set_option")
            self._dirty = True
        sig = self._weights_signature()
        if sig != self._weights_seen:
            self._dirty = True
        if self._dirty:
            sd = dict(self._weight_tensors())
            keep = []
            names = (C.c_char_p * len(sd))()
            ptrs = (C.c_void_p * len(sd))()
            shapes = (C.c_int64 * (4 * len(sd)))()
            for i, (k, v) in enumerate(sd.items()):
                t = v.detach().to(device=device, dtype=torch.float32).contiguous()
                keep.append(t)
                names[i] = k.encode()
                ptrs[i] = t.data_ptr()
                shp = list(t.shape) + [1] * (4 - t.dim())
                for j in range(4):
                    shapes[4 * i + j] = shp[j]
            stream = torch.cuda.current_stream(device).cuda_stream
            _capi.check(lib.ap_netg_load_weights(self._handle, len(sd), names, ptrs, shapes, 1, C.c_void_p(stream)),
                        "ap_netg_load_weights")
            self._dirty = False
            self._weights_seen = sig

    # ---- the reference interface ----------------------------------------------------------------
    supports_out = True  # forward(..., out=) writes the frames in place (frames.render_frames_sharded uses it)

    def _out_tensor(self, out, B, dev):
        if out is None:
            return torch.empty((B, self.output_nc, 256, 256), device=dev, dtype=torch.float32)
        # `out` may live on ANOTHER GPU of the box (a peer-mapped gather buffer, frames.PeerBuffer): the output kernel
        # stores through NVLink; the caller has enabled peer access
        if (tuple(out.shape) != (B, self.output_nc, 256, 256) or out.dtype != torch.float32 or not out.is_cuda
                or not out.is_contiguous()):
            raise RuntimeError(f"out: expected a contiguous CUDA fp32 tensor {(B, self.output_nc, 256, 256)}, got "
                               f"{tuple(out.shape)} {out.dtype} on {out.device}")
        return out

    def forward(self, input, land1, land2, motion, flow, ifmask, out=None):
        """networks.py:1315: returns a new [B, output_nc, 256, 256] fp32 tensor on input's device (or fills `out`, an
        extension the reference does not have)."""
        if not input.is_cuda:
            raise RuntimeError("the B200 generator runs on CUDA tensors only (no CPU fallback); "
                               "use forward_host() for host buffers")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("the B200 generator is inference-only: call it under torch.no_grad() "
                               "(the reference's BaseModel.test does, base_model.py:105-107)")
        dev = input.device
        B = input.shape[0]
        want = {"input": (B, 3, 256, 256), "land1": (B, 1, 256, 256), "land2": (B, 1, 256, 256),
                "motion": (B, 256, 256, 2), "flow": (B, 2, 256, 256), "ifmask": (B, 1, 256, 256)}
        ts = []
        for (name, shape), t in zip(want.items(), (input, land1, land2, motion, flow, ifmask)):
            if tuple(t.shape) != shape:
                raise RuntimeError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
            # the training path hands `motion` over on the CPU (geomgm_ifw_fore_model.py:455,530)
            ts.append(t.detach().to(device=dev, dtype=torch.float32).contiguous())
        with torch.cuda.device(dev):
            self._sync(dev)
            out = self._out_tensor(out, B, dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(_capi.lib().ap_netg_forward(self._handle, B, *[C.c_void_p(t.data_ptr()) for t in ts],
                                                    C.c_void_p(out.data_ptr()), C.c_void_p(stream)), "ap_netg_forward")
        return out

    @torch.no_grad()
    def forward_shared_photo(self, input, land1, land2, motion, flow, ifmask, out=None):
        """Clip form of forward(): `input` is ONE photo [1,3,256,256] and `land1` its landmark map [1,1,256,256], shared
        by the B frames the other four tensors describe (`ap_netg_forward_shared_photo`): the layers that depend on
        them alone run once per call.  Same result as forward(input.expand(B, ...), land1.expand(B, ...), ...)."""
        if not input.is_cuda:
            raise RuntimeError("the B200 generator runs on CUDA tensors only (no CPU fallback)")
        dev = input.device
        B = land2.shape[0]
        want = {"input": (1, 3, 256, 256), "land1": (1, 1, 256, 256), "land2": (B, 1, 256, 256),
                "motion": (B, 256, 256, 2), "flow": (B, 2, 256, 256), "ifmask": (B, 1, 256, 256)}
        ts = []
        for (name, shape), t in zip(want.items(), (input, land1, land2, motion, flow, ifmask)):
            if tuple(t.shape) != shape:
                raise RuntimeError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
            ts.append(t.detach().to(device=dev, dtype=torch.float32).contiguous())
        with torch.cuda.device(dev):
            self._sync(dev)
            out = self._out_tensor(out, B, dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(_capi.lib().ap_netg_forward_shared_photo(self._handle, B, *[C.c_void_p(t.data_ptr()) for t in ts],
                                                                 C.c_void_p(out.data_ptr()), C.c_void_p(stream)),
                        "ap_netg_forward_shared_photo")
        return out

    @torch.no_grad()
    def forward_host(self, input, land1, land2, motion, flow, ifmask, device: Optional[torch.device] = None,
                     out: Optional[torch.Tensor] = None):
        """End-to-end form: CPU tensors in (pinned for full copy speed), CPU frames out.
        H2D copies + forward + D2H copy + stream sync happen inside `ap_netg_forward_host`."""
        dev = torch.device(device if device is not None else next(self.parameters()).device)
        if dev.type != "cuda":
            raise RuntimeError("forward_host needs the module's parameters (or `device`) on a CUDA device")
        B = input.shape[0]
        ts = [t.detach().to(dtype=torch.float32).contiguous() for t in (input, land1, land2, motion, flow, ifmask)]
        if any(t.is_cuda for t in ts):
            raise RuntimeError("forward_host takes host tensors")
        if out is None:
            out = torch.empty((B, self.output_nc, 256, 256), dtype=torch.float32, pin_memory=True)
        with torch.cuda.device(dev):
            self._sync(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(_capi.lib().ap_netg_forward_host(self._handle, B, *[C.c_void_p(t.data_ptr()) for t in ts],
                                                         C.c_void_p(out.data_ptr()), C.c_void_p(stream)),
                        "ap_netg_forward_host")
        return out

    @torch.no_grad()
    def forward_host_async(self, input, land1, land2, motion, flow, ifmask, out: torch.Tensor,
                           device: Optional[torch.device] = None) -> torch.Tensor:
        """Pipelined forward_host (`ap_netg_forward_host_async`): returns at once; up to two calls are in flight, the
        uploads of one overlapping the forward of the other.  Keep the (pinned) host tensors alive and do not read `out`
        before host_sync()."""
        dev = torch.device(device if device is not None else next(self.parameters()).device)
        if dev.type != "cuda":
            raise RuntimeError("forward_host_async needs the module's parameters (or `device`) on a CUDA device")
        B = input.shape[0]
        ts = [t.detach().to(dtype=torch.float32).contiguous() for t in (input, land1, land2, motion, flow, ifmask)]
        if any(t.is_cuda for t in ts) or out.is_cuda:
            raise RuntimeError("forward_host_async takes host tensors")
        if tuple(out.shape) != (B, self.output_nc, 256, 256) or out.dtype != torch.float32 or not out.is_contiguous():
            raise RuntimeError(f"out: expected a contiguous fp32 host tensor {(B, self.output_nc, 256, 256)}")
        self._host_keep = getattr(self, "_host_keep", [])[-12:] + ts   # the copies read them after this call returns
        with torch.cuda.device(dev):
            self._sync(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(_capi.lib().ap_netg_forward_host_async(self._handle, B, *[C.c_void_p(t.data_ptr()) for t in ts],
                                                               C.c_void_p(out.data_ptr()), C.c_void_p(stream)),
                        "ap_netg_forward_host_async")
        return out

    def host_sync(self) -> None:
        """Waits until every forward_host_async call has delivered its frames."""
        if self._handle is not None:
            _capi.check(_capi.lib().ap_netg_host_sync(self._handle), "ap_netg_host_sync")
            self._host_keep = []

    # ---- introspection used by tests / bench ------------------------------------------------------
    def last_launch_count(self) -> int:
        n = C.c_int64(0)
        _capi.check(_capi.lib().ap_netg_last_launch_count(self._handle, C.byref(n)), "ap_netg_last_launch_count")
        return int(n.value)

    PROFILE_CLASSES = ("stem7x7", "landmark", "trunk_conv3x3", "strided_convs", "in_apply", "warp", "out_conv")

    def set_profiling(self, enable: bool) -> None:
        _capi.check(_capi.lib().ap_netg_set_profiling(self._handle, 1 if enable else 0), "ap_netg_set_profiling")

    def get_profile(self) -> Dict[str, Dict[str, float]]:
        """Per kernel class of the last profiled forward: {'ms', 'launches', 'flops'} (device time, CUDA events)."""
        n = len(self.PROFILE_CLASSES)
        ms, la, fl, nc = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)(), C.c_int(0)
        _capi.check(_capi.lib().ap_netg_get_profile(self._handle, n, ms, la, fl, C.byref(nc)), "ap_netg_get_profile")
        return {self.PROFILE_CLASSES[i]: {"ms": ms[i], "launches": int(la[i]), "flops": fl[i]} for i in range(nc.value)}

    def workspace_bytes(self, B: int) -> int:
        n = C.c_size_t(0)
        _capi.check(_capi.lib().ap_netg_workspace_bytes(self._handle, B, C.byref(n)), "ap_netg_workspace_bytes")
        return int(n.value)

    def debug_read(self, tap: str) -> torch.Tensor:
        """NCHW fp32 copy of a named intermediate of the last forward (names: oracle taps + up0/up1)."""
        dev = torch.device("cuda", self._handle_device)
        shape = (C.c_int64 * 4)()
        cap = 64 * 1024 * 1024
        buf = torch.empty(cap, device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(_capi.lib().ap_netg_debug_read(self._handle, tap.encode(), C.c_void_p(buf.data_ptr()), cap, shape,
                                                   C.c_void_p(stream)), "ap_netg_debug_read")
        n = shape[0] * shape[1] * shape[2] * shape[3]
        return buf[:n].view(shape[0], shape[1], shape[2], shape[3]).clone()


def get_norm_layer(norm_type="instance"):
    """networks.py:22-39, instance branch only."""
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def init_weights(net, init_type="normal", init_gain=0.02):
    """networks.py:71-102: N(0, gain) on every Conv weight, zero bias ('normal' only)."""
    if init_type != "normal":
        raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
    for m in net.modules():
        if hasattr(m, "weight") and m.__class__.__name__.find("Conv") != -1:
            nn.init.normal_(m.weight.data, 0.0, init_gain)
            if getattr(m, "bias", None) is not None:
                nn.init.constant_(m.bias.data, 0.0)
    for m in net.modules():
        if isinstance(m, ResnetConditionTriGenerator32_full_ifw):
            m.mark_weights_dirty()


def define_G(input_nc, output_nc, ngf, netG, norm="batch", use_dropout=False, init_type="normal", init_gain=0.02,
             gpu_ids=(), model0_res=0, model1_res=0, extra_channel=3, div=3, disp=1, regarch=4, precision="fp32"):
    """Same signature as networks.define_G (networks.py:123); only this generator is provided."""
    if netG != NETG_NAME:
        raise NotImplementedError("Generator model name [%s] is not recognized" % netG)
    net = ResnetConditionTriGenerator32_full_ifw(input_nc, output_nc, ngf, norm_layer=get_norm_layer(norm),
                                                 use_dropout=use_dropout, n_blocks=9, div=div, disp=disp,
                                                 precision=precision)
    gpu_ids = list(gpu_ids)
    if len(gpu_ids) > 0:  # init_net, networks.py:105-120
        assert torch.cuda.is_available()
        net.to(gpu_ids[0])
        net = torch.nn.DataParallel(net, gpu_ids[:1])
    init_weights(net, init_type, init_gain)
    return net


def install(networks_module, precision: str = "fp32") -> None:
    """Patch an imported reference `models.networks` so its own define_G builds the B200 generator."""
    cls = functools.partial(ResnetConditionTriGenerator32_full_ifw, precision=precision)
    networks_module.ResnetConditionTriGenerator32_full_ifw = cls


def conv2d_debug(x: torch.Tensor, w: torch.Tensor, stride=1, pad=1, pad_mode="zeros", transposed=False,
                 impl="fp32"):
    """One conv layer through the library's kernels (ap_conv2d_debug). Returns (raw output NCHW, stats [B,Cout,2])."""
    assert x.is_cuda and w.is_cuda
    B, Cin, H, W_ = x.shape
    Cout = w.shape[1] if transposed else w.shape[0]
    k = w.shape[2]
    Ho = 2 * H if transposed else H // stride
    y = torch.empty((B, Cout, Ho, Ho), device=x.device, dtype=torch.float32)
    st = torch.zeros((B, Cout, 2), device=x.device, dtype=torch.float64)
    x = x.contiguous().float()
    w = w.contiguous().float()
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _capi.check(_capi.lib().ap_conv2d_debug(_capi.PRECISIONS[impl], x.device.index or 0, B, H, W_, Cin, Cout, k, stride,
                                            pad, 1 if pad_mode == "reflect" else 0, 1 if transposed else 0,
                                            C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(y.data_ptr()),
                                            C.c_void_p(st.data_ptr()), C.c_void_p(stream)), "ap_conv2d_debug")
    return y, st
