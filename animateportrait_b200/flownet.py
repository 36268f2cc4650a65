"""Host-side mirror of the reference's intrinsic-flow network, backed by libapnetg.so (include/ap_flow.h).

`FlowUnet` has the constructor arguments, the submodule tree and therefore the state_dict keys of the reference class
(Module2/intrinsic_flow_models/networks.py:577-627 with `FlowUnetSkipConnectionBlock`, :510-575), so the checkpoint
`checkpoints/FlowReg_id_flow_faces/<epoch>_net_F.pth` loads with `load_state_dict` the day it is available; the children
only HOLD parameters, `forward` hands the key-point maps to the C ABI.  `flow_network_warp` is the caller's per-frame use
(Module2/models/geomcgt_ifw_test_model.py:62-76) as ONE library call from the landmarks (`ap_flow_warp_landmarks`): the
key-point discs are drawn on the GPU straight into the network's operand and the arg-max / mask / rescale / resize tail is
fused behind the network.  No CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _capi, conditioning


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder of the B200 flow network; call FlowUnet, not its children")


def _norm_layer(norm: str):
    if norm == "batch":
        return nn.BatchNorm2d
    if norm == "instance":
        return nn.InstanceNorm2d
    raise NotImplementedError()


class FlowUnetSkipConnectionBlock(_Holder):
    """Key layout of networks.py:510-558: down.{0|1} conv, down.{1|2} norm, up.1 transposed conv, up.2 norm, predict_flow.1."""

    def __init__(self, outer_nc, inner_nc, submodule=None, outermost=False, innermost=False, norm="batch"):
        super().__init__()
        nl = _norm_layer(norm)
        use_bias = norm == "instance"
        downconv = nn.Conv2d(outer_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
        if outermost:
            upconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1)
            down = [downconv, nl(inner_nc)]
        elif innermost:
            upconv = nn.ConvTranspose2d(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            down = [nn.LeakyReLU(0.2, True), downconv]
        else:
            upconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            down = [nn.LeakyReLU(0.2, True), downconv, nl(inner_nc)]
        self.down = nn.Sequential(*down)
        self.up = nn.Sequential(nn.ReLU(True), upconv, nl(outer_nc))
        self.submodule = submodule
        self.predict_flow = nn.Sequential(nn.LeakyReLU(0.1), nn.Conv2d(outer_nc, 2, kernel_size=3, stride=1, padding=1))


class FlowUnet(nn.Module):
    """Drop-in for networks.FlowUnet(input_nc, nf, start_scale, num_scale, norm, gpu_ids, max_nf) -- inference only.
    `size` (not in the reference) is the height = width of the key-point maps the caller feeds (224)."""

    def __init__(self, input_nc, nf=16, start_scale=2, num_scale=5, norm="batch", gpu_ids=(), max_nf=512, size=224):
        super().__init__()
        if norm not in ("batch", "instance"):
            raise NotImplementedError()
        self.input_nc, self.nf, self.start_scale, self.num_scale = int(input_nc), int(nf), int(start_scale), int(num_scale)
        self.norm, self.max_nf, self.size, self.gpu_ids = norm, int(max_nf), int(size), list(gpu_ids)
        nl = _norm_layer(norm)
        use_bias = norm == "instance"
        layers = [nn.Conv2d(input_nc, nf, kernel_size=7, padding=3, bias=use_bias), nl(nf), nn.LeakyReLU(0.1)]
        nc = nf
        for _ in range(int(math.log2(start_scale))):
            layers += [nn.Conv2d(nc, 2 * nc, kernel_size=3, stride=2, padding=1, bias=use_bias), nl(2 * nc), nn.LeakyReLU(0.1)]
            nc *= 2
        self.conv_downsample = nn.Sequential(*layers)
        block = None
        for l in reversed(range(num_scale)):
            block = FlowUnetSkipConnectionBlock(min(max_nf, nc * 2 ** l), min(max_nf, nc * 2 ** (l + 1)), submodule=block,
                                                innermost=(l == num_scale - 1), outermost=(l == 0), norm=norm)
        self.unet_block = block
        self.nf_out = min(max_nf, nc)
        self.predict_vis = nn.Sequential(nn.LeakyReLU(0.1), nn.Conv2d(min(max_nf, nc), 3, kernel_size=3, stride=1, padding=1))
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[int] = None
        self._finalizer = None
        self._seen = None

    @property
    def out_size(self) -> int:
        """Height = width of flow_out / vis_out: the reference up-samples by a hard-coded 2 (networks.py:583,641)."""
        return 2 * self.size // self.start_scale

    # ---- native handle -----------------------------------------------------------------------------------
    @staticmethod
    def _destroy(handle):
        try:
            _capi.lib().ap_flow_destroy(handle)
        except Exception:
            pass

    def _tensors(self):
        return [(k, v) for k, v in self.state_dict().items() if v.dtype.is_floating_point]

    def _sync(self, device: torch.device):
        lib = _capi.lib()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle_device != idx:
            if self._finalizer is not None:
                self._finalizer()
            h = C.c_void_p()
            _capi.check(lib.ap_flow_create(C.byref(h), self.input_nc, self.nf, self.start_scale, self.num_scale,
                                           0 if self.norm == "batch" else 1, self.max_nf, self.size, idx), "ap_flow_create")
            self._handle, self._handle_device = h, idx
            self._finalizer = weakref.finalize(self, FlowUnet._destroy, h)
            self._seen = None
        ts = self._tensors()
        sig = tuple((t.data_ptr(), t._version) for _, t in ts)
        if sig != self._seen:
            keep = [t.detach().to(device=device, dtype=torch.float32).contiguous() for _, t in ts]
            names = (C.c_char_p * len(ts))(*[k.encode() for k, _ in ts])
            ptrs = (C.c_void_p * len(ts))(*[t.data_ptr() for t in keep])
            shapes = (C.c_int64 * (4 * len(ts)))()
            for i, t in enumerate(keep):
                shp = list(t.shape) + [1] * (4 - t.dim())
                for j in range(4):
                    shapes[4 * i + j] = shp[j]
            stream = torch.cuda.current_stream(device).cuda_stream
            _capi.check(lib.ap_flow_load_weights(self._handle, len(ts), names, ptrs, shapes, 1, C.c_void_p(stream)),
                        "ap_flow_load_weights")
            self._seen = sig

    def _run(self, input, want_raw: bool, want_warp: bool):
        if not input.is_cuda:
            raise RuntimeError("the B200 flow network runs on CUDA tensors only (no CPU fallback)")
        if self.training and self.norm == "batch":
            raise RuntimeError("the B200 flow network is inference-only: call .eval() (the reference does, "
                               "geomcgt_ifw_test_model.py:216) -- BatchNorm uses its running statistics")
        B = input.shape[0]
        if tuple(input.shape) != (B, self.input_nc, self.size, self.size):
            raise RuntimeError(f"input: expected {(B, self.input_nc, self.size, self.size)}, got {tuple(input.shape)}")
        dev = input.device
        x = input.detach().to(dtype=torch.float32).contiguous()
        R = self.out_size
        flow = vis = iw = ifm = None
        with torch.cuda.device(dev):
            self._sync(dev)
            if want_raw:
                flow = torch.empty((B, 2, R, R), device=dev)
                vis = torch.empty((B, 3, R, R), device=dev)
            if want_warp:
                iw = torch.empty((B, 2, 256, 256), device=dev)
                ifm = torch.empty((B, 1, 256, 256), device=dev)
            p = lambda t: C.c_void_p(t.data_ptr() if t is not None else None)  # noqa: E731
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(_capi.lib().ap_flow_forward(self._handle, B, p(x), p(flow), p(vis), p(iw), p(ifm), C.c_void_p(stream)),
                        "ap_flow_forward")
        return flow, vis, iw, ifm

    @torch.no_grad()
    def forward(self, input, single_device=False):
        """networks.py:629-644: (flow_out, vis, flow_pyr, feat_out); the pyramid and the features, which the inference
        caller drops (`flow_out, vis_out, _, _ = netF(input_F)`), are returned as None."""
        flow, vis, _, _ = self._run(input, True, False)
        return flow, vis, None, None

    @torch.no_grad()
    def warp_tensors(self, input) -> Tuple[torch.Tensor, torch.Tensor]:
        """(iw_flow [B,2,256,256], real_A_if_mask [B,1,256,256]) of flow_network_warp, fused behind the network."""
        _, _, iw, ifm = self._run(input, False, True)
        return iw, ifm

    @torch.no_grad()
    def warp_landmarks(self, lm1: torch.Tensor, lm2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """flow_network_warp from landmark coordinates (`ap_flow_warp_landmarks`): lm2 [B,K,2], lm1 [B,K,2] or ONE set
        [K,2] shared by the batch; 256-pixel coordinates, scaled by 7/8 and drawn on the device."""
        if not (lm1.is_cuda and lm2.is_cuda):
            raise RuntimeError("the B200 flow network runs on CUDA tensors only (no CPU fallback)")
        if self.training and self.norm == "batch":
            raise RuntimeError("the B200 flow network is inference-only: call .eval()")
        K = self.input_nc // 2
        dev = lm2.device
        b = lm2.detach().to(dtype=torch.float32).contiguous()
        a = lm1.detach().to(device=dev, dtype=torch.float32).contiguous()
        B = b.shape[0]
        if tuple(b.shape) != (B, K, 2) or tuple(a.shape) not in ((B, K, 2), (K, 2)):
            raise RuntimeError(f"landmarks: expected lm2 [B,{K},2] and lm1 [B,{K},2] or [{K},2], got {tuple(lm2.shape)}, {tuple(lm1.shape)}")
        with torch.cuda.device(dev):
            self._sync(dev)
            iw = torch.empty((B, 2, 256, 256), device=dev)
            ifm = torch.empty((B, 1, 256, 256), device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(_capi.lib().ap_flow_warp_landmarks(self._handle, B, C.c_void_p(a.data_ptr()), int(a.dim() == 3),
                                                           C.c_void_p(b.data_ptr()), C.c_void_p(iw.data_ptr()),
                                                           C.c_void_p(ifm.data_ptr()), C.c_void_p(stream)),
                        "ap_flow_warp_landmarks")
        return iw, ifm

    def last_launch_count(self) -> int:
        n = C.c_int64(0)
        _capi.check(_capi.lib().ap_flow_last_launch_count(self._handle, C.byref(n)), "ap_flow_last_launch_count")
        return int(n.value)


@torch.no_grad()
def flow_network_warp(netF: FlowUnet, real_A, lm1: torch.Tensor, lm2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """geomcgt_ifw_test_model.py:62-76.  lm1 / lm2 [B,68,2] source / target landmarks in 256x256 pixel coordinates
    (device tensors; `real_A` is only resized and dropped by the reference and is ignored here).  The key-point maps are
    drawn on the GPU at 7/8 scale (224x224) straight into the network's operand (`ap_flow_warp_landmarks`)."""
    return netF.warp_landmarks(lm1, lm2)
