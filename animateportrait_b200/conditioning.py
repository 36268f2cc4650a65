"""Per-frame conditioning producers on the GPU: the tensors netG is called with, made from the
68x2 landmark coordinates of a whole batch of frames in one launch each.

Same names and argument meaning as the reference's CPU functions, but batched over a leading frame axis and on CUDA
tensors (there is no CPU or PyTorch fallback: every call goes to include/ap_cond.h or raises):

  draw2(height, width, lands, radius, thickness, op=0)   Module2/data/umlvdfw_test_dataset.py:34-41
  cal_motion256(lm2d0, lm2d)                             Module2/data/umlvdfw_test_dataset.py:67-81
  kp_to_map_some(img_sz, kps, mode='binary', radius=4)   Module2/models/geomcgt_ifw_test_model.py:12-44
  matte_photo(real_A, matte)                             Module2/models/geomcgt_ifw_test_model.py:280,292
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _capi

N_LANDMARKS = 68
MAX_TRIANGLES = 512


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _cuda_f32(name: str, t: torch.Tensor) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: CUDA tensor expected (the conditioning producers have no CPU fallback)")
    return t.detach().to(torch.float32).contiguous()


def _launch(dev: torch.device):
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return idx, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@torch.no_grad()
def draw2(height: int, width: int, lands: torch.Tensor, radius: int, thickness: int = 2, c=None, op: int = 0) -> torch.Tensor:
    """Landmark disc maps.  lands [T,N,2] or [N,2] (x,y) -> [T,1,height,width] (or [1,height,width] as the reference
    returns for one frame) float32 in {-1,+1}.  Only op=0 (the shipped configuration, `--draw_op 0`) is built."""
    if op != 0:
        raise NotImplementedError(f"draw2 op={op} is not implemented (only op=0, discs)")
    if c is not None and c != 255:
        raise NotImplementedError("draw2: only the default colour 255 is implemented")
    if height != width:
        raise NotImplementedError("draw2: square canvases only")
    single = lands.dim() == 2
    l = _cuda_f32("lands", lands[None] if single else lands)
    if l.dim() != 3 or l.shape[2] != 2:
        raise RuntimeError(f"lands: expected [T,N,2], got {tuple(lands.shape)}")
    T, n = l.shape[:2]
    out = torch.empty((T, 1, height, width), dtype=torch.float32, device=l.device)
    with torch.cuda.device(l.device):
        idx, st = _launch(l.device)
        _capi.check(_capi.lib().ap_cond_draw_landmarks(idx, T, n, height, int(radius), _ptr(l), _ptr(out), st),
                    "ap_cond_draw_landmarks")
    return out[0] if single else out


@torch.no_grad()
def cal_motion256(lm2d0: torch.Tensor, lm2d: torch.Tensor, return_triangle_count: bool = False):
    """Motion grid(s) from source landmarks lm2d0 ([68,2] shared by all frames, or [T,68,2]) to target landmarks lm2d
    ([T,68,2] or [68,2]) -> [T,256,256,2] (or [256,256,2]) float32, the `warp_motion` tensor of netG.forward."""
    single = lm2d.dim() == 2
    dst = _cuda_f32("lm2d", lm2d[None] if single else lm2d)
    src = _cuda_f32("lm2d0", lm2d0[None] if lm2d0.dim() == 2 else lm2d0)
    T = dst.shape[0]
    if tuple(dst.shape[1:]) != (N_LANDMARKS, 2) or tuple(src.shape[1:]) != (N_LANDMARKS, 2) or src.shape[0] not in (1, T):
        raise RuntimeError(f"cal_motion256: expected lm2d0 [68,2]|[T,68,2] and lm2d [T,68,2], got {tuple(lm2d0.shape)} "
                           f"and {tuple(lm2d.shape)}")
    if src.device != dst.device:
        raise RuntimeError("cal_motion256: lm2d0 and lm2d live on different devices")
    lib = _capi.lib()
    nbytes = C.c_size_t()
    _capi.check(lib.ap_cond_motion256_workspace_bytes(T, C.byref(nbytes)), "ap_cond_motion256_workspace_bytes")
    ws = torch.empty((nbytes.value + 15) // 16 * 16, dtype=torch.uint8, device=dst.device)
    motion = torch.empty((T, 256, 256, 2), dtype=torch.float32, device=dst.device)
    count = torch.empty((T,), dtype=torch.int32, device=dst.device) if return_triangle_count else None
    with torch.cuda.device(dst.device):
        idx, st = _launch(dst.device)
        _capi.check(lib.ap_cond_motion256(idx, T, _ptr(src), 1 if (src.shape[0] == T and T > 1) else 0, _ptr(dst),
                                          _ptr(motion), _ptr(ws), ws.numel(), _ptr(count), st), "ap_cond_motion256")
    ws.record_stream(torch.cuda.current_stream(dst.device))
    motion = motion[0] if single else motion
    return (motion, count) if return_triangle_count else motion


@torch.no_grad()
def kp_to_map_some(img_sz: Tuple[int, int], kps: torch.Tensor, mode: str = "binary", radius: float = 4) -> torch.Tensor:
    """Key-point maps for the flow network.  kps [T,K,2] (x,y) -> [T,K,h,w] float32 in {0,1}; x == -1 or y == -1 marks
    a missing point (empty map).  Only mode='binary' (what the reference's call sites use) is built."""
    if mode != "binary":
        raise NotImplementedError(f"kp_to_map mode={mode!r} is not implemented (only 'binary')")
    w, h = img_sz
    if w != h:
        raise NotImplementedError("kp_to_map: square maps only")
    k = _cuda_f32("kps", kps)
    if k.dim() != 3 or k.shape[2] != 2:
        raise RuntimeError(f"kps: expected [T,K,2], got {tuple(kps.shape)}")
    T, K = k.shape[:2]
    out = torch.empty((T, K, h, w), dtype=torch.float32, device=k.device)
    with torch.cuda.device(k.device):
        idx, st = _launch(k.device)
        _capi.check(_capi.lib().ap_cond_kp_to_map(idx, T, K, w, float(radius), _ptr(k), _ptr(out), st), "ap_cond_kp_to_map")
    return out


@torch.no_grad()
def matte_photo(real_A: torch.Tensor, matte: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(real_A matted onto white, mask) with mask = (matte > 0.5): real_A [B,C,H,W], matte [B,1,H,W]."""
    a, m = _cuda_f32("real_A", real_A), _cuda_f32("matte", matte)
    if a.dim() != 4 or m.dim() != 4 or m.shape[1] != 1 or m.shape[0] != a.shape[0] or m.shape[2:] != a.shape[2:]:
        raise RuntimeError(f"matte_photo: expected real_A [B,C,H,W] and matte [B,1,H,W], got {tuple(real_A.shape)} and "
                           f"{tuple(matte.shape)}")
    B, Cn, H, W = a.shape
    out, mask = torch.empty_like(a), torch.empty_like(m)
    with torch.cuda.device(a.device):
        idx, st = _launch(a.device)
        _capi.check(_capi.lib().ap_cond_matte_photo(idx, B, Cn, H * W, _ptr(a), _ptr(m), _ptr(out), _ptr(mask), st),
                    "ap_cond_matte_photo")
    return out, mask
