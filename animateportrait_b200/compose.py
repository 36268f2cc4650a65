"""Output stage after netG: foreground/background blend + uint8 image conversion, one fused CUDA kernel.

Mirrors, for CUDA tensors, what the reference does on the host side of every frame:
  * `GeomCGTIFWTestModel.forward` (Module2/models/geomcgt_ifw_test_model.py:297-300):
        mask1  = F.grid_sample(mask, warp_motion, align_corners=True)
        fake_B = ((fake_B/2+0.5)*mask1 + (fakeB_static/2+0.5)*(1-mask1))*2-1
  * `util.tensor2im` (Module2/util/util.py:9-29): ((x+1)/2*255).astype(uint8), HWC, grayscale tiled to RGB.
No PyTorch fallback: the call goes to `ap_netg_compose` (include/ap_netg.h) or raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _capi


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


@torch.no_grad()
def blend_and_convert(fake_B: torch.Tensor, mask: Optional[torch.Tensor] = None, warp_motion: Optional[torch.Tensor] = None,
                      fakeB_static: Optional[torch.Tensor] = None, want_blended: bool = True,
                      want_image: bool = True) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Returns (blended fp32 [B,onc,256,256] or None, uint8 image [B,256,256,3] or None) on fake_B's device.
    With mask/warp_motion/fakeB_static omitted only the image conversion runs (`tensor2im` semantics)."""
    if not fake_B.is_cuda:
        raise RuntimeError("blend_and_convert runs on CUDA tensors only (no CPU fallback)")
    B, onc, H, W = fake_B.shape
    if (H, W) != (256, 256) or onc not in (1, 3):
        raise RuntimeError(f"fake_B: expected [B,1|3,256,256], got {tuple(fake_B.shape)}")
    given = [t is not None for t in (mask, warp_motion, fakeB_static)]
    if any(given) and not all(given):
        raise RuntimeError("mask, warp_motion and fakeB_static come together (all or none)")
    dev = fake_B.device
    f = fake_B.detach().float().contiguous()
    m = g = s = None
    if all(given):
        want = {"mask": (mask, (B, 1, 256, 256)), "warp_motion": (warp_motion, (B, 256, 256, 2)),
                "fakeB_static": (fakeB_static, (B, onc, 256, 256))}
        for name, (t, shape) in want.items():
            if tuple(t.shape) != shape:
                raise RuntimeError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
        m, g, s = (t.detach().to(device=dev, dtype=torch.float32).contiguous() for t in (mask, warp_motion, fakeB_static))
    if not (want_blended or want_image):
        raise RuntimeError("nothing requested")
    blended = torch.empty_like(f) if want_blended else None
    image = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device=dev) if want_image else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        _capi.check(_capi.lib().ap_netg_compose(idx, B, onc, _ptr(f), _ptr(m), _ptr(g), _ptr(s), _ptr(blended), _ptr(image),
                                                C.c_void_p(stream)), "ap_netg_compose")
    return blended, image
