"""Rendering a whole clip (one photo, T target landmark sets) on the GPU: the reference's per-frame loop
(Module2/test.py:58-65 -> UMLVDFWTestDataset.__getitem__ -> GeomCGTIFWTestModel.set_input/forward -> save_images)
re-cut so that what does not depend on the frame is done once and what does is batched.

Per clip (hoisted):
  * the photo matting  real_A = ((real_A/2+.5)*mask + 1-mask)*2-1   (geomcgt_ifw_test_model.py:280,292)
  * the source landmark map  A_lm = draw2(A_lm_68)                  (umlvdfw_test_dataset.py:147)
  * inside netG, everything that depends on the photo and its landmark map alone: the three 7x7 stems, model_tri11,
    model_tri21, model_tri22, model_landmark_trans(A_lm) and their InstanceNorms (networks.py:1318-1331) run once per
    batch, not once per frame (`ap_netg_forward_shared_photo`)
  * the static drawing fakeB_static and the matte are INPUTS here: the reference recomputes them with MODNet and the
    512x512 static generator for every frame of the same photo (geomcgt_ifw_test_model.py:279-291); both networks need
    checkpoints that do not ship and are outside the path this repo rebuilds.
Per batch of frames (row f2 + the generator + rows f1/f4), all on the device, only 68x2 floats per frame come in:
  * tB_lm = draw2(tB_lm_68)                     (umlvdfw_test_dataset.py:148)
  * warp_motion = cal_motion256(A_lm_68, tB_lm_68)   (umlvdfw_test_dataset.py:161)
  * fake_B = netG(real_A, A_lm, tB_lm, warp_motion, iw_flow, real_A_if_mask)   (geomcgt_ifw_test_model.py:295)
  * blend with the static drawing + tensor2im -> uint8 HWC frames  (geomcgt_ifw_test_model.py:297-300, util/util.py:9-29)
iw_flow / real_A_if_mask come from the flow network netF (row f3): with `netF=` (flownet.FlowUnet) they are computed per
batch on the device from the same landmarks (flow_network_warp, geomcgt_ifw_test_model.py:62-76; the source key-point
maps once per photo); they can also be passed as tensors; with neither, the intrinsic-flow branch sees zero flow and a mask
of ones (every pixel visible).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import compose, conditioning
from .frames import _gather, _scatter, shard_range


class ClipRenderer:
    def __init__(self, netG, batch: int = 32, share_photo: bool = True, netF=None):
        """share_photo: run the photo-only part of the encoder once per batch (`forward_shared_photo`) instead of
        handing netG B copies of the photo (`forward`, the reference-shaped call); same frames either way.
        netF: the flow network (flownet.FlowUnet); render() then makes iw_flow / if_mask itself when they are not given."""
        self.netG = netG.module if isinstance(netG, torch.nn.DataParallel) else netG
        self.netF = netF
        self.batch = int(batch)
        self.share_photo = bool(share_photo)
        self._photo = None

    @torch.no_grad()
    def set_photo(self, real_A: torch.Tensor, A_lm_68: torch.Tensor, matte: Optional[torch.Tensor] = None,
                  fakeB_static: Optional[torch.Tensor] = None) -> None:
        """Frame-invariant work of a clip: real_A [1,3,256,256] in [-1,1], A_lm_68 [68,2]; optional matte [1,1,256,256]
        in [0,1] (MODNet's output) and static drawing fakeB_static [1,output_nc,256,256]."""
        if not real_A.is_cuda:
            raise RuntimeError("ClipRenderer works on CUDA tensors only (no CPU fallback)")
        dev = real_A.device
        if tuple(real_A.shape) != (1, 3, 256, 256) or tuple(A_lm_68.shape) != (68, 2):
            raise RuntimeError(f"set_photo: expected real_A [1,3,256,256] and A_lm_68 [68,2], got {tuple(real_A.shape)}, "
                               f"{tuple(A_lm_68.shape)}")
        if (matte is None) != (fakeB_static is None):
            raise RuntimeError("set_photo: matte and fakeB_static come together (the blend needs both) or not at all")
        lm = A_lm_68.to(dev, torch.float32).contiguous()
        mask = None
        if matte is not None:
            real_A, mask = conditioning.matte_photo(real_A, matte.to(dev))
        land1 = conditioning.draw2(256, 256, lm[None], 3)
        self._photo = {"real_A": real_A.float().contiguous(), "lm": lm, "land1": land1, "mask": mask,
                       "static": None if fakeB_static is None else fakeB_static.to(dev, torch.float32).contiguous(),
                       "expanded": {}}

    def _expanded(self, B: int):
        """The frame-invariant tensors repeated B times (netG's ABI takes one photo per frame), built once per B."""
        p = self._photo
        if B not in p["expanded"]:
            rep = lambda t: None if t is None else t.expand(B, *t.shape[1:]).contiguous()  # noqa: E731
            p["expanded"][B] = (rep(p["real_A"]), rep(p["land1"]), rep(p["mask"]), rep(p["static"]))
        return p["expanded"][B]

    @torch.no_grad()
    def render(self, tB_lm_68: torch.Tensor, iw_flow: Optional[torch.Tensor] = None,
               if_mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
               return_tensor: bool = False) -> torch.Tensor:
        """tB_lm_68 [T,68,2] (device, or host: copied) -> uint8 frames [T,256,256,3] on the device (`out` may supply
        the buffer), or with return_tensor the fp32 blended [T,output_nc,256,256] frames."""
        if self._photo is None:
            raise RuntimeError("render before set_photo")
        p = self._photo
        dev = p["real_A"].device
        lm = tB_lm_68.to(dev, torch.float32, non_blocking=True).contiguous()
        T = lm.shape[0]
        if tuple(lm.shape[1:]) != (68, 2):
            raise RuntimeError(f"tB_lm_68: expected [T,68,2], got {tuple(tB_lm_68.shape)}")
        for name, t, c in (("iw_flow", iw_flow, 2), ("if_mask", if_mask, 1)):
            if t is not None and (tuple(t.shape) != (T, c, 256, 256) or not t.is_cuda):
                raise RuntimeError(f"{name}: expected a CUDA tensor [T,{c},256,256], got {tuple(t.shape)}")
        onc = self.netG.output_nc
        if return_tensor:
            res = out if out is not None else torch.empty((T, onc, 256, 256), dtype=torch.float32, device=dev)
        else:
            res = out if out is not None else torch.empty((T, 256, 256, 3), dtype=torch.uint8, device=dev)
        zero_flow = ones_mask = None
        own_flow = iw_flow is None and if_mask is None and self.netF is not None
        for s in range(0, T, self.batch):
            e = min(s + self.batch, T)
            B = e - s
            photo, land1, mask, static = self._expanded(B)
            land2 = conditioning.draw2(256, 256, lm[s:e], 3)
            motion = conditioning.cal_motion256(p["lm"], lm[s:e])
            if own_flow:
                # flow_network_warp of the batch, one source landmark set for all its frames (ap_flow_warp_landmarks).  (Running
                # it on a second stream under the previous batch's generator was measured: no gain -- both are tensor-core
                # kernels that fill every SM's shared memory, they time-slice.)
                flow, ifm = self.netF.warp_landmarks(p["lm"], lm[s:e])
            else:
                if iw_flow is None:
                    if zero_flow is None or zero_flow.shape[0] != B:
                        zero_flow = torch.zeros((B, 2, 256, 256), device=dev)
                    flow = zero_flow
                else:
                    flow = iw_flow[s:e]
                if if_mask is None:
                    if ones_mask is None or ones_mask.shape[0] != B:
                        ones_mask = torch.ones((B, 1, 256, 256), device=dev)
                    ifm = ones_mask
                else:
                    ifm = if_mask[s:e]
            if self.share_photo:
                fake = self.netG.forward_shared_photo(p["real_A"], p["land1"], land2, motion, flow, ifm)
            else:
                fake = self.netG(photo, land1, land2, motion, flow, ifm)
            if mask is not None:
                blended, image = compose.blend_and_convert(fake, mask, motion, static, want_blended=return_tensor,
                                                           want_image=not return_tensor)
            elif return_tensor:
                blended, image = fake, None
            else:
                blended, image = compose.blend_and_convert(fake, want_blended=False)
            res[s:e].copy_(blended if return_tensor else image)
        return res


    @torch.no_grad()
    def render_to(self, sink, tB_lm_68: torch.Tensor, iw_flow: Optional[torch.Tensor] = None,
                  if_mask: Optional[torch.Tensor] = None, chunk: Optional[int] = None) -> int:
        """Render the clip chunk by chunk straight into a `sink.FrameSink` (row f4: the frames reach the encoder's stdin
        through a pinned ring while the next chunk renders; no PNG files, no frame directory).  Returns the frame count."""
        T = tB_lm_68.shape[0]
        step = int(chunk or self.batch)
        for s in range(0, T, step):
            e = min(s + step, T)
            sink.put(self.render(tB_lm_68[s:e], None if iw_flow is None else iw_flow[s:e],
                                 None if if_mask is None else if_mask[s:e]))
        return T


def render_clip_sharded(renderer: ClipRenderer, tB_lm_68_rank0: Optional[torch.Tensor], T: int, device,
                        iw_flow_rank0: Optional[torch.Tensor] = None, if_mask_rank0: Optional[torch.Tensor] = None,
                        group=None) -> Optional[torch.Tensor]:
    """Frames of one clip sharded over the ranks of `group` (every rank called set_photo with the same photo): rank 0
    scatters the target landmarks (68x2 floats per frame; plus flow/mask when given), every rank renders its contiguous
    chunk, rank 0 gathers the uint8 frames.  Returns [T,256,256,3] uint8 on rank 0, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lm = _scatter(tB_lm_68_rank0, (68, 2), T, device, group)
    has = torch.tensor([int(iw_flow_rank0 is not None), int(if_mask_rank0 is not None)], device=device)
    dist.broadcast(has, dist.get_global_rank(group, 0) if group else 0, group=group)
    flow = _scatter(iw_flow_rank0, (2, 256, 256), T, device, group) if int(has[0]) else None
    ifm = _scatter(if_mask_rank0, (1, 256, 256), T, device, group) if int(has[1]) else None
    lo, hi = shard_range(T, world, rank)
    if hi > lo:
        mine = renderer.render(lm, flow, ifm)
    else:
        mine = torch.empty((0, 256, 256, 3), dtype=torch.uint8, device=device)
    return _gather(mine, T, group)
