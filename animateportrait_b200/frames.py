"""Rendering the frames of a clip, on one GPU or sharded data-parallel over the GPUs of one box.

The reference renders one frame per iteration with batch size 1 (Module2/test.py:42,58-65).  Frames are
independent given the photo, so a clip of T frames splits into contiguous chunks, rank r
owning [r*T/G, (r+1)*T/G).  The only communication is a scatter of the per-frame conditioning tensors
from rank 0 and a gather of the finished frames back to rank 0 (torch.distributed: NCCL over NVLink on
GPUs, gloo in the CPU tests); there is no collective inside the generator.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

INPUT_NAMES = ("input", "land1", "land2", "motion", "flow", "ifmask")
FRAME_SHAPES = {"input": (3, 256, 256), "land1": (1, 256, 256), "land2": (1, 256, 256), "motion": (256, 256, 2),
                "flow": (2, 256, 256), "ifmask": (1, 256, 256)}


def shard_range(T: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous chunk of frames owned by `rank` (balanced to within one frame)."""
    if T < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad shard request T={T} world={world} rank={rank}")
    return (rank * T) // world, ((rank + 1) * T) // world


def render_frames(netG: Callable, inputs: Sequence[torch.Tensor], batch: int = 16) -> torch.Tensor:
    """Run the generator over T frames in batches; `inputs` are the six [T, ...] tensors on the GPU."""
    T = inputs[0].shape[0]
    outs: List[torch.Tensor] = []
    with torch.no_grad():
        for s in range(0, T, batch):
            outs.append(netG(*[t[s:s + batch] for t in inputs]))
    if not outs:
        raise ValueError("no frames to render")
    return torch.cat(outs, 0)


def _scatter(t_full: Optional[torch.Tensor], shape_tail, T: int, device, group, src: int = 0) -> torch.Tensor:
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(T, world, rank)
    mine = torch.empty((hi - lo,) + tuple(shape_tail), dtype=torch.float32, device=device)
    # ragged chunks: point-to-point sends (scatter needs equal sizes); NCCL batches them into one group
    if rank == src:
        ops = []
        for r in range(world):
            a, b = shard_range(T, world, r)
            if r == src:
                mine.copy_(t_full[a:b])
            elif b > a:
                ops.append(dist.P2POp(dist.isend, t_full[a:b].contiguous(), dist.get_global_rank(group, r) if group else r,
                                      group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    elif hi > lo:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.irecv, mine, dist.get_global_rank(group, src) if group else src,
                                                       group)]):
            req.wait()
    return mine


def _gather(mine: torch.Tensor, T: int, group, dst: int = 0) -> Optional[torch.Tensor]:
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank == dst:
        full = torch.empty((T,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        ops = []
        for r in range(world):
            a, b = shard_range(T, world, r)
            if r == dst:
                full[a:b].copy_(mine)
            elif b > a:
                ops.append(dist.P2POp(dist.irecv, full[a:b], dist.get_global_rank(group, r) if group else r, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return full
    if mine.shape[0] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine.contiguous(),
                                                       dist.get_global_rank(group, dst) if group else dst, group)]):
            req.wait()
    return None


def render_frames_sharded(netG: Callable, inputs_rank0: Optional[Sequence[torch.Tensor]], T: int, out_channels: int,
                          device, batch: int = 16, group=None) -> Optional[torch.Tensor]:
    """Scatter the six [T, ...] conditioning tensors from rank 0, render this rank's chunk, gather the
    frames on rank 0 (returns the [T, out_channels, 256, 256] tensor there, None elsewhere)."""
    rank = dist.get_rank(group)
    local = []
    for i, name in enumerate(INPUT_NAMES):
        src = inputs_rank0[i] if rank == 0 else None
        if rank == 0 and tuple(src.shape) != (T,) + FRAME_SHAPES[name]:
            raise ValueError(f"{name}: expected {(T,) + FRAME_SHAPES[name]}, got {tuple(src.shape)}")
        local.append(_scatter(src, FRAME_SHAPES[name], T, device, group))
    if local[0].shape[0] > 0:
        mine = render_frames(netG, local, batch)
    else:
        mine = torch.empty((0, out_channels, 256, 256), dtype=torch.float32, device=device)
    return _gather(mine, T, group)
