"""Rendering the frames of a clip, on one GPU or sharded data-parallel over the GPUs of one box.

The reference renders one frame per iteration with batch size 1 (Module2/test.py:42,58-65).  Frames are
independent given the photo, so a clip of T frames splits into contiguous shards, rank r
owning [r*T/G, (r+1)*T/G).  The only communication is a scatter of the per-frame conditioning tensors
from rank 0 and a gather of the finished frames back to rank 0; there is no collective inside the generator.

`render_frames_sharded` pipelines the three steps chunk by chunk (chunks of `batch` frames):

  scatter   NCCL grouped send/recv (torch.distributed point-to-point ops; gloo in the CPU tests), one group per
            round = the k-th chunk of every rank, all rounds queued up front on a communication stream: round k+1
            travels over NVLink while chunk k renders, only round 0 is exposed.  Rank 0 may hold the clip in pinned
            HOST memory: its upload runs round by round on a copy stream in front of the scatter.
  render    the generator on this rank's chunk.
  gather    "peer": every rank maps ONE frame buffer that lives in rank 0's HBM (CUDA IPC over NVLink peer access) and
            the generator's output stage stores its frames straight into it -- the epilogue of the last kernel IS the
            gather, no copy and no collective; one 4-byte all-reduce per call tells rank 0 that the stores have landed.
            "nccl": chunk k is sent back with NCCL behind the render of chunk k+1.
"""
from __future__ import annotations

import contextlib
import sys
from typing import Callable, Dict, List, Optional, Sequence, Tuple

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


def _call_into(netG: Callable, chunk: Sequence[torch.Tensor], dst: torch.Tensor) -> None:
    """dst <- netG(*chunk); generators that can write their frames in place (`out=`) do so (no copy)."""
    target = netG.module if isinstance(netG, torch.nn.DataParallel) else netG
    if getattr(target, "supports_out", False):
        target(*chunk, out=dst)
    else:
        dst.copy_(netG(*chunk))


def render_frames(netG: Callable, inputs: Sequence[torch.Tensor], batch: int = 16,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Run the generator over T frames in batches; `inputs` are the six [T, ...] tensors on the GPU."""
    T = inputs[0].shape[0]
    if T == 0:
        raise ValueError("no frames to render")
    with torch.no_grad():
        first = netG(*[t[:batch] for t in inputs])
        if out is None:
            out = torch.empty((T,) + tuple(first.shape[1:]), dtype=first.dtype, device=first.device)
        out[:first.shape[0]].copy_(first)
        for s in range(batch, T, batch):
            _call_into(netG, [t[s:s + batch] for t in inputs], out[s:s + batch])
    return out


# ----------------------------------------------------------------------------------------------------------
# point-to-point plumbing
# ----------------------------------------------------------------------------------------------------------
def _peer(group, r: int) -> int:
    return dist.get_global_rank(group, r) if group is not None else r


def _scatter(t_full: Optional[torch.Tensor], shape_tail, T: int, device, group, src: int = 0,
             dtype=torch.float32) -> torch.Tensor:
    """One-shot scatter of a [T, ...] tensor from rank `src` in ragged contiguous shards (small tensors: landmarks)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(T, world, rank)
    mine = torch.empty((hi - lo,) + tuple(shape_tail), dtype=dtype, device=device)
    # ragged chunks: point-to-point sends (scatter needs equal sizes); NCCL batches them into one group
    if rank == src:
        ops = []
        for r in range(world):
            a, b = shard_range(T, world, r)
            if r == src:
                mine.copy_(t_full[a:b])
            elif b > a:
                ops.append(dist.P2POp(dist.isend, t_full[a:b].contiguous(), _peer(group, r), group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    elif hi > lo:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.irecv, mine, _peer(group, src), group)]):
            req.wait()
    return mine


def _gather(mine: torch.Tensor, T: int, group, dst: int = 0, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """One-shot gather of ragged contiguous shards into a [T, ...] tensor on rank `dst`."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank == dst:
        full = out if out is not None else torch.empty((T,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        ops = []
        for r in range(world):
            a, b = shard_range(T, world, r)
            if r == dst:
                full[a:b].copy_(mine)
            elif b > a:
                ops.append(dist.P2POp(dist.irecv, full[a:b], _peer(group, r), group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return full
    if mine.shape[0] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine.contiguous(), _peer(group, dst), group)]):
            req.wait()
    return None


# ----------------------------------------------------------------------------------------------------------
# one buffer in rank 0's HBM, mapped by every rank of the box (CUDA IPC + NVLink peer access)
# ----------------------------------------------------------------------------------------------------------
class _RawCudaArray:
    """A raw device pointer dressed as a __cuda_array_interface__ object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


_TYPESTR = {torch.float32: "<f4", torch.uint8: "|u1"}


class PeerBuffer:
    """A buffer that lives on rank 0's GPU and is mapped into every rank's address space (CUDA IPC; NVLink peer access
    enabled by the mapping): kernels of any rank store into it directly.  `tensor` is the mapping of this rank (on rank 0:
    the buffer itself).  Collective: every rank of `group` constructs it together.  `ok` is False on every rank when any
    rank could not map it (include/ap_netg.h: ap_peer_alloc / ap_peer_open)."""

    def __init__(self, shape, dtype, device: torch.device, group=None):
        import ctypes as C
        from . import _capi
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.tensor: Optional[torch.Tensor] = None
        self._ptr = C.c_void_p()
        self._owner = self.rank == 0
        nbytes = int(torch.tensor([], dtype=dtype).element_size())
        for d in shape:
            nbytes *= int(d)
        err = ""
        payload = [None]
        lib = _capi.lib()
        try:
            if self._owner:
                handle = C.create_string_buffer(64)
                _capi.check(lib.ap_peer_alloc(device.index, max(nbytes, 256), C.byref(self._ptr), handle), "ap_peer_alloc")
                payload = [handle.raw]
        except Exception as e:  # pragma: no cover - environment dependent
            err = f"export: {e!r}"
        dist.broadcast_object_list(payload, src=_peer(group, 0), group=group)
        try:
            if not self._owner:
                if payload[0] is None:
                    raise RuntimeError("rank 0 exported nothing")
                _capi.check(lib.ap_peer_open(device.index, payload[0], C.byref(self._ptr)), "ap_peer_open")
            if self._ptr.value:
                with torch.cuda.device(device):
                    self.tensor = torch.as_tensor(_RawCudaArray(self._ptr.value, shape, _TYPESTR[dtype]))
        except Exception as e:  # pragma: no cover - environment dependent
            err = err or f"map: {e!r}"
            self.tensor = None
        flag = torch.tensor([0 if self.tensor is not None else 1], device=device, dtype=torch.int32)
        dist.all_reduce(flag, group=group)
        self.ok = int(flag.item()) == 0
        if not self.ok:
            if err:
                print(f"[frames] rank {self.rank}: peer buffer unavailable ({err})", file=sys.stderr)
            self.tensor = None

    def close(self) -> None:
        from . import _capi
        if self._ptr.value:
            self.tensor = None
            (_capi.lib().ap_peer_free if self._owner else _capi.lib().ap_peer_close)(self._ptr)
            self._ptr.value = None


_PEER_CACHE: Dict[tuple, PeerBuffer] = {}


def peer_frame_buffer(T: int, tail: Tuple[int, ...], dtype, device: torch.device, group=None) -> PeerBuffer:
    """The gather buffer of a clip shape, created once per (group, shape, dtype) and reused by every call."""
    key = (id(group), T, tuple(tail), dtype, device.index)
    if key not in _PEER_CACHE:
        _PEER_CACHE[key] = PeerBuffer((T,) + tuple(tail), dtype, device, group)
    return _PEER_CACHE[key]


_STREAMS: Dict[tuple, "torch.cuda.Stream"] = {}


def _stream(device: torch.device, name: str):
    if device.type != "cuda":
        return None
    key = (device.index, name)
    if key not in _STREAMS:
        _STREAMS[key] = torch.cuda.Stream(device)
    return _STREAMS[key]


def _on(stream):
    return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()


def render_frames_sharded(netG: Callable, inputs_rank0: Optional[Sequence[torch.Tensor]], T: int, out_channels: int,
                          device, batch: int = 16, group=None, gather: str = "auto",
                          out_host: Optional[torch.Tensor] = None, info: Optional[dict] = None) -> Optional[torch.Tensor]:
    """Scatter the six [T, ...] conditioning tensors from rank 0 chunk by chunk, render this rank's shard under the
    scatter of its next chunk, collect the frames on rank 0 (returns the [T, out_channels, 256, 256] tensor there, None
    elsewhere).  `inputs_rank0` may be device tensors or (pinned) host tensors; `out_host` (rank 0, pinned) also
    receives the frames.  gather: "peer" | "nccl" | "auto" (peer when the buffer can be mapped, else nccl)."""
    device = torch.device(device)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    cuda = device.type == "cuda"
    if gather not in ("auto", "peer", "nccl"):
        raise ValueError(f"gather={gather!r}")
    if rank == 0:
        for i, name in enumerate(INPUT_NAMES):
            if tuple(inputs_rank0[i].shape) != (T,) + FRAME_SHAPES[name]:
                raise ValueError(f"{name}: expected {(T,) + FRAME_SHAPES[name]}, got {tuple(inputs_rank0[i].shape)}")
    spans = [shard_range(T, world, r) for r in range(world)]
    lo, hi = spans[rank]
    rounds = max(-(-(b - a) // batch) for a, b in spans) if T > 0 else 0

    def chunk_of(r: int, k: int) -> Tuple[int, int]:
        a, b = spans[r]
        s = a + k * batch
        return s, max(s, min(s + batch, b))

    # ---- where the frames go ----
    peer = None
    if gather != "nccl" and cuda and world > 1:
        peer = peer_frame_buffer(T, (out_channels, 256, 256), torch.float32, device, group)
        if not peer.ok:
            if gather == "peer":
                raise RuntimeError("gather='peer': the frame buffer of rank 0 could not be mapped by every rank")
            peer = None
    if info is not None:
        info["gather"] = "peer-store" if peer is not None else "nccl"
    if peer is not None:
        full = peer.tensor                      # rank 0's buffer, on every rank
        mine = full[lo:hi]
    else:
        full = torch.empty((T, out_channels, 256, 256), dtype=torch.float32, device=device) if rank == 0 else None
        mine = full[lo:hi] if rank == 0 else torch.empty((hi - lo, out_channels, 256, 256), dtype=torch.float32, device=device)

    # ---- rank 0: the clip, on the device; host clips are uploaded round by round on a copy stream ----
    comm, h2d = _stream(device, "comm"), _stream(device, "h2d")
    cur = torch.cuda.current_stream(device) if cuda else None
    host_src = rank == 0 and cuda and not inputs_rank0[0].is_cuda
    uploaded: List[Optional["torch.cuda.Event"]] = [None] * rounds
    if rank == 0:
        if host_src:
            dev_in = [torch.empty(t.shape, dtype=torch.float32, device=device) for t in inputs_rank0]
            h2d.wait_stream(cur)
            with _on(h2d):
                for k in range(rounds):
                    for r in range(world):
                        s, e = chunk_of(r, k)
                        if e > s:
                            for i in range(6):
                                dev_in[i][s:e].copy_(inputs_rank0[i][s:e], non_blocking=True)
                    uploaded[k] = torch.cuda.Event()
                    uploaded[k].record(h2d)
        else:
            dev_in = list(inputs_rank0)
        local = [t[lo:hi] for t in dev_in]
    else:
        local = [torch.empty((hi - lo,) + FRAME_SHAPES[n], dtype=torch.float32, device=device) for n in INPUT_NAMES]

    # ---- scatter: every round queued up front, in round order, on the communication stream ----
    works: List[list] = []
    if comm is not None:
        comm.wait_stream(cur)
    with _on(comm):
        for k in range(rounds):
            ops = []
            if rank == 0:
                if uploaded[k] is not None:
                    comm.wait_event(uploaded[k])
                for r in range(1, world):
                    s, e = chunk_of(r, k)
                    if e > s:
                        ops += [dist.P2POp(dist.isend, dev_in[i][s:e], _peer(group, r), group) for i in range(6)]
            else:
                s, e = chunk_of(rank, k)
                if e > s:
                    ops += [dist.P2POp(dist.irecv, local[i][s - lo:e - lo], _peer(group, 0), group) for i in range(6)]
            works.append(dist.batch_isend_irecv(ops) if ops else [])

    # ---- render chunk k as soon as round k has landed; return its frames behind the render of chunk k+1 ----
    gather_works = []
    with torch.no_grad():
        for k in range(rounds):
            s, e = chunk_of(rank, k)
            if e > s:
                if rank == 0 and uploaded[k] is not None:
                    cur.wait_event(uploaded[k])
                for w in works[k]:
                    w.wait()
                _call_into(netG, [t[s - lo:e - lo] for t in local], mine[s - lo:e - lo])
            if peer is None and world > 1:
                ops = []
                if rank == 0:
                    for r in range(1, world):
                        a, b = chunk_of(r, k)
                        if b > a:
                            ops.append(dist.P2POp(dist.irecv, full[a:b], _peer(group, r), group))
                elif e > s:
                    ops.append(dist.P2POp(dist.isend, mine[s - lo:e - lo], _peer(group, 0), group))
                if ops:
                    gather_works += dist.batch_isend_irecv(ops)
    for w in gather_works:
        w.wait()
    if peer is not None:
        # the frames were stored into rank 0's buffer by the output kernels of every rank: one tiny all-reduce orders
        # rank 0's stream after the end of everybody's last kernel
        token = torch.zeros(1, device=device)
        dist.all_reduce(token, group=group)
    if rank == 0 and out_host is not None:
        out_host.copy_(full, non_blocking=True)
    return full if rank == 0 else None
