"""CPU (gloo, world_size 2 and 3): the frame-sharding plumbing of animateportrait_b200/frames.py.

The generator itself needs a B200; here `netG` is a cheap stand-in with the same signature so that the
scatter -> per-rank render -> gather path (the only multi-GPU logic of the hot path) is checked
for ragged chunk sizes, empty shards and frame order."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from animateportrait_b200.frames import FRAME_SHAPES, INPUT_NAMES, render_frames, render_frames_sharded, shard_range


def _stand_in_netg(input, land1, land2, motion, flow, ifmask):
    """Per-frame function of all six inputs (so a mis-routed tensor or frame changes the result)."""
    y = (input.mean(1, keepdim=True) + 2.0 * land1 - 3.0 * land2 + motion.permute(0, 3, 1, 2).sum(1, keepdim=True)
         + 0.5 * flow.sum(1, keepdim=True) + ifmask)
    return torch.tanh(y)


def _make_clip(T, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn((T,) + FRAME_SHAPES[n], generator=g) for n in INPUT_NAMES]


def _rendezvous(tmp_path):
    """file:// rendezvous in the test's own directory: no port to pick, so no race for one between back-to-back tests"""
    return "file://" + str(tmp_path / "rendezvous")


def _worker(rank, world, rendezvous, T, batch, out_path):
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    dist.init_process_group("gloo", init_method=rendezvous, rank=rank, world_size=world)
    try:
        clip = _make_clip(T) if rank == 0 else None
        frames = render_frames_sharded(_stand_in_netg, clip, T, 1, torch.device("cpu"), batch=batch)
        if rank == 0:
            torch.save(frames, out_path)
        else:
            assert frames is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T,batch", [(2, 5, 2), (2, 1, 4), (3, 7, 16)])
def test_sharded_render_equals_single_process(world, T, batch, tmp_path):
    out = str(tmp_path / "frames.pt")
    mp.spawn(_worker, args=(world, _rendezvous(tmp_path), T, batch, out), nprocs=world, join=True)
    got = torch.load(out)
    want = render_frames(_stand_in_netg, _make_clip(T), batch=3)
    assert got.shape == (T, 1, 256, 256)
    assert torch.equal(got, want)  # same per-frame arithmetic, frames back in clip order


def test_shard_ranges_partition_the_clip():
    for T in (0, 1, 5, 16, 733, 800):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(T, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == T
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
