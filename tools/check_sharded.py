"""Multi-GPU parity of frames.render_frames_sharded on real GPUs (launch with torchrun, one rank per GPU):
the frames rank 0 collects -- through the peer-mapped gather buffer and through NCCL, from device-resident and from
host-resident clips, ragged shard sizes -- must equal the same frames rendered on rank 0's GPU alone, bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import animateportrait_b200 as ap
from animateportrait_b200 import synth
from animateportrait_b200.frames import render_frames, render_frames_sharded


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    net = ap.define_G(3, 1, 64, ap.NETG_NAME, "instance", False, "normal", 0.02, [local], div=3, disp=3).module
    net.load_state_dict(synth.make_state_dict(1, seed=3, bias_std=0.1))
    ok = True
    for T, batch in ((world * 5 + 1, 2), (world * 8, 4)):
        host = [t.pin_memory() for t in synth.make_inputs(T, seed=50 + T, kind="smooth")] if rank == 0 else None
        devs = [t.to(dev) for t in host] if rank == 0 else None
        want = render_frames(net, devs, batch=batch) if rank == 0 else None
        for gather in (sys.argv[1:] or ["nccl", "peer"]):
            for src, name in ((devs, "device clip"), (host, "host clip")):
                info = {}
                out_host = torch.empty((T, 1, 256, 256), pin_memory=True) if rank == 0 else None
                for rep in range(2):  # second call reuses streams / peer buffer
                    got = render_frames_sharded(net, src, T, 1, dev, batch=batch, gather=gather, out_host=out_host, info=info)
                    torch.cuda.synchronize()
                    dist.barrier()
                    if rank == 0:
                        d = (got - want).abs().max().item()
                        dh = (out_host.to(dev) - want).abs().max().item()
                        good = d == 0.0 and dh == 0.0
                        ok &= good
                        print(f"T={T} batch={batch} gather={gather} ({info['gather']}) {name} rep={rep}: max diff {d:.2e} / host copy {dh:.2e} "
                              f"{'OK' if good else 'FAIL'}", flush=True)
    if rank == 0:
        print("SHARDED PARITY", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
