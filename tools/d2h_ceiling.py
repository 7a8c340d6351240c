#!/usr/bin/env python
"""Bare device->host copy ceiling of a box, N GPUs at once: the number bench.py's e2e leg runs into.

    python tools/d2h_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/d2h_ceiling.py

Every rank copies picture-sized blocks (1080p planar: 3,133,440 bytes, like the decoder's copy-out)
from its GPU into pinned host memory for a fixed time, first with the process bound to the CPUs /
memory of its GPU's NUMA node (jsmpeg_b200_bind_host_to_device, what bench.py does), then -- in a
fresh set of buffers allocated before binding -- unbound, as round 1 ran.  Rank 0 prints one JSON line
with the aggregate GB/s of both, so that `e2e` at N GPUs can be read against what the host can take.
Round 1's 8-GPU e2e went flat at ~124 GB/s aggregate (SCALE_r01: 39.6k frames/s x 3.13 MB at N = 4 and 8).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(torch, device, host_bufs, dev, seconds, streams):
    """Copy len(host_bufs) blocks round-robin on `streams` CUDA streams for `seconds`; returns bytes/s."""
    n = 0
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for i, h in enumerate(host_bufs):
            with torch.cuda.stream(streams[i % len(streams)]):
                h.copy_(dev, non_blocking=True)
        n += len(host_bufs)
        for s in streams:
            s.synchronize()
    dt = time.perf_counter() - t0
    return n * dev.numel() / dt


def main():
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    seconds = float(os.environ.get("D2H_SECONDS", 2.0))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    block = 1920 * 1088 * 3 // 2
    dev = torch.empty(block, dtype=torch.uint8, device=device)
    streams = [torch.cuda.Stream(device) for _ in range(2)]

    def allocate():
        bufs = [torch.empty(block, dtype=torch.uint8).pin_memory() for _ in range(64)]
        for b in bufs:
            b.fill_(0)  # first touch here
        return bufs

    def barrier():
        if world > 1:
            dist.barrier()

    unbound = allocate()  # pages placed wherever this process happens to run (round 1's situation)
    barrier()
    r_unbound = measure(torch, device, unbound, dev, seconds, streams)
    barrier()
    del unbound
    from jsmpeg_b200 import capi
    numa = capi.bind_host_to_device(local_rank)
    bound = allocate()
    barrier()
    r_bound = measure(torch, device, bound, dev, seconds, streams)
    barrier()
    vals = torch.tensor([r_unbound, r_bound], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(vals)
    if rank == 0:
        print(json.dumps({"what": "concurrent D2H of 3.13 MB blocks into pinned host memory, aggregate over all ranks",
                          "n_gpus": world, "unbound_GBps": vals[0].item() / 1e9, "numa_bound_GBps": vals[1].item() / 1e9,
                          "rank0_numa": numa, "seconds_each": seconds,
                          "frames_per_s_ceiling_1080p": vals[1].item() / block}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
