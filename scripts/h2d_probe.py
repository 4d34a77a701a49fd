#!/usr/bin/env python
"""Bare pinned-memory cudaMemcpyAsync probe: host->device (and device->host) bandwidth of ONE process per GPU, all
ranks copying at the same time (torchrun).  Tells whether the end-to-end scaling of bench.py's `e2e` (0.43 at N=8 in
round 1) is the platform (host memory / PCIe root complexes shared by the GPUs) or the library's host-buffer path."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 470 << 20                       # one cfg2 logits batch
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    res = {}
    for name, src, dst in (("h2d", h, d), ("d2h", d, h)):
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[name] = nbytes * n / dt / 1e9
    vals = torch.tensor([res["h2d"], res["d2h"]], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        h2d = [round(float(v[0]), 1) for v in allv]
        d2h = [round(float(v[1]), 1) for v in allv]
        print(f"H2D_PROBE world={world} pinned cudaMemcpyAsync of {nbytes >> 20} MiB, all ranks at once: "
              f"h2d GB/s per GPU {h2d} (sum {sum(h2d):.1f}); d2h {d2h} (sum {sum(d2h):.1f})")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
