#!/usr/bin/env python
"""What bounds the host-fed path when all GPUs of a box are fed at once: every rank
copies pinned host memory to its GPU at the same time --
  (a) from a large buffer (512 MiB, streams through host DRAM),
  (b) from a small one (4 MiB, stays in the last-level cache),
and the other way round, and the ranks' aggregate is compared with one rank alone and with what
the host cores copy among themselves (memcpy of 256 MiB per rank).  If (b) scales where (a) does
not, host DRAM is the ceiling; if neither does, it is the PCIe root / the hypervisor's IOMMU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/host_bw.py"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev_big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, nbytes, reps, solo):
        """GB/s of this rank; all ranks at once unless solo (then only rank 0 works)."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if not solo or rank == 0:
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        x = torch.tensor([nbytes * reps / dt / 1e9 if (not solo or rank == 0) else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(x)
        return float(x)

    rows = []
    for name, size, reps in (("512 MiB (DRAM)", 512 << 20, 8), ("4 MiB (cache)", 4 << 20, 1024)):
        host = torch.empty(size, dtype=torch.uint8).pin_memory()
        host.fill_(rank + 1)
        d = dev_big[:size]
        for dirn, fn in (("H2D", lambda: d.copy_(host, non_blocking=True)), ("D2H", lambda: host.copy_(d, non_blocking=True))):
            fn(); torch.cuda.synchronize()
            solo = timed(fn, size, reps, True)
            allr = timed(fn, size, reps, False)
            rows.append((f"{dirn} {name}", solo, allr))
    a = np.ones(256 << 20, np.uint8)
    b = np.empty_like(a)
    np.copyto(b, a)
    solo = timed(lambda: np.copyto(b, a), 2 * a.nbytes, 4, True)      # read + write
    allr = timed(lambda: np.copyto(b, a), 2 * a.nbytes, 4, False)
    rows.append(("host memcpy 256 MiB (read + write bytes)", solo, allr))
    if rank == 0:
        print(f"{world} ranks, one GPU each; GB/s: one rank alone | all ranks at once (sum)")
        for n, s, al in rows:
            print(f"  {n:44s} {s:8.1f} | {al:8.1f}   (x{al / s:.2f})")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
