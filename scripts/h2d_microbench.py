#!/usr/bin/env python
"""Pure host->device copy bandwidth of every rank at once (torchrun), next to bench.py's e2e numbers: is the host-buffer
path of the N-GPU run bound by the box (PCIe / host memory) or by the code?  Each rank copies a page-locked float32 image
batch of the bench's size to its GPU, K times, all ranks simultaneously; also a pageable->pinned staging copy by one
thread (numpy) for reference.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    cpus = bind_to_gpu_numa_node(torch, lr) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    B = 64 if world == 1 else 1024 // world
    n = B * 608 * 608 * 3
    pin = torch.empty(n, dtype=torch.float32, pin_memory=True); pin.uniform_(-1, 1)
    dev = torch.empty(n, dtype=torch.float32, device="cuda")
    page = np.array(pin.numpy(), copy=True)
    for _ in range(2):
        dev.copy_(pin, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    K = 5
    t0 = time.perf_counter()
    for _ in range(K):
        dev.copy_(pin, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    h2d = n * 4 * K / dt / 1e9
    t0 = time.perf_counter()
    np.copyto(pin.numpy(), page)
    stage1 = n * 4 / (time.perf_counter() - t0) / 1e9
    vals = torch.tensor([h2d, stage1], dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(vals) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        h = [float(v[0]) for v in allv]; s = [float(v[1]) for v in allv]
        print(json.dumps({"n_gpus": world, "batch_per_rank": B, "bytes_per_copy": n * 4, "h2d_gbs_per_rank": h, "h2d_gbs_aggregate": sum(h),
                          "one_thread_pageable_to_pinned_gbs_per_rank": s, "host_cpus": os.cpu_count(),
                          "affinity": f"{len(cpus)} cpus local to the GPU" if cpus else "unbound"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
