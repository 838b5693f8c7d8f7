#!/usr/bin/env python3
"""Host I/O ceiling of the box: pinned H2D and D2H throughput with N ranks copying AT THE SAME TIME (one process per
GPU, as bench.py runs), so the end-to-end scaling of the replicas can be read against what the host side can deliver.

    python tools/hostio_probe.py                                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/hostio_probe.py

Rank 0 prints one JSON object: per-rank and aggregate GB/s for H2D alone, D2H alone and both directions together
(30 MB copies, the size of one bench step's image block; best of 5 after a barrier)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 30 * 1024 * 1024
REP = 8                                     # copies per timed burst (per direction)
h_in = torch.empty(N, dtype=torch.uint8).pin_memory(); h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device="cuda"); d_out = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def burst(h2d, d2h):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(REP):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return REP * N / best / 1e9


res = {}
for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both_each_direction", True, True)):
    v = torch.tensor([burst(a, b)], device="cuda")
    if world > 1:
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        vals = [float(x.item()) for x in allv]
    else:
        vals = [float(v.item())]
    res[name] = {"per_rank_gbs": [round(x, 1) for x in vals], "aggregate_gbs": round(sum(vals), 1)}
if rank == 0:
    print(json.dumps({"ranks": world, "copy_bytes": N, "host_cores": os.cpu_count(), **res}))
if world > 1:
    dist.destroy_process_group()
