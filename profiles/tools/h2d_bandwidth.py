"""Host-to-device / device-to-host copy bandwidth from pinned memory, on one GPU or on all GPUs of the box at once
(run under torchrun: every rank copies at the same time, bracketed by barriers). One JSON line per configuration.

  python profiles/tools/h2d_bandwidth.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/tools/h2d_bandwidth.py
"""
import json
import os
import time

import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 4_000_000_000
host = [torch.empty(n // 4, dtype=torch.uint8).pin_memory() for _ in range(4)]
dev = [torch.empty(n // 4, dtype=torch.uint8, device="cuda") for _ in range(4)]


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()


def timed(fn):
    best = None
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        best = dt if best is None else min(best, dt)
    return best


for k in (1, 2):
    streams = [torch.cuda.Stream() for _ in range(k)]

    def h2d():
        for i in range(4):
            with torch.cuda.stream(streams[i % k]):
                dev[i].copy_(host[i], non_blocking=True)

    dt = timed(h2d)
    if rank == 0:
        print(json.dumps({"direction": "h2d", "gpus_at_once": world, "streams_per_gpu": k, "gb_per_gpu": n / 1e9,
                          "gbs_per_gpu": n / dt / 1e9, "gbs_aggregate": world * n / dt / 1e9}))


def d2h():
    for i in range(4):
        host[i].copy_(dev[i], non_blocking=True)


dt = timed(d2h)
if rank == 0:
    print(json.dumps({"direction": "d2h", "gpus_at_once": world, "streams_per_gpu": 1, "gb_per_gpu": n / 1e9,
                      "gbs_per_gpu": n / dt / 1e9, "gbs_aggregate": world * n / dt / 1e9}))
if dist is not None:
    dist.destroy_process_group()
