#!/usr/bin/env python
"""Concurrent pinned host->device copy ceiling of the box at N ranks (one process per GPU, plain cudaMemcpyAsync through torch): the denominator
for the e2e numbers of bench.py at N > 1 (every rank copies ~0.5 GB per step). Run:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/h2d_ceiling.py [--mb 512] [--reps 20]
Rank 0 prints one JSON line: per-rank GB/s (min / mean / max) and the aggregate (total bytes / max-over-ranks time)."""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--with-host-traffic", type=int, default=0, help="threads per rank that stream-copy host memory during the timed region")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(rank + 1)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    stop = False
    threads = []
    if args.with_host_traffic:
        import threading
        a = torch.empty(256 << 20, dtype=torch.uint8)
        b = torch.empty(256 << 20, dtype=torch.uint8)

        def churn():
            while not stop:
                b.copy_(a)
        threads = [threading.Thread(target=churn, daemon=True) for _ in range(args.with_host_traffic)]
        for t in threads:
            t.start()
    with torch.cuda.stream(s):
        for _ in range(3):
            d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(args.reps):
            d.copy_(h, non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    stop = True
    ms = e0.elapsed_time(e1)
    gbs = n * args.reps / (ms * 1e-3) / 1e9
    t = torch.tensor([ms, gbs], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = mn = sm = t
    if rank == 0:
        print(json.dumps({"n_gpus": world, "mb_per_copy": args.mb, "reps": args.reps, "host_traffic_threads_per_rank": args.with_host_traffic,
                          "per_rank_GBps": {"min": mn[1].item(), "mean": sm[1].item() / world, "max": mx[1].item()},
                          "aggregate_GBps": world * n * args.reps / (mx[0].item() * 1e-3) / 1e9, "host_cores": os.cpu_count()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
