#!/usr/bin/env python
"""Host -> device copy ceiling of this box at N concurrent ranks: the roofline of bench.py's `e2e` number.

    python tools/h2d_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_ceiling.py

Every rank pins one 1.92 GB window (bench.py's e2e window) after binding to its GPU's NUMA node when the platform
exposes one, then all ranks copy it to their device `--reps` times between barriers; no kernel runs.  Rank 0 prints
one JSON line: aggregate and per-rank GB/s, plus the same with a device -> host copy of the labels running alongside.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=1920000000)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bench import bind_to_gpu_numa
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    h = torch.empty(a.bytes // 2, dtype=torch.int16).pin_memory()
    h.zero_()
    d = torch.empty_like(h, device=dev)
    lab_d = torch.zeros(a.bytes // 320, dtype=torch.uint8, device=dev)
    lab_h = torch.empty(a.bytes // 320, dtype=torch.uint8).pin_memory()
    side = torch.cuda.Stream(device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(with_d2h):
        d.copy_(h, non_blocking=True)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            d.copy_(h, non_blocking=True)
            if with_d2h:
                with torch.cuda.stream(side):
                    lab_h.copy_(lab_d, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
        return a.reps * a.bytes * world / float(t.item()) / 1e9

    up = timed(False)
    both = timed(True)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "window_bytes": a.bytes, "reps": a.reps, "h2d_GBps_aggregate": up,
                          "h2d_GBps_per_gpu": up / world, "h2d_GBps_aggregate_with_label_d2h": both, "numa_rank0": numa,
                          "host_cpus": os.cpu_count(), "gpu": torch.cuda.get_device_name(local)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
