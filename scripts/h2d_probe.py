#!/usr/bin/env python
"""Host<->device copy bandwidth of every rank, alone and with all ranks copying at once.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_probe.py

Why: bench.py's e2e leg (host views in, labels out every step) ran at 52 GB/s on one GPU but at
15.5 GB/s per GPU with eight ranks.  This prints, per rank, the H2D / D2H rates
for ordinary pinned memory and for write-combined pinned memory, with 1 or all ranks active, plus the
topology facts that decide what can be done about it (NUMA nodes, CPU affinity, cores).
Rank 0 prints one JSON object.
"""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def wc_pinned(nbytes):
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(4))   # cudaHostAllocWriteCombined
    if rc != 0:
        return None
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


def rate(fn, nbytes, iters=8):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    from ups_b200.dp import init_from_env
    rank, local, world = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = 256 << 20
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    wc = wc_pinned(n)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream()
    res = {}

    def both():
        d.copy_(host, non_blocking=True)
        with torch.cuda.stream(s2):
            host2.copy_(d2, non_blocking=True)
    host2 = torch.empty(n, dtype=torch.uint8).pin_memory()

    def bar():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tests = {"h2d_pinned": lambda: d.copy_(host, non_blocking=True),
             "d2h_pinned": lambda: host.copy_(d, non_blocking=True),
             "h2d_and_d2h_pinned": both}
    if wc is not None:
        tests["h2d_write_combined"] = lambda: d.copy_(wc, non_blocking=True)
    # all ranks at once
    for name, fn in tests.items():
        bar()
        r = rate(fn, n * (2 if name.startswith("h2d_and") else 1))
        torch.cuda.synchronize()
        res[name + "_all_ranks"] = r
    # one rank at a time
    for name, fn in tests.items():
        for turn in range(world):
            bar()
            if turn == rank:
                res[name + "_alone"] = rate(fn, n * (2 if name.startswith("h2d_and") else 1))
        bar()
    allres = [None] * world
    if world > 1:
        dist.all_gather_object(allres, res)
    else:
        allres = [res]
    if rank == 0:
        info = {"world": world, "cpu_count": os.cpu_count(), "affinity": sorted(os.sched_getaffinity(0)),
                "numa_nodes": sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
                if os.path.isdir("/sys/devices/system/node") else None}
        try:
            info["meminfo"] = [l.strip() for l in open("/proc/meminfo").read().splitlines()[:3]]
            info["cpu_model"] = [l for l in open("/proc/cpuinfo").read().splitlines() if "model name" in l][0]
        except Exception:  # noqa: BLE001
            pass
        keys = sorted(allres[0])
        summary = {k: {"per_rank_GBs": [round(a[k], 1) for a in allres], "sum_GBs": round(sum(a[k] for a in allres), 1)}
                   for k in keys}
        print(json.dumps({"info": info, "copy_rates": summary}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
