#!/usr/bin/env python
"""Stand-alone timing of the gradient all-reduce (csrc/dp_allreduce.cu) against ncclAllReduce.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_allreduce.py [--mb 133.2]

For each variant (multicast / peer, number of CTAs) and for NCCL (SUM): CUDA-event time of `iters` back-to-back
all-reduces of the whole buffer on an otherwise idle GPU, max over ranks -> algorithmic bandwidth (buffer bytes / time)
and the NVLink-roofline fraction: every byte of the buffer must leave and enter each GPU once ((N-1)/N of it over
NVLink), so the floor is bytes*(N-1)/N / 770 GB/s (B200_PROFILING.md: measured peer copy rate per direction).
Rank 0 prints one JSON object.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=133.2)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ctas", default="8,16,32,64,128")
    args = ap.parse_args()
    from ups_b200.dp import GradAllReducer, init_from_env
    rank, local, world = init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n = (int(args.mb * 1e6 / 4) + 3) // 4 * 4

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"world": world, "buffer_mb": n * 4 / 1e6, "variants": {}}
    floor_ms = n * 4 * (world - 1) / world / 770e9 * 1e3

    def record(name, ms):
        out["variants"][name] = {"ms": round(ms, 4), "algbw_GBs": round(n * 4 / ms / 1e6, 1),
                                 "frac_of_nvlink_floor": round(floor_ms / ms, 3)}

    for mc in ("1", "0"):
        os.environ["UPS_DP_MULTICAST"] = mc
        red = GradAllReducer(n, dev, buckets=1, impl="peer")
        red.flat.fill_(1.0)
        cur = torch.cuda.current_stream()

        def run():
            red.launch()
            red.wait()
        for c in [int(x) for x in args.ctas.split(",")]:
            red.n_ctas = c
            record(f"{red.transport}_ctas{c}", timed(run))
        # correctness of the last variant: all ranks filled 1.0 -> mean stays 1.0 however often it is applied
        torch.cuda.synchronize()
        assert float((red.flat - 1.0).abs().max()) == 0.0, "mean of ones must be one"
        del red
    flat = torch.ones(n, device=dev)
    record("nccl_sum", timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.SUM)))
    out["nvlink_floor_ms"] = round(floor_ms, 4)
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
