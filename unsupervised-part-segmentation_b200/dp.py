"""Batch-sharded data-parallel wrapper of the per-step path (one process per GPU).

The path itself has no cross-sample dependency (TPS parameters, softmax, masks, pooling and
unpooling are all indexed by the sample), so ranks exchange nothing on the data path.  The
only collective of a training step is the gradient all-reduce of the encoder/decoder
parameters that surround the path (the reference trains on one GPU and has none; its IMM
baseline averages tower gradients on the CPU, baselines/imm/imm/train/cnn_train_multi.py:
75-118).  Here that all-reduce is NCCL over NVLink, bucketed, issued on a side stream so that
it overlaps the path's backward kernels.
"""
import os

import torch
import torch.distributed as dist


def shard_bounds(global_batch, rank, world_size):
    """Contiguous batch shard [lo, hi) of `rank`; the first (global_batch % world) ranks get one extra."""
    base, rem = divmod(int(global_batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed, rank):
    """Per-rank seed for the synthetic shard (rank-offset)."""
    return int(seed) * 1000003 + int(rank)


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


class GradAllReducer:
    """Bucketed mean all-reduce of a flat gradient buffer on a side stream (NCCL) or inline (gloo)."""

    def __init__(self, flat_grads, bucket_bytes=32 << 20, world_size=None):
        self.flat = flat_grads
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        n = max(1, int(bucket_bytes) // flat_grads.element_size())
        self.buckets = [flat_grads[i:i + n] for i in range(0, flat_grads.numel(), n)]
        self.cuda = flat_grads.is_cuda
        self.stream = torch.cuda.Stream(device=flat_grads.device) if self.cuda else None
        self._done = None

    def launch(self):
        """Enqueue the all-reduce behind everything already queued on the current stream."""
        if self.world == 1:
            return
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for b in self.buckets:
                    # SUM (not AVG): NCCL's in-switch NVLS algorithms exist for sum only, and on NVSwitch they
                    # need far fewer SM-resident channels than the 32-channel ring AVG falls back to
                    dist.all_reduce(b, op=dist.ReduceOp.SUM)
                    b.mul_(1.0 / self.world)
                self._done = torch.cuda.Event()
                self._done.record(self.stream)
        else:
            for b in self.buckets:
                dist.all_reduce(b, op=dist.ReduceOp.SUM)
                b.mul_(1.0 / self.world)

    def wait(self):
        """Make the current stream wait for the reduction (no host synchronisation)."""
        if self.cuda and self._done is not None:
            torch.cuda.current_stream().wait_event(self._done)
            self._done = None


class DataParallelPartStep:
    """PartStep on this rank's shard + overlapped all-reduce of the surrounding modules' gradients.

    `n_grad_params` sizes the stand-in gradient buffer (default 33.3 M fp32 = the reference's
    e_pi + e_alpha + dv + dd + discriminators)."""

    def __init__(self, per_gpu_batch, spatial_size, n_parts, local_app_size=64, n_views=3, use_tps=True,
                 views_grad=False, n_grad_params=33_300_000, bucket_bytes=256 << 20, device="cuda",
                 decode_bwd="auto"):
        from .step import PartStep
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.step = PartStep(per_gpu_batch, spatial_size, n_parts, local_app_size, n_views, use_tps, views_grad, device,
                             decode_bwd=decode_bwd)
        self.grads = torch.zeros(int(n_grad_params), dtype=torch.float32, device=device)
        self.reducer = GradAllReducer(self.grads, bucket_bytes, self.world)

    def forward(self, *a, **k):
        return self.step.forward(*a, **k)

    def backward(self, *a, **k):
        # the surrounding CNNs' gradients exist once their backward has run; the path's own
        # backward kernels (K4-K6) then overlap with the collective
        self.reducer.launch()
        out = self.step.backward(*a, **k)
        self.reducer.wait()
        return out
