#!/usr/bin/env python
"""Roofline numbers for the mask-statistics kernels at the CUB bench shape:
probs [256,128,128,16].  CUDA-event timing per C-ABI call, algorithmic bytes / time vs the measured HBM peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import ups_b200
    from ups_b200 import _cabi as C
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    B, H, W, K = 256, 128, 128, 16
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    probs = torch.softmax(torch.randn(B, H, W, K, device=dev, generator=g), -1)
    sf = torch.ones(B, K, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    mu, sigma, mom = torch.empty(B, K, 2, device=dev), torch.empty(B, K, 2, 2, device=dev), torch.empty(B, K, 5, device=dev)
    g_mu, g_sigma = torch.randn(B, K, 2, device=dev), torch.randn(B, K, 2, 2, device=dev)
    d = torch.empty_like(probs)
    out, gout = torch.empty((), device=dev), torch.ones((), device=dev)
    ws = torch.empty(max(C.workspace_bytes(C.OP_MOMENTS, B, H * W, K, 0), C.workspace_bytes(C.OP_KL, 0, 0, K, 0)),
                     dtype=torch.uint8, device=dev)
    nbytes = probs.numel() * 4
    calls = {
        "ups_mask_moments_fwd": (lambda: C.call("ups_mask_moments_fwd", probs.data_ptr(), sf.data_ptr(), mu.data_ptr(),
                                                sigma.data_ptr(), mom.data_ptr(), B, H, W, K, ws.data_ptr(), ws.numel(), st), nbytes),
        "ups_mask_moments_bwd": (lambda: C.call("ups_mask_moments_bwd", g_mu.data_ptr(), g_sigma.data_ptr(), sf.data_ptr(),
                                                mom.data_ptr(), d.data_ptr(), B, H, W, K, st), nbytes),
        "ups_categorical_kl_fwd": (lambda: C.call("ups_categorical_kl_fwd", probs.data_ptr(), out.data_ptr(), B * H * W, K,
                                                  ws.data_ptr(), ws.numel(), st), nbytes),
        "ups_categorical_kl_bwd": (lambda: C.call("ups_categorical_kl_bwd", probs.data_ptr(), gout.data_ptr(), d.data_ptr(),
                                                  B * H * W, K, st), 2 * nbytes),
    }
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2; READ to flush (clean lines)
    res = {}
    for name, (fn, by) in calls.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            flush.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        res[name] = dict(ms=ms, algorithmic_bytes=by, gbs=by / (ms * 1e-3) / 1e9, frac_of_measured_hbm=by / (ms * 1e-3) / 1e9 / peak)
        print(name, json.dumps(res[name]), flush=True)
    json.dump(dict(shape=[B, H, W, K], peak_gbs=peak, l2="flushed between timed calls (256 MB read)", calls=res),
              open(os.path.join(ROOT, "gpurun_out", "stats_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
