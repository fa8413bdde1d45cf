#!/usr/bin/env python
"""Roofline numbers for the mask-statistics / prior / sampling kernels at the CUB bench shape:
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
    ws = torch.empty(max(C.workspace_bytes(op, B, H * W, K, 0) for op in
                         (C.OP_MOMENTS, C.OP_KL, C.OP_MUMFORD_SHAH, C.OP_LOGIT_PRIORS, C.OP_WEAK_XENT)),
                     dtype=torch.uint8, device=dev)
    nbytes = probs.numel() * 4
    n_pix = B * H * W
    logits = torch.randn(B, H, W, K, device=dev, generator=g)
    eps = torch.randn(B, H, W, K, device=dev, generator=g)
    d2, d3 = torch.empty_like(probs), torch.empty_like(probs)
    labels = torch.empty(B, H, W, dtype=torch.int64, device=dev)
    sums, g_sums = torch.empty(B, 4, K, device=dev), torch.randn(B, 4, K, device=dev)
    pri, g_pri = torch.empty(3, device=dev), torch.ones(3, device=dev)
    table, rgb = torch.rand(K, 3, device=dev) * 2 - 1, torch.empty(B, H, W, 3, device=dev)
    lam = float(1.0 * (0.25 * (probs[:, :, :-1] - probs[:, :, 1:])).square().median() * 2)
    P = lambda t: t.data_ptr()  # noqa: E731
    calls = {
        "ups_mask_moments_fwd": (lambda: C.call("ups_mask_moments_fwd", probs.data_ptr(), sf.data_ptr(), mu.data_ptr(),
                                                sigma.data_ptr(), mom.data_ptr(), B, H, W, K, ws.data_ptr(), ws.numel(), st), nbytes),
        "ups_mask_moments_bwd": (lambda: C.call("ups_mask_moments_bwd", g_mu.data_ptr(), g_sigma.data_ptr(), sf.data_ptr(),
                                                mom.data_ptr(), d.data_ptr(), B, H, W, K, st), nbytes),
        "ups_categorical_kl_fwd": (lambda: C.call("ups_categorical_kl_fwd", probs.data_ptr(), out.data_ptr(), B * H * W, K,
                                                  ws.data_ptr(), ws.numel(), st), nbytes),
        "ups_categorical_kl_bwd": (lambda: C.call("ups_categorical_kl_bwd", probs.data_ptr(), gout.data_ptr(), d.data_ptr(),
                                                  B * H * W, K, st), 2 * nbytes),
        "ups_mumford_shah_fwd(sums)": (lambda: C.call("ups_mumford_shah_fwd", P(probs), 1.0, lam, None, None, None, None, P(sums),
                                                      B, H, W, K, P(ws), ws.numel(), st), nbytes),
        "ups_mumford_shah_fwd(maps)": (lambda: C.call("ups_mumford_shah_fwd", P(probs), 1.0, lam, P(d), P(d2), P(d3), None, None,
                                                      B, H, W, K, None, 0, st), 4 * nbytes),
        "ups_mumford_shah_bwd(sums)": (lambda: C.call("ups_mumford_shah_bwd", P(probs), 1.0, lam, None, None, None, P(g_sums), P(d),
                                                      B, H, W, K, st), 2 * nbytes),
        "ups_logit_priors_fwd": (lambda: C.call("ups_logit_priors_fwd", P(logits), P(pri), B, H, W, K, P(ws), ws.numel(), st), nbytes),
        "ups_logit_priors_bwd": (lambda: C.call("ups_logit_priors_bwd", P(logits), P(g_pri), P(d), B, H, W, K, st), 2 * nbytes),
        "ups_mean_field_sample_fwd": (lambda: C.call("ups_mean_field_sample_fwd", P(logits), P(eps), 0.7, P(d), logits.numel(), st),
                                      3 * nbytes),
        "ups_part_softmax_sampled_fwd": (lambda: C.call("ups_part_softmax_sampled_fwd", P(logits), P(eps), 0.7, P(d), P(d2),
                                                        P(labels), P(d3), n_pix, K, st), 5 * nbytes + 8 * n_pix),
        "ups_part_softmax_fwd": (lambda: C.call("ups_part_softmax_fwd", P(logits), P(d2), P(labels), P(d3), n_pix, K, st),
                                 3 * nbytes + 8 * n_pix),
        "ups_weak_xent_fwd": (lambda: C.call("ups_weak_xent_fwd", P(logits), 0, P(out), n_pix, K, P(ws), ws.numel(), st), nbytes),
        "ups_weak_xent_bwd": (lambda: C.call("ups_weak_xent_bwd", P(logits), 0, P(gout), P(d), n_pix, K, st), 2 * nbytes),
        "ups_mask2rgb_fwd(hot)": (lambda: C.call("ups_mask2rgb_fwd", P(probs), P(table), 1, P(rgb), n_pix, K, st),
                                  nbytes + 12 * n_pix),
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
