#!/usr/bin/env python
"""SURVEY.md 8f N4 at the CUB bench shape (B=256, 128x128, K=16, F=64, Co=32): the first decoder / encoder
convolutions computed on the part assignment versus materialising `injected` / the part images and running a
library convolution on them (cuDNN through torch, NHWC).  CUDA-event timing per call, L2 flushed between timed
calls; algorithmic bytes / time against the measured HBM peak.  Writes gpurun_out/<tag>_inject_conv.json."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(batch=256, size=128, parts=16, library=True, once=False, iters=10):
    """Returns {call: {ms, algorithmic_bytes, gbs, frac_of_measured_hbm}} (None with once=True)."""
    import types
    a = types.SimpleNamespace(batch=batch, size=size, parts=parts, no_library=not library, once=once)
    import torch
    import ups_b200  # noqa: F401
    from ups_b200 import _cabi as C
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(pk))["hbm_gbs"]) if os.path.exists(pk) else 6650.0
    B, H, W, K, F, Co = a.batch, a.size, a.size, a.parts, 64, 32
    P = H * W
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
    l0, feat = rn(B, H, W, K), rn(B, K, F)
    V, bias = rn(9, F + K, Co) * 0.04, rn(Co) * 0.04
    Ve, be = rn(9, 3, Co) * 0.2, rn(Co) * 0.2
    img = torch.rand(B, H, W, 3, device=dev, generator=g) * 2 - 1
    g_out, g_m0 = rn(B, H, W, Co), rn(B, H, W, K)
    m0, mh, dl0 = torch.empty_like(l0), torch.empty_like(l0), torch.empty_like(l0)
    labels = torch.empty(B, H, W, dtype=torch.int64, device=dev)
    G, dG = torch.empty(B, 9, K, Co, device=dev), torch.empty(B, 9, K, Co, device=dev)
    out = torch.empty(B, H, W, Co, device=dev)
    db, dfeat, dV = torch.empty(Co, device=dev), torch.empty_like(feat), torch.empty_like(V)
    inj = torch.empty(B, H, W, F + K, device=dev)
    out_pm = torch.empty(K * B, H, W, Co, device=dev)
    ws = torch.empty(max(C.inject_conv_workspace_bytes(B, H, W, K, Co), C.workspace_bytes(C.OP_STEP, B, P, K, F)),
                     dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr()  # noqa: E731
    n_pix = B * P
    calls = {
        # ---- this library, no injected map
        "ups_part_softmax_fwd (l0 -> m0, labels, mask)": (
            lambda: C.call("ups_part_softmax_fwd", p(l0), p(m0), p(labels), p(mh), n_pix, K, st), n_pix * (12 * K + 8)),
        "ups_inject_conv_table_fwd": (lambda: C.call("ups_inject_conv_table_fwd", p(feat), p(V), p(G), B, K, F, Co, st),
                                      4 * (B * K * F + B * 9 * K * Co)),
        "ups_inject_conv_fwd (mask -> h)": (
            lambda: C.call("ups_inject_conv_fwd", p(mh), p(G), p(bias), p(out), B, H, W, K, Co, st), n_pix * 4 * (K + Co)),
        "ups_inject_conv_bwd (g_h, g_m0 -> dl0, dG, db)": (
            lambda: C.call("ups_inject_conv_bwd", p(g_out), p(mh), p(G), p(m0), p(g_m0), p(dl0), p(dG), p(db), B, H, W, K, Co,
                           p(ws), ws.numel(), st), n_pix * 4 * (Co + 4 * K)),
        "ups_inject_conv_table_bwd": (
            lambda: C.call("ups_inject_conv_table_bwd", p(dG), p(feat), p(V), p(dfeat), p(dV), B, K, F, Co, st),
            4 * (2 * B * K * F + B * 9 * K * Co)),
        # ---- this library, materialising the injected map (K3 / K4 of the step)
        "ups_step_decode_fwd (l0 -> m0, labels, injected)": (
            lambda: C.call("ups_step_decode_fwd", p(l0), p(feat), p(m0), p(labels), p(inj), B, P, K, F, st),
            n_pix * (4 * (2 * K + F + K) + 8)),
        # ---- encoder side
        "ups_parts_conv_fwd (img, mask -> [K*B,P,Co])": (
            lambda: C.call("ups_parts_conv_fwd", p(img), p(mh), p(Ve), p(be), p(out_pm), B, H, W, K, 3, Co, st),
            n_pix * 4 * (3 + K + K * Co)),
    }
    g_pm = torch.empty(K * B, H, W, Co, device=dev).normal_(generator=g)
    dm1, dVe, dbe = torch.empty_like(l0), torch.empty_like(Ve), torch.empty_like(be)
    ws_pc = torch.empty(C.parts_conv_bwd_workspace_bytes(B, H, W, K, Co), dtype=torch.uint8, device=dev)
    calls["ups_parts_conv_bwd (g [K*B,P,Co], g_m1 -> dl1, dV, db)"] = (
        lambda: C.call("ups_parts_conv_bwd", p(g_pm), p(img), p(mh), p(Ve), p(m0), p(g_m0), p(dm1), p(dVe), p(dbe),
                       B, H, W, K, 3, Co, p(ws_pc), ws_pc.numel(), st), n_pix * 4 * (K * Co + 3 + 4 * K))
    C.call("ups_part_softmax_fwd", p(l0), p(m0), p(labels), p(mh), n_pix, K, st)
    C.call("ups_inject_conv_table_fwd", p(feat), p(V), p(G), B, K, F, Co, st)
    if not a.no_library:
        import torch.nn.functional as TF
        w_nchw = V.reshape(3, 3, F + K, Co).permute(3, 2, 0, 1).contiguous(memory_format=torch.channels_last)
        we_nchw = Ve.reshape(3, 3, 3, Co).permute(3, 2, 0, 1).contiguous(memory_format=torch.channels_last)
        x_inj = inj.permute(0, 3, 1, 2)            # NHWC storage viewed NCHW = channels_last
        C.call("ups_step_decode_fwd", p(l0), p(feat), p(m0), p(labels), p(inj), B, P, K, F, st)
        gy_nchw = g_out.permute(0, 3, 1, 2)
        parts_pm = torch.empty(K * B, H, W, 3, device=dev)
        C.call("ups_mask_parts_fwd", p(img), p(mh), p(parts_pm), B, P, K, 3, 1, st)
        x_parts = parts_pm.permute(0, 3, 1, 2)
        torch.backends.cudnn.benchmark = True
        torch.backends.cudnn.allow_tf32 = False
        calls["library: cuDNN conv2d fwd on injected [B,P,F+K] (fp32)"] = (
            lambda: TF.conv2d(x_inj, w_nchw, bias, padding=1), n_pix * 4 * (F + K + Co))
        calls["library: cuDNN conv2d bwd (dgrad + wgrad) on injected"] = (
            lambda: (torch.ops.aten.convolution_backward(gy_nchw, x_inj, w_nchw, [Co], [1, 1], [1, 1], [1, 1], False,
                                                         [0, 0], 1, [True, True, True])),
            n_pix * 4 * (2 * (F + K) + Co))
        calls["library: cuDNN conv2d fwd on part images [K*B,P,3] (fp32)"] = (
            lambda: TF.conv2d(x_parts, we_nchw, be, padding=1), n_pix * 4 * K * (3 + Co))
        gpm_nchw = g_pm.permute(0, 3, 1, 2)
        calls["library: cuDNN conv2d bwd (dgrad + wgrad) on part images"] = (
            lambda: (torch.ops.aten.convolution_backward(gpm_nchw, x_parts, we_nchw, [Co], [1, 1], [1, 1], [1, 1], False,
                                                         [0, 0], 1, [True, True, True])),
            n_pix * 4 * K * (2 * 3 + Co))
    if a.once:
        for name, (fn, by) in calls.items():
            fn()
        torch.cuda.synchronize()
        return None
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    res = {}
    for name, (fn, by) in calls.items():
        try:
            for _ in range(3):
                fn()
            ts = []
            for _ in range(iters):
                flush.sum()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            res[name] = dict(ms=ms, algorithmic_bytes=by, gbs=by / (ms * 1e-3) / 1e9,
                             frac_of_measured_hbm=by / (ms * 1e-3) / 1e9 / peak)
        except Exception as e:  # a library leg that cannot run must not hide our own numbers
            res[name] = dict(error=repr(e)[:300])
    return dict(shape=dict(B=B, H=H, W=W, K=K, F=F, Co=Co), peak_gbs=peak,
                l2="flushed between timed calls (256 MB read)", calls=res)


def measure_folded_step(batch=256, size=128, parts=16, iters=10):
    """The whole step with BOTH first convolutions folded in (PartStep(first_conv=32, encoder_conv=32)): ms per
    forward+backward, CUDA events over `iters` steps after 3 warm-up steps."""
    import torch
    import ups_b200
    from ups_b200.configs import CUB_TPS
    from ups_b200.step import PartStep
    B, S, K, F, Co = batch, size, parts, 64, 32
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
    views = torch.rand(3, B, S, S, 3, device=dev, generator=g) * 2 - 1
    l0, l1, feat = rn(B, S, S, K), rn(B, S, S, K), rn(B, K, F)
    Vd, bd, Ve, be = rn(3, 3, F + K, Co) * 0.04, rn(Co) * 0.04, rn(3, 3, 3, Co) * 0.2, rn(Co) * 0.2
    g_h0, g_e0 = rn(B, S, S, Co), rn(K * B, S, S, Co)
    g_m0, g_m1 = rn(B, S, S, K), rn(B, S, S, K)
    prm = ups_b200.tps_parameters(2 * B, generator=torch.Generator().manual_seed(1234), device=dev, **CUB_TPS)
    coord, tv = ups_b200.make_input_tps_param(prm)
    step = PartStep(B, S, K, F, n_views=3, first_conv=Co, encoder_conv=Co, device=dev)

    def one():
        step.forward(views, coord, tv, l0, l1, feat, Vd, bd, Ve, be)
        step.backward(g_h0, g_e0, None, g_m0, g_m1)
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"ms_per_step": ms, "images_per_s": B / (ms * 1e-3),
            "what": "fwd+bwd of the step with the decoder's and the appearance encoder's first 3x3 convolutions computed "
                    "from the part assignment (h0 and e0 instead of inj and parts)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01d")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--parts", type=int, default=16)
    ap.add_argument("--no-library", action="store_true", help="skip the cuDNN comparison legs")
    ap.add_argument("--once", action="store_true", help="call every entry once and exit (for ncu)")
    a = ap.parse_args()
    out = measure(a.batch, a.size, a.parts, library=not a.no_library, once=a.once)
    if out is None:
        return
    for name, r in out["calls"].items():
        print(name, json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_inject_conv.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
