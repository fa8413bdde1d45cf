#!/usr/bin/env python
"""BASELINE.json configs[4]: part-count / resolution sweep K in {8,16,32} x S in {128,256,512}
(F=64, V=2), timing the fused step with the CUDA-core K4 and, where it applies (K in {16,32}),
the TMA + tcgen05 K4; plus K=25 (the reference's shipped n_parts, train_cub_subset_tps.yaml:132), which is not a
power of two and runs padded to 32 on the fused kernels (and, for comparison, on the generic unfused kernels).  Prints one JSON line per cell and a markdown table.

    python scripts/sweep.py [--out gpurun_out/sweep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import ups_b200
    from ups_b200 import _cabi as C
    from ups_b200.step import PartStep
    from util import PENN_TPS

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--parts", type=int, nargs="*", default=[8, 16, 25, 32], help="part counts to run")
    ap.add_argument("--no-generic", action="store_true", help="skip the unfused K=25 variant (40 ms per step)")
    args = ap.parse_args()
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda", 0)
    F, V = 64, 2
    rows = []
    for S, B in ((128, 64), (256, 16), (512, 4)):      # SURVEY.md 8d config 5 batch sizes ...
        B *= 4                                           # ... x4 so that one step exceeds the 126 MB L2
        for K in args.parts:
            g = torch.Generator(device=dev).manual_seed(0)
            views = torch.rand(V, B, S, S, 3, device=dev, generator=g) * 2 - 1
            l0 = torch.randn(B, S, S, K, device=dev, generator=g)
            l1 = torch.randn(B, S, S, K, device=dev, generator=g)
            feat = torch.randn(B, K, F, device=dev, generator=g)
            g_inj = torch.randn(B, S, S, F + K, device=dev, generator=g)
            g_parts = torch.randn(K * B, S, S, 3, device=dev, generator=g)
            g_pooled = torch.randn(B, K, 3, device=dev, generator=g)
            g_m0 = torch.randn(B, S, S, K, device=dev, generator=g)
            g_m1 = torch.randn(B, S, S, K, device=dev, generator=g)
            prm = ups_b200.tps_parameters(2 * B, generator=torch.Generator().manual_seed(1234), device=dev, **PENN_TPS)
            coord, tv = ups_b200.make_input_tps_param(prm)
            for variant in (("padded", "pitched", "generic") if K == 25 else ("simt", "tc")):
                if (variant == "tc" and K == 8) or (variant == "generic" and args.no_generic):
                    continue
                a_l0, a_l1, a_g_inj, a_g_m0, a_g_m1 = l0, l1, g_inj, g_m0, g_m1
                # K = 25: "padded" = the fused kernels of Kp = 32, contiguous [.,25] inputs: logits and mask cotangents read
                # in place (row length 25, parts >= 25 are -inf / 0), g_inj copied into its row-pitched buffer (default); "pitched" = the producer writes into PartStep.pitched_inputs() (no copies);
                # "generic" = the unfused kernels
                os.environ["UPS_PAD_K"] = "0" if variant == "generic" else "1"
                step = PartStep(B, S, K, F, n_views=V, decode_bwd=variant if variant in ("simt", "tc") else "auto", device=dev)
                assert step.fused == (variant != "generic") and bool(step.Kp) == (variant in ("padded", "pitched"))
                a_l0, a_l1, a_g_inj, a_g_m0, a_g_m1 = l0, l1, g_inj, g_m0, g_m1
                if variant == "pitched":
                    pin = step.pitched_inputs()
                    for nm, src in (("l0", l0), ("l1", l1), ("g_inj", g_inj), ("g_m0", g_m0), ("g_m1", g_m1)):
                        pin[nm].copy_(src)
                    a_l0, a_l1, a_g_inj, a_g_m0, a_g_m1 = pin["l0"], pin["l1"], pin["g_inj"], pin["g_m0"], pin["g_m1"]

                def one():
                    step.forward(views, coord, tv, a_l0, a_l1, feat)
                    step.backward(a_g_inj, g_parts, g_pooled, a_g_m0, a_g_m1)
                for _ in range(3):
                    one()
                torch.cuda.synchronize()
                marks = []
                raw = C.call

                def timed(name, *a):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); raw(name, *a); e1.record()
                    marks.append((name, e0, e1))
                ups_b200.step.C.call = timed
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(args.steps):
                    one()
                t1.record()
                torch.cuda.synchronize()
                ups_b200.step.C.call = raw
                ms = t0.elapsed_time(t1) / args.steps
                k4 = [e0.elapsed_time(e1) for n, e0, e1 in marks
                      if n.startswith("ups_step_decode_bwd") or n in ("ups_part_inject_bwd",)]
                bytes_img = step.algorithmic_bytes_per_image()
                row = dict(S=S, K=K, B=B, k4=variant, ms_per_step=ms, images_per_s=B / (ms * 1e-3),
                           step_frac_of_measured_hbm=bytes_img * B / (ms * 1e-3) / 1e9 / peak,
                           k4_ms=sum(k4) / len(k4),
                           k4_gbs=4 * ((F + K) + 3 * K) * B * S * S / (sum(k4) / len(k4) * 1e-3) / 1e9)
                rows.append(row)
                print(json.dumps(row), flush=True)
                del step
            del views, l0, l1, feat, g_inj, g_parts, g_m0, g_m1
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(dict(peak_gbs=peak, rows=rows), open(args.out, "w"), indent=1)
    print("\n| S | K | B | K4 variant | img/s | step frac of measured HBM | K4 ms | K4 GB/s (algorithmic) |\n|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['S']} | {r['K']} | {r['B']} | {r['k4']} | {r['images_per_s']:.0f} | {r['step_frac_of_measured_hbm']:.3f} | "
              f"{r['k4_ms']:.3f} | {r['k4_gbs']:.0f} |")


if __name__ == "__main__":
    main()
