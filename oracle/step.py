"""TEST INFRASTRUCTURE — the per-step part-disentanglement path chained as the reference
chains it (cub/code/SB_model48i/model.py:337,426-485), forward + autograd backward.

Everything between the path's pieces that is NOT the path (the CNN encoders e_pi/e_alpha,
the mask decoder dv, the hourglass dd, the losses) is replaced by external tensors:
logits l0/l1 stand for dv's output, `feat` for e_alpha's output, and fixed cotangents for
what the losses/decoder would send back.  Also the CPU baseline that
bench.py times ("port").
"""
import torch

from . import parts as P
from . import tps as T


def step_forward(views, coord, t_vector, l0, l1, feat, use_tps=True):
    """views: list of V tensors [B,S,S,3] (V=3 CUB, 2 PennAction; use_tps=False DeepFashion).
    coord/t_vector: [2B,8,2] (make_input_tps_param output).  Returns dict of path outputs."""
    B = l0.shape[0]
    if use_tps:
        warped = T.make_tps_given(views, coord, t_vector)
    else:
        warped = list(views)
    m0 = P.softmax(l0)                                            # model.py:426
    m1 = P.softmax(l1)                                            # model.py:429
    m0_hard = P.straight_through_estimator(P.hard_max(m0, 3), m0)  # model.py:434-436
    m1_hard = P.straight_through_estimator(P.hard_max(m1, 3), m1)  # model.py:453-455
    labels0 = P.argmax_labels(m0)                                 # model.py:447,470
    view1_parts = P.mask_parts(warped[1], m1_hard)                # model.py:478
    K = l0.shape[3]
    # part-major fold of apply_partwise (nn.py:100-103): X[k*B+b] = parts[b,:,:,k,:]
    parts_pm = view1_parts.permute(3, 0, 1, 2, 4).reshape(K * B, *view1_parts.shape[1:3], -1)
    pooled = P.part_mean_pool(warped[1], m1_hard)                 # model.py:50-52 tail
    inj = P.inject(feat, m0_hard)                                 # model.py:482-484
    return dict(warped=warped, m0=m0, m1=m1, labels0=labels0, parts=parts_pm,
                pooled=pooled, inj=inj)


def step_forward_backward(views, coord, t_vector, l0, l1, feat, cot, use_tps=True,
                          views_grad=False):
    """cot: dict with g_inj, g_parts (part-major), g_pooled, g_m0, g_m1 [, g_warped list].
    Returns (forward outputs, dict(dl0, dl1, dfeat [, dviews]))."""
    l0 = l0.detach().requires_grad_(True)
    l1 = l1.detach().requires_grad_(True)
    feat = feat.detach().requires_grad_(True)
    if views_grad:
        views = [v.detach().requires_grad_(True) for v in views]
    out = step_forward(views, coord, t_vector, l0, l1, feat, use_tps=use_tps)
    outs = [out["inj"], out["parts"], out["pooled"], out["m0"], out["m1"]]
    gs = [cot["g_inj"], cot["g_parts"], cot["g_pooled"], cot["g_m0"], cot["g_m1"]]
    ins = [l0, l1, feat]
    if views_grad:
        ins += list(views)
        if cot.get("g_warped") is not None:
            outs += list(out["warped"])
            gs += list(cot["g_warped"])
    grads = torch.autograd.grad(outs, ins, gs, allow_unused=True)
    res = dict(dl0=grads[0], dl1=grads[1], dfeat=grads[2])
    if views_grad:
        res["dviews"] = [g if g is not None else torch.zeros_like(v)
                         for g, v in zip(grads[3:], views)]
    out = {k: ([t.detach() for t in v] if isinstance(v, list) else v.detach())
           for k, v in out.items()}
    return out, res


def reduction_refs_fp64(views, coord, t_vector, out, cot, use_tps=True, views_grad=False):
    """float64 values of the path's long sums, evaluated from the fp32 forward's own masks and sample positions (so the
    discrete decisions - argmax, floor - are the fp32 ones): pooled, dfeat and, with views_grad, dviews.  Tests bound
    the kernels by `1e-5 + 2 * |fp32 oracle - this|` (tests/util.py::own_error_atol) instead of an assumed sqrt(n) law.
      pooled[b,k,c] = mean_p parts[k*B+b, p, c]                     (model.py:50-52 tail on mask_parts' output)
      dfeat[b,k,f]  = sum_p m0_hard[b,p,k] * g_inj[b,p,f]           (autodiff of unpool_features, model.py:225-249)
      dviews        = scatter of the warped views' cotangents       (autodiff of _interpolate, transformations.py:114-169)
    """
    inj, parts = out["inj"].double(), out["parts"].double()
    B, S = inj.shape[0], inj.shape[1]
    K = out["m0"].shape[-1]
    F = inj.shape[-1] - K
    mh0 = inj[..., F:].reshape(B, S * S, K)
    res = dict(pooled=parts.reshape(K, B, S * S, -1).mean(2).permute(1, 0, 2),
               dfeat=torch.einsum("bpk,bpf->bkf", mh0, cot["g_inj"].double()[..., :F].reshape(B, S * S, F)))
    if views_grad:
        img1 = (out["warped"][1] if use_tps else views[1]).double()
        # mask of view 1 from parts = img * mask is not recoverable where img == 0: recompute it the oracle's way
        m1_hard = P.straight_through_estimator(P.hard_max(out["m1"], 3), out["m1"]).double()          # [B,S,S,K]
        gp = cot["g_parts"].double().reshape(K, B, S, S, -1).permute(1, 2, 3, 0, 4)                   # [B,S,S,K,3]
        if cot.get("g_pooled") is not None:
            gp = gp + cot["g_pooled"].double()[:, None, None] / (S * S)
        dimg1 = (m1_hard[..., None] * gp).sum(3)
        V = len(views)
        gw = [cot["g_warped"][i].double() if cot.get("g_warped") is not None else torch.zeros_like(img1) for i in range(V)]
        gw[1] = gw[1] + dimg1
        if use_tps:
            G01 = torch.cat(gw[:2], 0)
            U01 = torch.cat([v.detach() for v in views[:2]], 0)
            d01 = T.warp_grad_fp64(U01, coord, t_vector, S, G01)
            dv = list(torch.split(d01, B, 0))
            if V > 2:
                dv.append(T.warp_grad_fp64(views[2].detach(), coord[:B], t_vector[:B], S, gw[2]))
            res["dviews"] = dv
        else:
            res["dviews"] = gw
    return res
