"""-m gpu: the fused forward+backward step (K1..K6) against the chained oracle, plus
size-independent properties at BASELINE.json's full sizes."""
import pytest
import torch

from oracle import step as OS
from util import PENN_TPS, assert_bitexact, assert_close, cuda, make_inputs, own_error_atol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ups():
    import ups_b200
    return ups_b200


def _run(B, S, K, F, V, use_tps=True, views_grad=False, ties=False, seed=0, tps=None):
    from ups_b200.step import PartStep
    kw = dict(tps=tps) if tps else {}
    inp = make_inputs(B, S, K, F, V, seed=seed, ties=ties, **kw)
    c = inp["cot"]
    cot_o = dict(c, g_warped=c["g_warped"] if views_grad else None)
    out_o, grad_o = OS.step_forward_backward([v for v in inp["views"]], inp["coord"], inp["t_vector"], inp["l0"],
                                             inp["l1"], inp["feat"], cot_o, use_tps=use_tps, views_grad=views_grad)
    step = PartStep(B, S, K, F, n_views=V, use_tps=use_tps, views_grad=views_grad)
    d = cuda(inp)
    out = step.forward(d["views"], d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"])
    gw = torch.stack(d["cot"]["g_warped"]) if views_grad else None
    grad = step.backward(d["cot"]["g_inj"], d["cot"]["g_parts"], d["cot"]["g_pooled"], d["cot"]["g_m0"],
                         d["cot"]["g_m1"], gw)
    torch.cuda.synchronize()
    # float64 values of the long sums (pooled, dfeat, dviews) from the fp32 forward's own masks / sample positions
    grad_o["_ref64"] = OS.reduction_refs_fp64([v for v in inp["views"]], inp["coord"], inp["t_vector"], out_o, cot_o,
                                              use_tps=use_tps, views_grad=views_grad)
    return step, out, grad, out_o, grad_o


def _check(out, grad, out_o, grad_o, use_tps=True, views_grad=False):
    if use_tps:
        for i, w in enumerate(out_o["warped"]):
            assert_bitexact(out["warped"][i], w, f"warped[{i}]")
    assert_bitexact(out["m0"], out_o["m0"], "m0")
    assert_bitexact(out["m1"], out_o["m1"], "m1")
    assert out["labels0"].dtype == torch.int64
    assert torch.equal(out["labels0"].cpu(), out_o["labels0"]), "labels0 must be bit-exact"
    assert_bitexact(out["parts"], out_o["parts"], "parts (part-major)")
    # every fp32 value and gradient: 1e-4 rel / 1e-5 abs (BASELINE.json north_star).  The three quantities that are sums
    # over up to H*W pixels are compared with their float64 value; their absolute tolerance adds twice the fp32
    # oracle's own measured distance from it (util.own_error_atol) -- no assumed sqrt(n) scaling.
    r64 = grad_o["_ref64"]
    assert_close(out["pooled"], r64["pooled"], "pooled", atol=own_error_atol(out_o["pooled"], r64["pooled"]))
    assert_close(out["inj"], out_o["inj"], "inj")
    for k in ("dl0", "dl1"):
        assert_close(grad[k], grad_o[k], k)
    assert_close(grad["dfeat"], r64["dfeat"], "dfeat", atol=own_error_atol(grad_o["dfeat"], r64["dfeat"]))
    if views_grad:
        for i, w in enumerate(grad_o["dviews"]):
            assert_close(grad["dviews"][i], r64["dviews"][i], f"dviews[{i}]", atol=own_error_atol(w, r64["dviews"][i]))


def test_config1_cub_b8(ups):
    """BASELINE.json configs[0]: CUB 128x128, K=16, batch 8, 3 views, fixed TPS seed."""
    step, out, grad, out_o, grad_o = _run(8, 128, 16, 64, 3)
    assert step.fused
    _check(out, grad, out_o, grad_o)


@pytest.mark.parametrize("B,S,K,F,V", [(1, 32, 8, 16, 3), (3, 96, 32, 32, 2), (2, 64, 16, 64, 2), (5, 16, 16, 64, 3),
                                       (2, 32, 8, 64, 3), (2, 32, 32, 64, 3)])
def test_fused_shapes(ups, B, S, K, F, V):
    step, out, grad, out_o, grad_o = _run(B, S, K, F, V, ties=True, seed=B + K, tps=PENN_TPS if V == 2 else None)
    assert step.fused
    _check(out, grad, out_o, grad_o)


@pytest.mark.parametrize("B,S,K,F", [(1, 20, 16, 10), (2, 32, 4, 8), (1, 18, 25, 64)])
def test_unfused_fallback_shapes(ups, B, S, K, F):
    """Shapes outside the fused kernels (feature sizes other than 16/32/64, pixel counts that are not a multiple of
    32): the generic kernels, same parity bar."""
    step, out, grad, out_o, grad_o = _run(B, S, K, F, 3, ties=True, seed=K)
    assert not step.fused and not step.Kp
    _check(out, grad, out_o, grad_o)


@pytest.mark.parametrize("B,S,K,F,V", [(2, 24, 25, 64, 3), (8, 128, 25, 64, 3), (3, 32, 12, 32, 2), (2, 64, 5, 16, 3),
                                       (2, 32, 31, 64, 2)])
def test_padded_part_counts(ups, B, S, K, F, V):
    """n_parts = 25 is what the reference ships (train_cub_subset_tps.yaml:132).  Part counts that are not a power of
    two run on the fused kernels of the next power of two (logits padded with -inf): same parity bar as the fused path,
    probabilities / masks / labels bit-exact."""
    step, out, grad, out_o, grad_o = _run(B, S, K, F, V, ties=True, seed=K, tps=PENN_TPS if V == 2 else None)
    assert step.fused and step.Kp in (8, 16, 32) and step.Kp >= K
    assert out["parts"].shape[0] == K * B and out["inj"].shape[-1] == F + K
    _check(out, grad, out_o, grad_o)


def test_padded_part_count_with_views_grad(ups):
    step, out, grad, out_o, grad_o = _run(2, 64, 25, 64, 3, views_grad=True, seed=8)
    assert step.Kp == 32
    _check(out, grad, out_o, grad_o, views_grad=True)


def test_views_grad_tps_backward(ups):
    step, out, grad, out_o, grad_o = _run(3, 64, 16, 64, 3, views_grad=True, seed=4)
    _check(out, grad, out_o, grad_o, views_grad=True)


def test_no_tps_deepfashion(ups):
    """DeepFashion SB_model48c has no TPS (deepfashion/code/SB_model48c/model.py:266-280)."""
    step, out, grad, out_o, grad_o = _run(2, 64, 16, 64, 2, use_tps=False, seed=6)
    _check(out, grad, out_o, grad_o, use_tps=False)


def test_optional_cotangents(ups):
    from ups_b200.step import PartStep
    B, S, K, F = 2, 32, 16, 64
    inp = make_inputs(B, S, K, F, 3, seed=9)
    cot = dict(inp["cot"], g_pooled=torch.zeros(B, K, 3), g_m0=torch.zeros(B, S, S, K), g_m1=torch.zeros(B, S, S, K),
               g_warped=None)
    _, grad_o = OS.step_forward_backward([v for v in inp["views"]], inp["coord"], inp["t_vector"], inp["l0"],
                                         inp["l1"], inp["feat"], cot)
    d = cuda(inp)
    step = PartStep(B, S, K, F)
    step.forward(d["views"], d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"])
    grad = step.backward(d["cot"]["g_inj"], d["cot"]["g_parts"])
    for k in ("dl0", "dl1", "dfeat"):
        assert_close(grad[k], grad_o[k], k)


def test_run_to_run_determinism(ups):
    from ups_b200.step import PartStep
    B, S, K, F = 4, 64, 16, 64
    d = cuda(make_inputs(B, S, K, F, 3, seed=1))
    step = PartStep(B, S, K, F)
    res = []
    for _ in range(2):
        step.forward(d["views"], d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"])
        g = step.backward(d["cot"]["g_inj"], d["cot"]["g_parts"], d["cot"]["g_pooled"], d["cot"]["g_m0"], d["cot"]["g_m1"])
        torch.cuda.synchronize()
        res.append({k: v.clone() for k, v in dict(g, pooled=step.pooled, inj=step.inj).items()})
    for k in res[0]:
        assert torch.equal(res[0][k], res[1][k]), f"{k} differs between two runs (no atomics on this path)"


def test_full_size_properties_cub_b256(ups):
    """BASELINE.json configs[1] (CUB 128x128, K=16, batch 256): too big for the CPU oracle, so
    check size-independent identities on the device, and exact agreement of a B=8 slice with a
    B=8 run (every op on the path is per-sample)."""
    from ups_b200.step import PartStep
    B, S, K, F = 256, 128, 16, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    views = torch.rand(3, B, S, S, 3, device="cuda", generator=g) * 2 - 1
    l0 = torch.randn(B, S, S, K, device="cuda", generator=g)
    l1 = torch.randn(B, S, S, K, device="cuda", generator=g)
    feat = torch.randn(B, K, F, device="cuda", generator=g)
    small = make_inputs(B, 8, K, F, 3, seed=0)          # only for 2B sets of TPS parameters
    coord, tv = small["coord"].cuda(), small["t_vector"].cuda()
    step = PartStep(B, S, K, F)
    out = step.forward(views, coord, tv, l0, l1, feat)
    g_inj = torch.randn(B, S, S, F + K, device="cuda", generator=g)
    g_parts = torch.randn(K * B, S, S, 3, device="cuda", generator=g)
    grad = step.backward(g_inj, g_parts)
    torch.cuda.synchronize()
    m0, m1, lab = out["m0"], out["m1"], out["labels0"]
    assert torch.equal(lab, torch.argmax(m0, 3))                       # labels = first argmax of probs
    assert (m0.sum(-1) - 1).abs().max() < 1e-6 and (m1.sum(-1) - 1).abs().max() < 1e-6
    mh0 = out["inj"][..., F:]
    assert torch.equal(mh0 != 0, m0 == m0.max(-1, keepdim=True).values)  # hard mask marks exactly the maxima
    sel = feat[torch.arange(B, device="cuda")[:, None, None], lab] * mh0.max(-1, keepdim=True).values
    single = (mh0 != 0).sum(-1) == 1
    assert torch.equal(out["inj"][..., :F][single], sel[single])       # one-hot rows: inj == mon*feat[label]
    parts = out["parts"].view(K, B, S, S, 3)
    mh1 = (m1 == m1.max(-1, keepdim=True).values)
    assert torch.equal(parts != 0, (mh1.permute(3, 0, 1, 2)[..., None] & (out["warped"][1] != 0)[None]))
    assert (grad["dl0"].sum(-1).abs().max() < 1e-3) and (grad["dl1"].sum(-1).abs().max() < 1e-3)  # softmax-bwd rows sum to 0
    # per-sample independence: rerun the first 8 samples alone, results must be bit-identical
    s8 = PartStep(8, S, K, F)
    c8 = torch.cat([coord[:8], coord[B:B + 8]])
    t8 = torch.cat([tv[:8], tv[B:B + 8]])
    o8 = s8.forward(views[:, :8].contiguous(), c8, t8, l0[:8].contiguous(), l1[:8].contiguous(), feat[:8].contiguous())
    g8 = s8.backward(g_inj[:8].contiguous(), g_parts.view(K, B, S, S, 3)[:, :8].reshape(K * 8, S, S, 3).contiguous())
    torch.cuda.synchronize()
    for k in ("m0", "m1", "labels0", "inj"):
        assert torch.equal(o8[k], out[k][:8]), k
    assert torch.equal(o8["warped"][0], out["warped"][0][:8])
    assert torch.equal(g8["dl0"], grad["dl0"][:8]) and torch.equal(g8["dl1"], grad["dl1"][:8])
    # the number of pixel splits (hence the fp32 summation order of ~1000 O(1) terms per entry)
    # depends on B: entries agree to summation rounding, ~1e-6 of the column scale (~30)
    assert_close(g8["dfeat"], grad["dfeat"][:8].cpu(), "dfeat slice", atol=1e-4)
    assert_close(o8["pooled"], out["pooled"][:8].cpu(), "pooled slice")


@pytest.mark.parametrize("B,S,K,seed", [(2, 32, 16, 0), (3, 64, 16, 1), (1, 128, 16, 2), (2, 32, 32, 3), (40, 64, 16, 4),
                                        (100, 64, 16, 5), (30, 128, 16, 6), (90, 64, 32, 7), (1, 16, 16, 8)])
def test_decode_bwd_tensor_core_path(ups, B, S, K, seed):
    """K4 on tcgen05 (3xTF32, TMEM accumulators) against the oracle and against the SIMT kernel."""
    from oracle import parts as OP
    from ups_b200 import _cabi as C
    F, P = 64, S * S
    g = torch.Generator().manual_seed(seed)
    l0 = torch.randn(B, S, S, K, generator=g)
    l0[:, 0] = torch.round(l0[:, 0])                      # tied maxima -> several non-zeros in the hard mask
    feat = torch.randn(B, K, F, generator=g)
    g_inj = torch.randn(B, S, S, F + K, generator=g)
    g_m0 = torch.randn(B, S, S, K, generator=g)
    lo = l0.clone().requires_grad_(True)
    fo = feat.clone().requires_grad_(True)
    m0 = OP.softmax(lo)
    inj = OP.inject(fo, OP.straight_through_estimator(OP.hard_max(m0, 3), m0))
    dl0_o, dfeat_o = torch.autograd.grad([inj, m0], [lo, fo], [g_inj, g_m0])
    # float64 value of dfeat from the fp32 hard mask (the last K channels of inj): the bound of util.own_error_atol
    dfeat64 = torch.einsum("bpk,bpf->bkf", inj.detach()[..., F:].double().reshape(B, P, K),
                           g_inj[..., :F].double().reshape(B, P, F))
    m0c, featc, g_injc, g_m0c = m0.detach().cuda(), feat.cuda(), g_inj.cuda(), g_m0.cuda()
    ws = torch.empty(C.workspace_bytes(C.OP_STEP, B, P, K, F), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for name in ("ups_step_decode_bwd_tc", "ups_step_decode_bwd"):
        dl0 = torch.full((B, S, S, K), float("nan"), device="cuda")
        dfeat = torch.full((B, K, F), float("nan"), device="cuda")
        C.call(name, g_injc.data_ptr(), m0c.data_ptr(), g_m0c.data_ptr(), featc.data_ptr(), dl0.data_ptr(),
               dfeat.data_ptr(), B, P, K, F, ws.data_ptr(), ws.numel(), st)
        torch.cuda.synchronize()
        res[name] = (dl0, dfeat)
        assert_close(dl0, dl0_o, f"{name} dl0")
        assert_close(dfeat, dfeat64, f"{name} dfeat", atol=own_error_atol(dfeat_o, dfeat64))
    # without the external cotangent
    dl0 = torch.empty(B, S, S, K, device="cuda")
    dfeat = torch.empty(B, K, F, device="cuda")
    C.call("ups_step_decode_bwd_tc", g_injc.data_ptr(), m0c.data_ptr(), None, featc.data_ptr(), dl0.data_ptr(),
           dfeat.data_ptr(), B, P, K, F, ws.data_ptr(), ws.numel(), st)
    m0b = OP.softmax(lo)
    (dl0_o2,) = torch.autograd.grad(OP.inject(feat, OP.straight_through_estimator(OP.hard_max(m0b, 3), m0b)), lo, g_inj)
    assert_close(dl0, dl0_o2, "tc dl0 (no g_m0)")


@pytest.mark.parametrize("B,S,K,F,Co", [(2, 64, 16, 64, 32), (3, 32, 25, 16, 16), (1, 48, 8, 32, 64)])
def test_first_conv_step(ups, B, S, K, F, Co):
    """PartStep(first_conv=Co) ends the decode side in the decoder's first convolution
    (cub/code/SB_model48i/model.py:96,485) without forming `inj`; everything else is the same step."""
    import math
    from oracle import inject_conv as IC
    from oracle import parts as OP
    from ups_b200.step import PartStep
    inp = make_inputs(B, S, K, F, 3, seed=B + Co, ties=True)
    g = torch.Generator().manual_seed(Co)
    stdv = math.sqrt(1.0 / ((F + K) * 9))
    V = (torch.rand(3, 3, F + K, Co, generator=g) * 2 - 1) * stdv
    b = (torch.rand(Co, generator=g) * 2 - 1) * stdv
    g_h0 = torch.randn(B, S, S, Co, generator=g)
    c = inp["cot"]
    out_o, grad_o = OS.step_forward_backward([v for v in inp["views"]], inp["coord"], inp["t_vector"], inp["l0"],
                                             inp["l1"], inp["feat"], dict(c, g_warped=None))
    lo = [t.clone().requires_grad_(True) for t in (inp["l0"], inp["feat"], V, b)]
    m0_o = OP.softmax(lo[0])
    h0_o = IC.inject_conv2d(lo[1], OP.hard_max_straight_through(m0_o, 3), lo[2], lo[3])
    dl0_o, dfeat_o, dV_o, db_o = torch.autograd.grad([h0_o, m0_o], lo, [g_h0, c["g_m0"]])
    step = PartStep(B, S, K, F, n_views=3, first_conv=Co)
    d = cuda(inp)
    out = step.forward(d["views"], d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"], V.cuda(), b.cuda())
    grad = step.backward(g_h0.cuda(), d["cot"]["g_parts"], d["cot"]["g_pooled"], d["cot"]["g_m0"], d["cot"]["g_m1"])
    torch.cuda.synchronize()
    assert "inj" not in out
    assert_bitexact(out["m0"], out_o["m0"], "m0")
    assert torch.equal(out["labels0"].cpu(), out_o["labels0"])
    assert_bitexact(out["parts"], out_o["parts"], "parts")
    assert_close(out["h0"], h0_o.detach(), "h0")
    assert_close(grad["dl1"], grad_o["dl1"], "dl1")
    assert_close(grad["dl0"], dl0_o, "dl0")
    # dfeat / dV / db are sums over up to B*S*S*9 products: float64 autograd of the same convolution on the fp32 hard mask
    # gives their exact value; tolerance = north star + twice the fp32 oracle's own distance from it
    mask32 = OP.hard_max_straight_through(m0_o, 3).detach().double()
    x64 = [t.detach().double().requires_grad_(True) for t in (inp["feat"], V, b)]
    g64 = torch.autograd.grad(IC.inject_conv2d(x64[0], mask32, x64[1], x64[2]), x64, g_h0.double())
    for name, got, o32, o64 in (("dfeat", grad["dfeat"], dfeat_o, g64[0]), ("dV", grad["dV"], dV_o, g64[1]),
                                ("db", grad["db"], db_o, g64[2])):
        assert_close(got, o64, name, atol=own_error_atol(o32, o64))


@pytest.mark.parametrize("B,S,K,F,V", [(4, 128, 16, 64, 3), (3, 96, 8, 16, 2), (2, 64, 32, 64, 3)])
def test_fused_forward_launch_equals_two_kernels(ups, B, S, K, F, V):
    """K1 + K3 in one launch (ups_step_warp_decode_fwd) is bit-identical to K1 and K3 as two kernels."""
    from ups_b200.step import PartStep
    d = cuda(make_inputs(B, S, K, F, V, seed=5, ties=True))
    one = PartStep(B, S, K, F, n_views=V)
    assert one.fuse_fwd
    a = one.forward(d["views"], d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"])
    two = PartStep(B, S, K, F, n_views=V)
    two.forward_warp(d["views"], d["coord"], d["t_vector"])
    b = two.forward_parts(d["l0"], d["l1"], d["feat"])
    torch.cuda.synchronize()
    for k in ("warped", "m0", "m1", "labels0", "parts", "pooled", "inj"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("B,S,K,F,V", [(2, 256, 16, 64, 2), (1, 512, 32, 64, 2), (1, 256, 32, 64, 3)])
def test_big_shapes_vs_oracle(ups, B, S, K, F, V):
    """The resolutions of BASELINE.json configs[2] (DeepFashion 256x256) and of the sweep's large corner (512 px, K=32)
    against the chained oracle: same bar as config 1 (labels and masks bit-exact, fp32 1e-4 / 1e-5)."""
    step, out, grad, out_o, grad_o = _run(B, S, K, F, V, seed=S + K, tps=PENN_TPS if V == 2 else None)
    assert step.fused and step.decode_bwd == "tc"
    _check(out, grad, out_o, grad_o)


def test_full_size_pennaction_b512_slice_equality(ups):
    """BASELINE.json configs[3] (PennAction 128x128, K=16, batch 512, one warp per sample): too big for the CPU oracle
    as a whole, so (i) a slice of 4 samples spread over the batch is compared with the oracle and (ii) the same slice
    run alone (batch 4) gives bit-identical per-sample results (every op on the path is per-sample)."""
    from ups_b200.step import PartStep
    B, S, K, F, V = 512, 128, 16, 64, 2
    g = torch.Generator(device="cuda").manual_seed(3)
    views = torch.rand(V, B, S, S, 3, device="cuda", generator=g) * 2 - 1
    l0 = torch.randn(B, S, S, K, device="cuda", generator=g)
    l1 = torch.randn(B, S, S, K, device="cuda", generator=g)
    feat = torch.randn(B, K, F, device="cuda", generator=g)
    g_inj = torch.randn(B, S, S, F + K, device="cuda", generator=g)
    g_parts = torch.randn(K * B, S, S, 3, device="cuda", generator=g)
    prm = ups.tps_parameters(2 * B, generator=torch.Generator().manual_seed(7), device="cuda", **PENN_TPS)
    coord, tv = ups.make_input_tps_param(prm)
    step = PartStep(B, S, K, F, n_views=V)
    out = step.forward(views, coord, tv, l0, l1, feat)
    grad = step.backward(g_inj, g_parts)
    torch.cuda.synchronize()
    idx = torch.tensor([0, 171, 340, 511], device="cuda")
    n = idx.numel()
    rows = torch.cat([idx, idx + B])
    sl = dict(views=views[:, idx].contiguous(), coord=coord[rows].contiguous(), tv=tv[rows].contiguous(),
              l0=l0[idx].contiguous(), l1=l1[idx].contiguous(), feat=feat[idx].contiguous(), g_inj=g_inj[idx].contiguous(),
              g_parts=g_parts.view(K, B, S, S, 3)[:, idx].reshape(K * n, S, S, 3).contiguous())
    small = PartStep(n, S, K, F, n_views=V)
    o4 = small.forward(sl["views"], sl["coord"], sl["tv"], sl["l0"], sl["l1"], sl["feat"])
    g4 = small.backward(sl["g_inj"], sl["g_parts"])
    torch.cuda.synchronize()
    for k in ("m0", "m1", "labels0", "inj"):
        assert torch.equal(o4[k], out[k][idx]), k
    assert torch.equal(o4["warped"], out["warped"][:, idx])
    assert torch.equal(o4["parts"], out["parts"].view(K, B, S, S, 3)[:, idx].reshape(K * n, S, S, 3))
    assert torch.equal(g4["dl0"], grad["dl0"][idx]) and torch.equal(g4["dl1"], grad["dl1"][idx])
    # (i) the slice against the oracle
    c = {k: v.cpu() for k, v in sl.items()}
    cot = dict(g_inj=c["g_inj"], g_parts=c["g_parts"], g_pooled=torch.zeros(n, K, 3), g_m0=torch.zeros(n, S, S, K),
               g_m1=torch.zeros(n, S, S, K), g_warped=None)
    out_o, grad_o = OS.step_forward_backward([v for v in c["views"]], c["coord"], c["tv"], c["l0"], c["l1"], c["feat"], cot)
    r64 = OS.reduction_refs_fp64([v for v in c["views"]], c["coord"], c["tv"], out_o, cot)
    for i in range(V):
        assert_bitexact(out["warped"][i][idx], out_o["warped"][i], f"warped[{i}]")
    assert_bitexact(out["m0"][idx], out_o["m0"], "m0")
    assert torch.equal(out["labels0"][idx].cpu(), out_o["labels0"])
    assert_close(grad["dl0"][idx], grad_o["dl0"], "dl0")
    assert_close(grad["dl1"][idx], grad_o["dl1"], "dl1")
    assert_close(grad["dfeat"][idx], r64["dfeat"], "dfeat", atol=own_error_atol(grad_o["dfeat"], r64["dfeat"]))
    assert_close(out["pooled"][idx], r64["pooled"], "pooled", atol=own_error_atol(out_o["pooled"], r64["pooled"]))


def test_padded_part_count_pitched_inputs_are_zero_copy(ups):
    """A producer that writes logits / cotangents into PartStep.pitched_inputs() (row pitch Kp floats) gets the same bits
    as the contiguous-tensor path (whose kernels read the ragged rows in place), without the one pad copy left."""
    from ups_b200 import _cabi as C
    from ups_b200.step import PartStep
    B, S, K, F, V = 3, 64, 25, 64, 3
    d = cuda(make_inputs(B, S, K, F, V, seed=2))
    c = d["cot"]
    a = PartStep(B, S, K, F, n_views=V)
    oa = a.forward(d["views"], d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"])
    ga = a.backward(c["g_inj"], c["g_parts"], c["g_pooled"], c["g_m0"], c["g_m1"])
    b = PartStep(B, S, K, F, n_views=V)
    pin = b.pitched_inputs()
    for name, src in (("l0", d["l0"]), ("l1", d["l1"]), ("g_inj", c["g_inj"]), ("g_m0", c["g_m0"]), ("g_m1", c["g_m1"])):
        assert not pin[name].is_contiguous() and pin[name].shape == src.shape
        pin[name].copy_(src)
    C.launch_count_reset()
    ob = b.forward(d["views"], d["coord"], d["t_vector"], pin["l0"], pin["l1"], d["feat"])
    gb = b.backward(pin["g_inj"], c["g_parts"], c["g_pooled"], pin["g_m0"], pin["g_m1"])
    n_pitched = C.launch_count()
    torch.cuda.synchronize()
    for k in ("m0", "m1", "labels0", "parts", "pooled", "inj"):
        assert torch.equal(oa[k], ob[k]), k
    for k in ("dl0", "dl1", "dfeat"):
        assert torch.equal(ga[k], gb[k]), k
    C.launch_count_reset()
    a.forward(d["views"], d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"])
    a.backward(c["g_inj"], c["g_parts"], c["g_pooled"], c["g_m0"], c["g_m1"])
    # contiguous [.,K] logits and mask cotangents are read in place (row length K); only g_inj is re-pitched, because
    # the TMA tensor map of the decode backward needs a 16-byte row pitch
    assert C.launch_count() == n_pitched + 1, "the contiguous path adds exactly one pad copy (g_inj)"


@pytest.mark.parametrize("V", [3, 2])
def test_views_grad_without_external_cotangent(ups, V):
    """views_grad with g_warped = None: only view 1 receives a cotangent (dimg1 from the encode side); K6 forms it on
    the fly and skips the samples whose cotangent is identically zero."""
    from ups_b200.step import PartStep
    B, S, K, F = 2, 64, 16, 64
    inp = make_inputs(B, S, K, F, V, seed=12, **(dict(tps=PENN_TPS) if V == 2 else {}))
    c = inp["cot"]
    cot_o = dict(c, g_warped=[torch.zeros(B, S, S, 3) for _ in range(V)])
    out_o, grad_o = OS.step_forward_backward([v for v in inp["views"]], inp["coord"], inp["t_vector"], inp["l0"], inp["l1"],
                                             inp["feat"], cot_o, views_grad=True)
    r64 = OS.reduction_refs_fp64([v for v in inp["views"]], inp["coord"], inp["t_vector"], out_o, cot_o, views_grad=True)
    d = cuda(inp)
    step = PartStep(B, S, K, F, n_views=V, views_grad=True)
    step.forward(d["views"], d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"])
    grad = step.backward(d["cot"]["g_inj"], d["cot"]["g_parts"], d["cot"]["g_pooled"], d["cot"]["g_m0"], d["cot"]["g_m1"], None)
    torch.cuda.synchronize()
    for i in range(V):
        assert_close(grad["dviews"][i], r64["dviews"][i], f"dviews[{i}]", atol=own_error_atol(grad_o["dviews"][i], r64["dviews"][i]))
    assert float(grad["dviews"][0].abs().max()) == 0.0


@pytest.mark.parametrize("B,S,K,F,Co,Ce", [(2, 128, 16, 64, 32, 32), (2, 32, 8, 16, 16, 8)])
def test_first_conv_on_both_sides(ups, B, S, K, F, Co, Ce):
    """SURVEY.md 8f N4, both halves in one step: the decode side ends in `dd`'s first convolution (h0), the encode side in
    `e_alpha`'s first convolution on the part images (e0 [K*B,S,S,Ce], model.py:40,478); neither `inj` nor `parts` exists."""
    import math
    from oracle import inject_conv as IC
    from oracle import parts as OP
    from oracle import parts_conv as PC
    from oracle import tps as OT
    from ups_b200.step import PartStep
    inp = make_inputs(B, S, K, F, 3, seed=Co + Ce)
    g = torch.Generator().manual_seed(Ce)
    Vd = (torch.rand(3, 3, F + K, Co, generator=g) * 2 - 1) * math.sqrt(1.0 / ((F + K) * 9))
    bd = (torch.rand(Co, generator=g) * 2 - 1) * 0.1
    Ve = (torch.rand(3, 3, 3, Ce, generator=g) * 2 - 1) * math.sqrt(1.0 / 27)
    be = (torch.rand(Ce, generator=g) * 2 - 1) * 0.1
    g_h0 = torch.randn(B, S, S, Co, generator=g)
    g_e0 = torch.randn(K * B, S, S, Ce, generator=g)
    c = inp["cot"]
    # oracle
    warped = OT.make_tps_given([v for v in inp["views"]], inp["coord"], inp["t_vector"])
    xs = [t.clone().requires_grad_(True) for t in (inp["l0"], inp["l1"], inp["feat"], Vd, bd, Ve, be)]
    m0_o, m1_o = OP.softmax(xs[0]), OP.softmax(xs[1])
    h0_o = IC.inject_conv2d(xs[2], OP.hard_max_straight_through(m0_o, 3), xs[3], xs[4])
    mh1_o = OP.hard_max_straight_through(m1_o, 3)
    e0_o = PC.parts_conv2d(warped[1], mh1_o, xs[5], xs[6])
    grads = torch.autograd.grad([h0_o, e0_o, m0_o, m1_o], xs, [g_h0, g_e0, c["g_m0"], c["g_m1"]])
    # float64 values of the batch-summed filter gradients on the fp32 hard masks
    x64 = [t.detach().double().requires_grad_(True) for t in (inp["feat"], Vd, bd, Ve, be)]
    h64 = IC.inject_conv2d(x64[0], OP.hard_max_straight_through(m0_o, 3).detach().double(), x64[1], x64[2])
    e64 = PC.parts_conv2d(warped[1].double(), mh1_o.detach().double(), x64[3], x64[4])
    g64 = torch.autograd.grad([h64, e64], x64, [g_h0.double(), g_e0.double()])
    d = cuda(inp)
    step = PartStep(B, S, K, F, n_views=3, first_conv=Co, encoder_conv=Ce)
    out = step.forward(d["views"], d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"], Vd.cuda(), bd.cuda(), Ve.cuda(),
                       be.cuda())
    grad = step.backward(g_h0.cuda(), g_e0.cuda(), None, d["cot"]["g_m0"], d["cot"]["g_m1"])
    torch.cuda.synchronize()
    assert "inj" not in out and "parts" not in out
    assert_bitexact(out["m1"], m1_o.detach(), "m1")
    assert_close(out["h0"], h0_o.detach(), "h0")
    assert_close(out["e0"], e0_o.detach(), "e0")
    assert_close(grad["dl0"], grads[0], "dl0")
    assert_close(grad["dl1"], grads[1], "dl1")
    for name, got, o32, o64 in (("dfeat", grad["dfeat"], grads[2], g64[0]), ("dV", grad["dV"], grads[3], g64[1]),
                                ("db", grad["db"], grads[4], g64[2]), ("dVe", grad["dVe"], grads[5], g64[3]),
                                ("dbe", grad["dbe"], grads[6], g64[4])):
        assert_close(got, o64, name, atol=own_error_atol(o32, o64))
