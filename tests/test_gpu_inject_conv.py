"""-m gpu: SURVEY.md 8f N4 — the decoder's first convolution on the part assignment
(cub/code/SB_model48i/model.py:482-485 + cub/code/nn.py:617-664) against the oracle and the committed
reference fixture, forward and all gradients, through the reference-named helper."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import inject_conv as IC
from oracle import parts as OP
from util import ATOL, assert_bitexact, assert_close, own_error_atol

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inject_conv.npz")


@pytest.fixture(scope="module")
def ups():
    import ups_b200
    return ups_b200


def _case(B, H, W, K, F, Co, kind, seed):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, H, W, K, generator=g)
    if kind == "ties":          # exact ties (several ones per pixel) and empty pixels among the hard ones
        logits[:, 0] = torch.round(logits[:, 0])
    p = OP.softmax(logits)
    mask = p if kind == "soft" else OP.hard_max_straight_through(p, 3)
    if kind == "ties":
        mask = mask.clone()
        mask[:, -1, ::2] = 0.0
    feat = torch.randn(B, K, F, generator=g)
    stdv = math.sqrt(1.0 / ((F + K) * 9))
    V = (torch.rand(3, 3, F + K, Co, generator=g) * 2 - 1) * stdv
    b = (torch.rand(Co, generator=g) * 2 - 1) * stdv
    gy = torch.randn(B, H, W, Co, generator=g)
    return logits, mask.detach(), feat, V, b, gy


def _sum_atol(n_terms, want32, want64):
    """Absolute tolerance for batch-summed gradients (dfeat, dV, db: sums over up to B*H*W products).  Entries that
    cancel to ~0 cannot agree to 1e-5 between two fp32 summation orders: the oracle's own fp32 result is that far
    from the fp64 value of the same expression.  So: the north-star 1e-5 plus twice the oracle's own measured fp32
    error (util.own_error_atol); no assumed growth with the number of terms."""
    return own_error_atol(want32, want64) if want64 is not None else ATOL


def _check_grads(got, want, H, W, B, want64=None, names=("dmask", "dfeat", "dV", "db")):
    n_terms = (9, H * W * 9, B * H * W, B * H * W)
    for i, (a, o, name, n) in enumerate(zip(got, want, names, n_terms)):
        w64 = None if want64 is None else want64[i]
        assert_close(a, o, name, atol=ATOL if i == 0 else _sum_atol(n, o, w64))


def _oracle_grads(fn, inputs, cots, dtype):
    xs = [t.detach().to(dtype).requires_grad_(True) for t in inputs]
    ys = fn(*xs)
    ys = ys if isinstance(ys, (list, tuple)) else [ys]
    return [y.detach() for y in ys], torch.autograd.grad(ys, xs, [c.to(dtype) for c in cots])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_reference_fixture(ups, tag):
    """Values and gradients produced by the reference's own unpool_features / _conv2d bodies (tests/golden/
    make_golden_inject_conv.py): hard mask with a tie and an empty pixel (a), the CUB sizes K=16 F=64 Co=32 (b),
    a soft mask with K=5 (c)."""
    g = np.load(GOLDEN)
    mask, feat, V, b = (torch.from_numpy(g[f"{tag}_{n}"]).cuda().requires_grad_(True) for n in ("mask", "feat", "V", "b"))
    y = ups.model.inject_conv2d(feat, mask, V, b)
    assert_close(y, torch.from_numpy(g[f"{tag}_out"]), "out")
    grads = torch.autograd.grad(y, [mask, feat, V, b], torch.from_numpy(g[f"{tag}_g_out"]).cuda())
    B, H, W, _ = mask.shape
    _check_grads(grads, [torch.from_numpy(g[f"{tag}_{n}"]) for n in ("dmask", "dfeat", "dV", "db")], H, W, B)


SHAPES = [(2, 64, 64, 16, 64, 32, "hard"), (1, 128, 128, 16, 64, 32, "hard"), (3, 40, 50, 25, 16, 32, "ties"),
          (2, 33, 31, 8, 8, 64, "hard"), (1, 16, 48, 32, 64, 64, "ties"), (1, 16, 48, 8, 64, 64, "hard"), (2, 20, 70, 4, 4, 4, "soft"),
          (1, 1, 1, 16, 64, 32, "hard"), (1, 3, 37, 5, 6, 8, "soft"), (5, 17, 16, 12, 32, 16, "ties"),
          (2, 64, 64, 16, 64, 32, "soft")]


@pytest.mark.parametrize("B,H,W,K,F,Co,kind", SHAPES)
def test_inject_conv2d_vs_oracle(ups, B, H, W, K, F, Co, kind):
    _, mask, feat, V, b, gy = _case(B, H, W, K, F, Co, kind, seed=B * 100 + K + Co)
    fn = lambda m, f, v, bb: IC.inject_conv2d(f, m, v, bb)  # noqa: E731
    (y_o,), g_o = _oracle_grads(fn, (mask, feat, V, b), [gy], torch.float32)
    _, g_o64 = _oracle_grads(fn, (mask, feat, V, b), [gy], torch.float64)
    lc = [t.cuda().requires_grad_(True) for t in (mask, feat, V, b)]
    y = ups.model.inject_conv2d(lc[1], lc[0], lc[2], lc[3])
    assert y.shape == (B, H, W, Co)
    assert_close(y, y_o.detach(), "out")
    # equals the un-fused path of this library too: conv over the materialised injected map
    inj = ups.model.inject_features(lc[1].detach(), lc[0].detach())
    assert_close(IC.conv2d_same(inj.cpu(), V, b), y_o.detach(), "conv(inject_features)")
    got = torch.autograd.grad(y, lc, gy.cuda())
    _check_grads(got, g_o, H, W, B, g_o64)


def test_table(ups):
    B, K, F, Co = 3, 16, 64, 32
    g = torch.Generator().manual_seed(5)
    feat, V = torch.randn(B, K, F, generator=g), torch.randn(3, 3, F + K, Co, generator=g)
    G = ups.ops.inject_conv_table(feat.cuda(), V.reshape(9, F + K, Co).cuda())
    G32, G64 = IC.inject_conv_table(feat, V), IC.inject_conv_table(feat.double(), V.double())
    assert_close(G, G64, "G", atol=own_error_atol(G32, G64))


@pytest.mark.parametrize("B,H,W,K,F,Co", [(2, 64, 64, 16, 64, 32), (1, 48, 40, 8, 16, 32), (2, 32, 32, 25, 8, 16)])
def test_decode_conv2d_vs_oracle(ups, B, H, W, K, F, Co):
    """softmax -> argmax -> ST(hard_max) -> inject -> conv, and its backward with cotangents on both the conv output
    and the probabilities (model.py:426,434-436,447,470-473,482-485)."""
    logits, _, feat, V, b, gy = _case(B, H, W, K, F, Co, "ties", seed=K + Co)
    g = torch.Generator().manual_seed(3)
    gm = torch.randn(B, H, W, K, generator=g)
    def fn(l, f, v, bb):
        m = OP.softmax(l)
        mh_ = OP.hard_max_straight_through(m, 3)
        return IC.inject_conv2d(f, mh_, v, bb), m, mh_

    (y_o, m0_o), g_o = _oracle_grads(lambda *a: fn(*a)[:2], (logits, feat, V, b), [gy, gm], torch.float32)
    mh_o = fn(logits, feat, V, b)[2]
    # the fp64 value of the same expression, on the fp32 oracle's own hard assignment
    def fn64(l, f, v, bb):
        m = torch.softmax(l, -1)
        return IC.inject_conv2d(f, (mh_o.double() - m).detach() + m, v, bb), m

    _, g_o64 = _oracle_grads(fn64, (logits, feat, V, b), [gy, gm], torch.float64)
    lc = [t.cuda().requires_grad_(True) for t in (logits, feat, V, b)]
    m0, labels, mh, y = ups.model.decode_conv2d(*lc)
    assert_bitexact(m0, m0_o.detach(), "m0")
    assert_bitexact(mh, mh_o.detach(), "mask")
    assert torch.equal(labels.cpu(), OP.argmax_labels(m0_o.detach()))
    assert_close(y, y_o.detach(), "out")
    got = torch.autograd.grad([y, m0], lc, [gy.cuda(), gm.cuda()])
    assert_close(got[0], g_o[0], "dlogits")
    _check_grads(got, g_o, H, W, B, g_o64, names=("dlogits", "dfeat", "dV", "db"))


def test_run_to_run_determinism(ups):
    _, mask, feat, V, b, gy = _case(4, 64, 64, 16, 64, 32, "ties", seed=11)
    outs = []
    for _ in range(2):
        lc = [t.cuda().requires_grad_(True) for t in (mask, feat, V, b)]
        y = ups.model.inject_conv2d(lc[1], lc[0], lc[2], lc[3])
        outs.append([y] + list(torch.autograd.grad(y, lc, gy.cuda())))
    for a, c in zip(*outs):
        assert_bitexact(a, c.cpu(), "run-to-run")


def test_full_size_linearity(ups):
    """CUB B=64 slice of BASELINE configs[1] (the oracle would take minutes): out(feat_a + feat_b) - bias terms is
    additive in the features, and the hard-mask conv equals a gather of table rows."""
    B, H, W, K, F, Co = 64, 128, 128, 16, 64, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn(B, H, W, K, device="cuda", generator=g)
    mh = ups.nn.hard_max_straight_through(ups.nn.softmax(logits), 3)
    fa, fb = torch.randn(B, K, F, device="cuda", generator=g), torch.randn(B, K, F, device="cuda", generator=g)
    V = torch.randn(3, 3, F + K, Co, device="cuda", generator=g) * 0.05
    b = torch.randn(Co, device="cuda", generator=g)
    zero_b = torch.zeros_like(b)
    ya = ups.model.inject_conv2d(fa, mh, V, b)
    yb = ups.model.inject_conv2d(fb, mh, V, zero_b)
    Vm = V.clone()
    Vm[:, :, F:] = 0                                    # drop the mask channels so that they are not counted twice
    y0 = ups.model.inject_conv2d(fb * 0, mh, V - Vm, zero_b)
    yab = ups.model.inject_conv2d(fa + fb, mh, V, b)
    assert_close(yab, (ya + yb - y0).cpu(), "additivity in feat", rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------- encoder side: conv on the masked part images
PC_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "parts_conv.npz")


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_parts_conv_reference_fixture(ups, tag):
    g = np.load(PC_GOLDEN)
    mask, image, V, b = (torch.from_numpy(g[f"{tag}_{n}"]).cuda() for n in ("mask", "image", "V", "b"))
    mask, V, b = mask.requires_grad_(True), V.requires_grad_(True), b.requires_grad_(True)
    y = ups.model.parts_conv2d(image, mask, V, b)
    assert_close(y, torch.from_numpy(g[f"{tag}_out"]), "out")
    if V.shape[-1] in (8, 16, 32, 64):
        got = torch.autograd.grad(y, [mask, V, b], torch.from_numpy(g[f"{tag}_g_out"]).cuda())
        # the fixture holds the reference's fp32 values; the float64 value of the same gradients (oracle autograd in
        # double on the same inputs) gives the measured fp32 error of the reference order, hence the tolerance
        from oracle import parts_conv as PC
        x64 = [torch.from_numpy(g[f"{tag}_{n}"]).double().requires_grad_(True) for n in ("mask", "V", "b")]
        y64 = PC.parts_conv2d(torch.from_numpy(g[f"{tag}_image"]).double(), *x64)
        g64 = torch.autograd.grad(y64, x64, torch.from_numpy(g[f"{tag}_g_out"]).double())
        for a, name, o64 in zip(got, ("dmask", "dV", "db"), g64):
            ref32 = torch.from_numpy(g[f"{tag}_{name}"])
            assert_close(a, o64, name, atol=own_error_atol(ref32, o64))


@pytest.mark.parametrize("B,H,W,K,Co,kind", [(2, 64, 64, 16, 32, "hard"), (1, 128, 128, 16, 32, "ties"),
                                             (3, 40, 50, 25, 32, "ties"), (2, 33, 31, 8, 64, "soft"),
                                             (1, 16, 160, 4, 16, "hard"), (2, 9, 70, 5, 8, "soft"), (1, 1, 1, 16, 32, "hard"),
                                             # Co = 32 and W in {128, 256}: the tcgen05 kernel (parts_conv_bwd_tc.cu)
                                             (3, 33, 128, 5, 32, "ties"), (1, 40, 256, 8, 32, "soft"), (2, 1, 128, 3, 32, "hard"),
                                             (5, 7, 256, 32, 32, "hard")])
def test_parts_conv2d_backward_vs_oracle(ups, B, H, W, K, Co, kind):
    from oracle import parts_conv as PC
    _, mask, _, _, _, _ = _case(B, H, W, K, 4, Co, kind, seed=B * 10 + K + 1)
    g = torch.Generator().manual_seed(Co + 1)
    image = torch.rand(B, H, W, 3, generator=g) * 2 - 1
    V = (torch.rand(3, 3, 3, Co, generator=g) * 2 - 1) * math.sqrt(1.0 / 27)
    b = (torch.rand(Co, generator=g) * 2 - 1) * math.sqrt(1.0 / 27)
    gy = torch.randn(K * B, H, W, Co, generator=g)
    fn = lambda m, v, bb: PC.parts_conv2d(image.to(m.dtype), m, v, bb)  # noqa: E731
    _, g_o = _oracle_grads(fn, (mask, V, b), [gy], torch.float32)
    _, g_o64 = _oracle_grads(fn, (mask, V, b), [gy], torch.float64)
    lc = [t.cuda().requires_grad_(True) for t in (mask, V, b)]
    y = ups.model.parts_conv2d(image.cuda(), *lc)
    got = torch.autograd.grad(y, lc, gy.cuda())
    for a, o, o64, name, n in zip(got, g_o, g_o64, ("dmask", "dV", "db"), (27 * Co, B * H * W, K * B * H * W)):
        assert_close(a, o, name, atol=_sum_atol(n, o, o64))
    with pytest.raises(Exception, match="no gradient with respect to the image"):
        im = image.cuda().requires_grad_(True)
        torch.autograd.grad(ups.model.parts_conv2d(im, *lc).sum(), [im])


@pytest.mark.parametrize("B,H,W,K,Co,kind", [(2, 64, 64, 16, 32, "hard"), (1, 128, 128, 16, 32, "ties"),
                                             (3, 40, 50, 25, 32, "ties"), (2, 33, 31, 8, 64, "soft"),
                                             (1, 16, 48, 32, 128, "hard"), (2, 9, 70, 4, 4, "soft"),
                                             (1, 1, 1, 16, 32, "hard")])
def test_parts_conv2d_vs_oracle(ups, B, H, W, K, Co, kind):
    from oracle import parts_conv as PC
    _, mask, _, _, _, _ = _case(B, H, W, K, 4, Co, kind, seed=B * 10 + K)
    g = torch.Generator().manual_seed(Co)
    image = torch.rand(B, H, W, 3, generator=g) * 2 - 1
    V = (torch.rand(3, 3, 3, Co, generator=g) * 2 - 1) * math.sqrt(1.0 / 27)
    b = (torch.rand(Co, generator=g) * 2 - 1) * math.sqrt(1.0 / 27)
    y_o = PC.parts_conv2d(image, mask, V, b)
    y = ups.model.parts_conv2d(image.cuda(), mask.cuda(), V.cuda(), b.cuda())
    assert y.shape == (K * B, H, W, Co)
    assert_close(y, y_o, "out")
    # equals this library's own un-fused path: conv over the materialised part-major part images
    parts_pm = ups.model.mask_parts_partmajor(image.cuda(), mask.cuda())
    assert_close(IC.conv2d_same(parts_pm.cpu(), V, b), y_o, "conv(mask_parts_partmajor)")


def test_full_size_backward_properties(ups):
    """CUB B=64 slice of BASELINE configs[1]: properties of the decode+conv backward that need no oracle —
    softmax-backward rows sum to zero, db is the plain sum of the cotangent, the backward is linear in the
    cotangent, and the hard-mask conv equals the un-fused inject -> conv path of this library on a sample."""
    B, H, W, K, F, Co = 64, 128, 128, 16, 64, 32
    g = torch.Generator(device="cuda").manual_seed(1)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    logits, feat = rn(B, H, W, K).requires_grad_(True), rn(B, K, F).requires_grad_(True)
    V, b = (rn(3, 3, F + K, Co) * 0.04).requires_grad_(True), (rn(Co) * 0.04).requires_grad_(True)
    m0, labels, mh, y = ups.model.decode_conv2d(logits, feat, V, b)
    assert labels.dtype == torch.int64 and torch.equal(labels, m0.argmax(-1))
    g1, g2 = rn(B, H, W, Co), rn(B, H, W, Co)
    d1 = torch.autograd.grad(y, [logits, feat, V, b], g1, retain_graph=True)
    d2 = torch.autograd.grad(y, [logits, feat, V, b], g2, retain_graph=True)
    d12 = torch.autograd.grad(y, [logits, feat, V, b], g1 + g2)
    assert float(d1[0].sum(-1).abs().max()) < 2e-4                       # sum_k p_k (g_k - <g,p>) = 0
    db64 = g1.double().sum((0, 1, 2)).cpu()
    assert_close(d1[3], db64, "db", rtol=1e-4, atol=own_error_atol(g1.sum((0, 1, 2)).cpu(), db64))
    for a, c, e, name in zip(d1, d2, d12, ("dlogits", "dfeat", "dV", "db")):
        scale = float(e.abs().max())
        assert float((a + c - e).abs().max()) <= 2e-5 * max(1.0, scale), name
    inj = ups.model.inject_features(feat.detach()[:2], mh.detach()[:2])
    assert_close(IC.conv2d_same(inj.cpu(), V.detach().cpu(), b.detach().cpu()), y.detach()[:2].cpu(), "conv(inject)")
