"""-m gpu: the stand-alone part-map operators (through the reference-named helpers, the torch
custom ops and the C ABI) against the oracle on identical seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import parts as OP
from util import assert_bitexact, assert_close, cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ups():
    import ups_b200
    return ups_b200


def _logits(n, K, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, K, generator=g) * 3
    x[: n // 8] = torch.round(x[: n // 8])        # exact ties
    x[n // 8: n // 4] *= 40                       # logits out to +-120
    x[n // 4: n // 4 + 7] = 0.0                   # all-equal rows
    return x


@pytest.mark.parametrize("K", [1, 2, 3, 4, 8, 16, 25, 32, 64])
def test_softmax_labels_hard_bitexact(ups, K):
    x = _logits(4099, K, K).reshape(1, 4099, 1, K)
    probs, labels, hard = ups.ops.part_softmax_full(x.cuda())
    p = OP.softmax(x)
    assert_bitexact(probs, p, f"softmax K={K}")
    assert labels.dtype == torch.int64
    assert torch.equal(labels.cpu(), OP.argmax_labels(p)), "labels must be bit-exact"
    assert_bitexact(hard, OP.straight_through_estimator(OP.hard_max(p, 3), p), f"hard/ST K={K}")
    assert_bitexact(ups.softmax(x.cuda()), p, "nn.softmax")


@pytest.mark.parametrize("K", [4, 16, 25])
def test_softmax_backward(ups, K):
    x = _logits(1031, K, 100 + K).reshape(1, 1031, 1, K)
    g = torch.randn(x.shape, generator=torch.Generator().manual_seed(5))
    xo = x.clone().requires_grad_(True)
    (do,) = torch.autograd.grad(OP.softmax(xo), xo, g)
    xc = x.cuda().requires_grad_(True)
    (dc,) = torch.autograd.grad(ups.softmax(xc), xc, g.cuda())
    assert_close(dc, do, f"softmax bwd K={K}")


def test_hardmax_st_argmax_onehot(ups):
    y = torch.tensor([[[[0.2, 0.5, 0.5, 0.1]]]])
    assert ups.hard_max(y.cuda(), 3).cpu().flatten().tolist() == [0, 1, 1, 0]
    assert ups.nn.argmax(y.cuda(), 3).item() == 1
    st = ups.straight_through_estimator(torch.tensor([1.0]).cuda(), torch.tensor([0.3]).cuda())
    assert st.item() == float(np.float32(np.float32(1.0) - np.float32(0.3)) + np.float32(0.3))  # 8c(5)
    g = torch.Generator().manual_seed(3)
    m = torch.softmax(torch.randn(2, 5, 7, 6, generator=g), -1)
    assert_bitexact(ups.hard_max(m.cuda(), 3), OP.hard_max(m, 3), "hard_max")
    assert_bitexact(ups.hard_max(m.cuda(), 1), OP.hard_max(m, 1), "hard_max axis=1")
    assert_bitexact(ups.hard_max_straight_through(m.cuda(), 3), OP.hard_max_straight_through(m, 3), "hmst")
    assert_bitexact(ups.mask2hotmask(m.cuda(), 6), OP.mask2hotmask(m, 6), "mask2hotmask")
    # ST gradient: identity into y, nothing into y_hard
    yc = m.cuda().requires_grad_(True)
    out = ups.straight_through_estimator(ups.hard_max(yc, 3), yc)
    (gy,) = torch.autograd.grad(out, yc, torch.ones_like(out))
    assert torch.equal(gy.cpu(), torch.ones_like(m))


def test_spatial_softmax(ups):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 12, 12, 16, generator=g) * 2
    xo = x.clone().requires_grad_(True)
    po = OP.spatial_softmax(xo)
    xc = x.cuda().requires_grad_(True)
    pc = ups.softmax(xc, spatial=True)
    assert_close(pc, po, "spatial softmax", atol=1e-8)
    assert_close(pc.sum((1, 2)), torch.ones(3, 16), "sums to 1 over HW")
    cot = torch.randn(x.shape, generator=g)
    (do,) = torch.autograd.grad(po, xo, cot)
    (dc,) = torch.autograd.grad(pc, xc, cot.cuda())
    assert_close(dc, do, "spatial softmax bwd", atol=1e-7)


def _masks(B, S, K, seed, hard):
    g = torch.Generator().manual_seed(seed)
    p = OP.softmax(torch.randn(B, S, S, K, generator=g))
    return OP.straight_through_estimator(OP.hard_max(p, 3), p) if hard else p


@pytest.mark.parametrize("K,hard", [(16, True), (5, False)])
def test_mask_parts_and_partwise(ups, K, hard):
    B, S = 3, 12
    g = torch.Generator().manual_seed(K)
    img = torch.rand(B, S, S, 3, generator=g) * 2 - 1
    m = _masks(B, S, K, 11, hard)
    io, mo = img.clone().requires_grad_(True), m.clone().requires_grad_(True)
    ic, mc = img.cuda().requires_grad_(True), m.cuda().requires_grad_(True)
    po, pc = OP.mask_parts(io, mo), ups.mask_parts(ic, mc)
    assert list(pc.shape) == [B, S, S, K, 3]
    assert_bitexact(pc, po, "mask_parts")
    cot = torch.randn(po.shape, generator=g)
    go = torch.autograd.grad(po, [io, mo], cot)
    gc = torch.autograd.grad(pc, [ic, mc], cot.cuda())
    assert_close(gc[0], go[0], "mask_parts dimage")
    assert_close(gc[1], go[1], "mask_parts dmask")
    # part-major layout == apply_partwise's fold; fold index is k*B + b              SURVEY 8c(7)
    pm = ups.model.mask_parts_partmajor(ic, mc)
    assert_bitexact(pm, po.permute(3, 0, 1, 2, 4).reshape(K * B, S, S, 3), "part-major")
    seen = {}
    out = ups.apply_partwise(pc, lambda x: seen.setdefault("x", x))
    assert_bitexact(seen["x"], po.permute(3, 0, 1, 2, 4).reshape(K * B, S, S, 3), "fold")
    assert_bitexact(out, po, "apply_partwise(identity)")
    with pytest.raises(AssertionError):
        ups.mask_parts(ic, mc[:, :-1])                                                   # model.py:180


def test_encode_parts_matches_oracle(ups):
    B, S, K, F = 2, 8, 4, 6
    g = torch.Generator().manual_seed(2)
    img = torch.rand(B, S, S, 3, generator=g)
    m = _masks(B, S, K, 4, True)
    Wl = torch.randn(3, F, generator=g)
    enc_o = OP.encode_parts(OP.mask_parts(img, m), lambda x: (x.mean((1, 2)) @ Wl).reshape(-1, 1, 1, F))
    Wc = Wl.cuda()
    enc_c = ups.encode_parts(ups.mask_parts(img.cuda(), m.cuda()),
                             lambda x: (x.mean((1, 2)) @ Wc).reshape(-1, 1, 1, F))
    assert list(enc_c.shape) == [B, K, F]
    assert_close(enc_c, enc_o, "encode_parts")


@pytest.mark.parametrize("K,Fg,hard", [(4, 3, False), (16, 4, True), (25, 2, False)])
def test_pool_features(ups, K, Fg, hard):
    B, S = 3, 16
    g = torch.Generator().manual_seed(K + Fg)
    fm = torch.randn(B, S, S, K * Fg, generator=g)
    m = _masks(B, S, K, 21, hard)
    fo, mo = fm.clone().requires_grad_(True), m.clone().requires_grad_(True)
    fc, mc = fm.cuda().requires_grad_(True), m.cuda().requires_grad_(True)
    oo, oc = OP.pool_features(fo, mo), ups.pool_features(fc, mc)
    assert_close(oc, oo, "pool_features")
    cot = torch.randn(oo.shape, generator=g)
    go = torch.autograd.grad(oo, [fo, mo], cot)
    gc = torch.autograd.grad(oc, [fc, mc], cot.cuda())
    assert_close(gc[0], go[0], "pool dfmap")
    assert_close(gc[1], go[1], "pool dmask")
    ones = torch.ones(B, S, S, K)
    assert_close(ups.pool_features(fc, ones.cuda()), fm.reshape(B, S * S, K, Fg).mean(1), "all-ones mask")  # 8c(8)
    with pytest.raises(AssertionError):
        ups.pool_features(fc[:, :, :, :-1], mc)


def test_get_features_and_mean_pool(ups):
    B, S, K, Cf = 2, 16, 8, 5
    g = torch.Generator().manual_seed(8)
    fm = torch.randn(B, S, S, Cf, generator=g)
    m = _masks(B, S, K, 5, False)
    fo, mo = fm.clone().requires_grad_(True), m.clone().requires_grad_(True)
    fc, mc = fm.cuda().requires_grad_(True), m.cuda().requires_grad_(True)
    oo, oc = OP.get_features(fo, mo, True), ups.get_features(fc, mc, True)
    assert_close(oc, oo, "get_features slim")
    cot = torch.randn(oo.shape, generator=g)
    go = torch.autograd.grad(oo, [fo, mo], cot)
    gc = torch.autograd.grad(oc, [fc, mc], cot.cuda())
    assert_close(gc[0], go[0], "get_features dfeatures")
    assert_close(gc[1], go[1], "get_features dmask")
    f5 = torch.randn(B, S, S, K, Cf, generator=g)
    assert_close(ups.get_features(f5.cuda(), mc, False), OP.get_features(f5, m, False), "get_features full")
    assert_close(ups.part_mean_pool(fc, mc), OP.part_mean_pool(fm, m), "part_mean_pool")


@pytest.mark.parametrize("K,F,hard", [(16, 64, True), (25, 5, False), (8, 16, False)])
def test_unpool_inject_gather(ups, K, F, hard):
    B, S = 2, 12
    g = torch.Generator().manual_seed(K * F)
    feat = torch.randn(B, K, F, generator=g)
    m = _masks(B, S, K, 31, hard)
    fo, mo = feat.clone().requires_grad_(True), m.clone().requires_grad_(True)
    fc, mc = feat.cuda().requires_grad_(True), m.cuda().requires_grad_(True)
    uo, uc = OP.unpool_features(fo, mo), ups.unpool_features(fc, mc)
    assert list(uc.shape) == [B, S, S, K, F]
    assert_bitexact(uc, uo, "unpool_features")
    assert_bitexact(ups.unpool_features(fc, mc, reshape=True), uo.reshape(B, S, S, K * F), "reshape=True")
    cot = torch.randn(uo.shape, generator=g)
    go = torch.autograd.grad(uo, [fo, mo], cot)
    gc = torch.autograd.grad(uc, [fc, mc], cot.cuda())
    assert_close(gc[0], go[0], "unpool dfeat")
    assert_close(gc[1], go[1], "unpool dmask")
    io, ic = OP.inject(fo, mo), ups.inject_features(fc, mc)
    assert list(ic.shape) == [B, S, S, F + K]
    assert_close(ic, io, "inject")
    assert_bitexact(ic[..., F:], m, "inject tail == mask")
    cot = torch.randn(io.shape, generator=g)
    go = torch.autograd.grad(io, [fo, mo], cot)
    gc = torch.autograd.grad(ic, [fc, mc], cot.cuda())
    assert_close(gc[0], go[0], "inject dfeat")
    assert_close(gc[1], go[1], "inject dmask")
    labels = torch.argmax(m, 3)
    assert_bitexact(ups.unpool_features_gathered(fc, labels.cuda()), OP.unpool_features_gathered(feat, labels), "gather")
    if hard:  # one-hot mask: inject == mask-scaled gather
        mon = m.max(-1, keepdim=True).values
        assert_close(ic[..., :F], OP.unpool_features_gathered(feat, labels) * mon, "one-hot inject == gather")
    with pytest.raises(AssertionError):
        ups.unpool_features(fc[:, :-1], mc)


def test_pool_unpool_block(ups):
    B, S, K, Fg = 2, 8, 4, 3
    g = torch.Generator().manual_seed(77)
    fm = torch.randn(B, S, S, K * Fg, generator=g)
    m1, m0 = _masks(B, S, K, 1, False), _masks(B, S, K, 2, False)
    la_o, inj_o = OP.pool_unpool_block(fm, m1, m0, reshape=True)
    la_c, inj_c = ups.pool_unpool_block(fm.cuda(), m1.cuda(), m0.cuda(), reshape=True)
    assert_close(la_c, la_o, "pool_unpool_block features")
    assert_close(inj_c, inj_o, "pool_unpool_block injected")


def test_golden_chain_through_public_api(ups, golden):
    """The fixtures produced by the reference's own code, replayed through ups_b200's helpers."""
    for name in ("parts_k4.npz", "parts_k16.npz", "parts_k25.npz"):
        gd = golden(name)
        t = lambda k: torch.from_numpy(gd[k]).cuda()  # noqa: E731
        l0, l1, img, feat = (t(k).requires_grad_(True) for k in ("l0", "l1", "img", "feat"))
        m0, m1 = ups.softmax(l0), ups.softmax(l1)
        m0h = ups.straight_through_estimator(ups.hard_max(m0, 3), m0)
        m1h = ups.straight_through_estimator(ups.hard_max(m1, 3), m1)
        parts = ups.mask_parts(img, m1h)
        u5 = ups.unpool_features(feat, m0h)
        inj = ups.inject_features(feat, m0h)
        pooled = ups.part_mean_pool(img, m1h)
        lab = ups.nn.argmax(m0, 3)
        assert np.array_equal(lab.cpu().numpy(), gd["labels0"])
        for got, k in ((m0, "m0"), (m1, "m1"), (m0h, "m0_hard"), (parts, "view1_parts"), (u5, "u5"), (inj, "inj"),
                       (pooled, "pooled")):
            assert_close(got, torch.from_numpy(gd[k]), f"{name}:{k}", rtol=1e-4, atol=1e-5)
        grads = torch.autograd.grad([inj, parts, pooled, m0, m1], [l0, l1, feat, img],
                                    [t(k) for k in ("g_inj", "g_parts", "g_pooled", "g_m0", "g_m1")])
        for got, k in zip(grads, ("dl0", "dl1", "dfeat", "dimg")):
            assert_close(got, torch.from_numpy(gd[k]), f"{name}:{k}")


def test_cpu_tensors_are_rejected(ups):
    with pytest.raises(Exception):
        ups.softmax(torch.randn(1, 2, 2, 4))


def test_unpool_features_gathered_backward(ups):
    """tf.gather is differentiable in feature_vectors (cub/code/nn.py:2469-2487): dfeat = segment sum of the cotangent."""
    B, S, K, F = 3, 24, 25, 16
    g = torch.Generator().manual_seed(2)
    feat = torch.randn(B, K, F, generator=g)
    labels = torch.randint(0, K, (B, S, S), generator=g)
    G = torch.randn(B, S, S, F, generator=g)
    fo = feat.clone().requires_grad_(True)
    (d_o,) = torch.autograd.grad(OP.unpool_features_gathered(fo, labels), fo, G)
    fc = feat.cuda().requires_grad_(True)
    (d_c,) = torch.autograd.grad(ups.unpool_features_gathered(fc, labels.cuda()), fc, G.cuda())
    assert_close(d_c, d_o, "dfeat of the gathered unpooling")


@pytest.mark.parametrize("n_rows,n_cols,n_fill,ss,ds", [(1000, 25, 7, 25, 32), (1000, 25, 0, 32, 25), (77, 89, 7, 89, 96),
                                                      (77, 89, 0, 96, 89), (5, 3, 2, 6, 7), (64, 48, 0, 96, 48)])
def test_copy_rows(n_rows, n_cols, n_fill, ss, ds):
    """ups_copy_rows (padding / cutting the [.,K] rows of a part count that is not a power of two): vector and scalar paths."""
    from ups_b200 import _cabi as C
    g = torch.Generator().manual_seed(n_cols)
    src = torch.randn(n_rows, ss, generator=g).cuda()
    dst = torch.full((n_rows, ds), 7.0, device="cuda")
    C.call("ups_copy_rows", src.data_ptr(), ss, dst.data_ptr(), ds, n_rows, n_cols, n_fill, float("-inf"),
           torch.cuda.current_stream().cuda_stream)
    want = torch.full((n_rows, ds), 7.0)
    want[:, :n_cols] = src.cpu()[:, :n_cols]
    want[:, n_cols:n_cols + n_fill] = float("-inf")
    assert torch.equal(dst.cpu(), want)
