"""CPU tier: the known-answer vectors SURVEY.md 8c lists for the oracle (the reference's own tests hold none for this
path), each derivable by hand from the cited reference lines.  The GPU tier then compares the kernels with the oracle."""
import math

import torch

from oracle import inject_conv as IC
from oracle import parts as OP
from oracle import parts_conv as PC


def test_hard_max_ties_and_argmax():
    """(4) cub/code/nn.py:134-136: ties give several ones; tf.argmax returns the first index, int64."""
    y = torch.tensor([0.2, 0.5, 0.5, 0.1]).reshape(1, 1, 1, 4)
    assert OP.hard_max(y, 3).flatten().tolist() == [0.0, 1.0, 1.0, 0.0]
    lab = OP.argmax_labels(y)
    assert lab.dtype == torch.int64 and lab.item() == 1


def test_straight_through_value_and_gradient():
    """(5) cub/code/nn.py:154-168: forward fl(fl(h - y) + y), gradient 1 into y and 0 into h."""
    h = torch.tensor([1.0], requires_grad=True)
    y = torch.tensor([0.3], requires_grad=True)
    out = OP.straight_through_estimator(h, y)
    want = (torch.tensor(1.0) - torch.tensor(0.3)) + torch.tensor(0.3)
    assert out.item() == want.item()
    gh, gy = torch.autograd.grad(out.sum(), [h, y], allow_unused=True)
    assert gy.item() == 1.0 and (gh is None or gh.item() == 0.0)


def test_unpool_one_hot_is_a_gather_and_inject_appends_the_mask():
    """(6) cub/code/SB_model48i/model.py:225-249,482-484 vs cub/code/nn.py:2469-2487."""
    g = torch.Generator().manual_seed(0)
    B, H, W, K, F = 2, 5, 6, 4, 3
    labels = torch.randint(0, K, (B, H, W), generator=g)
    mask = torch.nn.functional.one_hot(labels, K).float()
    feat = torch.randn(B, K, F, generator=g)
    inj = OP.inject(feat, mask)
    assert inj.shape == (B, H, W, F + K)
    assert torch.equal(inj[..., :F], OP.unpool_features_gathered(feat, labels))
    assert torch.equal(inj[..., F:], mask)


def test_apply_partwise_identity_and_part_major_batch_index():
    """(7) cub/code/nn.py:81-113: the function sees [K*B,h,w,f] with row k*B+b."""
    B, H, W, K, C = 3, 2, 2, 4, 1
    x = torch.zeros(B, H, W, K, C)
    for b in range(B):
        for k in range(K):
            x[b, :, :, k, :] = 10 * k + b
    seen = {}

    def f(t):
        seen["rows"] = t[:, 0, 0, 0].tolist()
        return t

    assert torch.equal(OP.apply_partwise(x, f), x)
    assert seen["rows"] == [float(10 * k + b) for k in range(K) for b in range(B)]


def test_pool_features_all_ones_mask_is_the_group_mean():
    """(8) deepfashion/code/foo.py:287-307."""
    g = torch.Generator().manual_seed(1)
    B, H, W, K, Fg = 2, 4, 4, 3, 2
    fmap = torch.randn(B, H, W, K * Fg, generator=g)
    out = OP.pool_features(fmap, torch.ones(B, H, W, K))
    assert torch.allclose(out, fmap.reshape(B, H * W, K, Fg).mean(1), atol=1e-6)


def test_softmax_normalisation():
    """(9) cub/code/nn.py:58-71: rows sum to 1 over K; spatial_softmax sums to 1 over H*W per channel."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 6, 5, 16, generator=g) * 3
    assert float((OP.softmax(x).sum(-1) - 1).abs().max()) <= 16 * 2 ** -23
    assert float((OP.spatial_softmax(x).sum((1, 2)) - 1).abs().max()) <= 1e-5


def test_mask_parts_shape_and_values():
    """(10) cub/code/SB_model48i/model.py:176-187."""
    g = torch.Generator().manual_seed(3)
    img, mask = torch.randn(2, 4, 4, 3, generator=g), torch.rand(2, 4, 4, 5, generator=g)
    parts = OP.mask_parts(img, mask)
    assert parts.shape == (2, 4, 4, 5, 3)
    for k in range(5):
        assert torch.equal(parts[..., k, :], img * mask[..., k:k + 1])


def test_first_decoder_conv_known_answers():
    """with an all-zero mask the first decoder layer is its bias; with a one-hot mask it is the bias plus,
    per tap, one row of the per-sample table G[b,t,k,:] = feat[b,k,:] . V[t,:F,:] + V[t,F+k,:] (model.py:482-485)."""
    g = torch.Generator().manual_seed(4)
    B, H, W, K, F, Co = 1, 4, 5, 3, 2, 4
    feat, V, b = torch.randn(B, K, F, generator=g), torch.randn(3, 3, F + K, Co, generator=g), torch.randn(Co, generator=g)
    y0 = IC.inject_conv2d(feat, torch.zeros(B, H, W, K), V, b)
    assert torch.allclose(y0, b.expand(B, H, W, Co), atol=1e-6)
    labels = torch.randint(0, K, (B, H, W), generator=g)
    y = IC.inject_conv2d(feat, torch.nn.functional.one_hot(labels, K).float(), V, b)
    G = IC.inject_conv_table(feat, V)
    for (yy, xx) in ((0, 0), (2, 3), (3, 4)):
        want = b.clone()
        for i in range(3):
            for j in range(3):
                qy, qx = yy + i - 1, xx + j - 1
                if 0 <= qy < H and 0 <= qx < W:
                    want = want + G[0, 3 * i + j, labels[0, qy, qx]]
        assert torch.allclose(y[0, yy, xx], want, atol=1e-5)


def test_first_encoder_conv_known_answers():
    """plane k of the part-major output differs from the bias only within one pixel of part k."""
    g = torch.Generator().manual_seed(5)
    B, H, W, K, Co = 2, 6, 6, 4, 3
    img = torch.rand(B, H, W, 3, generator=g) + 0.5
    V, b = torch.randn(3, 3, 3, Co, generator=g), torch.randn(Co, generator=g)
    labels = torch.zeros(B, H, W, dtype=torch.long)
    labels[:, 4, 4] = 2                                     # a single pixel of part 2
    out = PC.parts_conv2d(img, torch.nn.functional.one_hot(labels, K).float(), V, b)
    assert out.shape == (K * B, H, W, Co)
    for bb in range(B):
        for k in (1, 3):                                    # empty parts: bias everywhere
            assert torch.allclose(out[k * B + bb], b.expand(H, W, Co), atol=1e-6)
        plane = out[2 * B + bb] - b
        far = torch.ones(H, W, dtype=torch.bool)
        far[3:6, 3:6] = False
        assert float(plane[far].abs().max()) <= 1e-6 and float(plane[4, 4].abs().max()) > 0
        # centre tap (t = 4) at the pixel itself
        assert torch.allclose(plane[4, 4], img[bb, 4, 4] @ V[1, 1], atol=1e-5)
