"""TEST INFRASTRUCTURE — CPU restatement of the reference's part-map helpers (fp32, NHWC).

Each function cites the reference lines it follows.  Softmax over the part axis uses the
canonical exp / sum order of oracle/canon.py so that `argmax`, `hard_max` (which marks
EVERY tied maximum) and the straight-through masks are bit-reproducible by the kernels.
Pinned against tests/golden/parts_*.npz (reference function bodies run under tf1_shim).
"""
import torch

from .canon import exp_canon, sum_tree

PARTS_DIM = 3      # cub/code/SB_model48i/model.py:12
FEATURE_DIM = 4    # cub/code/SB_model48i/model.py:13


class _SoftmaxCanon(torch.autograd.Function):
    """tf.nn.softmax over the last axis: p = exp(x - max) * (1 / sum) — the reciprocal-multiply
    form of TF's Eigen CPU kernel (its CUDA kernel divides; a 1-ulp matter).  Backward is
    TF's SoftmaxGrad formula dx = p * (g - sum_j g_j p_j) evaluated on the saved output."""

    @staticmethod
    def forward(ctx, x):
        m = x.max(dim=-1, keepdim=True).values
        e = exp_canon(x - m)
        s = sum_tree(e)
        r = torch.ones_like(s) / s                  # correctly rounded reciprocal
        p = e * r[..., None]
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        return p * (g - (g * p).sum(-1, keepdim=True))


def softmax(x, spatial=False):
    """cub/code/nn.py:58-62."""
    if spatial:
        return spatial_softmax(x)
    return _SoftmaxCanon.apply(x)


def spatial_softmax(features):
    """cub/code/nn.py:65-71 — softmax over H*W for every (n, c)."""
    N, H, W, C = features.shape
    f = features.permute(0, 3, 1, 2).reshape(N * C, H * W)
    p = torch.softmax(f, dim=-1)
    return p.reshape(N, C, H, W).permute(0, 2, 3, 1)


def hard_max(y, axis):
    """cub/code/nn.py:134-136 — 1.0 at every position equal to the max (ties -> several)."""
    return (y == y.max(dim=axis, keepdim=True).values).to(y.dtype)


def straight_through_estimator(y_hard, y):
    """cub/code/nn.py:154-168 — value fl(fl(y_hard - y) + y), gradient identity into y."""
    return (y_hard - y).detach() + y


def hard_max_straight_through(y, axis):
    """cub/code/nn.py:117-131."""
    return straight_through_estimator(hard_max(y, axis), y)


def argmax_labels(p):
    """tf.argmax(., 3) at cub/code/SB_model48i/model.py:447,465,470 — int64, first index."""
    return torch.argmax(p, dim=3)


def mask2hotmask(mask, n_parts):
    """cub/code/nn.py:2086-2089."""
    return torch.nn.functional.one_hot(torch.argmax(mask, dim=3), n_parts).to(torch.float32)


def apply_partwise(input_, func):
    """cub/code/nn.py:81-113 — [b,h,w,K,f] -> part-major [K*b,h,w,f] -> func -> back."""
    b, h, w, parts, f = input_.shape
    x = input_.permute(3, 0, 1, 2, 4).reshape(b * parts, h, w, f)
    y = func(x)
    _, h_out, w_out, c_out = y.shape
    out = y.reshape(parts, b, h_out, w_out, c_out)
    return out.permute(1, 2, 3, 0, 4)


def mask_parts(image, mask):
    """cub/code/SB_model48i/model.py:176-187 — [B,H,W,3],[B,H,W,K] -> [B,H,W,K,3]."""
    bs, h, w, _ = image.shape
    mshape = list(mask.shape)
    assert mshape[0] == bs and mshape[1] == h and mshape[2] == w, mshape
    return image.unsqueeze(PARTS_DIM) * mask.unsqueeze(FEATURE_DIM)


def encode_parts(part_image, encoder):
    """cub/code/SB_model48i/model.py:214-222 — [B,H,W,K,3] -> [B,K,F]."""
    b, h, w, parts, channels = part_image.shape
    enc = apply_partwise(part_image, encoder).reshape(b, parts, -1)
    assert enc.shape[0] == b and enc.shape[1] == parts
    return enc


def unpool_features(feature_vectors, mask, reshape=False):
    """cub/code/SB_model48i/model.py:225-249 ; deepfashion/code/foo.py:462-498 (`reshape`).
    out[b,h,w,k,f] = mask[b,h,w,k] * feat[b,k,f]."""
    bs, h, w, n_parts = mask.shape
    fshape = list(feature_vectors.shape)
    assert len(fshape) == 3, fshape
    assert fshape[0] == bs and fshape[1] == n_parts, fshape
    out = mask.unsqueeze(4) * feature_vectors[:, None, None, :, :]
    if reshape:
        out = out.reshape(bs, h, w, n_parts * fshape[2])
    return out


def inject(feature_vectors, mask):
    """cub/code/SB_model48i/model.py:482-484 — reduce_sum over parts (ascending k) then
    concat with the mask -> [B,h,w,F+K]."""
    u = unpool_features(feature_vectors, mask)
    acc = u[:, :, :, 0, :]
    for k in range(1, u.shape[3]):
        acc = acc + u[:, :, :, k, :]
    return torch.cat([acc, mask], dim=3)


def unpool_features_gathered(feature_vectors, mask):
    """cub/code/nn.py:2469-2487 — integer-label gather feat[b, label[b,h,w]] -> [B,h,w,F]."""
    bs, h, w = mask.shape
    fshape = list(feature_vectors.shape)
    assert len(fshape) == 3, fshape
    assert fshape[0] == bs, fshape
    b = torch.arange(bs)[:, None, None]
    return feature_vectors[b, mask.long()]


def pool_features(feature_map, mask):
    """deepfashion/code/foo.py:287-307 — [bs,h,w,K*f'],[bs,h,w,K] -> mean_hw -> [bs,K,f']."""
    bs, h, w, n_features = feature_map.shape
    mshape = list(mask.shape)
    assert mshape[0] == bs and mshape[1] == h and mshape[2] == w, mshape
    n_parts = mshape[3]
    assert n_features % n_parts == 0, (n_features, n_parts)
    fm = feature_map.reshape(bs, h, w, n_parts, -1)
    out = (fm * mask.unsqueeze(4)).mean(dim=(1, 2))
    assert out.shape[0] == bs and out.shape[1] == n_parts
    return out


def pool_unpool_block(feature_map, pool_mask, unpool_mask, reshape=False):
    """deepfashion/code/foo.py:574-578."""
    local_app_features = pool_features(feature_map, pool_mask)
    injected_mask = unpool_features(local_app_features, unpool_mask, reshape=reshape)
    return local_app_features, injected_mask


def get_features(features, part_map, slim):
    """baselines/unsupervised-disentangling/ops.py:182-193 — dense (KxHW).(HWxC) contraction."""
    if slim:
        return torch.einsum('bijf,bijk->bkf', features, part_map)
    return torch.einsum('bijkf,bijk->bkf', features, part_map)


def part_mean_pool(image, mask):
    """Pooling tail of e_alpha (cub/code/SB_model48i/model.py:50-52, reduce_mean over H,W)
    applied directly to mask_parts(image, mask): pooled[b,k,c] = mean_hw image*mask."""
    return get_features(image, mask, True) / float(image.shape[1] * image.shape[2])
