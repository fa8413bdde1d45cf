"""Drop-in for the part-map helpers of the reference's `nips19.nn` module
(cub/code/nn.py == pennaction/code/nn.py; deepfashion/code/nn.py): same names, argument
order and defaults; torch CUDA tensors (NHWC, fp32) instead of TF graph tensors."""
import torch

from . import ops


def softmax(x, spatial=False):
    """cub/code/nn.py:58-62 — K-way softmax over the last axis (spatial=True: over H*W)."""
    if spatial:
        return spatial_softmax(x)
    return ops.part_softmax(x)


def spatial_softmax(features):
    """cub/code/nn.py:65-71 — [N,H,W,C] -> softmax over H*W for every (n, c)."""
    assert features.dim() == 4, list(features.shape)
    return ops.spatial_softmax(features)


def _to_last(y, axis):
    axis = axis % y.dim()
    return (y, None) if axis == y.dim() - 1 else (y.movedim(axis, -1).contiguous(), axis)


def hard_max(y, axis):
    """cub/code/nn.py:134-136 — float(y == reduce_max(y, axis, keep_dims=True)); every tied
    maximum is marked.  Not differentiable (tf.equal)."""
    yl, moved = _to_last(y.detach(), axis)
    out = ops.hard_max(yl)
    return out if moved is None else out.movedim(-1, moved)


def straight_through_estimator(y_hard, y):
    """cub/code/nn.py:154-168 — tf.stop_gradient(y_hard - y) + y."""
    return ops.straight_through(y_hard.detach(), y)


def hard_max_straight_through(y, axis):
    """cub/code/nn.py:117-131 (deprecated in the reference in favour of the two calls)."""
    return straight_through_estimator(hard_max(y, axis), y)


def apply_partwise(input_, func):
    """cub/code/nn.py:81-113 — [b,h,w,parts,f] -> part-major [parts*b,h,w,f] -> func ->
    [b,h_out,w_out,parts,c_out]."""
    b, h, w, parts, f = input_.shape
    x = ops.partwise_fold(input_)
    y = func(x)
    assert y.dim() == 4 and y.shape[0] == parts * b, list(y.shape)
    return ops.partwise_unfold(y, parts)


def mask2hotmask(mask, n_parts):
    """cub/code/nn.py:2086-2089 — one_hot(argmax(mask, 3), n_parts)."""
    return ops.one_hot(ops.argmax(mask.detach()), n_parts)


def argmax(y, axis=3):
    """tf.argmax(y, 3) as used at cub/code/SB_model48i/model.py:447,465,470 — int64, first index."""
    yl, _ = _to_last(y.detach(), axis)
    return ops.argmax(yl)


def unpool_features_gathered(feature_vectors, mask):
    """cub/code/nn.py:2469-2487 — feature_vectors [B,parts,F], integer mask [B,h,w] -> [B,h,w,F]."""
    bs, h, w = mask.shape
    fshape = list(feature_vectors.shape)
    assert len(fshape) == 3, fshape
    assert fshape[0] == bs, fshape
    return ops.part_gather(feature_vectors, mask)


def probs_to_mu_sigma(probs, scaling_factor):
    """cub/code/nn.py:1541-1587 — per-part mean (y, x) and 2x2 covariance of the [b,h,w,k] densities on
    the linspace(-1, 1) grid; scaling_factor [b,k].  Returns (mu [b,k,2], sigma [b,k,2,2])."""
    bn, h, w, nk = probs.shape
    sf = scaling_factor.reshape(bn, nk)
    mu, sigma, _ = ops.mask_moments(probs, sf)
    return mu, sigma
