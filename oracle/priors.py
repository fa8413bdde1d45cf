"""TEST INFRASTRUCTURE — CPU restatement of the mask priors and of the mean-field sampling step that sit on
either side of the part-map softmax.  Pinned against tests/golden/priors.npz (the
reference's function bodies run under tf1_shim; tests/test_oracle_golden.py).

Every elementwise map here is a chain of single IEEE operations with one possible order, so the CUDA kernels
reproduce them bit for bit; the spatial sums carry the 1e-4 / 1e-5 tolerance.
"""
import torch

from . import parts as OP


# ------------------------------------------------------------------ Mumford-Shah on the probabilities
def tf_squared_grad(x):
    """cub/code/nn.py:1357-1378 (fd_kernel, tf_grad, tf_squared_grad).  The 3x3 SAME cross-correlation with
    0.5*[0, .5, -.5] along the centre row / centre column has two non-zero taps, i.e. for x [b,h,w,k]
        gx = 0.25*x[i,j] - 0.25*x[i,j+1]     gy = 0.25*x[i,j] - 0.25*x[i+1,j]     (zero beyond the border)
    and g = gx*gx + gy*gy.  Scaling by 0.25 is exact, so fl(0.25a - 0.25b) = 0.25*fl(a - b) whatever order the
    convolution adds its taps in."""
    right = torch.cat([x[:, :, 1:], torch.zeros_like(x[:, :, :1])], dim=2)
    down = torch.cat([x[:, 1:], torch.zeros_like(x[:, :1])], dim=1)
    gx = 0.25 * (x - right)
    gy = 0.25 * (x - down)
    return gx * gx + gy * gy


def mumford_shah(x, alpha, lambda_):
    """cub/code/nn.py:1381-1386 -> (r, smoothness_cost, contour_cost), each [b,h,w,k].  tf.minimum routes the
    gradient to its first argument where alpha*g <= lambda (MinimumGrad)."""
    g = tf_squared_grad(x)
    ag = alpha * g
    lam = torch.full_like(ag, lambda_)
    r = torch.where(ag <= lam, ag, lam)
    zero = torch.zeros_like(g)
    smooth = torch.where(ag < lam, r, zero)
    contour = torch.where(ag >= lam, r, zero)
    return r, smooth, contour


def edge_set(x, alpha, lambda_):
    """cub/code/nn.py:1389-1392."""
    g = tf_squared_grad(x)
    thr = torch.tensor(lambda_ / alpha, dtype=torch.float32)      # Python-float division, then one fp32 rounding
    return (g > thr).to(x.dtype)


def mumford_shah_sums(x, alpha, lambda_):
    """The spatial sums the training step squares (cub/code/SB_model48i/model.py:744-769):
    [b,4,k] = sum over (h,w) of (r, smoothness_cost, contour_cost, x)."""
    r, s, c = mumford_shah(x, alpha, lambda_)
    return torch.stack([v.sum(dim=(1, 2)) for v in (r, s, c, x)], dim=1)


# ------------------------------------------------------------------ mean-field distribution on the logits
def mean_field_sample(mean, eps, noise_level=1.0):
    """MeanFieldDistribution.sample (cub/code/nn.py:1421-1427): mean + noise_level*N(0,1).  The normal draw
    `eps` is an input (TF's RNG stream cannot be reproduced)."""
    return mean + torch.tensor(noise_level, dtype=torch.float32) * eps


def mean_field_kl(mean):
    """MeanFieldDistribution.kl (cub/code/nn.py:1429-1436)."""
    return (0.5 * (mean * mean).sum(dim=(1, 2, 3))).mean()


def _image_gradients(x):
    dy = torch.cat([x[:, 1:] - x[:, :-1], torch.zeros_like(x[:, :1])], dim=1)
    dx = torch.cat([x[:, :, 1:] - x[:, :, :-1], torch.zeros_like(x[:, :, :1])], dim=2)
    return dy, dx


def kl_improper_gmrf(mean):
    """MeanFieldDistribution.kl_improper_gmrf (cub/code/nn.py:1438-1446; call site
    cub/code/SB_model48i/model.py:1071): tf.image.image_gradients = forward differences, zero last row/column."""
    dy, dx = _image_gradients(mean)
    return (0.5 * (dy * dy + dx * dx)).sum(dim=(1, 2, 3)).mean()


def kl_tv(mean):
    """MeanFieldDistribution.kl_tv (cub/code/nn.py:1448-1451): mean over the batch of tf.image.total_variation."""
    dy = mean[:, 1:] - mean[:, :-1]
    dx = mean[:, :, 1:] - mean[:, :, :-1]
    return (dy.abs().sum(dim=(1, 2, 3)) + dx.abs().sum(dim=(1, 2, 3))).mean()


def logit_priors(mean):
    """[3] = (kl, kl_improper_gmrf, kl_tv) of one logits tensor."""
    return torch.stack([mean_field_kl(mean), kl_improper_gmrf(mean), kl_tv(mean)])


# ------------------------------------------------------------------ weak cross entropy
class _XentV2(torch.autograd.Function):
    """tf.nn.softmax_cross_entropy_with_logits_v2 over the last axis with TF's registered first-order gradient
    (nn_grad._SoftmaxCrossEntropyWithLogitsGrad): d/dlogits = g*(softmax - labels), d/dlabels = -g*log_softmax."""

    @staticmethod
    def forward(ctx, labels, logits):
        lsm = torch.log_softmax(logits, dim=-1)
        ctx.save_for_backward(labels, lsm)
        return -(labels * lsm).sum(dim=-1)

    @staticmethod
    def backward(ctx, g):
        labels, lsm = ctx.saved_tensors
        g = g.unsqueeze(-1)
        return -g * lsm, g * (torch.exp(lsm) - labels)


def weak_cross_entropy(logits, entropy_func="cross_entropy"):
    """cub/code/SB_model48i/model.py:667-681: the part logits supervised by their own hard (or soft) assignment."""
    p_labels = OP.softmax(logits, spatial=False)
    if entropy_func == "cross_entropy":
        labels = OP.straight_through_estimator(OP.hard_max(p_labels, 3), p_labels)
    elif entropy_func == "entropy":
        labels = p_labels
    else:
        raise ValueError("unkown entropy_func")
    return _XentV2.apply(labels, logits).mean()


# ------------------------------------------------------------------ colouring for logging / evaluation dumps
def mask2hotmask(mask, n_parts):
    """cub/code/nn.py:2086-2089."""
    return torch.nn.functional.one_hot(mask.argmax(dim=3), n_parts).to(mask.dtype)


def mask2rgb(mask, colors, make_hot=True):
    """cub/code/nn.py:2067-2083 with the colour table as an input: `colors` [k,3] in [0,1] (the reference takes it
    from matplotlib's inferno LUT, make_mask_colors, which is not part of the path)."""
    n_parts = mask.shape[3]
    hot = mask2hotmask(mask, n_parts) if make_hot else mask
    c = ((colors.to(torch.float64) - 0.5) * 2).to(torch.float32)
    return (hot[..., None] * c[None, None, None]).sum(dim=3)
