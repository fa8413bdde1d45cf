"""TEST INFRASTRUCTURE — CPU restatement of the data-parallel wrapper's stand-in modules.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.

encoder tail: the end of `encoder_model` — mean over (H,W) then a 1x1 conv to F
    (cub/code/SB_model48i/model.py:50-52); the mean is the path's `pooled`, the conv is `pooled @ Wlin + blin`.
decoder head: a 1x1 conv F+K -> 3 on the injection concat(unpool_features_gathered(feat, labels), one_hot(labels))
    (cub/code/nn.py:2469-2487 gathers feat[b, label]; model.py:482-484 concatenates the mask; the first layer of `dd`,
    model.py:96,485, is a convolution on that map).
Gradients come from torch autograd in float64 (the kernels are checked against them to 1e-4 / 1e-5 scaled by the
number of summed terms).
"""
import torch


def tail_fwd(pooled, Wlin, blin):
    return pooled @ Wlin + blin


def head_fwd(labels, feat, Whead, bhead):
    B, K, F = feat.shape
    lab = labels.reshape(B, -1)
    gathered = torch.gather(feat, 1, lab[..., None].expand(B, lab.shape[1], F))        # feat[b, label[b,p], :]
    onehot = torch.nn.functional.one_hot(lab, K).to(feat.dtype)
    inj = torch.cat([gathered, onehot], -1)
    return inj @ Whead + bhead


def tail_grads(pooled, dfeat, Wlin, blin):
    """[dWlin (3*F), dblin (F)] flattened, float64."""
    W = Wlin.double().requires_grad_(True)
    b = blin.double().requires_grad_(True)
    (tail_fwd(pooled.double(), W, b) * dfeat.double()).sum().backward()
    return torch.cat([W.grad.reshape(-1), b.grad.reshape(-1)])


def head_grads(g_recon, labels, feat, Whead, bhead):
    """[dWhead ((F+K)*3), dbhead (3)] flattened, float64."""
    W = Whead.double().requires_grad_(True)
    b = bhead.double().requires_grad_(True)
    B = feat.shape[0]
    (head_fwd(labels, feat.double(), W, b) * g_recon.double().reshape(B, -1, 3)).sum().backward()
    return torch.cat([W.grad.reshape(-1), b.grad.reshape(-1)])
