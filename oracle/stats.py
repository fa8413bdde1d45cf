"""TEST INFRASTRUCTURE — CPU restatement of the mask statistics the training step computes on the
part probabilities.  Pinned against tests/golden/stats.npz (the reference's
function bodies run under tf1_shim)."""
import torch


def probs_to_mu_sigma(probs, scaling_factor):
    """cub/code/nn.py:1541-1587.  probs [b,h,w,k]; scaling_factor [b,k] -> mu [b,k,2] (y, x),
    sigma [b,k,2,2]."""
    bn, h, w, nk = probs.shape
    y_t = torch.linspace(-1.0, 1.0, h).reshape(h, 1).repeat(1, w)
    x_t = torch.linspace(-1.0, 1.0, w).reshape(1, w).repeat(h, 1)
    meshgrid = torch.stack([y_t, x_t], dim=-1)                               # [h,w,2]
    mu = torch.einsum("ijl,aijk->akl", meshgrid, probs)
    mu_out_prod = torch.einsum("akm,akn->akmn", mu, mu)
    mesh_out_prod = torch.einsum("ijm,ijn->ijmn", meshgrid, meshgrid)
    stddev = torch.einsum("ijmn,aijk->akmn", mesh_out_prod, probs) - mu_out_prod
    sigma = (scaling_factor ** 2)[..., None, None] * stddev
    mu = mu * scaling_factor[..., None]
    return mu, sigma


def categorical_kl(probs):
    """cub/code/SB_model48i/model.py:21-25."""
    k = float(probs.shape[-1])
    logkp = torch.log(k * probs + 1e-20)
    kl = (probs * logkp).sum(dim=-1)
    return kl.mean()
