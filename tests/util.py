"""Shared helpers for the parity tests: seeded synthetic inputs on the CPU,
identical tensors are fed to the oracle and (copied to cuda:0) to the kernels."""
import numpy as np
import torch

from oracle import tps as OT

RTOL, ATOL = 1e-4, 1e-5       # BASELINE.json north_star: fp32 within 1e-4 relative / 1e-5 absolute
CUB_TPS = dict(scal=0.8, tps_scal=0.15, rot_scal=0.2, off_scal=0.2, scal_var=0.1, augm_scal=1.0)
PENN_TPS = dict(scal=0.95, tps_scal=0.08, rot_scal=0.05, off_scal=0.2, scal_var=0.05, augm_scal=1.0)


def make_inputs(B, S, K, F, V=3, seed=0, tps=CUB_TPS, ties=False, tps_seed=1234):
    g = torch.Generator().manual_seed(seed)
    views = torch.rand(V, B, S, S, 3, generator=g) * 2 - 1
    l0 = torch.randn(B, S, S, K, generator=g)
    l1 = torch.randn(B, S, S, K, generator=g)
    if ties:  # exact ties and extreme logits in the first rows
        l0[:, 0] = torch.round(l0[:, 0])
        l1[:, 0] = torch.round(l1[:, 0])
        l0[:, 1] = l0[:, 1] * 40
        l1[:, 1, :, :] = 0.0
    feat = torch.randn(B, K, F, generator=g)
    gt = torch.Generator().manual_seed(tps_seed)
    prm = OT.tps_parameters(2 * B, generator=gt, **tps)
    coord, t_vector = OT.make_input_tps_param(prm)
    cot = dict(g_inj=torch.randn(B, S, S, F + K, generator=g), g_parts=torch.randn(K * B, S, S, 3, generator=g),
               g_pooled=torch.randn(B, K, 3, generator=g), g_m0=torch.randn(B, S, S, K, generator=g),
               g_m1=torch.randn(B, S, S, K, generator=g),
               g_warped=[torch.randn(B, S, S, 3, generator=g) for _ in range(V)])
    return dict(views=views, l0=l0, l1=l1, feat=feat, coord=coord, t_vector=t_vector, prm=prm, cot=cot)


def reduce_atol(n_terms, atol=ATOL):
    """Absolute tolerance for an fp32 reduction over `n_terms` O(1) terms (dfeat sums P/K pixels
    per entry): two valid fp32 summation orders - the oracle's and the kernel's - differ by
    ~eps*sqrt(n)*|term| on entries that cancel to ~0, where a relative bound is meaningless.
    1e-5 up to 1024 terms, growing with sqrt(n) beyond."""
    return atol * max(1.0, (n_terms / 1024.0) ** 0.5)


def cuda(x):
    if isinstance(x, torch.Tensor):
        return x.cuda().contiguous()
    if isinstance(x, dict):
        return {k: cuda(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [cuda(v) for v in x]
    return x


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int32)


def assert_bitexact(got, want, what=""):
    a, b = bits(got), bits(want)
    if not torch.equal(a, b):
        d = (got.detach().cpu().float() - want.float()).abs()
        n = (a != b).sum().item()
        raise AssertionError(f"{what}: {n}/{a.numel()} elements differ bitwise, max abs diff {d.max().item():.3e}")


def assert_close(got, want, what="", rtol=RTOL, atol=ATOL):
    g = got.detach().cpu().double()
    w = want.detach().double()
    assert g.shape == w.shape, (what, g.shape, w.shape)
    err = (g - w).abs()
    tol = atol + rtol * w.abs()
    if not bool((err <= tol).all()):
        i = torch.argmax(err - tol)
        raise AssertionError(f"{what}: {(err > tol).sum().item()}/{err.numel()} outside rtol={rtol} atol={atol}; "
                             f"worst |d|={err.flatten()[i].item():.3e} at ref={w.flatten()[i].item():.6e}")
