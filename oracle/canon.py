"""TEST INFRASTRUCTURE — canonical fp32 evaluation order shared by oracle and CUDA kernels.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (ups_b200) never does.

Why this file exists
--------------------
The reference (TensorFlow 1.14) leaves the fp32 rounding sequence of `exp`, `log`,
`matmul` and `reduce_sum` to Eigen / cuBLAS, and differs between its own CPU and GPU
kernels by a few ulp.  A few ulp in a softmax is enough to flip `argmax` at a near-tie,
and a few ulp in a TPS source coordinate is enough to move a bilinear sample across a
`floor` boundary.  To make "labels bit-exact" and "identical sample positions" testable,
oracle and kernels evaluate those three places in ONE spelled-out order that uses only
IEEE-754 correctly rounded operations (+, -, *, /, floor, integer ops), never a fused
multiply-add and never a library transcendental:

* `exp_canon`  — Cephes `expf` (the polynomial Eigen's `pexp<float>` also uses), each
  multiply and add rounded separately.
* `log_canon`  — Cephes `logf` (Eigen `plog<float>`), same rule.
* softmax normalisation — p_k = e_k * fl(1/sum): ONE correctly rounded reciprocal per pixel,
  then one multiply per part (oracle/parts.py::_SoftmaxCanon).
* `sum_tree`   — K-way sum as a balanced adjacent-pair tree over the terms zero padded
  to a power of two (invariant under XOR permutations of the leaves, so 1, 2, 4 or 8
  lanes per pixel with a shuffle butterfly all produce the same bits).
* `solve_canon` — 11x11 Gauss-Jordan with first-max partial pivoting in float64,
  each multiply/subtract/divide rounded separately.

The CUDA side is unsupervised-part-segmentation_b200/csrc/canon_math.cuh; tests compare
the two bit for bit (host-compiled on CPU, device-compiled on the GPU box).

Every torch CPU elementwise kernel used here performs exactly one IEEE operation per
element (torch eager does not contract a*b+c), so the spelled-out order is what runs.
"""
import numpy as np
import torch

# ---------------------------------------------------------------- exp
_EXP_LO = -80.0  # below this exp_canon returns exactly 0 (keeps every intermediate normal)
_EXP_HI = 88.0
_LOG2EF = 1.44269504088896341
_EXP_C1 = 0.693359375
_EXP_C2 = -2.12194440e-4
_EXP_P = (1.9875691500e-4, 1.3981999507e-3, 8.3334519073e-3,
          4.1665795894e-2, 1.6666665459e-1, 5.0000001201e-1)


def _f32(v):
    return torch.tensor(v, dtype=torch.float32)


def exp_canon(x: torch.Tensor) -> torch.Tensor:
    """Canonical fp32 exp.  Spec (all ops fp32, rounded separately):
        zero = x < -80;  x = min(x, 88); x = max(x, -80)
        n  = floor(x*LOG2EF + 0.5)
        r  = (x - n*C1) - n*C2
        z  = r*r
        y  = ((((P0*r + P1)*r + P2)*r + P3)*r + P4)*r + P5
        y  = (y*z + r) + 1
        out = zero ? 0 : y * 2^n          (2^n built from exponent bits, product exact)
    """
    assert x.dtype == torch.float32
    zero = x < _EXP_LO
    x = torch.clamp(x, min=_EXP_LO, max=_EXP_HI)
    n = torch.floor(x * _f32(_LOG2EF) + _f32(0.5))
    r = x - n * _f32(_EXP_C1)
    r = r - n * _f32(_EXP_C2)
    z = r * r
    y = _f32(_EXP_P[0]) * r + _f32(_EXP_P[1])
    for c in _EXP_P[2:]:
        y = y * r + _f32(c)
    y = y * z + r
    y = y + _f32(1.0)
    two_n = ((n.to(torch.int32) + 127) << 23).view(torch.float32)
    out = y * two_n
    return torch.where(zero, torch.zeros_like(out), out)


# ---------------------------------------------------------------- log
_SQRTHF = 0.707106781186547524
_LOG_P = (7.0376836292e-2, -1.1514610310e-1, 1.1676998740e-1, -1.2420140846e-1,
          1.4249322787e-1, -1.6668057665e-1, 2.0000714765e-1, -2.4999993993e-1,
          3.3333331174e-1)
_LOG_Q1 = -2.12194440e-4
_LOG_Q2 = 0.693359375


def log_canon(x: torch.Tensor) -> torch.Tensor:
    """Canonical fp32 natural log for normal positive x.  Spec:
        x = m * 2^e, m in [0.5, 1)                       (bit extraction, exact)
        if m < SQRTHF: e -= 1; m = (m + m) - 1   else   m = m - 1
        z = m*m
        y = P0; for c in P1..P8: y = y*m + c
        y = (y*m)*z
        y = y + e*Q1
        y = y - 0.5*z
        out = (m + y) + e*Q2
    """
    assert x.dtype == torch.float32
    bits = x.view(torch.int32)
    e = ((bits >> 23) & 0xFF) - 126
    m = ((bits & 0x007FFFFF) | 0x3F000000).view(torch.float32)
    small = m < _f32(_SQRTHF)
    e = torch.where(small, e - 1, e)
    m = torch.where(small, (m + m) - _f32(1.0), m - _f32(1.0))
    ef = e.to(torch.float32)
    z = m * m
    y = torch.full_like(m, _LOG_P[0])
    for c in _LOG_P[1:]:
        y = y * m + _f32(c)
    y = (y * m) * z
    y = y + ef * _f32(_LOG_Q1)
    y = y - _f32(0.5) * z
    out = (m + y) + ef * _f32(_LOG_Q2)
    return out


# ---------------------------------------------------------------- K-way sum
def sum_tree(e: torch.Tensor) -> torch.Tensor:
    """Canonical sum over the last axis: zero pad to a power of two, add adjacent pairs
    until one value is left:  ((e0+e1)+(e2+e3)) + ((e4+e5)+(e6+e7)) ..."""
    K = e.shape[-1]
    p2 = 1
    while p2 < K:
        p2 *= 2
    if p2 != K:
        e = torch.cat([e, e.new_zeros(e.shape[:-1] + (p2 - K,))], dim=-1)
    while e.shape[-1] > 1:
        e = e[..., 0::2] + e[..., 1::2]
    return e[..., 0]


# ---------------------------------------------------------------- 11x11 solve (float64)
def solve_canon(A: np.ndarray, Bm: np.ndarray) -> np.ndarray:
    """Canonical Gauss-Jordan.  A [N,n,n], Bm [N,n,m] float64 -> X [N,n,m] float64.
    for c in 0..n-1:
        p = first argmax_{r>=c} |A[r,c]|; swap rows c,p (A and B)
        for every row r != c:  f = A[r,c] / A[c,c]
            A[r,j] = A[r,j] - f*A[c,j]  for j > c ;  B[r,:] = B[r,:] - f*B[c,:]
    X[c,:] = B[c,:] / A[c,c]
    """
    A = np.array(A, dtype=np.float64, copy=True)
    Bm = np.array(Bm, dtype=np.float64, copy=True)
    N, n, _ = A.shape
    idx = np.arange(N)
    for c in range(n):
        p = np.argmax(np.abs(A[:, c:, c]), axis=1) + c
        rc, rp = A[idx, c, :].copy(), A[idx, p, :].copy()
        A[idx, c, :], A[idx, p, :] = rp, rc
        bc, bp = Bm[idx, c, :].copy(), Bm[idx, p, :].copy()
        Bm[idx, c, :], Bm[idx, p, :] = bp, bc
        f = A[:, :, c] / A[:, c, c][:, None]          # [N,n]
        f[:, c] = 0.0
        if c + 1 < n:
            A[:, :, c + 1:] = A[:, :, c + 1:] - f[:, :, None] * A[:, c, c + 1:][:, None, :]
        Bm[:, :, :] = Bm - f[:, :, None] * Bm[:, c, :][:, None, :]
    d = np.stack([A[:, c, c] for c in range(n)], axis=1)
    return Bm / d[:, :, None]
