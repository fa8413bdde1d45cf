"""canon_math.cuh (host build) must reproduce oracle/canon.py and oracle/tps.py BIT FOR BIT.
The same header is what the CUDA kernels compile, so this pins the kernel arithmetic on CPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import canon, parts as P, tps as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "unsupervised-part-segmentation_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(CSRC, "libups_canon_host.so")
    src = os.path.join(CSRC, "canon_host.cpp")
    hdr = os.path.join(CSRC, "canon_math.cuh")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so])
    return ctypes.CDLL(so)


def fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def test_exp_bitexact(lib):
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.randn(200000, generator=g) * 20, torch.linspace(-90, 90, 50001),
                   torch.tensor([0.0, -0.0, -80.0, -80.00001, 88.0, 100.0, -1e30])]).numpy()
    y = np.empty_like(x)
    lib.ups_host_exp(fp(x), fp(y), ctypes.c_longlong(x.size))
    ref = canon.exp_canon(torch.from_numpy(x)).numpy()
    assert np.array_equal(bits(y), bits(ref))


def test_log_bitexact(lib):
    g = torch.Generator().manual_seed(1)
    x = torch.cat([torch.rand(200000, generator=g) * 16 + 1e-6, torch.logspace(-6, 3, 50001)]).numpy()
    y = np.empty_like(x)
    lib.ups_host_log(fp(x), fp(y), ctypes.c_longlong(x.size))
    ref = canon.log_canon(torch.from_numpy(x)).numpy()
    assert np.array_equal(bits(y), bits(ref))
    assert np.abs(y - np.log(x.astype(np.float64))).max() < 1e-6


@pytest.mark.parametrize("K", [1, 2, 3, 4, 8, 16, 25, 32, 64])
def test_softmax_bitexact(lib, K):
    g = torch.Generator().manual_seed(K)
    x = (torch.randn(4096, K, generator=g) * 3)
    x[:64] = torch.round(x[:64])            # plenty of exact ties
    x[64:96] = x[64:96] * 40                # extreme logits
    xn = x.numpy()
    p = np.empty_like(xn)
    hard = np.empty_like(xn)
    lab = np.empty(xn.shape[0], np.int64)
    lib.ups_host_softmax(fp(xn), fp(p), fp(lab), fp(hard), ctypes.c_longlong(xn.shape[0]), K)
    pr = P.softmax(x)
    assert np.array_equal(bits(p), bits(pr.numpy()))
    assert np.array_equal(lab, torch.argmax(pr, -1).numpy())
    hr = P.straight_through_estimator(P.hard_max(pr, 1), pr)
    assert np.array_equal(bits(hard), bits(hr.numpy()))


@pytest.mark.parametrize("K", [3, 5, 12, 25, 31])
def test_minus_infinity_padding_gives_the_same_bits(lib, K):
    """What the padded part counts rest on: the canonical softmax of [.,K] logits and of the same
    logits padded with -inf to the next power of two agree bit for bit on the first K parts (exp_canon(-inf) == 0
    exactly, the pair-tree sum is defined on zero-padded terms), the padding parts get probability 0 and are never the
    arg-max -- in the oracle and in the host build of the header the kernels compile."""
    Kp = 1 << (K - 1).bit_length()
    g = torch.Generator().manual_seed(100 + K)
    x = torch.randn(2048, K, generator=g) * 3
    x[:64] = torch.round(x[:64])
    x[64:96] = x[64:96] * 40
    xp = torch.full((2048, Kp), float("-inf"))
    xp[:, :K] = x
    assert float(canon.exp_canon(torch.tensor([float("-inf")]))[0]) == 0.0
    pr, prp = P.softmax(x), P.softmax(xp)
    assert np.array_equal(bits(prp[:, :K].numpy()), bits(pr.numpy())) and bool((prp[:, K:] == 0).all())
    assert torch.equal(torch.argmax(prp, -1), torch.argmax(pr, -1))
    xn = np.ascontiguousarray(xp.numpy())
    p = np.empty_like(xn)
    hard = np.empty_like(xn)
    lab = np.empty(xn.shape[0], np.int64)
    lib.ups_host_softmax(fp(xn), fp(p), fp(lab), fp(hard), ctypes.c_longlong(xn.shape[0]), Kp)
    assert np.array_equal(bits(p[:, :K]), bits(pr.numpy())) and not p[:, K:].any() and not hard[:, K:].any()
    assert np.array_equal(lab, torch.argmax(pr, -1).numpy())
    hr = P.straight_through_estimator(P.hard_max(pr, 1), pr)
    assert np.array_equal(bits(hard[:, :K]), bits(hr.numpy()))


def _params(N, seed, **kw):
    g = torch.Generator().manual_seed(seed)
    base = dict(scal=0.8, tps_scal=0.15, rot_scal=0.2, off_scal=0.2, scal_var=0.1, rescal=1.0)
    base.update(kw)
    return T.tps_parameters(N, generator=g, **base), g


def test_tps_input_param_bitexact(lib):
    prm, _ = _params(64, 3)
    coord, tv = T.make_input_tps_param(prm)
    out = np.empty((64, 8, 2), np.float32)
    a = {k: np.ascontiguousarray(v.numpy()) for k, v in prm.items()}
    lib.ups_host_tps_input_param(fp(a["coord"]), fp(a["vector"]), fp(a["offset"]), fp(a["offset_2"]),
                                 fp(a["t_scal"]), fp(a["rot_mat"]), fp(out), 64)
    assert np.array_equal(bits(out), bits(tv.numpy()))


@pytest.mark.parametrize("kw", [dict(), dict(tps_scal=0.6, off_scal=0.6), dict(scal=0.95, tps_scal=0.08, rot_scal=0.05)])
def test_tps_solve_and_warp_bitexact(lib, kw):
    N, S, C = 16, 32, 3
    prm, g = _params(N, 7, **kw)
    coord, tv = T.make_input_tps_param(prm)
    Tm, _ = T.tps_system(coord.flip(-1), tv.flip(-1))
    cn, vn = np.ascontiguousarray(coord.numpy()), np.ascontiguousarray(tv.numpy())
    Th = np.empty((N, 2, 11), np.float32)
    lib.ups_host_tps_solve(fp(cn), fp(vn), fp(Th), N)
    assert np.array_equal(bits(Th), bits(Tm.numpy()))
    U = (torch.rand(N, S, S, C, generator=g) * 2 - 1)
    out_o, mesh_o = T.ThinPlateSpline(U, coord, tv, S, C)
    Un = np.ascontiguousarray(U.numpy())
    out = np.empty((N, S, S, C), np.float32)
    mesh = np.empty((N, S, S, 2), np.float32)
    lib.ups_host_tps_warp(fp(Un), fp(cn), fp(Th), fp(out), fp(mesh), N, S, S, C, S, S)
    assert np.array_equal(bits(mesh), bits(mesh_o.numpy()))
    assert np.array_equal(bits(out), bits(out_o.numpy()))
