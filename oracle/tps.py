"""TEST INFRASTRUCTURE — CPU restatement of the reference's thin-plate-spline warp.

Follows baselines/unsupervised-disentangling/transformations.py (the vendored copy of the
un-vendored `eddata.utils.tps` that cub/code/SB_model48i/model.py:5,300-309 calls):
  tps_parameters        transformations.py:17-39
  make_input_tps_param  transformations.py:59-77
  ThinPlateSpline       transformations.py:93-244
      _solve_system :215-235, _meshgrid :171-188, _transform :190-213, _interpolate :114-169

Arithmetic is fp32 in the canonical order of oracle/canon.py (the reference leaves the
order to TF's kernels); the 11x11 system is solved in float64 (the reference inverts in
fp32; cond(W) is 15-460 so the fp32 inverse carries ~1e-5 relative noise that no second
implementation can reproduce).  Pinned against tests/golden/tps_*.npz, which were produced
by executing the reference's own function bodies under tests/golden/tf1_shim.py.
"""
import math
import numpy as np
import torch

from .canon import log_canon, solve_canon, _f32

_BASE = [[-0.5, -0.5], [0.5, -0.5], [-0.5, 0.5], [0.5, 0.5],
         [0.2, -0.2], [-0.2, 0.2], [0.2, 0.2], [-0.2, -0.2]]


class AttrDict(dict):
    """attribute-access dict (the reference returns a DotMap, transformations.py:38)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def tps_parameters(batch_size, scal, tps_scal, rot_scal, off_scal, scal_var, rescal=1,
                   augm_scal=None, generator=None):
    """transformations.py:17-39.  `augm_scal` is the kwarg name the shipped config uses
    (cub/code/SB_model48i/train_cub_subset_tps.yaml:194); it aliases `rescal`."""
    if augm_scal is not None:
        rescal = augm_scal
    g = generator

    def U(shape, lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    coord = torch.tensor([_BASE], dtype=torch.float32).repeat(batch_size, 1, 1)
    coord = coord + U(coord.shape, -0.2, 0.2)
    vector = U(coord.shape, -tps_scal, tps_scal)
    offset = U((batch_size, 1, 2), -off_scal, off_scal)
    offset_2 = U((batch_size, 1, 2), -off_scal, off_scal)
    t_scal = U((batch_size, 2), scal * (1.0 - scal_var), scal * (1.0 + scal_var))
    t_scal = t_scal * rescal
    rot = U((batch_size, 1), -rot_scal, rot_scal)
    a, b = torch.cos(rot), torch.sin(rot)
    rot_mat = torch.stack([torch.cat([a, -b], 1), torch.cat([b, a], 1)], 1)  # [B,2,2]
    return AttrDict(coord=coord, vector=vector, offset=offset, offset_2=offset_2,
                    t_scal=t_scal, rot_mat=rot_mat)


def make_input_tps_param(tps_param, move_point=None, scal_point=None):
    """transformations.py:59-77.  Canonical order:
        s  = t_scal[b,k] * ((coord + vector) - offset) + offset
        u  = s - offset_2
        tv = ((rot[b,l,0]*u[...,0] + rot[b,l,1]*u[...,1]) + offset_2[...,l]) - coord[...,l]
    """
    coord, vector = tps_param.coord, tps_param.vector
    offset, offset_2 = tps_param.offset, tps_param.offset_2
    rot_mat, t_scal = tps_param.rot_mat, tps_param.t_scal
    s = t_scal[:, None, :] * ((coord + vector) - offset) + offset
    u = s - offset_2
    tv = (rot_mat[:, None, :, 0] * u[..., 0:1] + rot_mat[:, None, :, 1] * u[..., 1:2])
    t_vector = (tv + offset_2) - coord
    if move_point is not None and scal_point is not None:
        coord = scal_point[:, None, :] * (coord + move_point)
        t_vector = scal_point[:, None, :] * t_vector
    else:
        assert move_point is None and scal_point is None
    return coord, t_vector


def tps_system(coord, vector):
    """_solve_system (:215-235) on already flipped coord/vector.  Returns T [B,2,n+3] fp32.
    W is built in fp32 exactly as the reference builds it (with log_canon for tf.log);
    the solve is the canonical float64 Gauss-Jordan, rounded to fp32 once at the end."""
    B, n, _ = coord.shape
    ones = torch.ones(B, n, 1, dtype=torch.float32)
    p = torch.cat([ones, coord], 2)                                     # [B,n,3]
    diff = p[:, :, None, :] - p[:, None, :, :]                          # [B,n,n,3]
    sq = diff * diff
    d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    r = d2 * log_canon(d2 + _f32(1e-6))
    W0 = torch.cat([p, r], 2)                                           # [B,n,n+3]
    W1 = torch.cat([torch.zeros(B, 3, 3), p.transpose(1, 2)], 2)        # [B,3,n+3]
    W = torch.cat([W0, W1], 1)                                          # [B,n+3,n+3]
    tp = torch.cat([coord + vector, torch.zeros(B, 3, 2)], 1)           # [B,n+3,2]
    X = solve_canon(W.double().numpy(), tp.double().numpy())            # [B,n+3,2]
    T = torch.from_numpy(X).to(torch.float32).transpose(1, 2).contiguous()
    return T, W


def tps_grid_coords(T, coord, out_h, out_w):
    """_meshgrid (:171-188) + T.grid (:196-198), canonical left-to-right accumulation:
        x_j = -1 + j*step_w, y_i = -1 + i*step_h, step = 2/(n-1)        (tf.linspace)
        r_n = d2*log(d2 + 1e-6),  d2 = (x-px)^2 + (y-py)^2
        acc = T[.,0]; acc += T[.,1]*x; acc += T[.,2]*y; acc += T[.,3+n]*r_n  (n ascending)
    Returns x_s, y_s [B, out_h, out_w]."""
    B, n, _ = coord.shape
    step_w = _f32(2.0) / _f32(float(out_w - 1))
    step_h = _f32(2.0) / _f32(float(out_h - 1))
    xs = _f32(-1.0) + torch.arange(out_w, dtype=torch.float32) * step_w
    ys = _f32(-1.0) + torch.arange(out_h, dtype=torch.float32) * step_h
    x_t = xs[None, None, :].expand(1, out_h, out_w)
    y_t = ys[None, :, None].expand(1, out_h, out_w)
    outs = []
    for d in range(2):
        acc = T[:, d, 0][:, None, None] + T[:, d, 1][:, None, None] * x_t
        acc = acc + T[:, d, 2][:, None, None] * y_t
        outs.append(acc)
    for k in range(n):
        dx = x_t - coord[:, k, 0][:, None, None]
        dy = y_t - coord[:, k, 1][:, None, None]
        d2 = dx * dx + dy * dy
        r = d2 * log_canon(d2 + _f32(1e-6))
        for d in range(2):
            outs[d] = outs[d] + T[:, d, 3 + k][:, None, None] * r
    return outs[0], outs[1]


def bilinear_sample(U, x_s, y_s):
    """_interpolate (:114-169): pixel coords (v+1)*size/2, floor, +1, clip all four indices,
    weights from the CLIPPED indices, out = ((wa*Ia + wb*Ib) + wc*Ic) + wd*Id."""
    B, H, W, C = U.shape
    X = ((x_s + _f32(1.0)) * _f32(float(W))) / _f32(2.0)
    Y = ((y_s + _f32(1.0)) * _f32(float(H))) / _f32(2.0)
    x0 = torch.floor(X).to(torch.int64)
    y0 = torch.floor(Y).to(torch.int64)
    x1, y1 = x0 + 1, y0 + 1
    x0, x1 = x0.clamp(0, W - 1), x1.clamp(0, W - 1)
    y0, y1 = y0.clamp(0, H - 1), y1.clamp(0, H - 1)
    b = torch.arange(B)[:, None, None]
    Ia, Ib = U[b, y0, x0], U[b, y1, x0]
    Ic, Id = U[b, y0, x1], U[b, y1, x1]
    x0f, x1f, y0f, y1f = (t.to(torch.float32) for t in (x0, x1, y0, y1))
    wa = ((x1f - X) * (y1f - Y))[..., None]
    wb = ((x1f - X) * (Y - y0f))[..., None]
    wc = ((X - x0f) * (y1f - Y))[..., None]
    wd = ((X - x0f) * (Y - y0f))[..., None]
    return ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id


def ThinPlateSpline(U, coord, vector, out_size, n_c, move=None, scal=None):
    """transformations.py:93-244.  Returns (output [B,S,S,C], t_arr [B,S,S,2] = (y, x))."""
    coord = coord.flip(-1)
    vector = vector.flip(-1)
    out_size = int(out_size)
    assert U.shape[-1] == int(n_c)
    T, _ = tps_system(coord.detach(), vector.detach())
    x_s, y_s = tps_grid_coords(T, coord.detach(), out_size, out_size)
    if move is not None and scal is not None:                      # :202-208
        y_s = y_s * scal[:, 0][:, None, None] + move[:, :, 0][:, :, None]
        x_s = x_s * scal[:, 1][:, None, None] + move[:, :, 1][:, :, None]
    else:
        assert move is None and scal is None
    out = bilinear_sample(U, x_s, y_s)
    t_arr = torch.stack([y_s, x_s], -1)
    return out, t_arr


def warp_grad_fp64(U, coord, vector, out_size, G, move=None, scal=None):
    """dL/dU of ThinPlateSpline for the cotangent G, with the SAME fp32 sample positions and bilinear weights as the
    fp32 path but float64 products and float64 scatter-add: the exact value that every fp32 accumulation order (the
    oracle's autograd, the kernel's atomics) approximates.  Tests derive their scatter tolerance from it."""
    coord_f, vector_f = coord.flip(-1), vector.flip(-1)
    T, _ = tps_system(coord_f, vector_f)
    x_s, y_s = tps_grid_coords(T, coord_f, int(out_size), int(out_size))
    if move is not None and scal is not None:
        y_s = y_s * scal[:, 0][:, None, None] + move[:, :, 0][:, :, None]
        x_s = x_s * scal[:, 1][:, None, None] + move[:, :, 1][:, :, None]
    U64 = U.detach().double().requires_grad_(True)
    out = bilinear_sample(U64, x_s, y_s)
    (dU,) = torch.autograd.grad(out, U64, G.double())
    return dU


def make_tps_given(views, coord, vector):
    """The warp half of TrainModel.make_tps for given (coord, t_vector) of 2B samples:
    views[0:2] concatenated use rows [0,2B); views[2] (the target) re-uses rows [0,B)
    (cub/code/SB_model48i/model.py:298-310)."""
    bs = views[0].shape[0]
    img_batch = torch.cat(views[:2], 0)
    t_images, _ = ThinPlateSpline(img_batch, coord, vector, img_batch.shape[1],
                                  img_batch.shape[-1])
    out = list(torch.split(t_images, bs, 0))
    if len(views) > 2:
        t3, _ = ThinPlateSpline(views[2], coord[:bs], vector[:bs], views[2].shape[1],
                                views[2].shape[-1])
        out.append(t3)
    return out


def make_tps(views, tps_parameters_kwargs, generator=None):
    """TrainModel.make_tps — cub/code/SB_model48i/model.py:282-311 (3 views: the target view
    re-uses the first-half parameters) and pennaction/code/SB_model48i/model.py:281-303 (2)."""
    n = views[0].shape[0] * 2
    tps_params = tps_parameters(n, generator=generator, **tps_parameters_kwargs)
    coord, vector = make_input_tps_param(tps_params)
    return make_tps_given(views, coord, vector)
