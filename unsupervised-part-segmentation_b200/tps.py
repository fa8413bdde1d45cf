"""Drop-in for `eddata.utils.tps` as the reference calls it (cub/code/SB_model48i/model.py:5,
300-309); the vendored source of that module is
baselines/unsupervised-disentangling/transformations.py:17-77,93-244."""
import torch

from . import ops

_BASE = [[-0.5, -0.5], [0.5, -0.5], [-0.5, 0.5], [0.5, 0.5],
         [0.2, -0.2], [-0.2, 0.2], [0.2, 0.2], [-0.2, -0.2]]


class DotMap(dict):
    """attribute-access dict standing in for dotmap.DotMap (transformations.py:3,38)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def tps_parameters(batch_size, scal, tps_scal, rot_scal, off_scal, scal_var, rescal=1, augm_scal=None,
                   generator=None, device="cuda"):
    """transformations.py:17-39.  ~50 random floats per sample, drawn on the host from
    `generator` (the reference draws them with tf.random_uniform) and moved to `device`.
    `augm_scal` is the kwarg the shipped config uses for `rescal`
    (cub/code/SB_model48i/train_cub_subset_tps.yaml:194)."""
    if augm_scal is not None:
        rescal = augm_scal

    def U(shape, lo, hi):
        return torch.rand(shape, generator=generator, dtype=torch.float32) * (hi - lo) + lo

    coord = torch.tensor([_BASE], dtype=torch.float32).repeat(batch_size, 1, 1)
    coord = coord + U(coord.shape, -0.2, 0.2)
    vector = U(coord.shape, -tps_scal, tps_scal)
    offset = U((batch_size, 1, 2), -off_scal, off_scal)
    offset_2 = U((batch_size, 1, 2), -off_scal, off_scal)
    t_scal = U((batch_size, 2), scal * (1.0 - scal_var), scal * (1.0 + scal_var)) * rescal
    rot = U((batch_size, 1), -rot_scal, rot_scal)
    a, b = torch.cos(rot), torch.sin(rot)
    rot_mat = torch.stack([torch.cat([a, -b], 1), torch.cat([b, a], 1)], 1)
    d = dict(coord=coord, vector=vector, offset=offset, offset_2=offset_2, t_scal=t_scal, rot_mat=rot_mat)
    return DotMap({k: v.to(device, non_blocking=True) for k, v in d.items()})


def make_input_tps_param(tps_param, move_point=None, scal_point=None):
    """transformations.py:59-77 -> (coord, t_vector)."""
    coord = tps_param.coord
    t_vector = ops.tps_input_param(coord, tps_param.vector, tps_param.offset, tps_param.offset_2,
                                   tps_param.t_scal, tps_param.rot_mat)
    if move_point is not None and scal_point is not None:
        # crop branch (:70-72): two [B,8,2] broadcasts, outside the hot path -> plain torch
        coord = scal_point[:, None, :] * (coord + move_point)
        t_vector = scal_point[:, None, :] * t_vector
    else:
        assert move_point is None and scal_point is None
    return coord, t_vector


def ThinPlateSpline(U, coord, vector, out_size, n_c, move=None, scal=None):
    """transformations.py:93-244 -> (output [B,out,out,C], t_arr [B,out,out,2] = (y, x))."""
    assert U.dim() == 4 and U.shape[-1] == int(n_c), (list(U.shape), n_c)
    assert coord.shape == vector.shape and coord.shape[0] == U.shape[0] and tuple(coord.shape[1:]) == (8, 2), \
        list(coord.shape)
    assert (move is None) == (scal is None)
    # Gradients flow to the images U only.  The reference draws coord / vector / move / scal at random every step
    # (cub/code/SB_model48i/model.py:298-300): they are data, never variables.  A caller that asks for their gradient
    # gets an error here rather than a silently missing one (the TF graph would have propagated it).
    if any(t is not None and t.requires_grad for t in (coord, vector, move, scal)):
        raise ops.C.UpsError("ups_b200.ThinPlateSpline: coord / vector / move / scal must not require grad "
                             "(no gradient with respect to the warp parameters is implemented); detach them")
    T = ops.tps_solve(coord, vector)
    out, mesh = ops.tps_warp(U, coord, T, int(out_size), move, scal)
    return out, mesh
