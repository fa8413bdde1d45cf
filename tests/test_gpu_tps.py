"""-m gpu: thin-plate-spline warp kernels against the oracle (bit-exact forward, thanks to the
shared canonical arithmetic) and against the reference-generated fixtures."""
import numpy as np
import pytest
import torch

from oracle import tps as OT
from util import own_error_atol, CUB_TPS, PENN_TPS, assert_bitexact, assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ups():
    import ups_b200
    return ups_b200


def _case(N, S, C, kw, seed):
    g = torch.Generator().manual_seed(seed)
    prm = OT.tps_parameters(N, generator=g, **kw)
    U = torch.rand(N, S, S, C, generator=g) * 2 - 1
    return prm, U, g


def test_make_input_tps_param_bitexact(ups):
    prm, _, _ = _case(67, 8, 3, CUB_TPS, 1)
    coord_o, tv_o = OT.make_input_tps_param(prm)
    dprm = ups.tps.DotMap({k: v.cuda() for k, v in prm.items()})
    coord_c, tv_c = ups.make_input_tps_param(dprm)
    assert_bitexact(coord_c, coord_o, "coord")
    assert_bitexact(tv_c, tv_o, "t_vector")


@pytest.mark.parametrize("kw,N,S,C", [(CUB_TPS, 24, 128, 3), (PENN_TPS, 5, 96, 3),
                                      (dict(CUB_TPS, tps_scal=0.6, off_scal=0.6), 7, 64, 3),
                                      (CUB_TPS, 3, 33, 1), (CUB_TPS, 2, 40, 5)])
def test_warp_forward_bitexact(ups, kw, N, S, C):
    prm, U, _ = _case(N, S, C, kw, 7)
    coord, tv = OT.make_input_tps_param(prm)
    T_o, _ = OT.tps_system(coord.flip(-1), tv.flip(-1))
    assert_bitexact(ups.ops.tps_solve(coord.cuda(), tv.cuda()), T_o, "T (11x11 solve)")
    out_o, mesh_o = OT.ThinPlateSpline(U, coord, tv, S, C)
    out_c, mesh_c = ups.ThinPlateSpline(U.cuda(), coord.cuda(), tv.cuda(), S, C)
    assert_bitexact(mesh_c, mesh_o, "t_arr")
    assert_bitexact(out_c, out_o, "warped output")


def test_warp_out_size_differs_from_input(ups):
    prm, U, _ = _case(3, 48, 3, CUB_TPS, 11)
    coord, tv = OT.make_input_tps_param(prm)
    out_o, mesh_o = OT.ThinPlateSpline(U, coord, tv, 32, 3)
    out_c, mesh_c = ups.ThinPlateSpline(U.cuda(), coord.cuda(), tv.cuda(), 32, 3)
    assert list(out_c.shape) == [3, 32, 32, 3]
    assert_bitexact(mesh_c, mesh_o, "t_arr")
    assert_bitexact(out_c, out_o, "warped output")


def test_warp_move_scal_branch(ups):
    prm, U, g = _case(4, 32, 3, CUB_TPS, 5)
    mp = torch.rand(4, 1, 2, generator=g) * 0.2 - 0.1
    sp = torch.rand(4, 2, generator=g) * 0.2 + 0.9
    coord, tv = OT.make_input_tps_param(prm)
    out_o, mesh_o = OT.ThinPlateSpline(U, coord, tv, 32, 3, move=mp, scal=sp)
    out_c, mesh_c = ups.ThinPlateSpline(U.cuda(), coord.cuda(), tv.cuda(), 32, 3, move=mp.cuda(), scal=sp.cuda())
    assert_bitexact(mesh_c, mesh_o, "t_arr (move/scal)")
    assert_bitexact(out_c, out_o, "warped output (move/scal)")


def test_identity_and_out_of_range_known_answers(ups):
    base = torch.tensor([OT._BASE]).repeat(2, 1, 1)
    U = torch.rand(2, 16, 16, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1
    out, mesh = ups.ThinPlateSpline(U.cuda(), base.cuda(), torch.zeros(2, 8, 2).cuda(), 16, 3)
    out, mesh = out.cpu(), mesh.cpu()
    assert out[:, -1].abs().max() < 1e-5 and out[:, :, -1].abs().max() < 1e-5
    lin = torch.linspace(-1, 1, 16)
    assert (mesh[..., 1] - lin[None, None, :]).abs().max() < 1e-5
    # a warp that throws every sample far outside the image -> ~0 everywhere          SURVEY 8c(2)
    far = torch.full((2, 8, 2), 5.0)
    out2, _ = ups.ThinPlateSpline(U.cuda(), base.cuda(), far.cuda(), 16, 3)
    assert out2.abs().max().item() < 1e-5


def test_warp_backward_matches_autograd(ups):
    for kw, N, S in ((CUB_TPS, 6, 64), (dict(CUB_TPS, tps_scal=0.6, off_scal=0.6), 4, 32)):
        prm, U, g = _case(N, S, 3, kw, 13)
        coord, tv = OT.make_input_tps_param(prm)
        G = torch.randn(N, S, S, 3, generator=g)
        Uo = U.clone().requires_grad_(True)
        out_o, _ = OT.ThinPlateSpline(Uo, coord, tv, S, 3)
        (dU_o,) = torch.autograd.grad(out_o, Uo, G)
        Uc = U.cuda().requires_grad_(True)
        out_c, _ = ups.ThinPlateSpline(Uc, coord.cuda(), tv.cuda(), S, 3)
        (dU_c,) = torch.autograd.grad(out_c, Uc, G.cuda())
        # scatter-add order differs (atomics): tolerance, not bits.  The absolute part is the north-star 1e-5 plus
        # twice the fp32 oracle's own distance from the float64 scatter of the same weights (util.own_error_atol)
        dU64 = OT.warp_grad_fp64(U, coord, tv, S, G)
        assert_close(dU_c, dU64, "dU", rtol=1e-4, atol=own_error_atol(dU_o, dU64))


def test_against_reference_generated_fixtures(ups, golden):
    """tests/golden/tps_*.npz come from the reference's own ThinPlateSpline under the shim, with its fp32 matrix inverse
    and (``_inv64``) with the inverse in float64.  Shipped parameter ranges: every pixel within 1e-4 rel / 1e-5 abs of
    both; exaggerated / identity warps: the derived bound of util.check_tps_against_fixture."""
    from util import check_tps_against_fixture
    for tag, strict in (("cub", True), ("penn", True), ("big", False), ("identity", False)):
        gd = golden(f"tps_{tag}.npz")
        U = torch.from_numpy(gd["U"])
        S = U.shape[1]
        out, mesh = ups.ThinPlateSpline(U.cuda(), torch.from_numpy(gd["coord"]).cuda(),
                                        torch.from_numpy(gd["t_vector"]).cuda(), S, U.shape[3])
        refs = [("fp32 inverse", gd)] + ([] if tag == "identity" else [("fp64 inverse", golden(f"tps_{tag}_inv64.npz"))])
        for name, r in refs:
            check_tps_against_fixture(out, mesh, torch.from_numpy(r["out"]), torch.from_numpy(r["t_arr"]), S, strict,
                                      f"{tag} / {name}")
            if "dU" in r and tag != "big":       # backward through the bilinear sampling (scatter-add)
                Uc = U.cuda().requires_grad_(True)
                o2, m2 = ups.ThinPlateSpline(Uc, torch.from_numpy(gd["coord"]).cuda(), torch.from_numpy(gd["t_vector"]).cuda(),
                                             S, U.shape[3])
                G = torch.from_numpy(gd["G"])
                (dU,) = torch.autograd.grad(o2, Uc, G.cuda())
                # float64 inverse: north-star tolerance on every element.  fp32 inverse (the reference's own,
                # transformations.py:228): the bilinear weights move with the sample position, so the measured
                # position difference (in pixels) times the largest cotangent is added -- 2 of 3072 elements need it
                extra = 0.0 if name == "fp64 inverse" else \
                    4.0 * float(G.abs().max()) * float((m2.detach().cpu() - torch.from_numpy(r["t_arr"])).abs().max()) * S / 2
                assert_close(dU, torch.from_numpy(r["dU"]), f"{tag} / {name} dU", rtol=1e-4, atol=1e-5 + extra)


def test_make_tps_three_and_two_views(ups):
    B, S = 4, 32
    g = torch.Generator().manual_seed(3)
    views = [torch.rand(B, S, S, 3, generator=g) * 2 - 1 for _ in range(3)]
    kw = {k: v for k, v in CUB_TPS.items()}
    out_c = ups.make_tps([v.cuda() for v in views], kw, generator=torch.Generator().manual_seed(42))
    out_o = OT.make_tps(views, kw, generator=torch.Generator().manual_seed(42))
    assert len(out_c) == 3
    for a, b in zip(out_c, out_o):
        assert_bitexact(a, b, "make_tps view")
    out_c2 = ups.make_tps([v.cuda() for v in views[:2]], kw, generator=torch.Generator().manual_seed(42))
    assert len(out_c2) == 2
    assert_bitexact(out_c2[1], out_o[1], "make_tps (2 views)")


def test_warp_parameter_gradients_raise(ups):
    """gradients with respect to the warp parameters or through the mesh are an error, not a silent zero."""
    from ups_b200._cabi import UpsError
    prm, U, g = _case(2, 16, 3, CUB_TPS, 1)
    coord, tv = OT.make_input_tps_param(prm)
    with pytest.raises(UpsError, match="must not require grad"):
        ups.ThinPlateSpline(U.cuda(), coord.cuda().requires_grad_(True), tv.cuda(), 16, 3)
    Uc = U.cuda().requires_grad_(True)
    out, mesh = ups.ThinPlateSpline(Uc, coord.cuda(), tv.cuda(), 16, 3)
    with pytest.raises(Exception, match="sampling mesh"):
        torch.autograd.grad(mesh.sum() + out.sum(), Uc)
    (dU,) = torch.autograd.grad(ups.ThinPlateSpline(Uc, coord.cuda(), tv.cuda(), 16, 3)[0].sum(), Uc)
    assert dU.shape == Uc.shape
