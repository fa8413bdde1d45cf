"""CPU tier: the C-ABI library loads and exports every symbol include/ups_b200.h declares, the
Python mirror keeps the reference's helper signatures, and shape errors are AssertionErrors
raised before any kernel is touched.  No compute calls (there is no GPU here)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ups():
    import __graft_entry__ as ge
    ge.build()
    import ups_b200
    return ups_b200


def test_every_declared_symbol_is_exported(ups):
    hdr = open(os.path.join(ROOT, "include", "ups_b200.h")).read()
    names = sorted(set(re.findall(r"\b(ups_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30, names
    lib = ctypes.CDLL(ups._cabi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ups_b200.h but not exported"
    bound = set(ups._cabi._SIGS) | {"ups_version", "ups_last_error_string", "ups_launch_count",
                                    "ups_launch_count_reset", "ups_workspace_bytes",
                                    "ups_inject_conv_workspace_bytes", "ups_parts_conv_bwd_workspace_bytes",
                                    "ups_standin_workspace_bytes", "ups_dp_allreduce_signal_bytes"}
    assert set(names) == bound, set(names) ^ bound
    assert "sm_100a" in ups._cabi.version()


def test_library_records_the_hash_of_its_sources(ups):
    """the binary that travels to the GPU box must be the one built from the tree's sources."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "unsupervised-part-segmentation_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.built_hash() == b.source_hash()
    assert f"src={b.source_hash()}" in ups._cabi.version()


def test_library_is_sm100a_only(ups):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", ups._cabi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_workspace_query_and_error_paths(ups):
    C = ups._cabi
    assert C.workspace_bytes(C.OP_STEP, 256, 16384, 16, 64) >= 256 * 16 * 64 * 4
    assert C.workspace_bytes(C.OP_TPS_SOLVE, 8, 1, 1, 1) == 0
    with pytest.raises(C.UpsError, match="null pointer"):
        C.call("ups_part_softmax_fwd", None, None, None, None, 4, 16, None)
    with pytest.raises(C.UpsError, match="K=0"):
        C.call("ups_part_softmax_fwd", 16, 16, None, None, 4, 0, None)
    with pytest.raises(C.UpsError, match="fused path needs K"):
        C.call("ups_step_decode_fwd", 16, 16, 16, 16, 16, 1, 1024, 25, 64, None)
    with pytest.raises(C.UpsError, match="P % 32"):
        C.call("ups_step_encode_bwd", 16, None, 16, 16, None, 16, None, 1, 100, 16, None)
    assert C.launch_count() == 0


def test_row_length_arguments_are_validated_without_a_gpu(ups):
    """The `_rows` forms of the fused entry points (ragged [.,K'] rows read in place, K' <= K) reject a row length outside
    1..K, and a tensor of full rows (K' == K) that is not 16-byte aligned, before touching the device."""
    C = ups._cabi
    with pytest.raises(C.UpsError, match="rows of 33 floats for K=32"):
        C.call("ups_step_decode_fwd_rows", 16, 33, 16, 16, 16, 16, 1, 1024, 32, 64, None)
    with pytest.raises(C.UpsError, match="rows of 0 floats"):
        C.call("ups_step_encode_fwd_rows", 16, 0, 16, 16, 16, 16, 1, 1024, 32, 25, None, 0, None)
    with pytest.raises(C.UpsError, match="16-byte alignment"):         # full rows keep the vector-load contract
        C.call("ups_step_decode_fwd_rows", 20, 32, 16, 16, 16, 16, 1, 1024, 32, 64, None)
    with pytest.raises(C.UpsError, match="workspace"):                  # ragged rows may start anywhere: next check fires
        C.call("ups_step_encode_fwd_rows", 20, 25, 16, 16, 16, 16, 1, 1024, 32, 25, None, 0, None)
    assert C.launch_count() == 0


def test_n4_error_paths_without_a_gpu(ups):
    """The first-convolution entry points validate sizes before touching the device."""
    C = ups._cabi
    assert C.inject_conv_workspace_bytes(256, 128, 128, 16, 32) > 256 * 9 * 16 * 32 * 4
    assert C.parts_conv_bwd_workspace_bytes(256, 128, 128, 16, 32) > 256 * 16 * 128 * 128 * 4
    with pytest.raises(C.UpsError, match="null pointer"):
        C.call("ups_inject_conv_fwd", None, None, None, None, 1, 8, 8, 4, 8, None)
    with pytest.raises(C.UpsError, match="Co=6"):
        C.call("ups_inject_conv_fwd", 16, 16, 16, 16, 1, 8, 8, 4, 6, None)
    with pytest.raises(C.UpsError, match="K=33"):
        C.call("ups_inject_conv_bwd", 16, 16, 16, None, None, 16, 16, None, 1, 8, 8, 33, 8, None, 0, None)
    with pytest.raises(C.UpsError, match="workspace"):
        C.call("ups_inject_conv_bwd", 16, 16, 16, None, None, 16, 16, None, 1, 8, 8, 4, 8, None, 0, None)
    with pytest.raises(C.UpsError, match="does not fit shared memory"):
        C.call("ups_inject_conv_bwd", 16, 16, 16, None, None, 16, 16, None, 1, 8, 8, 32, 128, 16, 1 << 30, None)
    with pytest.raises(C.UpsError, match="3-channel"):
        C.call("ups_parts_conv_fwd", 16, 16, 16, 16, 16, 1, 8, 8, 4, 4, 8, None)
    with pytest.raises(C.UpsError, match="Co=12"):
        C.call("ups_parts_conv_bwd", 16, 16, 16, 16, None, None, 16, None, None, 1, 8, 8, 4, 3, 12, None, 0, None)
    assert C.launch_count() == 0
    with pytest.raises(AssertionError):                                 # feature / mask part counts differ
        ups.inject_conv2d(torch.zeros(2, 4, 6), torch.zeros(2, 8, 8, 5), torch.zeros(3, 3, 11, 8), torch.zeros(8))
    with pytest.raises(AssertionError):                                 # filter input channels != F + parts
        ups.inject_conv2d(torch.zeros(2, 5, 6), torch.zeros(2, 8, 8, 5), torch.zeros(3, 3, 10, 8), torch.zeros(8))
    with pytest.raises(AssertionError):                                 # image / mask spatial sizes differ (model.py:180)
        ups.parts_conv2d(torch.zeros(2, 8, 8, 3), torch.zeros(2, 8, 4, 5), torch.zeros(3, 3, 3, 8), torch.zeros(8))


def test_inject_conv_backward_plan_fits_the_sm(ups):
    """Every (K, Co) the entry point accepts either gets a tiling that fits one SM (227 KB dynamic shared memory per
    CTA; two CTAs per SM need <= 113 KB each) or is refused up front; the workspace covers the per-CTA partials."""
    C = ups._cabi
    for K in (1, 4, 8, 16, 25, 32):
        for Co in (4, 8, 12, 16, 32, 64, 128):
            for (B, H, W) in ((256, 128, 128), (2, 8, 8), (128, 256, 256)):
                p = C.inject_conv_bwd_plan(B, H, W, K, Co)
                if p["variant"] == 0:
                    assert 9 * K * Co * 4 * 2 > 100 * 1024, (K, Co, p)        # only big tables are refused
                    continue
                assert p["smem_bytes"] <= 227 * 1024, (K, Co, p)
                assert p["tile_rows"] in (4, 8, 16) and p["tiles"] == -(-W // 32) * -(-H // p["tile_rows"])
                assert p["ctas_per_sample"] * p["tiles_per_cta"] >= p["tiles"]
                if p["variant"] == 2:
                    assert Co in (8, 16, 32, 64, 128) and p["tile_rows"] in (4, 8)
                need = B * p["ctas_per_sample"] * (9 * K * Co + Co) * 4
                assert C.inject_conv_workspace_bytes(B, H, W, K, Co) >= need
    cub = C.inject_conv_bwd_plan(256, 128, 128, 16, 32)                       # the bench shape: tensor cores, 2 CTAs/SM
    assert cub["variant"] == 2 and cub["tile_rows"] == 8 and cub["smem_bytes"] <= 113 * 1024


REF_SIGNATURES = {
    # helper: parameter names (and defaults) in the reference, with file:line
    "softmax": "(x, spatial=False)",                                  # cub/code/nn.py:58
    "spatial_softmax": "(features)",                                  # cub/code/nn.py:65
    "apply_partwise": "(input_, func)",                               # cub/code/nn.py:81
    "hard_max_straight_through": "(y, axis)",                         # cub/code/nn.py:118
    "hard_max": "(y, axis)",                                          # cub/code/nn.py:134
    "straight_through_estimator": "(y_hard, y)",                      # cub/code/nn.py:154
    "mask2hotmask": "(mask, n_parts)",                                # cub/code/nn.py:2086
    "unpool_features_gathered": "(feature_vectors, mask)",            # cub/code/nn.py:2469
    "mask_parts": "(image, mask)",                                    # cub/code/SB_model48i/model.py:176
    "encode_parts": "(part_image, encoder)",                          # cub/code/SB_model48i/model.py:214
    "unpool_features": "(feature_vectors, mask, reshape=False)",      # model.py:225 ; deepfashion/code/foo.py:462
    "pool_features": "(feature_map, mask)",                           # deepfashion/code/foo.py:287
    "pool_unpool_block": "(feature_map, pool_mask, unpool_mask, reshape=False)",   # foo.py:574
    "get_features": "(features, part_map, slim)",                     # baselines/unsupervised-disentangling/ops.py:182
    "make_input_tps_param": "(tps_param, move_point=None, scal_point=None)",       # transformations.py:59
    "ThinPlateSpline": "(U, coord, vector, out_size, n_c, move=None, scal=None)",  # transformations.py:93
    "probs_to_mu_sigma": "(probs, scaling_factor)",                   # cub/code/nn.py:1541
    "categorical_kl": "(probs)",                                      # cub/code/SB_model48i/model.py:21
    "mumford_shah": "(x, alpha, lambda_)",                            # cub/code/nn.py:1381
    "edge_set": "(x, alpha, lambda_)",                                # cub/code/nn.py:1389
}


def test_reference_signatures(ups):
    for name, sig in REF_SIGNATURES.items():
        assert str(inspect.signature(getattr(ups, name))) == sig, name
    p = inspect.signature(ups.tps_parameters).parameters     # transformations.py:17 (+ config's augm_scal)
    assert list(p)[:7] == ["batch_size", "scal", "tps_scal", "rot_scal", "off_scal", "scal_var", "rescal"]
    assert p["rescal"].default == 1 and "augm_scal" in p
    assert list(inspect.signature(ups.make_tps).parameters)[:2] == ["views", "tps_parameters"]  # model.py:282
    p = inspect.signature(ups.mask2rgb).parameters           # cub/code/nn.py:2067 (+ the colour table as an input)
    assert list(p)[:2] == ["mask", "make_hot"] and p["make_hot"].default is True
    mf = ups.MeanFieldDistribution                            # cub/code/nn.py:1395-1457
    assert str(inspect.signature(mf.__init__)) == "(self, parameters, dim, stochastic=True)"
    assert list(inspect.signature(mf.sample).parameters)[:2] == ["self", "noise_level"]
    for m in ("kl", "kl_improper_gmrf", "kl_tv", "kl_mumford_sha", "n_parameters"):
        assert hasattr(mf, m), m


def test_tps_parameters_ranges_and_structure(ups):
    prm = ups.tps_parameters(64, scal=0.8, tps_scal=0.15, rot_scal=0.2, off_scal=0.2, scal_var=0.1, augm_scal=1.0,
                             generator=torch.Generator().manual_seed(0), device="cpu")
    assert set(prm) == {"coord", "vector", "offset", "offset_2", "t_scal", "rot_mat"}
    assert prm.coord.shape == (64, 8, 2) and prm.offset.shape == (64, 1, 2) and prm.t_scal.shape == (64, 2)
    assert prm.rot_mat.shape == (64, 2, 2)
    assert prm.vector.abs().max() <= 0.15 and prm.offset.abs().max() <= 0.2
    assert prm.t_scal.min() >= 0.8 * 0.9 - 1e-6 and prm.t_scal.max() <= 0.8 * 1.1 + 1e-6
    r = prm.rot_mat
    assert torch.allclose(r[:, 0, 0], r[:, 1, 1]) and torch.allclose(r[:, 0, 1], -r[:, 1, 0])
    assert torch.allclose(r[:, 0, 0] ** 2 + r[:, 1, 0] ** 2, torch.ones(64), atol=1e-6)
    # same draws as the oracle for the same seed (the parameters are INPUTS to both sides)
    from oracle import tps as OT
    po = OT.tps_parameters(64, 0.8, 0.15, 0.2, 0.2, 0.1, augm_scal=1.0, generator=torch.Generator().manual_seed(0))
    for k in prm:
        assert torch.equal(prm[k], po[k]), k


def test_shape_asserts_fire_before_any_kernel(ups):
    img, m = torch.zeros(2, 8, 8, 3), torch.zeros(2, 8, 4, 5)
    with pytest.raises(AssertionError):
        ups.mask_parts(img, m)                                     # model.py:180
    with pytest.raises(AssertionError):
        ups.unpool_features(torch.zeros(2, 4, 6), torch.zeros(2, 8, 8, 5))      # foo.py:466-467
    with pytest.raises(AssertionError):
        ups.pool_features(torch.zeros(2, 8, 8, 13), torch.zeros(2, 8, 8, 4))    # foo.py:290 (13 % 4 != 0)
    with pytest.raises(AssertionError):
        ups.unpool_features_gathered(torch.zeros(2, 4), torch.zeros(2, 8, 8))   # nn.py:2478


def test_no_cpu_fallback(ups):
    with pytest.raises(ups._cabi.UpsError, match="CUDA tensor"):
        ups.softmax(torch.randn(1, 2, 2, 4))
    from ups_b200.step import PartStep
    with pytest.raises(ups._cabi.UpsError, match="no CPU path"):
        PartStep(1, 32, 16, 64, device="cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unsupervised-part-segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_bench_reference_arm_prints_one_json_line():
    """bench.py's contract: stdout carries exactly one JSON line (the reference arm runs on the host cores, so this
    runs here); everything else, including whatever a library writes to file descriptor 1, goes to stderr."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
