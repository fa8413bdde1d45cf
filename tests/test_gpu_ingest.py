"""-m gpu: device-side image ingest (uint8 -> [-1, 1] fp32, cub/code/data/data.py:134) against the
oracle and the reference-generated fixture, bit for bit; and the fused step fed with uint8 views."""
import os

import numpy as np
import pytest
import torch

from oracle import ingest as OI
from util import assert_bitexact, make_inputs, cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ups():
    import ups_b200
    return ups_b200


def test_all_bytes_and_fixture_bitexact(ups):
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "ingest.npz"))
    for src, want in ((d["all_bytes"], d["out_bytes"]), (d["img"], d["out_img"])):
        got = ups.images_from_uint8(torch.from_numpy(src).cuda())
        assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
        assert_bitexact(got, torch.from_numpy(want), "images_from_uint8 vs reference expression")


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 4096 + 5, 3 * 8 * 128 * 128 * 3])
def test_ragged_sizes_bitexact(ups, n):
    g = torch.Generator().manual_seed(n)
    src = torch.randint(0, 256, (n,), dtype=torch.uint8, generator=g)
    got = ups.images_from_uint8(src.cuda())
    assert_bitexact(got, OI.images_from_uint8(src), f"n={n}")


def test_rejects_wrong_inputs(ups):
    with pytest.raises(ups._cabi.UpsError, match="uint8"):
        ups.images_from_uint8(torch.zeros(4, device="cuda"))
    with pytest.raises(ups._cabi.UpsError, match="CUDA"):
        ups.images_from_uint8(torch.zeros(4, dtype=torch.uint8))


def test_step_accepts_uint8_views(ups):
    """PartStep.forward(uint8 views) == PartStep.forward(oracle-normalised fp32 views), bit for bit."""
    from ups_b200.step import PartStep
    B, S, K, F, V = 2, 64, 16, 64, 3
    inp = make_inputs(B, S, K, F, V, seed=3)
    g = torch.Generator().manual_seed(7)
    v8 = torch.randint(0, 256, (V, B, S, S, 3), dtype=torch.uint8, generator=g)
    d = cuda(inp)
    step = PartStep(B, S, K, F, n_views=V)
    out = step.forward(OI.images_from_uint8(v8).cuda(), d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"])
    want = {k: [t.clone() for t in out[k]] if k == "warped" else out[k].clone() for k in ("warped", "parts", "pooled")}
    out = step.forward(v8.cuda(), d["coord"], d["t_vector"], d["l0"], d["l1"], d["feat"])
    for i in range(V):
        assert_bitexact(out["warped"][i], want["warped"][i].cpu(), f"warped[{i}]")
    assert_bitexact(out["parts"], want["parts"].cpu(), "parts")
    assert_bitexact(out["pooled"], want["pooled"].cpu(), "pooled")


def test_labels_to_u8():
    from ups_b200 import _cabi as C
    for n in (1, 3, 4, 1000, 128 * 128 * 8 + 3):
        lab = torch.randint(0, 25, (n,), dtype=torch.int64, device="cuda")
        out = torch.full((n + 4,), 77, dtype=torch.uint8, device="cuda")
        C.call("ups_labels_i64_to_u8", lab.data_ptr(), out.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        assert torch.equal(out[:n].cpu(), lab.cpu().to(torch.uint8)) and bool((out[n:] == 77).all())
