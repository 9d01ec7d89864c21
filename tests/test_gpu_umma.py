"""GPU parity of the tcgen05 (tensor-core) kernels against the oracle and the exact-fp32 CUDA-core path.
Tolerances (BASELINE.json): fp32-parity mode "bf16x3" <= 1e-4 relative to max|ref|; the single-bf16 fast
mode is stated separately (<= 1e-2)."""
import numpy as np
import pytest
import torch

import cases
from oracle import hotpath_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def A():
    import anystereo_b200 as a
    yield a
    a.set_corr_mode("fp32")
    a.set_update_engine("fp32")


SHAPES = [
    # B, D, H, W, L
    (2, 16, 5, 23, 4),      # golden RAFT case: ragged everything, BN=32
    (2, 24, 4, 20, 2),      # golden IGEV case
    (1, 96, 3, 312, 2),     # IGEV config-2 row: 3 M tiles x 2 N tiles (BN=160), K = 64 + 32
    (1, 256, 2, 184, 4),    # RAFT config-1 row: BN=192, 4 K blocks, 4 pooled levels
    (1, 64, 2, 720, 4),     # config-3 width: 6 x 4 tiles
    (3, 40, 7, 50, 6),      # D not a multiple of 16, 6 levels (pool down to 1 column)
]


@pytest.mark.parametrize("mode,tol", [("bf16x3", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("shape", SHAPES)
def test_corr_umma_vs_oracle(A, mode, tol, shape):
    B, D, H, W, L = shape
    rng = np.random.RandomState(D * 7 + W)
    f1 = torch.from_numpy(rng.standard_normal((B, D, H, W))).float()
    f2 = torch.from_numpy(rng.standard_normal((B, D, H, W))).float()
    ref = O.corr_pyramid(O.all_pairs_corr(f1, f2), L)
    A.set_corr_mode(mode)
    blk = A.CorrBlock1D(f1.cuda(), f2.cuda(), num_levels=L, radius=4)
    torch.cuda.synchronize()
    prev = None
    for l, lvl in enumerate(blk.init_corr_pyramid):
        assert tuple(lvl.shape) == tuple(ref[l].shape)
        if lvl.numel():
            assert rel(lvl, ref[l]) < tol, (mode, l)
        if prev is not None and lvl.numel():
            # fused pooling is bit-exact against (a+b)*0.5 of OUR finer level
            assert torch.equal(lvl.cpu(), O.halve_last(prev.cpu()))
        prev = lvl
    # and against the exact-fp32 CUDA-core kernel
    A.set_corr_mode("fp32")
    blk32 = A.CorrBlock1D(f1.cuda(), f2.cuda(), num_levels=L, radius=4)
    assert rel(blk.init_corr_pyramid[0], blk32.init_corr_pyramid[0]) < tol


def test_corr_umma_golden(A, golden):
    g = golden("raft_corrblock")
    c = cases.raft_corr_case()
    A.set_corr_mode("bf16x3")
    corr = A.CorrBlock1D.corr(c["f1"].cuda(), c["f2"].cuda())
    assert rel(corr, g["corr"]) < 1e-4
    gi = golden("igev_geovolume")
    ci = cases.igev_geo_case()
    B, _, H, W = ci["f1"].shape
    blk = A.Combined_Geo_Encoding_Volume(ci["f1"].cuda(), ci["f2"].cuda(), ci["geo"].cuda(), num_levels=2, radius=4)
    for dist in ("uniform", "edge"):
        out = blk(ci["disps"][dist].cuda(), O.pixel_coords(B, H, W).cuda())
        assert rel(out, gi["lookup_" + dist]) < 1e-4
    A.set_corr_mode("fp32")


def test_corr_umma_fullsize_property(A):
    """config-2 size: split-bf16 tensor-core volume vs the fp32 CUDA-core volume, whole tensor."""
    torch.manual_seed(1)
    B, D, H, W = 2, 96, 96, 312
    f1 = torch.randn(B, D, H, W, device="cuda")
    f2 = torch.randn(B, D, H, W, device="cuda")
    A.set_corr_mode("fp32")
    ref = A.CorrBlock1D(f1, f2, num_levels=2, radius=4)
    A.set_corr_mode("bf16x3")
    got = A.CorrBlock1D(f1, f2, num_levels=2, radius=4)
    for l in range(2):
        assert rel(got.init_corr_pyramid[l], ref.init_corr_pyramid[l]) < 1e-4
    A.set_corr_mode("fp32")
