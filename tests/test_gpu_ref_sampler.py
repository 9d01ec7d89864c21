"""GPU: our corr_sampler against the REFERENCE's own CUDA kernels (sampler/sampler_kernel.cu compiled for
sm_100a into oracle/_ref by oracle/build_ref_sampler.py).  Skipped when that binary was not built."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_sampler
    m = ref_sampler.load()
    if m is None:
        pytest.skip("oracle/_ref/corr_sampler_ref.so not built")
    return m


@pytest.mark.parametrize("shape", [(2, 5, 23, 23), (1, 8, 184, 184), (2, 6, 312, 156)])
def test_against_reference_cuda_sampler(ref, shape):
    import anystereo_b200 as A
    B, H, W1, W2 = shape
    rng = np.random.RandomState(W2)
    vol = torch.from_numpy(rng.standard_normal(shape)).float().cuda()
    # the reference kernel reads (and ignores) coords channel 1: give it two channels
    coords = torch.from_numpy(rng.uniform(-6, W2 + 6, size=(B, 2, H, W1))).float().cuda()
    g = torch.from_numpy(rng.standard_normal((B, 9, H, W1))).float().cuda()
    want, = ref.forward(vol, coords, 4)
    got, = A.corr_sampler.forward(vol, coords, 4)
    torch.cuda.synchronize()
    assert got.shape == want.shape
    # same index math (bit-exact taps); values differ only by fma contraction
    assert float((got - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max()))
    wantg, = ref.backward(vol, coords, g, 4)
    gotg, = A.corr_sampler.backward(vol, coords, g, 4)
    torch.cuda.synchronize()
    assert float((gotg - wantg).abs().max()) <= 2e-6 * max(1.0, float(wantg.abs().max()))
    # integer coordinates: both are exact gathers -> bit-identical
    ci = torch.floor(coords)
    a, = ref.forward(vol, ci, 4)
    b, = A.corr_sampler.forward(vol, ci, 4)
    assert torch.equal(a, b)
