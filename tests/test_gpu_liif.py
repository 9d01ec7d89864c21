"""GPU parity of the LIIF arbitrary-scale upsampler against the committed reference outputs
(tests/golden/liif_upsample.npz) and the oracle.  Tolerance: 1e-4 relative to max|ref| in the fp32-parity mode
(split bf16 on the tensor cores), 2e-2 in the single-bf16 mode; nearest-pixel index math bit-exact."""
import numpy as np
import pytest
import torch

import cases
from oracle import liif_oracle as LO

pytestmark = pytest.mark.gpu

AFF = {"win_w": 3, "win_h": 3, "dilation": [1, 2, 4, 8]}


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def A():
    import anystereo_b200 as a
    yield a
    a.set_update_engine("fp32")


def make_module(A, c, seed):
    chanels = [f.shape[1] for f in c["feats"]]
    m = A.liif_out_multi_scale_Training(encoder_dim=sum(chanels), mlphidden_list=[128, 64, 64], pos_dim=0,
                                        unfold="with_v2ISU", affinity_settings=AFF, number_input=c["n_in"], chanels=chanels)
    m.load_state_dict(LO.make_liif_params(c["in_dim"], seed=seed), strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("shape", [(2, 176, 6, 10), (1, 32, 33, 47), (3, 8, 16, 16), (1, 5, 1, 1), (1, 40, 17, 2)])
def test_isu_affinity_vs_oracle(A, shape):
    rng = np.random.RandomState(shape[1])
    x = torch.from_numpy(rng.standard_normal(shape).astype("float32"))
    if shape[2] > 2:
        x[0, :, 1, 1] = 0.0                                   # zero vector: F.normalize eps path
    ref = LO.isu_affinity(x)
    got = A.liif.isu_affinity(x.cuda())
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    assert float((got.cpu() - ref).abs().max()) < 2e-6
    assert float(got.min()) >= 0.0


def test_isu_affinity_golden(A, golden):
    g = golden("liif_upsample")
    c = cases.liif_case(2)
    got = A.liif.isu_affinity(c["feats"][0].cuda())
    assert float((got.cpu() - torch.from_numpy(g["n2_affinity0"])).abs().max()) < 2e-6


@pytest.mark.parametrize("engine,tol", [("bf16x3", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("n_in", [2, 3])
def test_liif_logits_and_upsample_golden(A, golden, n_in, engine, tol):
    g = golden("liif_upsample")
    c = cases.liif_case(n_in)
    m = make_module(A, c, 40 + n_in)
    A.set_update_engine(engine)
    feats = [f.cuda() for f in c["feats"]]
    logits = m(feats, c["coords"].cuda(), c["scale"].cuda())
    up = m.upsample(feats, c["coords"].cuda(), c["disp"].cuda(), 4.0 * c["scale"].cuda())
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    t = "n%d_" % n_in
    assert tuple(logits.shape) == g[t + "logits"].shape
    assert rel(logits, g[t + "logits"]) < tol
    assert rel(up.unsqueeze(1), g[t + "up_disp"]) < tol
    # the two-step API (softmax outside, context_upsample_multiscale_train) agrees with the fused tail
    d = c["disp"].cuda() * 4.0 * c["scale"].cuda().view(-1, 1, 1, 1)
    up2 = A.context_upsample_multiscale_train(d, torch.softmax(logits, 1), c["coords"].cuda())
    assert rel(up2, up) < 1e-5


def test_upsample_disp_matches_oracle_ragged(A):
    """continuous_IGEVStereo.upsample_disp call shape: x = cat(stem_4x, hidden); Q not a multiple of the 64-query tile,
    per-sample scales, fp32-parity engine."""
    rng = np.random.RandomState(5)
    B, h, w, Q = 2, 9, 13, 1000
    stem4 = torch.from_numpy(rng.standard_normal((B, 48, h, w)).astype("float32"))
    hid = torch.from_numpy(np.tanh(rng.standard_normal((B, 128, h, w))).astype("float32"))
    stem2 = torch.from_numpy(rng.standard_normal((B, 32, 2 * h, 2 * w)).astype("float32"))
    coords = torch.from_numpy(rng.uniform(-1, 1, (B, Q, 2)).astype("float32"))
    disp = torch.from_numpy(rng.uniform(0, 30, (B, 1, h, w)).astype("float32"))
    scale = torch.tensor([2.5, 3.7])
    c = dict(feats=[torch.cat([stem4, hid], 1), stem2], n_in=2, in_dim=228)
    m = make_module(A, c, 3)
    A.set_update_engine("bf16x3")
    got = A.upsample_disp(m, disp.cuda(), hid.cuda(), stem4.cuda(), stem2.cuda(), None, hr_coord=coords.cuda(), scale=scale.cuda())
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    ref = LO.upsample_disp_multiscale(LO.make_liif_params(228, seed=3), disp, c["feats"], coords, scale)
    assert tuple(got.shape) == (B, 1, Q)
    assert rel(got, ref) < 1e-4


def test_liif_unsupported_variants_raise(A):
    with pytest.raises(NotImplementedError):
        A.liif_out_multi_scale_Training(encoder_dim=208, pos_dim=24, unfold="with_v2ISU", affinity_settings=AFF,
                                        number_input=2, chanels=[176, 32])
    with pytest.raises(NotImplementedError):
        A.liif_out_multi_scale_Training(encoder_dim=208, pos_dim=0, unfold="with_Dila_ISU", affinity_settings=AFF,
                                        number_input=2, chanels=[176, 32])


@pytest.mark.parametrize("engine,tol", [("bf16x3", 1e-4), ("bf16", 2e-2)])
def test_model_level_igev_upsample(A, golden, engine, tol):
    """Inputs and output captured inside the reference's continuous_IGEVStereo.forward (x2.5 query grid)."""
    g = golden("model_igev_upsample")
    params = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("liif.")}
    m = A.liif_out_multi_scale_Training(encoder_dim=208, mlphidden_list=[128, 64, 64], pos_dim=0, unfold="with_v2ISU",
                                        affinity_settings=AFF, number_input=2, chanels=[176, 32])
    m.load_state_dict(params, strict=True)
    m = m.cuda().eval()
    Ho, Wo = [int(v) for v in g["out_hw"]]
    coords = torch.stack(torch.meshgrid(LO.make_coord_axis(Ho), LO.make_coord_axis(Wo), indexing="ij"), -1).reshape(1, -1, 2)
    A.set_update_engine(engine)
    t = lambda k: torch.from_numpy(g[k]).cuda()   # noqa: E731
    up = A.upsample_disp(m, t("disp"), t("hidden"), t("stem4"), t("stem2"), None, hr_coord=coords.cuda(), scale=t("scale"))
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    assert rel(up, g["up_disp"]) < tol
    # mean absolute error in pixels of the full-resolution disparity
    assert float((up.cpu() - torch.from_numpy(g["up_disp"])).abs().mean()) < (0.01 if engine == "bf16x3" else 0.1)


def test_liif_fullsize_config4_vs_torch_same_gpu(A):
    """BASELINE config 4 size: one 384x1248 pair queried on the x2.5 grid (3.0 M queries) against the oracle's
    arithmetic run by torch on the same GPU (strict fp32); also the 64-bit indexing / ragged last tile at full size."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    dev = "cuda"
    h, w, scale = 96, 312, 2.5
    stem4 = torch.randn(1, 48, h, w, device=dev)
    hid = torch.tanh(torch.randn(1, 128, h, w, device=dev))
    stem2 = torch.randn(1, 32, 2 * h, 2 * w, device=dev)
    disp = torch.rand(1, 1, h, w, device=dev) * 48
    H, W = int(h * 4 * scale), int(w * 4 * scale)
    coords = torch.stack(torch.meshgrid(LO.make_coord_axis(H, dev), LO.make_coord_axis(W, dev), indexing="ij"), -1).reshape(1, -1, 2)
    coords = coords[:, :-37].contiguous()                   # Q not a multiple of the 128-query tile
    params = LO.make_liif_params(228, seed=9)
    c = dict(feats=[torch.cat([stem4, hid], 1), stem2], n_in=2, in_dim=228)
    m = make_module(A, c, 9)
    sc = torch.tensor([scale], device=dev)
    A.set_update_engine("bf16x3")
    got = A.upsample_disp(m, disp, hid, stem4, stem2, None, hr_coord=coords, scale=sc)
    A.set_update_engine("fp32")
    ref = LO.upsample_disp_multiscale({k: v.to(dev) for k, v in params.items()}, disp, c["feats"], coords, sc)
    torch.cuda.synchronize()
    assert rel(got, ref) < 1e-4
    assert float((got - ref).abs().mean()) < 1e-3          # pixels of full-resolution disparity


def test_liif_single_input_raft_style(A):
    """continuous_RaftStereo.upsample_disp with neither stem (prune_raft_stereo.py:221-222): one feature map."""
    rng = np.random.RandomState(8)
    B, h, w, Q = 2, 7, 11, 500
    hid = torch.from_numpy(np.tanh(rng.standard_normal((B, 128, h, w))).astype("float32"))
    coords = torch.from_numpy(rng.uniform(-1, 1, (B, Q, 2)).astype("float32"))
    disp = torch.from_numpy(rng.uniform(0, 30, (B, 1, h, w)).astype("float32"))
    scale = torch.tensor([1.5, 2.0])
    c = dict(feats=[hid], n_in=1, in_dim=128 + 8 + 2)
    m = make_module(A, c, 5)
    A.set_update_engine("bf16x3")
    got = A.upsample_disp(m, disp.cuda(), hid.cuda(), None, None, None, hr_coord=coords.cuda(), scale=scale.cuda())
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    ref = LO.upsample_disp_multiscale(LO.make_liif_params(138, seed=5), disp, [hid], coords, scale)
    assert rel(got, ref) < 1e-4


@pytest.mark.parametrize("mode", ["disparity_norm", "disparity_norm2"])
def test_upsample_disp_normalised_variants(A, mode):
    """args.disparity_norm / disparity_norm2 branches of upsample_disp (continuous_IGEVstereo.py:198-201, :226-235)."""
    rng = np.random.RandomState(12)
    B, h, w, Q = 2, 6, 9, 300
    stem4 = torch.from_numpy(rng.standard_normal((B, 48, h, w)).astype("float32"))
    hid = torch.from_numpy(np.tanh(rng.standard_normal((B, 128, h, w))).astype("float32"))
    stem2 = torch.from_numpy(rng.standard_normal((B, 32, 2 * h, 2 * w)).astype("float32"))
    coords = torch.from_numpy(rng.uniform(-1, 1, (B, Q, 2)).astype("float32"))
    disp = torch.from_numpy(rng.uniform(0, 30, (B, 1, h, w)).astype("float32"))
    scale = torch.tensor([2.5, 3.7])
    feats = [torch.cat([stem4, hid], 1), stem2]
    m = make_module(A, dict(feats=feats, n_in=2, in_dim=228), 6)
    A.set_update_engine("bf16x3")
    got = A.upsample_disp(m, disp.cuda(), hid.cuda(), stem4.cuda(), stem2.cuda(), None, hr_coord=coords.cuda(),
                          scale=scale.cuda(), **{mode: True})
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    params = LO.make_liif_params(228, seed=6)
    mask = torch.softmax(LO.liif_logits(params, feats, coords), dim=1)
    k = 1.0 if mode == "disparity_norm" else 1024.0
    up = LO.context_upsample_multiscale(disp / w * k, mask, coords).unsqueeze(1)
    ref = up / k * torch.round(w * 4.0 * scale.view(-1, 1, 1))
    assert rel(got, ref) < 1e-4


@pytest.mark.parametrize("scale", [1.0, 1.7, 3.3])
def test_upsampler_training_kernels_match_oracle_autograd(A, scale):
    """The upsampler in TRAINING on the GPU (first Linear at source resolution, fused gather + relative-coordinate terms +
    bias + ReLU kernel, context-upsample kernel, and their hand-written adjoints, csrc/liif_train.cu) against autograd over
    the oracle restatement of the reference on the CPU: value, d disp, d features, d parameters."""
    torch.manual_seed(0)
    lc = cases.liif_case(2, h=9, w=13, scale=scale, extra_q=7)
    chanels = [f.shape[1] for f in lc["feats"]]
    m = A.liif_out_multi_scale_Training(encoder_dim=sum(chanels), mlphidden_list=[128, 64, 64], pos_dim=0,
                                        unfold="with_v2ISU", affinity_settings=AFF, number_input=2, chanels=chanels)
    lp = LO.make_liif_params(lc["in_dim"], seed=2)
    m.load_state_dict(lp, strict=True)
    m = m.cuda().train()
    feats = [f.clone().cuda().requires_grad_(True) for f in lc["feats"]]
    disp = lc["disp"].clone().cuda().requires_grad_(True)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = A.upsample_disp(m, disp, feats[0][:, 48:], feats[0][:, :48], feats[1], None, hr_coord=lc["coords"].cuda(),
                              scale=lc["scale"].cuda())
        assert out.grad_fn is not None
        wgt = torch.linspace(0.5, 1.5, out.numel()).view_as(out)
        (out * wgt.cuda()).sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    p2 = {k: v.clone().requires_grad_(True) for k, v in lp.items()}
    f2 = [f.clone().requires_grad_(True) for f in lc["feats"]]
    d2 = lc["disp"].clone().requires_grad_(True)
    ref = LO.upsample_disp_multiscale(p2, d2, f2, lc["coords"], lc["scale"])
    (ref * wgt).sum().backward()
    assert rel(out, ref) < 1e-5
    assert rel(disp.grad, d2.grad) < 1e-5
    for a, b in zip(feats, f2):
        assert rel(a.grad, b.grad) < 2e-5
    for k, v in m.state_dict(keep_vars=True).items():
        assert rel(v.grad, p2[k].grad) < 2e-5, k
