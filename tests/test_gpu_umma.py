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


# ---------------------------------------------------------------------------------------------------
# update block on tensor cores
# ---------------------------------------------------------------------------------------------------
import types  # noqa: E402


def make_block(A, family, seed):
    cls = A.BasicMultiUpdateBlock if family == "igev" else A.BasicMultiUpdateBlockRAFT
    args = types.SimpleNamespace(corr_levels=2 if family == "igev" else 4, corr_radius=4, n_gru_layers=3)
    m = cls(args, hidden_dims=[128, 128, 128])
    m.load_state_dict(O.make_update_block_params(162 if family == "igev" else 36, seed=seed), strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("engine,tol", [("bf16x3", 2e-4), ("f16f8", 2e-4), ("bf16", 5e-2), ("fp16", 6e-3)])
@pytest.mark.parametrize("family", ["igev", "raft"])
def test_update_block_umma_golden(A, golden, family, engine, tol):
    g = golden("update_block_" + family)
    c = cases.update_block_case(family)
    m = make_block(A, family, 11)
    A.set_update_engine(engine)
    inp = [[t.cuda() for t in lst] for lst in c["inp"]]
    with torch.no_grad():
        net, delta = m([t.cuda() for t in c["net"]], inp, c["corr"].cuda(), c["disp"].cuda())
        torch.cuda.synchronize()
        for i in range(3):
            assert rel(net[i], g["full_net%d" % i]) < tol, i
        assert rel(delta, g["full_delta"]) < tol
        net = m([t.cuda() for t in c["net"]], inp, iter16=True, iter08=True, iter04=False, update=False)
        for i in range(3):
            assert rel(net[i], g["lowres_net%d" % i]) < tol
    A.set_update_engine("fp32")


@pytest.mark.parametrize("engine", ["bf16x3", "f16f8", "fp16"])
@pytest.mark.parametrize("family", ["igev", "raft"])
def test_iteration_loop_epe_umma(A, golden, family, engine):
    """fp32-parity tensor-core mode (and the IEEE-half fast mode): final disparity within 0.01 px of the reference
    after equal iterations."""
    g = golden("loop_" + family)
    c = cases.loop_case(family)
    iters = int(g["iters"])
    m = make_block(A, family, 12 if family == "igev" else 13)
    A.set_update_engine(engine)
    A.set_corr_mode("bf16x3")
    net = [t.cuda() for t in c["net"]]
    inp = [[t.cuda() for t in lst] for lst in c["inp"]]
    if family == "igev":
        disp, net, hist = A.igev_iterations(m, c["f1"].cuda(), c["f2"].cuda(), c["geo"].cuda(), net, inp,
                                            c["init_disp"].cuda(), iters, keep_all=True)
    else:
        disp, net, hist = A.raft_iterations(m, c["f1"].cuda(), c["f2"].cuda(), net, inp, iters, keep_all=True)
    ref = torch.from_numpy(g["disps"])
    err = (torch.stack([h.cpu() for h in hist]) - ref).abs()
    epe = float(err[-1].mean()) * 4
    A.set_update_engine("fp32")
    A.set_corr_mode("fp32")
    assert epe < 0.01, epe


def test_update_block_umma_vs_fp32_medium(A):
    """1/4-KITTI-sized maps with ragged tiles (W not a multiple of 16, H not of 8): tensor-core engine vs the
    exact-fp32 CUDA-core engine on the GPU."""
    torch.manual_seed(3)
    B, H, W = 2, 46, 75
    m = make_block(A, "igev", 5)
    sizes = [(H, W), ((H + 1) // 2, (W + 1) // 2), (((H + 1) // 2 + 1) // 2, ((W + 1) // 2 + 1) // 2)]
    net = [torch.tanh(torch.randn(B, 128, h, w, device="cuda")) for h, w in sizes]
    inp = [[0.5 * torch.randn(B, 128, h, w, device="cuda") for _ in range(3)] for h, w in sizes]
    corr = torch.randn(B, 162, H, W, device="cuda")
    disp = torch.rand(B, 1, H, W, device="cuda") * 20
    with torch.no_grad():
        A.set_update_engine("fp32")
        n32, d32 = m([t.clone() for t in net], inp, corr, disp)
        A.set_update_engine("bf16x3")
        n3, d3 = m([t.clone() for t in net], inp, corr, disp)
        # second call on its own outputs exercises the cached hi/lo planes
        n3b, d3b = m(list(n3), inp, corr, disp)
        A.set_update_engine("fp32")
        n32b, d32b = m(list(n32), inp, corr, disp)
    for i in range(3):
        assert rel(n3[i], n32[i]) < 2e-4
        assert rel(n3b[i], n32b[i]) < 4e-4
    assert rel(d3, d32) < 2e-4
    assert rel(d3b, d32b) < 4e-4


def test_hot_loop_cuda_graph(A):
    """The whole IGEV hot-path step (volume build + iterations) captured in a CUDA graph replays to the same
    result as eager execution, and follows new inputs loaded into its static buffers."""
    c = cases.loop_case("igev", seed=55, B=1, H=16, W=24)
    m = make_block(A, "igev", 7)
    A.set_update_engine("bf16x3")
    A.set_corr_mode("bf16x3")
    dev = "cuda"
    args = [c["f1"].to(dev), c["f2"].to(dev), c["geo"].to(dev), [t.to(dev) for t in c["net"]],
            [[t.to(dev) for t in l] for l in c["inp"]], c["init_disp"].to(dev)]
    eager, _ = A.igev_iterations(m, *args, 4)
    g = A.HotLoopGraph(m, args[0], args[1], args[2], args[3], args[4], args[5], iters=4)
    d1, _ = g.replay()
    torch.cuda.synchronize()
    assert torch.equal(d1, eager)
    # new inputs through the static buffers
    c2 = cases.loop_case("igev", seed=56, B=1, H=16, W=24)
    args2 = [c2["f1"].to(dev), c2["f2"].to(dev), c2["geo"].to(dev), [t.to(dev) for t in c2["net"]],
             [[t.to(dev) for t in l] for l in c2["inp"]], c2["init_disp"].to(dev)]
    eager2, _ = A.igev_iterations(m, *args2, 4)
    g.load(*args2)
    d2, _ = g.replay()
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    A.set_corr_mode("fp32")
    assert torch.equal(d2, eager2)


def test_raft_fullsize_config1(A):
    """BASELINE config 1 shapes (320x736 -> 80x184, D=256, L=4): tensor-core pyramid + fused lookup against the
    exact-fp32 CUDA-core path, and integer-disparity lookups are exact gathers."""
    torch.manual_seed(2)
    B, D, H, W = 1, 256, 80, 184
    f1 = torch.randn(B, D, H, W, device="cuda") / 4
    f2 = torch.randn(B, D, H, W, device="cuda") / 4
    A.set_corr_mode("fp32")
    ref = A.CorrBlock1D(f1, f2, num_levels=4, radius=4)
    A.set_corr_mode("bf16x3")
    got = A.CorrBlock1D(f1, f2, num_levels=4, radius=4)
    A.set_corr_mode("fp32")
    for l in range(4):
        assert rel(got.init_corr_pyramid[l], ref.init_corr_pyramid[l]) < 1e-4
    coords = O.pixel_coords(B, H, W).cuda()
    disp = torch.randint(0, 60, (B, 1, H, W), device="cuda").float()
    out = ref(disp, coords)
    N = B * H * W
    for l in range(4):
        Wl = W >> l
        lvl = ref.init_corr_pyramid[l].reshape(N, Wl)
        x = (coords.reshape(N, 1) - disp.reshape(N, 1)) / 2 ** l
        fl = torch.floor(x)
        fr = (x - fl)
        idx = fl.long() + torch.arange(-4, 6, device="cuda").view(1, 10)
        ok = (idx >= 0) & (idx < Wl)
        g = torch.gather(lvl, 1, idx.clamp(0, Wl - 1)) * ok
        want = g[:, :9] * (1 - fr) + g[:, 1:] * fr
        have = out[:, 9 * l:9 * l + 9].permute(0, 2, 3, 1).reshape(N, 9)
        assert float((have - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max()))


def test_config3_middlebury_full_res_smoke(A):
    """BASELINE config 3: 1984x2880 -> 496x720, RAFT family, 4-level pyramid (1.9 GB).  Two iterations of the hot
    path on the tensor-core engines; the fp32 CUDA-core engine must agree on the first iteration's update."""
    torch.manual_seed(4)
    B, D, H, W = 1, 256, 496, 720
    dev = "cuda"
    f1 = torch.randn(B, D, H, W, device=dev) / 8
    f2 = torch.randn(B, D, H, W, device=dev) / 8
    m = make_block(A, "raft", 21)
    sizes = [(H, W), (H // 2, W // 2), (H // 4, W // 4)]
    net = [torch.tanh(torch.randn(B, 128, h, w, device=dev)) for h, w in sizes]
    inp = [[0.5 * torch.randn(B, 128, h, w, device=dev) for _ in range(3)] for h, w in sizes]
    A.set_update_engine("bf16x3")
    A.set_corr_mode("bf16x3")
    disp, net2, hist = A.raft_iterations(m, f1, f2, [t.clone() for t in net], inp, 2, keep_all=True)
    A.set_update_engine("fp32")
    disp32, _, hist32 = A.raft_iterations(m, f1, f2, [t.clone() for t in net], inp, 1, keep_all=True)
    A.set_corr_mode("fp32")
    torch.cuda.synchronize()
    assert tuple(disp.shape) == (B, 1, H, W) and bool(torch.isfinite(disp).all())
    assert float((hist[0] - hist32[0]).abs().max()) < 2e-4 * max(1.0, float(hist32[0].abs().max()))


@pytest.mark.parametrize("engine", ["fp32", "bf16x3", "f16f8"])
def test_model_level_raft_epe(A, golden, engine):
    """Drop-in at the model level: the tensors the real reference RAFT model feeds into its hot path (captured on
    CPU, tests/golden/make_model_golden.py) replayed through our operators for the same 32 iterations give the
    reference model's low-resolution disparity to within 0.01 px (x4 = full-resolution pixels)."""
    g = golden("model_raft_boundary")
    net = [torch.from_numpy(g["net%d" % i]).cuda() for i in range(3)]
    inp = [[torch.from_numpy(g["inp%d_%d" % (i, j)]).cuda() for j in range(3)] for i in range(3)]
    args = types.SimpleNamespace(corr_levels=4, corr_radius=4, n_gru_layers=3)
    m = A.BasicMultiUpdateBlockRAFT(args, hidden_dims=[128, 128, 128])
    m.load_state_dict(O.make_update_block_params(36, seed=77), strict=True)
    m = m.cuda().eval()
    A.set_update_engine(engine)
    A.set_corr_mode("bf16x3" if engine == "f16f8" else engine)
    disp, _ = A.raft_iterations(m, torch.from_numpy(g["f1"]).cuda(), torch.from_numpy(g["f2"]).cuda(), net, inp,
                                int(g["iters"]))
    A.set_update_engine("fp32")
    A.set_corr_mode("fp32")
    epe = float((disp.cpu() - torch.from_numpy(g["disp_lowres"])).abs().mean()) * 4
    assert epe < 0.01, epe


@pytest.mark.parametrize("engine", ["fp32", "bf16x3", "f16f8"])
def test_model_level_igev_epe(A, golden, engine):
    """IGEV family at the model level: build_gwc_volume on the real model's matching features, then the combined
    volume + 32 iterations from the real model's boundary tensors -> within 0.01 px of the reference model."""
    g = golden("model_igev_boundary")
    net = [torch.from_numpy(g["net%d" % i]).cuda() for i in range(3)]
    inp = [[torch.from_numpy(g["inp%d_%d" % (i, j)]).cuda() for j in range(3)] for i in range(3)]
    f1, f2 = torch.from_numpy(g["f1"]).cuda(), torch.from_numpy(g["f2"]).cuda()
    gwc = A.build_gwc_volume(f1, f2, 48, 8)
    assert rel(gwc, g["gwc"]) < 1e-5
    args = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=3)
    m = A.BasicMultiUpdateBlock(args, hidden_dims=[128, 128, 128])
    m.load_state_dict(O.make_update_block_params(162, seed=78), strict=True)
    m = m.cuda().eval()
    A.set_update_engine(engine)
    A.set_corr_mode("bf16x3" if engine == "f16f8" else engine)
    disp, _ = A.igev_iterations(m, f1, f2, torch.from_numpy(g["geo"]).cuda(), net, inp,
                                torch.from_numpy(g["init_disp"]).cuda(), int(g["iters"]))
    A.set_update_engine("fp32")
    A.set_corr_mode("fp32")
    epe = float((disp.cpu() - torch.from_numpy(g["disp_lowres"])).abs().mean()) * 4
    assert epe < 0.01, epe


@pytest.mark.parametrize("engine,tol", [("fp32", 1e-4), ("bf16x3", 3e-4), ("f16f8", 3e-4)])
@pytest.mark.parametrize("n_layers", [1, 2])
def test_update_block_fewer_gru_layers(A, engine, tol, n_layers):
    """n_gru_layers = 1 / 2 variants of BasicMultiUpdateBlock (update.py:111-112,121-130) and half-precision inputs."""
    import torch.nn as nn
    torch.manual_seed(n_layers)
    B, H, W = 1, 12, 20
    args = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=n_layers)
    m = A.BasicMultiUpdateBlock(args, hidden_dims=[128, 128, 128]).cuda().eval()
    p = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sizes = [(H, W), (H // 2, W // 2), (H // 4, W // 4)][:n_layers]
    net = [torch.tanh(torch.randn(B, 128, h, w)) for h, w in sizes]
    inp = [[0.5 * torch.randn(B, 128, h, w) for _ in range(3)] for h, w in sizes]
    corr = torch.randn(B, 162, H, W)
    disp = torch.rand(B, 1, H, W) * 10
    ref_net, ref_delta = O.update_block(p, net, inp, corr, disp, iter16=n_layers == 3, iter08=n_layers >= 2,
                                        n_gru_layers=n_layers)
    A.set_update_engine(engine)
    with torch.no_grad():
        got_net, got_delta = m([t.cuda() for t in net], [[t.cuda() for t in l] for l in inp], corr.cuda(), disp.cuda(),
                               iter16=n_layers == 3, iter08=n_layers >= 2)
        for i in range(n_layers):
            assert rel(got_net[i], ref_net[i]) < tol
        assert rel(got_delta, ref_delta) < tol
        if engine == "fp32":   # half inputs are accepted and widened
            h_net, h_delta = m([t.cuda().half() for t in net], [[t.cuda() for t in l] for l in inp], corr.cuda().half(),
                               disp.cuda(), iter16=n_layers == 3, iter08=n_layers >= 2)
            assert rel(h_delta, ref_delta) < 2e-2
    A.set_update_engine("fp32")


@pytest.mark.parametrize("engine,tol", [("bf16x3", 1e-4), ("bf16", 2e-2), ("fp16", 3e-3)])
@pytest.mark.parametrize("shape", [(2, 5, 23, 16, 2), (1, 16, 24, 48, 2), (1, 7, 150, 24, 1), (3, 9, 40, 20, 2)])
@pytest.mark.parametrize("tap", [False, True])
def test_fused_lookup_convc1_vs_oracle(A, engine, tol, shape, tap):
    """lookup fused with BasicMotionEncoder.convc1 + ReLU on the tensor cores against
    relu(conv1x1(oracle lookup)); ragged 128-pixel tiles, out-of-range disparities, 1 and 2 levels; both kernel
    generations (tap: as_geo_lookup_convc1_tap, tap-major K order, 2 levels only)."""
    B, H, W, Dg, Lv = shape
    if tap and Lv != 2:
        pytest.skip("the tap-major kernel is built for 2 levels")
    c = cases.igev_geo_case(seed=31 + H, B=B, D=16, H=H, W=W, Dg=Dg)
    rng = np.random.RandomState(W)
    disp = torch.from_numpy(rng.uniform(-6, Dg + 6, (B, 1, H, W)).astype("float32"))
    disp[0, 0, 0, :3] = torch.tensor([0.0, Dg - 1.0, 3.0])          # integer positions: frac == 0
    coords = torch.arange(W).float().reshape(1, 1, W, 1).repeat(B, H, 1, 1)
    w = torch.from_numpy(rng.standard_normal((64, Lv * 81, 1, 1)).astype("float32")) * 0.2
    b = torch.from_numpy(rng.standard_normal(64).astype("float32")) * 0.1
    feat = O.geo_lookup(O.geo_pyramid(c["geo"], Lv), O.corr_pyramid(O.all_pairs_corr(c["f1"], c["f2"]), Lv), disp, coords,
                        4, exact=True)
    ref = torch.relu(torch.nn.functional.conv2d(feat.double(), w.double(), b.double())).float()
    A.set_corr_mode("fp32")
    A.set_update_engine(engine)                       # selects the 16-bit operand format (bf16 / IEEE half)
    vol = A.Combined_Geo_Encoding_Volume(c["f1"].cuda(), c["f2"].cuda(), c["geo"].cuda(), num_levels=Lv, radius=4)
    split = engine == "bf16x3"
    w_hi, w_lo = A.geometry.DeferredGeoLookup.pack_convc1_weight(w.cuda(), split, tap, b.cuda() if tap else None)
    out_hi = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    out_lo = torch.full_like(out_hi, float("nan")) if split else None
    d = vol.deferred(disp.cuda(), coords.cuda())
    assert d.fusable and d.shape == (B, Lv * 81, H, W)
    assert d.tap_major == (Lv == 2)
    d.convc1_planes(w_hi, w_lo, b.cuda(), out_hi, out_lo, tap)
    torch.cuda.synchronize()
    widen = (lambda t: t.view(torch.float16).float()) if engine == "fp16" else (lambda t: t.float())
    got = widen(out_hi) + (widen(out_lo) if split else 0)
    # coords=None means the default pixel grid
    out2 = torch.empty_like(out_hi)
    vol.deferred(disp.cuda(), None).convc1_planes(w_hi, w_lo, b.cuda(), out2, torch.empty_like(out_hi) if split else None, tap)
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    assert rel(got.permute(0, 3, 1, 2), ref) < tol
    assert torch.equal(out2.view(torch.int16), out_hi.view(torch.int16))


@pytest.mark.parametrize("engine", ["bf16x3", "f16f8", "bf16"])
def test_fused_lookup_matches_unfused_loop(A, engine):
    """igev_iterations with the lookup fused into the update block == the same loop with the 162-channel tensor
    materialised (identical arithmetic: fp32 interpolation, same bf16 split, same K order on the MMA)."""
    c = cases.loop_case("igev", seed=77, B=2, H=19, W=45)
    m = make_block(A, "igev", 9)
    A.set_update_engine(engine)
    A.set_corr_mode("bf16x3")
    args = [c["f1"].cuda(), c["f2"].cuda(), c["geo"].cuda(), [t.cuda() for t in c["net"]],
            [[t.cuda() for t in l] for l in c["inp"]], c["init_disp"].cuda()]
    prev = A.set_lookup_fusion(True)
    for on in (False, True):                        # both paths pack their weights / fill the context caches
        A.set_lookup_fusion(on)
        A.igev_iterations(m, *args, 1)
    n0 = A._lib.launch_count
    d_f, net_f = A.igev_iterations(m, *args, 6)
    n1 = A._lib.launch_count
    A.set_lookup_fusion(False)
    d_u, net_u = A.igev_iterations(m, *args, 6)
    n2 = A._lib.launch_count
    A.set_lookup_fusion(prev)
    A.set_update_engine("fp32")
    A.set_corr_mode("fp32")
    assert (n2 - n1) - (n1 - n0) >= 6, (n0, n1, n2)         # at least one launch fewer per iteration
    tol = {"bf16x3": 1e-5, "f16f8": 1e-4}.get(engine, 2e-3)       # K order differs: fp32 accumulation order, amplified by 6 GRU steps
    assert rel(d_f, d_u) < tol
    for a, b in zip(net_f, net_u):
        assert rel(a, b) < tol


def test_fused_lookup_falls_back(A):
    """Deferred lookups are materialised by the exact-fp32 engine and by unsupported shapes (G != 8)."""
    c = cases.loop_case("igev", seed=78, B=1, H=8, W=20)
    m = make_block(A, "igev", 9)
    args = [c["f1"].cuda(), c["f2"].cuda(), c["geo"].cuda()]
    vol = A.Combined_Geo_Encoding_Volume(*args, num_levels=2, radius=4)
    disp = c["init_disp"].cuda()
    net = [t.cuda() for t in c["net"]]
    inp = [[t.cuda() for t in l] for l in c["inp"]]
    with torch.no_grad():
        A.set_update_engine("fp32")
        n_a, d_a = m(list(net), inp, vol.deferred(disp, None), disp)
        n_b, d_b = m(list(net), inp, vol(disp, None), disp)
    assert torch.equal(d_a, d_b)
    for fn in (A._lib.lib().as_geo_lookup_convc1, A._lib.lib().as_geo_lookup_convc1_tap):
        rc = fn(None, 4, 8, None, None, None, 2, None, None, None, None, None, 3, None, None, 1, 1, 1, 4, None)
        assert rc == -1      # AS_ERR_BAD_ARG


@pytest.mark.parametrize("tap", [False, True])
def test_fused_lookup_fullsize_config2(A, tap):
    """Config-2 size (8 x 96x312, Dg=48): the fused lookup+convc1 kernels against the unfused kernels
    (lookup -> bf16 split -> 1x1 tcgen05 conv), same arithmetic up to the K order of the fp32 accumulation."""
    torch.manual_seed(2)
    dev = "cuda"
    B, D, H, W, Dg = 8, 96, 96, 312, 48
    f1 = torch.randn(B, D, H, W, device=dev) * 0.2
    f2 = torch.randn(B, D, H, W, device=dev) * 0.2
    geo = torch.randn(B, 8, Dg, H, W, device=dev)
    A.set_corr_mode("bf16x3")
    A.set_update_engine("bf16x3")
    vol = A.Combined_Geo_Encoding_Volume(f1, f2, geo, num_levels=2, radius=4)
    disp = (torch.rand(B, 1, H, W, device=dev) * (Dg + 8) - 4).contiguous()
    m = make_block(A, "igev", 4)
    from anystereo_b200 import update_umma as U
    wf = U._fused_c1_weights(m, True, tap_major=tap)
    out_hi = torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16)
    out_lo = torch.empty_like(out_hi)
    vol.deferred(disp, None).convc1_planes(wf["hi"], wf["lo"], wf["bias"], out_hi, out_lo, tap)
    feat = vol(disp, None)
    ref = torch.relu(torch.nn.functional.conv2d(feat.double(), m.encoder.convc1.weight.double(), m.encoder.convc1.bias.double()))
    torch.cuda.synchronize()
    A.set_corr_mode("fp32")
    A.set_update_engine("fp32")
    got = (out_hi.float() + out_lo.float()).permute(0, 3, 1, 2)
    assert rel(got, ref.float()) < 1e-4


def test_fp16_operand_format_saturates(A):
    """IEEE-half operand planes clamp to +-65504 instead of overflowing to inf (the fp32 accumulators never see inf)."""
    L = A._lib
    x = torch.tensor([1.0e6, -3.0e5, 65504.0, 0.333251953125, -1.5, 0.0, 7.0e-8, 1.0], device="cuda")
    hi = torch.empty(8, device="cuda", dtype=torch.bfloat16)
    A.set_update_engine("fp16")
    assert L.lib().as_get_operand_format() == 1
    L.call("as_split_f32", x.data_ptr(), hi.data_ptr(), None, 8, L.stream_ptr())
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    assert L.lib().as_get_operand_format() == 0
    got = hi.view(torch.float16).float().cpu()
    ref = x.cpu().clamp(-65504.0, 65504.0).half().float()
    assert torch.equal(got, ref)
    assert L.lib().as_set_operand_format(7) == -1


@pytest.mark.parametrize("engine,tol", [("bf16x3", 1e-4), ("fp16", 3e-3)])
@pytest.mark.parametrize("shape", [(2, 5, 23, 4), (1, 9, 150, 4), (1, 16, 24, 2), (3, 7, 40, 4)])
def test_fused_raft_lookup_convc1_vs_oracle(A, engine, tol, shape):
    """RAFT twin of the fused kernel: CorrBlock1D lookup (L = 4 or 2 levels) + convc1 + ReLU against
    relu(conv1x1(oracle lookup)); ragged tiles, out-of-range disparities, coarse levels narrower than the window."""
    B, H, W, Lv = shape
    c = cases.raft_corr_case(seed=41 + H, B=B, D=16, H=H, W=W, L=Lv)
    rng = np.random.RandomState(W + Lv)
    disp = torch.from_numpy(rng.uniform(-8, W + 8, (B, 1, H, W)).astype("float32"))
    disp[0, 0, 0, :3] = torch.tensor([0.0, 1.0, 3.0])
    coords = torch.arange(W).float().reshape(1, 1, W, 1).repeat(B, H, 1, 1)
    w = torch.from_numpy(rng.standard_normal((64, Lv * 9, 1, 1)).astype("float32")) * 0.3
    b = torch.from_numpy(rng.standard_normal(64).astype("float32")) * 0.1
    feat = O.corrblock1d_lookup(O.corr_pyramid(O.all_pairs_corr(c["f1"], c["f2"]), Lv), disp, coords, 4, exact=True)
    ref = torch.relu(torch.nn.functional.conv2d(feat.double(), w.double(), b.double())).float()
    A.set_corr_mode("fp32")
    A.set_update_engine(engine)
    blk = A.CorrBlock1D(c["f1"].cuda(), c["f2"].cuda(), num_levels=Lv, radius=4)
    split = engine == "bf16x3"
    d = blk.deferred(disp.cuda(), coords.cuda())
    assert d.fusable and d.shape == (B, Lv * 9, H, W)
    w_hi, w_lo = type(d).pack_convc1_weight(w.cuda(), split)
    out_hi = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    out_lo = torch.full_like(out_hi, float("nan")) if split else None
    d.convc1_planes(w_hi, w_lo, b.cuda(), out_hi, out_lo)
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    widen = (lambda t: t.view(torch.float16).float()) if engine == "fp16" else (lambda t: t.float())
    got = widen(out_hi) + (widen(out_lo) if split else 0)
    assert rel(got.permute(0, 3, 1, 2), ref) < tol


def test_fused_raft_lookup_matches_unfused_loop(A):
    c = cases.loop_case("raft", seed=79, B=2, H=13, W=37)
    m = make_block(A, "raft", 9)
    A.set_update_engine("bf16x3")
    A.set_corr_mode("bf16x3")
    args = [c["f1"].cuda(), c["f2"].cuda(), [t.cuda() for t in c["net"]], [[t.cuda() for t in l] for l in c["inp"]]]
    prev = A.set_lookup_fusion(True)
    d_f, net_f = A.raft_iterations(m, *args, 6)
    A.set_lookup_fusion(False)
    d_u, net_u = A.raft_iterations(m, *args, 6)
    A.set_lookup_fusion(prev)
    A.set_update_engine("fp32")
    A.set_corr_mode("fp32")
    assert rel(d_f, d_u) < 1e-5
    for a, b in zip(net_f, net_u):
        assert rel(a, b) < 1e-5


@pytest.mark.parametrize("family", ["igev", "raft"])
def test_encoder_side_stream_overlap_is_exact(A, family):
    """The motion encoder forked onto a side stream (default) produces bit-identical results to the single-stream
    schedule, eagerly and across repeated calls (stream/event ordering, record_stream lifetime)."""
    c = cases.loop_case(family, seed=91, B=2, H=24, W=40)
    m = make_block(A, family, 3)
    A.set_update_engine("bf16x3")
    A.set_corr_mode("bf16x3")
    net = [t.cuda() for t in c["net"]]
    inp = [[t.cuda() for t in l] for l in c["inp"]]

    def run():
        if family == "igev":
            return A.igev_iterations(m, c["f1"].cuda(), c["f2"].cuda(), c["geo"].cuda(), net, inp, c["init_disp"].cuda(), 5)[0]
        return A.raft_iterations(m, c["f1"].cuda(), c["f2"].cuda(), net, inp, 5)[0]

    A.update_umma.set_encoder_overlap(False)
    ref = run()
    A.update_umma.set_encoder_overlap(True)
    outs = [run() for _ in range(3)]
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    A.set_corr_mode("fp32")
    for o in outs:
        assert torch.equal(o, ref)


@pytest.mark.parametrize("engine,tol", [("fp32", 1e-4), ("bf16x3", 1e-3), ("f16f8", 1e-3)])
def test_slow_fast_gru_loop_vs_reference(A, golden, engine, tol):
    """args.slow_fast_gru (continuous_IGEVstereo.py:288-291): two extra low-resolution GRU passes per iteration."""
    c = cases.loop_case("igev", seed=61, B=1, H=16, W=24)
    g = golden("loop_igev_slowfast")                         # the reference's own classes, tests/golden/make_golden.py
    iters = int(g["iters"])
    ref = torch.from_numpy(g["disps"][-1])
    m = make_block(A, "igev", 8)
    A.set_update_engine(engine)
    A.set_corr_mode("fp32" if engine == "fp32" else "bf16x3")
    disp, _ = A.igev_iterations(m, c["f1"].cuda(), c["f2"].cuda(), c["geo"].cuda(), [t.cuda() for t in c["net"]],
                                [[t.cuda() for t in l] for l in c["inp"]], c["init_disp"].cuda(), iters, slow_fast_gru=True)
    torch.cuda.synchronize()
    A.set_update_engine("fp32")
    A.set_corr_mode("fp32")
    assert rel(disp, ref) < tol


def _decode_x8(lo_plane, C, weights=False):
    """AS_FMT_F16F8 "lo" plane (csrc/common.cuh): per 64-channel chunk 128 bytes [first | second] of e5m2 values.
    Activations: first = lo * 2^6, second = hi * 2^-8; weights: first = hi * 2^-6, second = lo * 2^8."""
    raw = lo_plane.contiguous().view(torch.uint8).reshape(-1, C // 64, 2, 64)
    f8 = raw.view(torch.float8_e5m2).float()
    first, second = f8[:, :, 0, :].reshape(-1, C), f8[:, :, 1, :].reshape(-1, C)
    if weights:
        return first * 64.0, second / 256.0           # hi, lo
    return second * 256.0, first / 64.0               # hi, lo


def test_f16f8_plane_and_weight_encoding(A):
    """The 2-pass parity format: hi = IEEE half, lo plane = e5m2 pairs with power-of-two scales (bit-exact encoding)."""
    L = A._lib
    torch.manual_seed(3)
    x = (torch.randn(2, 5, 7, 128, device="cuda") * torch.logspace(-3, 1, 128, device="cuda")).contiguous()
    A.set_update_engine("f16f8")
    try:
        hi = torch.empty(x.shape, device="cuda", dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        L.call("as_split_f32", x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), L.stream_ptr())
        w = torch.randn(96, 64, 3, 3, device="cuda") * 0.05
        whi = torch.empty((96, 9 * 64), device="cuda", dtype=torch.bfloat16)
        wlo = torch.empty_like(whi)
        L.call("as_pack_conv_weight_bf16", w.data_ptr(), whi.data_ptr(), wlo.data_ptr(), 96, 64, 3, 3, 96, 64, L.stream_ptr())
        torch.cuda.synchronize()
    finally:
        A.set_update_engine("fp32")
    xh = x.half().float()
    assert torch.equal(hi.view(torch.float16).float(), xh)
    dec_hi, dec_lo = _decode_x8(lo, 128)
    exp_lo = ((x - xh) * 64.0).to(torch.float8_e5m2).float() / 64.0
    exp_hi = (xh / 256.0).to(torch.float8_e5m2).float() * 256.0
    assert torch.equal(dec_lo, exp_lo.reshape(-1, 128))
    assert torch.equal(dec_hi, exp_hi.reshape(-1, 128))
    wk = w.permute(0, 2, 3, 1).reshape(96, 9 * 64)                    # [n][tap][c]
    wh = wk.half().float()
    assert torch.equal(whi.view(torch.float16).float(), wh)
    dwh, dwl = _decode_x8(wlo, 9 * 64, weights=True)
    assert torch.equal(dwh, ((wh / 64.0).to(torch.float8_e5m2).float() * 64.0))
    assert torch.equal(dwl, (((wk - wh) * 256.0).to(torch.float8_e5m2).float() / 256.0))


@pytest.mark.parametrize("shape", [(1, 16, 24), (2, 40, 72)])
def test_f16f8_conv_matches_fp32(A, shape):
    """One 3x3 convolution (N = 256, z|r shape) on the 2-pass engine against torch fp64: error at the level of the 3-pass
    split (<= 1e-4 of max|ref|), far below a single half pass."""
    import types
    B, H, W = shape
    torch.manual_seed(5)
    args = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=3)
    ub = A.BasicMultiUpdateBlock(args, hidden_dims=[128, 128, 128]).cuda().eval()
    from anystereo_b200 import update_umma as U
    x = torch.randn(B, H, W, 128, device="cuda")
    errs = {}
    for engine in ("f16f8", "bf16x3", "fp16"):
        A.set_update_engine(engine)
        split = engine != "fp16"
        ns = {"f16f8": 2, "bf16x3": 3, "fp16": 1}[engine]
        pl = U._Planes(x.shape, "cuda", split)
        A._lib.call("as_split_f32", x.data_ptr(), pl.hi.data_ptr(), A._lib.ptr(pl.lo), x.numel(), A._lib.stream_ptr())
        wt = U._weights(ub, "t." + engine, [ub.disp_head.conv1], split=split)
        out = torch.empty(B, H, W, 256, device="cuda")
        U._conv(B, H, W, [pl], wt, ns, A._lib.UEPI_LINEAR_F32, out_f32=out)
        torch.cuda.synchronize()
        ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), ub.disp_head.conv1.weight.double(),
                                         ub.disp_head.conv1.bias.double(), padding=1).permute(0, 2, 3, 1)
        errs[engine] = float((out.double() - ref).abs().max() / ref.abs().max())
    A.set_update_engine("fp32")
    assert errs["f16f8"] < 1e-4 and errs["bf16x3"] < 1e-4, errs
    assert errs["fp16"] > 3 * errs["f16f8"], errs


def test_iterations_replay_their_captured_step_by_default(A):
    """igev_iterations / raft_iterations replay a CUDA graph of the step by default: same bits as the
    eager launches, rebuilt when a parameter changes, new inputs honoured, launches accounted for."""
    c = cases.loop_case("igev", seed=9, B=1, H=16, W=24)
    args = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=3)
    m = A.BasicMultiUpdateBlock(args, hidden_dims=[128, 128, 128])
    m.load_state_dict(O.make_update_block_params(162, seed=5), strict=True)
    m = m.cuda().eval()
    cu = lambda t: t.cuda()                                                      # noqa: E731
    A.set_update_engine("f16f8")
    A.set_corr_mode("bf16x3")

    def run(scale=1.0):
        return A.igev_iterations(m, cu(c["f1"]) * scale, cu(c["f2"]), cu(c["geo"]), [cu(t) for t in c["net"]],
                                 [[cu(t) for t in l] for l in c["inp"]], cu(c["init_disp"]), 4)
    try:
        prev = A.set_graph_replay(False)
        d_eager, n_eager = run()
        d_eager2, _ = run(0.5)
        A.set_graph_replay(True)
        n0 = A._lib.launch_count
        d1, n1 = run()
        built = A._lib.launch_count - n0
        d2, n2 = run()                                   # pure replay
        replayed = A._lib.launch_count - n0 - built
        assert len(A.hotpath._GRAPH["cache"]) == 1 and replayed > 20
        assert torch.equal(d1, d_eager) and torch.equal(d2, d_eager)
        for a, b in zip(n2, n_eager):
            assert torch.equal(a, b)
        d3, _ = run(0.5)                                 # same graph, different inputs
        assert torch.equal(d3, d_eager2) and len(A.hotpath._GRAPH["cache"]) == 1
        with torch.no_grad():
            m.disp_head.conv2.bias.add_(0.25)            # parameter version changes -> the step is captured again
        d4, _ = run()
        assert len(A.hotpath._GRAPH["cache"]) == 2
        A.set_graph_replay(False)
        d4e, _ = run()
        assert torch.equal(d4, d4e) and not torch.equal(d4, d1)
    finally:
        A.set_graph_replay(prev)
        A.hotpath.graph_cache_clear()
        A.set_update_engine("fp32")
        A.set_corr_mode("fp32")


@pytest.mark.parametrize("family", ["igev", "raft"])
def test_lowres_single_pass_operator_errors(A, golden, family):
    """The default "f16f8" engine runs the 1/8- and 1/16-resolution GRUs in ONE half pass (set_lowres_single_pass): one
    update-block call against the reference golden stays inside the 2e-4 operator tolerance on every output (measured
    net0 1.9e-5, net1 1.1e-4, net2 8e-5, delta 2e-5), and two passes everywhere (the knob off) differ measurably."""
    g = golden("update_block_" + family)
    c = cases.update_block_case(family)
    m = make_block(A, family, 11)
    A.set_update_engine("f16f8")
    prev = A.set_lowres_single_pass(True)
    try:
        inp = [[t.cuda() for t in lst] for lst in c["inp"]]
        with torch.no_grad():
            net, delta = m([t.cuda() for t in c["net"]], inp, c["corr"].cuda(), c["disp"].cuda())
            A.set_lowres_single_pass(False)
            m.reset_caches()
            net2, delta2 = m([t.cuda() for t in c["net"]], inp, c["corr"].cuda(), c["disp"].cuda())
        errs = [rel(net[i], g["full_net%d" % i]) for i in range(3)] + [rel(delta, g["full_delta"])]
        errs2 = [rel(net2[i], g["full_net%d" % i]) for i in range(3)] + [rel(delta2, g["full_delta"])]
    finally:
        A.set_lowres_single_pass(prev)
        A.set_update_engine("fp32")
    assert max(errs) < 2e-4, errs
    assert max(errs2) < 2e-4 and not torch.equal(net[1], net2[1]), errs2
    print("%s: one pass at low resolution %s | two passes everywhere %s" % (family, ["%.1e" % e for e in errs], ["%.1e" % e for e in errs2]))


def test_gate_weight_residual_only_operator_errors(A, golden):
    """The IGEV update block's 1/4-resolution gates under "f16f8": hi*hi + the weight-residual cross term only (kernel mode
    nsplit = 4).  One call against the reference golden stays inside the 2e-4 operator tolerance; both cross terms
    (set_gate_weight_residual_only(False)) differ measurably; the RAFT block's default is both."""
    g = golden("update_block_igev")
    c = cases.update_block_case("igev")
    m = make_block(A, "igev", 11)
    assert m.gate_weight_residual_only and not make_block(A, "raft", 11).gate_weight_residual_only
    A.set_update_engine("f16f8")
    prev = A.set_gate_weight_residual_only(None)
    try:
        inp = [[t.cuda() for t in lst] for lst in c["inp"]]
        with torch.no_grad():
            net, delta = m([t.cuda() for t in c["net"]], inp, c["corr"].cuda(), c["disp"].cuda())
            A.set_gate_weight_residual_only(False)
            m.reset_caches()
            net2, delta2 = m([t.cuda() for t in c["net"]], inp, c["corr"].cuda(), c["disp"].cuda())
        errs = [rel(net[i], g["full_net%d" % i]) for i in range(3)] + [rel(delta, g["full_delta"])]
        errs2 = [rel(net2[i], g["full_net%d" % i]) for i in range(3)] + [rel(delta2, g["full_delta"])]
    finally:
        A.set_gate_weight_residual_only(prev)
        A.set_update_engine("fp32")
    print("weight residual only %s | both cross terms %s" % (["%.1e" % e for e in errs], ["%.1e" % e for e in errs2]))
    assert max(errs) < 2e-4, errs
    assert max(errs2) < 2e-4 and not torch.equal(net[0], net2[0]), errs2
