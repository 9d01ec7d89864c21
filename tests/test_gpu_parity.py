"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors made from
the reference, on the same seeded inputs.  Tolerances follow BASELINE.json:
  * lookup indexing / pooling index math: bit-exact;
  * correlation and lookup values: <= 1e-4 relative (max|ours-ref| / max|ref|) in fp32;
  * final disparity after equal iterations: mean |delta| <= 0.01 px (x4: low-res -> full-res pixels).
"""
import types

import numpy as np
import pytest
import torch

import cases
from oracle import hotpath_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-4


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def A():
    import anystereo_b200 as a
    assert a._lib.lib().as_compiled_sm() == 100
    a.set_corr_mode("fp32")
    a.set_update_engine("fp32")
    return a


def cu(t):
    return t.cuda()


def make_block(A, family, seed):
    cls = A.BasicMultiUpdateBlock if family == "igev" else A.BasicMultiUpdateBlockRAFT
    args = types.SimpleNamespace(corr_levels=2 if family == "igev" else 4, corr_radius=4, n_gru_layers=3)
    m = cls(args, hidden_dims=[128, 128, 128])
    p = O.make_update_block_params(162 if family == "igev" else 36, seed=seed)
    m.load_state_dict(p, strict=True)
    return m.cuda().eval(), p


# ---------------------------------------------------------------------------------------------------
# a6 corr_sampler
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float64, 1e-12), (torch.float16, 2e-3)])
@pytest.mark.parametrize("shape", [(2, 3, 7, 9), (1, 5, 23, 23), (2, 4, 40, 37), (1, 2, 130, 64)])
def test_sampler_fwd_bwd(A, dtype, tol, shape):
    B, H, W1, W2 = shape
    rng = np.random.RandomState(B * 1000 + W2)
    vol = torch.from_numpy(rng.standard_normal(shape)).to(dtype)
    coords = torch.from_numpy(rng.uniform(-6, W2 + 6, size=(B, 2, H, W1))).float()
    coords[0, 0, 0, :3] = torch.tensor([0.0, float(W2 - 1), -4.0])[:min(3, W1)]
    g = torch.from_numpy(rng.standard_normal((B, 9, H, W1))).to(dtype)
    out, = A.corr_sampler.forward(cu(vol), cu(coords), 4)
    ref = O.sampler_forward(vol.double(), coords, 4)
    assert out.dtype == dtype and tuple(out.shape) == (B, 9, H, W1)
    assert rel(out, ref) < tol
    gv, = A.corr_sampler.backward(cu(vol), cu(coords), cu(g), 4)
    refg = O.sampler_backward(vol.double(), coords, g.double(), 4)
    assert gv.dtype == dtype and gv.shape == vol.shape
    assert rel(gv, refg) < tol
    # autograd wrapper
    v = cu(vol).requires_grad_(True)
    o = A.corr_sampler.CorrSampler.apply(v, cu(coords), 4)
    o.backward(cu(g))
    assert rel(v.grad, refg) < tol


def test_sampler_radius_variants(A):
    rng = np.random.RandomState(5)
    vol = torch.from_numpy(rng.standard_normal((1, 3, 11, 13))).float()
    coords = torch.from_numpy(rng.uniform(-3, 15, size=(1, 1, 3, 11))).float()
    for r in (0, 1, 2, 7):
        out, = A.corr_sampler.forward(cu(vol), cu(coords), r)
        assert rel(out, O.sampler_forward(vol, coords, r)) < 2e-6


def test_sampler_errors(A):
    v = torch.zeros(1, 2, 3, 4)
    c = torch.zeros(1, 1, 2, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        A.corr_sampler.forward(v, c, 4)
    with pytest.raises(RuntimeError, match="must be contiguous"):
        A.corr_sampler.forward(cu(torch.zeros(1, 2, 4, 3)).transpose(2, 3), cu(c), 4)


# ---------------------------------------------------------------------------------------------------
# index math: bit-exact
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("level", [0, 1, 3])
def test_tap_indices_bit_exact(A, kind, level):
    import ctypes
    L = A._lib
    rng = np.random.RandomState(level * 2 + kind)
    B, H, W = 2, 5, 23
    fields = cases.disparity_fields(rng, B, H, W, 12)
    for name, d in fields.items():
        disp = torch.from_numpy(d.astype("float32")).cuda()
        coords = O.pixel_coords(B, H, W).cuda()
        tap = torch.empty(B * H * W, dtype=torch.int32, device="cuda")
        frac = torch.empty(B * H * W, dtype=torch.float32, device="cuda")
        L.call("as_lookup_taps", disp.data_ptr(), coords.data_ptr(), B, H, W, 4, level, kind, tap.data_ptr(),
               frac.data_ptr(), L.stream_ptr())
        xg, xc = O._level_positions(disp.cpu(), coords.cpu(), level)
        t_ref, f_ref = O.tap_indices(xg if kind == 0 else xc, 4)
        assert torch.equal(tap.cpu(), t_ref), name
        assert torch.equal(frac.cpu(), f_ref), name


# ---------------------------------------------------------------------------------------------------
# a1/a2/a3 RAFT
# ---------------------------------------------------------------------------------------------------
def test_corr_and_pyramid_raft(A, golden):
    g = golden("raft_corrblock")
    c = cases.raft_corr_case()
    corr = A.CorrBlock1D.corr(cu(c["f1"]), cu(c["f2"]))
    assert tuple(corr.shape) == g["corr"].shape and corr.is_contiguous()
    assert rel(corr, g["corr"]) < REL
    blk = A.CorrBlock1D(cu(c["f1"]), cu(c["f2"]), num_levels=c["L"], radius=c["r"])
    assert blk.num_levels == c["L"] and blk.radius == c["r"]
    prev = None
    for i, lvl in enumerate(blk.init_corr_pyramid):
        assert tuple(lvl.shape) == g["pyr%d" % i].shape
        assert rel(lvl, g["pyr%d" % i]) < REL
        if prev is not None:   # pooling index math: bit-exact against (a+b)*0.5 of OUR finer level
            assert torch.equal(lvl.cpu(), O.halve_last(prev.cpu()))
        prev = lvl


@pytest.mark.parametrize("dist", ["uniform", "smooth", "integer", "negative", "far_oob", "edge"])
def test_lookup_raft(A, golden, dist):
    g = golden("raft_corrblock")
    c = cases.raft_corr_case()
    B, _, H, W = c["f1"].shape
    blk = A.CorrBlock1D(cu(c["f1"]), cu(c["f2"]), num_levels=c["L"], radius=c["r"])
    out = blk(cu(c["disps"][dist]), O.pixel_coords(B, H, W).cuda())
    ref = g["lookup_" + dist]
    assert tuple(out.shape) == ref.shape and out.dtype == torch.float32 and out.is_contiguous()
    if np.abs(ref).max() == 0:
        assert float(out.abs().max()) == 0
    else:
        assert rel(out, ref) < REL
    # against the exact-index oracle on OUR pyramid: only fma rounding apart
    pyr = [p.cpu().contiguous() for p in blk.init_corr_pyramid]
    ex = O.corrblock1d_lookup(pyr, c["disps"][dist], O.pixel_coords(B, H, W), c["r"], exact=True)
    assert float((out.cpu() - ex).abs().max()) <= 2e-6 * max(1.0, float(ex.abs().max()))


# ---------------------------------------------------------------------------------------------------
# a2/a4 IGEV
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dist", ["uniform", "smooth", "integer", "negative", "far_oob", "edge"])
def test_lookup_igev(A, golden, dist):
    g = golden("igev_geovolume")
    c = cases.igev_geo_case()
    B, _, H, W = c["f1"].shape
    blk = A.Combined_Geo_Encoding_Volume(cu(c["f1"]), cu(c["f2"]), cu(c["geo"]), num_levels=c["L"], radius=c["r"])
    for i in range(c["L"]):
        assert tuple(blk.geo_volume_pyramid[i].shape) == g["geo_pyr%d" % i].shape
        assert np.array_equal(blk.geo_volume_pyramid[i].cpu().numpy(), g["geo_pyr%d" % i])   # bit-exact
        assert rel(blk.init_corr_pyramid[i], g["corr_pyr%d" % i]) < REL
    out = blk(cu(c["disps"][dist]), O.pixel_coords(B, H, W).cuda())
    ref = g["lookup_" + dist]
    assert tuple(out.shape) == ref.shape == (B, 162, H, W) and out.is_contiguous()
    assert rel(out, ref) < REL


@pytest.mark.parametrize("G,Dg,L,r", [(4, 10, 3, 2), (8, 16, 1, 4), (8, 24, 3, 4), (2, 7, 2, 3)])
def test_lookup_igev_generic_shapes(A, G, Dg, L, r):
    """Non-default group counts / radii / level counts go through the generic kernel."""
    rng = np.random.RandomState(G * 100 + Dg)
    B, D, H, W = 1, 8, 3, 21
    f1 = torch.from_numpy(rng.standard_normal((B, D, H, W))).float()
    f2 = torch.from_numpy(rng.standard_normal((B, D, H, W))).float()
    geo = torch.from_numpy(rng.standard_normal((B, G, Dg, H, W))).float()
    disp = torch.from_numpy(rng.uniform(-2, Dg + 2, size=(B, 1, H, W))).float()
    blk = A.Combined_Geo_Encoding_Volume(cu(f1), cu(f2), cu(geo), num_levels=L, radius=r)
    out = blk(cu(disp), O.pixel_coords(B, H, W).cuda())
    cp = O.corr_pyramid(O.all_pairs_corr(f1, f2), L)
    gp = O.geo_pyramid(geo, L)
    ref = O.geo_lookup(gp, cp, disp, O.pixel_coords(B, H, W), r, exact=False)
    assert rel(out, ref) < REL


def test_lookup_medium_vs_oracle(A):
    """IGEV shape at 1/4 of KITTI (oracle finishes in seconds): fast kernel, ragged tail CTA."""
    rng = np.random.RandomState(77)
    B, D, H, W, Dg = 2, 96, 11, 78, 48
    f1 = torch.from_numpy(rng.standard_normal((B, D, H, W))).float()
    f2 = torch.from_numpy(rng.standard_normal((B, D, H, W))).float()
    geo = torch.from_numpy(rng.standard_normal((B, 8, Dg, H, W))).float()
    blk = A.Combined_Geo_Encoding_Volume(cu(f1), cu(f2), cu(geo), num_levels=2, radius=4)
    cp = O.corr_pyramid(O.all_pairs_corr(f1, f2), 2)
    gp = O.geo_pyramid(geo, 2)
    coords = O.pixel_coords(B, H, W)
    for name, d in cases.disparity_fields(rng, B, H, W, Dg).items():
        disp = torch.from_numpy(d.astype("float32"))
        out = blk(cu(disp), coords.cuda())
        ref = O.geo_lookup(gp, cp, disp, coords, 4, exact=False)
        assert rel(out, ref) < REL, name


# ---------------------------------------------------------------------------------------------------
# a7 GWC
# ---------------------------------------------------------------------------------------------------
def test_gwc(A, golden):
    g = golden("gwc_volume")
    for name, c in cases.gwc_cases().items():
        out = A.build_gwc_volume(cu(c["left"]), cu(c["right"]), c["maxdisp"], c["groups"])
        assert tuple(out.shape) == g[name].shape and out.dtype == torch.float32
        assert rel(out, g[name]) < 1e-5, name
        # structural zeros (x < d) are exact zeros
        ref0 = torch.from_numpy(g[name]) == 0
        assert bool((out.cpu()[ref0] == 0).all())
    h = A.build_gwc_volume(cu(c["left"]).half(), cu(c["right"]).half(), c["maxdisp"], c["groups"])
    assert h.dtype == torch.float16


# ---------------------------------------------------------------------------------------------------
# a8-a11 update block
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("family", ["igev", "raft"])
def test_update_block(A, golden, family):
    g = golden("update_block_" + family)
    c = cases.update_block_case(family)
    m, _ = make_block(A, family, 11)
    names = [n for n, _ in m.named_parameters()]
    assert names[0] == "encoder.convc1.weight" and "gru04.convz.weight" in names and names[-1] == "disp_head.conv2.bias"
    inp = [[cu(t) for t in lst] for lst in c["inp"]]
    with torch.no_grad():
        net, delta = m([cu(t) for t in c["net"]], inp, cu(c["corr"]), cu(c["disp"]))
        for i in range(3):
            assert tuple(net[i].shape) == g["full_net%d" % i].shape
            assert rel(net[i], g["full_net%d" % i]) < REL
        assert rel(delta, g["full_delta"]) < REL
        net = m([cu(t) for t in c["net"]], inp, iter16=True, iter08=False, iter04=False, update=False)
        for i in range(3):
            assert rel(net[i], g["only16_net%d" % i]) < REL
        net = m([cu(t) for t in c["net"]], inp, iter16=True, iter08=True, iter04=False, update=False)
        for i in range(3):
            assert rel(net[i], g["lowres_net%d" % i]) < REL


@pytest.mark.parametrize("family", ["igev", "raft"])
def test_iteration_loop_epe(A, golden, family):
    g = golden("loop_" + family)
    c = cases.loop_case(family)
    iters = int(g["iters"])
    m, _ = make_block(A, family, 12 if family == "igev" else 13)
    net = [cu(t) for t in c["net"]]
    inp = [[cu(t) for t in lst] for lst in c["inp"]]
    if family == "igev":
        disp, net, hist = A.igev_iterations(m, cu(c["f1"]), cu(c["f2"]), cu(c["geo"]), net, inp, cu(c["init_disp"]),
                                            iters, keep_all=True)
    else:
        disp, net, hist = A.raft_iterations(m, cu(c["f1"]), cu(c["f2"]), net, inp, iters, keep_all=True)
    ref = torch.from_numpy(g["disps"])
    err = (torch.stack([h.cpu() for h in hist]) - ref).abs()
    epe_fullres = float(err[-1].mean()) * 4
    assert epe_fullres < 0.01, epe_fullres
    assert epe_fullres < 1e-3          # the fp32 engine is far inside the gate
    for i in range(3):
        assert rel(net[i], g["net%d" % i]) < 1e-3


# ---------------------------------------------------------------------------------------------------
# a13 adjoints
# ---------------------------------------------------------------------------------------------------
def test_adjoints(A, golden):
    g = golden("adjoints")
    c = cases.igev_geo_case(seed=8, B=1, D=24, H=3, W=14, Dg=16)
    f1 = cu(c["f1"]).requires_grad_(True)
    f2 = cu(c["f2"]).requires_grad_(True)
    geo = cu(c["geo"]).requires_grad_(True)
    blk = A.Combined_Geo_Encoding_Volume(f1, f2, geo, num_levels=2, radius=4)
    out = blk(cu(c["disps"]["uniform"]), O.pixel_coords(1, 3, 14).cuda())
    (out * cu(torch.from_numpy(g["cot"]))).sum().backward()
    assert rel(geo.grad, g["g_geo"]) < REL
    assert rel(f1.grad, g["g_f1"]) < REL
    assert rel(f2.grad, g["g_f2"]) < REL
    gc = cases.gwc_cases()["odd"]
    Lf = cu(gc["left"]).requires_grad_(True)
    Rf = cu(gc["right"]).requires_grad_(True)
    vol = A.build_gwc_volume(Lf, Rf, gc["maxdisp"], gc["groups"])
    (vol * cu(torch.from_numpy(g["gwc_cot"]))).sum().backward()
    assert rel(Lf.grad, g["gwc_gL"]) < 1e-5
    assert rel(Rf.grad, g["gwc_gR"]) < 1e-5


def test_raft_lookup_adjoint(A):
    rng = np.random.RandomState(3)
    B, D, H, W = 1, 16, 3, 26
    f1 = torch.from_numpy(rng.standard_normal((B, D, H, W))).float()
    f2 = torch.from_numpy(rng.standard_normal((B, D, H, W))).float()
    disp = torch.from_numpy(rng.uniform(0, 12, size=(B, 1, H, W))).float()
    cot = torch.from_numpy(rng.standard_normal((B, 36, H, W))).float()
    a1, a2 = cu(f1).requires_grad_(True), cu(f2).requires_grad_(True)
    blk = A.CorrBlock1D(a1, a2, num_levels=4, radius=4)
    (blk(cu(disp), O.pixel_coords(B, H, W).cuda()) * cu(cot)).sum().backward()
    r1, r2 = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
    pyr = O.corr_pyramid(O.all_pairs_corr(r1, r2), 4)
    (O.corrblock1d_lookup(pyr, disp, O.pixel_coords(B, H, W), 4) * cot).sum().backward()
    assert rel(a1.grad, r1.grad) < REL
    assert rel(a2.grad, r2.grad) < REL


# ---------------------------------------------------------------------------------------------------
# full-size (BASELINE.json config 2 shapes) size-independent properties
# ---------------------------------------------------------------------------------------------------
def test_fullsize_properties_igev(A):
    torch.manual_seed(0)
    B, D, H, W, Dg = 2, 96, 96, 312, 48          # 384x1248 at 1/4; B=2 keeps the test short
    dev = "cuda"
    f1 = torch.randn(B, D, H, W, device=dev)
    f2 = torch.randn(B, D, H, W, device=dev)
    gwc = A.build_gwc_volume(f1, f2, Dg, 8)
    # GWC: d = 0 slice is the per-group mean of the product; x < d is zero
    ref0 = (f1 * f2).view(B, 8, 12, H, W).mean(2)
    assert rel(gwc[:, :, 0], ref0) < 1e-5
    assert float(gwc[:, :, 5, :, :5].abs().max()) == 0.0
    ref7 = (f1[..., 7:] * f2[..., :-7]).view(B, 8, 12, H, W - 7).mean(2)
    assert rel(gwc[:, :, 7, :, 7:], ref7) < 1e-5
    blk = A.Combined_Geo_Encoding_Volume(f1, f2, gwc, num_levels=2, radius=4)
    coords = O.pixel_coords(B, H, W).to(dev)
    # (1) integer disparities: interpolation weights are exactly (1,0) -> outputs are exact gathers
    disp = torch.randint(0, Dg, (B, 1, H, W), device=dev).float()
    out = blk(disp, coords)
    geo0 = blk.geo_volume_pyramid[0].squeeze(2)            # [N,8,Dg] view
    N = B * H * W
    idx = (disp.reshape(N, 1, 1).long() + torch.arange(-4, 5, device=dev).view(1, 1, 9))
    ok = (idx >= 0) & (idx < Dg)
    gath = torch.gather(geo0, 2, idx.clamp(0, Dg - 1).expand(N, 8, 9)) * ok
    got = out[:, :72].permute(0, 2, 3, 1).reshape(N, 8, 9)
    assert torch.equal(got, gath)
    corr0 = blk.init_corr_pyramid[0].reshape(N, W)
    xs = torch.arange(W, device=dev).view(1, 1, W).expand(B, H, W).reshape(N, 1)
    idc = xs - disp.reshape(N, 1).long() + torch.arange(-4, 5, device=dev).view(1, 9)
    okc = (idc >= 0) & (idc < W)
    gc = torch.gather(corr0, 1, idc.clamp(0, W - 1)) * okc
    assert torch.equal(out[:, 72:81].permute(0, 2, 3, 1).reshape(N, 9), gc)
    # (2) linearity in the volumes: lookup(2*V) == 2*lookup(V) exactly (power of two)
    blk2 = A.Combined_Geo_Encoding_Volume(f1 * 2, f2, gwc * 2, num_levels=2, radius=4)
    d2 = torch.rand(B, 1, H, W, device=dev) * Dg
    assert torch.equal(blk2(d2, coords), 2 * blk(d2, coords))
    # (3) correlation symmetry: corr(f1,f2)[x1,x2] == corr(f2,f1)[x2,x1]
    c12 = A.CorrBlock1D.corr(f1[:1], f2[:1]).squeeze(3)
    c21 = A.CorrBlock1D.corr(f2[:1], f1[:1]).squeeze(3)
    assert rel(c12, c21.transpose(2, 3)) < 1e-6
    # (4) sampler == level-0 corr channels of the fused lookup
    x0 = (coords.reshape(B, 1, H, W) - d2).contiguous()
    smp, = A.corr_sampler.forward(blk.init_corr_pyramid[0].reshape(B, H, W, W).contiguous(), x0, 4)
    assert rel(smp, blk(d2, coords)[:, 72:81]) < 1e-6


# ---- initial-disparity head (classifier Conv3d + softmax + disparity_regression) ----
def test_init_disparity_golden(A, golden):
    g = golden("init_disparity")
    c = cases.init_disp_case()
    disp, prob = A.init_disparity(c["geo"].cuda(), c["weight"].cuda(), return_prob=True)
    torch.cuda.synchronize()
    assert tuple(disp.shape) == g["init_disp"].shape and tuple(prob.shape) == g["prob"].shape
    assert float((prob.cpu() - torch.from_numpy(g["prob"])).abs().max()) < 5e-6
    assert float((disp.cpu() - torch.from_numpy(g["init_disp"])).abs().max()) < 1e-4 * float(np.abs(g["init_disp"]).max())
    d2 = A.disparity_regression(torch.from_numpy(g["prob"]).cuda(), 12)
    assert float((d2.cpu() - torch.from_numpy(g["init_disp"])).abs().max()) < 1e-5


@pytest.mark.parametrize("shape", [(1, 8, 48, 7, 70), (2, 8, 17, 3, 33), (1, 4, 64, 2, 5), (1, 8, 1, 4, 4)])
def test_init_disparity_vs_oracle(A, shape):
    B, G, D, H, W = shape
    rng = np.random.RandomState(D + W)
    geo = torch.from_numpy(rng.standard_normal(shape).astype("float32")) * 3
    w = torch.from_numpy(rng.standard_normal((1, G, 3, 3, 3)).astype("float32")) * 0.2
    ref_d, ref_p = O.init_disparity(geo, w)
    disp, prob = A.init_disparity(geo.cuda(), w.cuda(), return_prob=True)
    torch.cuda.synchronize()
    assert float((prob.cpu() - ref_p).abs().max()) < 5e-6
    assert float((disp.cpu() - ref_d).abs().max()) < 1e-4 * max(1.0, float(ref_d.abs().max()))
    only = A.init_disparity(geo.cuda(), w.cuda())
    assert torch.equal(only, disp)
    with pytest.raises(RuntimeError):
        A.init_disparity(geo.cuda(), w.cuda()[:, :, :2])


def test_init_disparity_fullsize_config2(A):
    """8 x [8,48,96,312] (368 MB): against torch conv3d + softmax on the same GPU (strict fp32)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(4)
    geo = torch.randn(8, 8, 48, 96, 312, device="cuda")
    w = torch.randn(1, 8, 3, 3, 3, device="cuda") * 0.2
    disp = A.init_disparity(geo, w)
    ref = O.disparity_regression(torch.softmax(torch.nn.functional.conv3d(geo, w, padding=1).squeeze(1), dim=1), 48)
    torch.cuda.synchronize()
    assert float((disp - ref).abs().max()) < 2e-3           # disparities up to 47; 1e-4 relative


def test_model_level_init_disparity(A, golden):
    """Initial disparity of the real reference model graph (tests/golden/model_igev_boundary.npz) from its own
    geometry-encoding volume and classifier weight, through the fused kernel."""
    g = golden("model_igev_boundary")
    disp = A.init_disparity(torch.from_numpy(g["geo"]).cuda(), torch.from_numpy(g["classifier_weight"]).cuda())
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["init_disp"])
    assert float((disp.cpu() - ref).abs().max()) < 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize("shape", [(1, 96, 5, 45, 48), (2, 96, 7, 32, 20), (1, 48, 3, 70, 48), (1, 96, 4, 9, 13)])
@pytest.mark.parametrize("with_att", [False, True])
def test_gwc_corr_stem_matches_unfused(A, shape, with_att):
    """build_gwc_volume + Conv3d(8,8,3) + eval BatchNorm3d + LeakyReLU (+ FeatureAtt multiply) in one kernel
    against the same chain in torch (fp64) over the oracle's volume; ragged widths, maxdisp > W, 6 / 12 channels per group."""
    B, C, H, W, D = shape
    g = torch.Generator().manual_seed(H * 100 + W)
    f1 = torch.randn(B, C, H, W, generator=g)
    f2 = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(8, 8, 3, 3, 3, generator=g) * 0.2
    scale = torch.rand(8, generator=g) + 0.5
    shift = torch.randn(8, generator=g) * 0.3
    att = torch.sigmoid(torch.randn(B, 8, H, W, generator=g)) if with_att else None
    vol = O.gwc_volume(f1, f2, D, 8).double()
    ref = torch.nn.functional.conv3d(vol, w.double(), padding=1)
    ref = ref * scale.double().view(1, 8, 1, 1, 1) + shift.double().view(1, 8, 1, 1, 1)
    ref = torch.nn.functional.leaky_relu(ref, 0.01)
    if with_att:
        ref = ref * att.double().unsqueeze(2)
    got = A.gwc_corr_stem(f1.cuda(), f2.cuda(), D, 8, w.cuda(), scale.cuda(), shift.cuda(), 0.01,
                          att.cuda() if with_att else None)
    assert got.shape == (B, 8, D, H, W)
    assert rel(got, ref.float()) < 1e-5
    # unsupported shapes are refused loudly (the deferring build_gwc_volume never produces them)
    with pytest.raises(RuntimeError):
        A.gwc_corr_stem(f1.cuda(), f2.cuda(), 64, 8, w.cuda(), scale.cuda(), shift.cuda())
