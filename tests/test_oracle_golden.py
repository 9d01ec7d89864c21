"""CPU: pin the oracle restatement (oracle/hotpath_oracle.py) against outputs of the
reference itself (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import cases
from oracle import hotpath_oracle as O


def rel(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_corr_and_pyramid_raft(golden):
    g = golden("raft_corrblock")
    c = cases.raft_corr_case()
    corr = O.all_pairs_corr(c["f1"], c["f2"])
    assert corr.shape == g["corr"].shape
    assert rel(corr, g["corr"]) < 2e-6
    # pooling chain is bit-exact given the same level-0 input
    pyr = O.corr_pyramid(torch.from_numpy(g["corr"]), c["L"])
    for i, lvl in enumerate(pyr):
        assert lvl.shape == g["pyr%d" % i].shape
        assert np.array_equal(lvl.numpy(), g["pyr%d" % i]), "level %d not bit-exact" % i
    assert [p.shape[-1] for p in pyr] == [23, 11, 5, 2]


@pytest.mark.parametrize("dist", ["uniform", "smooth", "integer", "negative", "far_oob", "edge"])
def test_lookup_raft(golden, dist):
    g = golden("raft_corrblock")
    c = cases.raft_corr_case()
    pyr = [torch.from_numpy(g["pyr%d" % i]) for i in range(c["L"])]
    B, _, H, W = c["f1"].shape
    coords = O.pixel_coords(B, H, W)
    ref = g["lookup_" + dist]
    out = O.corrblock1d_lookup(pyr, c["disps"][dist], coords, c["r"], exact=False)
    assert out.shape == ref.shape == (B, 36, H, W)
    assert rel(out, ref) < 1e-6 or np.abs(ref).max() == 0
    out_x = O.corrblock1d_lookup(pyr, c["disps"][dist], coords, c["r"], exact=True)
    if np.abs(ref).max() == 0:
        assert float(out_x.abs().max()) == 0.0
    else:
        # exact-index route vs the grid_sample route: oracle noise ~1e-5
        assert rel(out_x, ref) < 1e-4


@pytest.mark.parametrize("dist", ["uniform", "smooth", "integer", "negative", "far_oob", "edge"])
def test_lookup_igev(golden, dist):
    g = golden("igev_geovolume")
    c = cases.igev_geo_case()
    B, _, H, W = c["f1"].shape
    corr = O.all_pairs_corr(c["f1"], c["f2"])
    cp = O.corr_pyramid(corr, c["L"])
    gp = O.geo_pyramid(c["geo"], c["L"])
    for i in range(c["L"]):
        assert rel(cp[i], g["corr_pyr%d" % i]) < 2e-6
        assert np.array_equal(gp[i].numpy(), g["geo_pyr%d" % i])
    coords = O.pixel_coords(B, H, W)
    ref = g["lookup_" + dist]
    cp_ref = [torch.from_numpy(g["corr_pyr%d" % i]) for i in range(c["L"])]
    out = O.geo_lookup(gp, cp_ref, c["disps"][dist], coords, c["r"], exact=False)
    assert out.shape == ref.shape == (B, 162, H, W)
    assert rel(out, ref) < 1e-6
    out_x = O.geo_lookup(gp, cp_ref, c["disps"][dist], coords, c["r"], exact=True)
    assert rel(out_x, ref) < 1e-4


def test_sampler_matches_python_lookup(golden):
    """corr_sampler.forward (exact-index CUDA semantics) equals level-0 of the Python lookup."""
    g = golden("raft_corrblock")
    c = cases.raft_corr_case()
    B, _, H, W = c["f1"].shape
    vol = torch.from_numpy(g["corr"]).reshape(B, H, W, W)
    for dist in ("uniform", "edge", "negative"):
        x = (O.pixel_coords(B, H, W).reshape(B, 1, H, W) - c["disps"][dist])
        out = O.sampler_forward(vol, x, 4)
        ref = g["lookup_" + dist][:, :9]
        assert rel(out, ref) < 1e-4


def test_sampler_backward_is_adjoint():
    rng = np.random.RandomState(0)
    vol = torch.from_numpy(rng.standard_normal((2, 3, 7, 9))).float()
    coords = torch.from_numpy(rng.uniform(-3, 12, size=(2, 1, 3, 7))).float()
    gout = torch.from_numpy(rng.standard_normal((2, 9, 3, 7))).float()
    lhs = (O.sampler_forward(vol, coords, 4).double() * gout.double()).sum()
    rhs = (O.sampler_backward(vol, coords, gout, 4).double() * vol.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-4 * abs(float(lhs))


def test_gwc(golden):
    g = golden("gwc_volume")
    for name, c in cases.gwc_cases().items():
        out = O.gwc_volume(c["left"], c["right"], c["maxdisp"], c["groups"])
        assert out.shape == g[name].shape
        assert np.array_equal(out.numpy(), g[name]), name


@pytest.mark.parametrize("family", ["igev", "raft"])
def test_update_block(golden, family):
    g = golden("update_block_" + family)
    c = cases.update_block_case(family)
    p = O.make_update_block_params(c["cor_planes"], seed=11)
    net, delta = O.update_block(p, c["net"], c["inp"], c["corr"], c["disp"])
    for i in range(3):
        assert rel(net[i], g["full_net%d" % i]) < 1e-6
    assert rel(delta, g["full_delta"]) < 1e-6
    net = O.update_block(p, c["net"], c["inp"], iter16=True, iter08=False, iter04=False, update=False)
    for i in range(3):
        assert rel(net[i], g["only16_net%d" % i]) < 1e-6
    net = O.update_block(p, c["net"], c["inp"], iter16=True, iter08=True, iter04=False, update=False)
    for i in range(3):
        assert rel(net[i], g["lowres_net%d" % i]) < 1e-6


@pytest.mark.parametrize("family", ["igev", "raft"])
def test_iteration_loop(golden, family):
    g = golden("loop_" + family)
    c = cases.loop_case(family)
    iters = int(g["iters"])
    if family == "igev":
        p = O.make_update_block_params(162, seed=12)
        disp, net, hist = O.igev_iterations(p, c["f1"], c["f2"], c["geo"], c["net"], c["inp"],
                                            c["init_disp"], iters, keep_all=True)
    else:
        p = O.make_update_block_params(36, seed=13)
        disp, net, hist = O.raft_iterations(p, c["f1"], c["f2"], c["net"], c["inp"], iters, keep_all=True)
    ref = torch.from_numpy(g["disps"])
    err = (torch.stack(hist) - ref).abs()
    # low-res disparity; x4 = full-res pixels.  Gate from BASELINE.json: 0.01 px EPE.
    assert float(err[-1].mean()) * 4 < 1e-3
    for i in range(3):
        assert rel(net[i], g["net%d" % i]) < 1e-4


def test_adjoints(golden):
    g = golden("adjoints")
    c = cases.igev_geo_case(seed=8, B=1, D=24, H=3, W=14, Dg=16)
    B, H, W, r, L, G = 1, 3, 14, 4, 2, 8
    N = B * H * W
    cot = torch.from_numpy(g["cot"]).permute(0, 2, 3, 1).reshape(N, -1)   # [N,162]
    d = c["disps"]["uniform"]
    coords = O.pixel_coords(B, H, W)
    Dg = c["geo"].shape[2]
    g_geo_lvls, g_corr_lvls = [], []
    for i in range(L):
        xg, xc = O._level_positions(d, coords, i)
        base = i * (G + 1) * 9
        gg = cot[:, base:base + G * 9].reshape(N, G, 9)
        gc = cot[:, base + G * 9:base + (G + 1) * 9].reshape(N, 1, 9)
        g_geo_lvls.append(O.lookup_rows_bwd(gg, xg, r, Dg >> i))
        g_corr_lvls.append(O.lookup_rows_bwd(gc, xc, r, W >> i))
    g_geo0 = g_geo_lvls[0] + O.halve_last_bwd(g_geo_lvls[1], Dg)
    g_corr0 = g_corr_lvls[0] + O.halve_last_bwd(g_corr_lvls[1], W)
    g_geo = g_geo0.reshape(B, H, W, G, Dg).permute(0, 3, 4, 1, 2)
    assert rel(g_geo, g["g_geo"]) < 1e-4
    d1, d2 = O.all_pairs_corr_bwd(g_corr0.reshape(B, H, W, W), c["f1"], c["f2"])
    assert rel(d1, g["g_f1"]) < 1e-4
    assert rel(d2, g["g_f2"]) < 1e-4
    gc = cases.gwc_cases()["odd"]
    dL, dR = O.gwc_volume_bwd(torch.from_numpy(g["gwc_cot"]), gc["left"], gc["right"], gc["groups"])
    assert rel(dL, g["gwc_gL"]) < 1e-5
    assert rel(dR, g["gwc_gR"]) < 1e-5


def _model_case(golden):
    g = golden("model_raft_boundary")
    net = [torch.from_numpy(g["net%d" % i]) for i in range(3)]
    inp = [[torch.from_numpy(g["inp%d_%d" % (i, j)]) for j in range(3)] for i in range(3)]
    return g, torch.from_numpy(g["f1"]), torch.from_numpy(g["f2"]), net, inp


def test_model_level_raft(golden):
    """Oracle loop fed with the tensors the REAL reference model graph passes into the hot path reproduces the
    low-res disparity that model computed (tests/golden/make_model_golden.py)."""
    g, f1, f2, net, inp = _model_case(golden)
    p = O.make_update_block_params(36, seed=77)
    disp, _ = O.raft_iterations(p, f1, f2, net, inp, int(g["iters"]))
    err = (disp - torch.from_numpy(g["disp_lowres"])).abs()
    assert float(err.mean()) * 4 < 1e-3


def test_model_level_igev(golden):
    """Same at the IGEV family: tensors captured inside the real continuous_IGEVStereo graph (timm backbone replaced
    by a shape-identical torchvision MobileNetV2, off the hot path)."""
    g = golden("model_igev_boundary")
    net = [torch.from_numpy(g["net%d" % i]) for i in range(3)]
    inp = [[torch.from_numpy(g["inp%d_%d" % (i, j)]) for j in range(3)] for i in range(3)]
    f1, f2 = torch.from_numpy(g["f1"]), torch.from_numpy(g["f2"])
    assert np.array_equal(O.gwc_volume(f1, f2, 48, 8).numpy(), g["gwc"])
    p = O.make_update_block_params(162, seed=78)
    disp, _ = O.igev_iterations(p, f1, f2, torch.from_numpy(g["geo"]), net, inp, torch.from_numpy(g["init_disp"]),
                                int(g["iters"]))
    err = (disp - torch.from_numpy(g["disp_lowres"])).abs()
    assert float(err.mean()) * 4 < 1e-3


# ---- LIIF arbitrary-scale upsampler ---------------------------------------------
@pytest.mark.parametrize("n_in", [2, 3])
def test_liif_oracle_vs_reference(golden, n_in):
    from oracle import liif_oracle as LO
    g = golden("liif_upsample")
    c = cases.liif_case(n_in)
    t = "n%d_" % n_in
    params = LO.make_liif_params(c["in_dim"], seed=40 + n_in)
    aff = LO.isu_affinity(c["feats"][0])
    assert aff.shape == g[t + "affinity0"].shape
    assert rel(aff, g[t + "affinity0"]) < 2e-6
    assert float(aff.min()) >= 0.0
    q, r = LO.liif_query(LO.structure_feature(c["feats"][-1]), c["coords"])
    # nearest-pixel index math is bit-exact: the gathered features are copies, rel_coord is the same fp32 arithmetic
    assert np.array_equal(q[:, :, :4].numpy(), g[t + "qfeat_last_head"])
    assert np.array_equal(r.numpy(), g[t + "rel_last"])
    logits = LO.liif_logits(params, c["feats"], c["coords"])
    assert logits.shape == g[t + "logits"].shape
    assert rel(logits, g[t + "logits"]) < 1e-5
    up = LO.upsample_disp_multiscale(params, c["disp"], c["feats"], c["coords"], c["scale"])
    assert up.shape == g[t + "up_disp"].shape
    assert rel(up, g[t + "up_disp"]) < 1e-5


def test_liif_oracle_vs_model_graph(golden):
    """The upsampler restatement against the output of the reference's own model graph
    (continuous_IGEVStereo.forward -> upsample_disp at a x2.5 query grid, tests/golden/make_model_golden.py)."""
    from oracle import liif_oracle as LO
    g = golden("model_igev_upsample")
    params = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("liif.")}
    Ho, Wo = [int(v) for v in g["out_hw"]]
    coords = torch.stack(torch.meshgrid(LO.make_coord_axis(Ho), LO.make_coord_axis(Wo), indexing="ij"), -1).reshape(1, -1, 2)
    x = torch.cat([torch.from_numpy(g["stem4"]), torch.from_numpy(g["hidden"])], 1)
    up = LO.upsample_disp_multiscale(params, torch.from_numpy(g["disp"]), [x, torch.from_numpy(g["stem2"])], coords,
                                     torch.from_numpy(g["scale"]))
    assert up.shape == g["up_disp"].shape
    assert rel(up, g["up_disp"]) < 1e-5


def test_init_disparity_oracle_vs_reference(golden):
    """classifier Conv3d + softmax + disparity_regression (continuous_IGEVstereo.py:267-268)."""
    g = golden("init_disparity")
    c = cases.init_disp_case()
    disp, prob = O.init_disparity(c["geo"], c["weight"])
    assert rel(prob, g["prob"]) < 1e-5
    assert rel(disp, g["init_disp"]) < 1e-5
    assert np.array_equal(O.disparity_regression(torch.from_numpy(g["prob"]), 12).numpy(), g["init_disp"])


def test_init_disparity_oracle_vs_model_graph(golden):
    """The init-disparity restatement against the tensor the real continuous_IGEVStereo.forward computes
    (classifier -> softmax -> disparity_regression, continuous_IGEVstereo.py:267-268) from its own geometry volume."""
    g = golden("model_igev_boundary")
    disp, _ = O.init_disparity(torch.from_numpy(g["geo"]), torch.from_numpy(g["classifier_weight"]))
    assert disp.shape == g["init_disp"].shape
    assert rel(disp, g["init_disp"]) < 1e-5


def test_slow_fast_gru_loop_oracle_vs_reference(golden):
    g = golden("loop_igev_slowfast")
    c = cases.loop_case("igev", seed=61, B=1, H=16, W=24)
    params = O.make_update_block_params(162, seed=8)
    disp, _, hist = O.igev_iterations(params, c["f1"], c["f2"], c["geo"], c["net"], c["inp"], c["init_disp"], int(g["iters"]),
                                      slow_fast_gru=True, keep_all=True)
    assert rel(torch.stack(hist), g["disps"]) < 2e-5
