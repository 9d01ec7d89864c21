"""Generate golden vectors by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_loader.py) on the seeded inputs of cases.py.

Run in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs: tests/golden/*.npz (committed).  CPU, fp32, torch %s-independent inputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle import hotpath_oracle as O  # noqa: E402  (only for the portable weight generator)


def coords_of(B, H, W):
    # continuous_IGEVstereo.py:280 / prune_raft_stereo.py:272
    return torch.arange(W).float().reshape(1, 1, W, 1).repeat(B, H, 1, 1)


def save(name, **arrs):
    out = {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def load_params_into(module, params):
    sd = module.state_dict()
    assert set(sd.keys()) == set(params.keys()), (set(sd.keys()) ^ set(params.keys()))
    for k in sd:
        assert tuple(sd[k].shape) == tuple(params[k].shape), k
    module.load_state_dict(params, strict=True)


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    R = ref_loader.load()

    # ---- RAFT CorrBlock1D: corr, pyramid, lookup -------------------------------------
    c = cases.raft_corr_case()
    blk = R.CorrBlock1D(c["f1"], c["f2"], num_levels=c["L"], radius=c["r"])
    B, _, H, W = c["f1"].shape
    arrs = {"corr": R.CorrBlock1D.corr(c["f1"], c["f2"])}
    for i, lvl in enumerate(blk.init_corr_pyramid):
        arrs["pyr%d" % i] = lvl
    for k, d in c["disps"].items():
        arrs["lookup_" + k] = blk(d, coords_of(B, H, W))
    save("raft_corrblock", **arrs)

    # ---- IGEV Combined_Geo_Encoding_Volume -------------------------------------------
    c = cases.igev_geo_case()
    blk = R.Combined_Geo_Encoding_Volume(c["f1"], c["f2"], c["geo"], num_levels=c["L"], radius=c["r"])
    B, _, H, W = c["f1"].shape
    arrs = {}
    for i, lvl in enumerate(blk.init_corr_pyramid):
        arrs["corr_pyr%d" % i] = lvl
    for i, lvl in enumerate(blk.geo_volume_pyramid):
        arrs["geo_pyr%d" % i] = lvl
    for k, d in c["disps"].items():
        arrs["lookup_" + k] = blk(d, coords_of(B, H, W))
    save("igev_geovolume", **arrs)

    # ---- build_gwc_volume --------------------------------------------------------------
    arrs = {}
    for name, g in cases.gwc_cases().items():
        arrs[name] = R.build_gwc_volume(g["left"], g["right"], g["maxdisp"], g["groups"])
    save("gwc_volume", **arrs)

    # ---- initial-disparity head: the reference's classifier Conv3d + softmax + regression ----
    from models.coreContinuous_IGEV.submodule import disparity_regression
    c = cases.init_disp_case()
    classifier = torch.nn.Conv3d(8, 1, 3, 1, 1, bias=False)          # continuous_IGEVstereo.py:176
    classifier.weight.data.copy_(c["weight"])
    prob = torch.nn.functional.softmax(classifier(c["geo"]).squeeze(1), dim=1)     # :267
    save("init_disparity", prob=prob, init_disp=disparity_regression(prob, c["geo"].shape[2]))

    # ---- update block (both families, all flag combinations the models use) -------------
    for fam, cls in (("igev", R.IGEVUpdateBlock), ("raft", R.RAFTUpdateBlock)):
        c = cases.update_block_case(fam)
        args = ref_loader.update_block_args(fam)
        mod = cls(args, hidden_dims=[128, 128, 128]).eval()
        load_params_into(mod, O.make_update_block_params(c["cor_planes"], seed=11))
        arrs = {}
        net, delta = mod([t.clone() for t in c["net"]], c["inp"], c["corr"], c["disp"])
        for i in range(3):
            arrs["full_net%d" % i] = net[i]
        arrs["full_delta"] = delta
        # slow_fast_gru pre-passes (continuous_IGEVstereo.py:288-291)
        net = mod([t.clone() for t in c["net"]], c["inp"], iter16=True, iter08=False, iter04=False, update=False)
        for i in range(3):
            arrs["only16_net%d" % i] = net[i]
        net = mod([t.clone() for t in c["net"]], c["inp"], iter16=True, iter08=True, iter04=False, update=False)
        for i in range(3):
            arrs["lowres_net%d" % i] = net[i]
        save("update_block_" + fam, **arrs)

    # ---- whole iterative loop (reference classes driven exactly like the model forward) ---
    ITERS = 32
    c = cases.loop_case("igev")
    args = ref_loader.update_block_args("igev")
    mod = R.IGEVUpdateBlock(args, hidden_dims=[128, 128, 128]).eval()
    load_params_into(mod, O.make_update_block_params(162, seed=12))
    B, _, H, W = c["f1"].shape
    geo_fn = R.Combined_Geo_Encoding_Volume(c["f1"].float(), c["f2"].float(), c["geo"].float(), radius=4, num_levels=2)
    coords = coords_of(B, H, W)
    disp = c["init_disp"]
    net = [t.clone() for t in c["net"]]
    hist = []
    for _ in range(ITERS):
        feat = geo_fn(disp, coords)
        net, delta = mod(net, c["inp"], feat, disp, iter16=True, iter08=True)
        disp = disp + delta
        hist.append(disp)
    save("loop_igev", disps=torch.stack(hist), net0=net[0], net1=net[1], net2=net[2], iters=ITERS)

    c = cases.loop_case("raft")
    args = ref_loader.update_block_args("raft")
    mod = R.RAFTUpdateBlock(args, hidden_dims=[128, 128, 128]).eval()
    load_params_into(mod, O.make_update_block_params(36, seed=13))
    corr_fn = R.CorrBlock1D(c["f1"].float(), c["f2"].float(), radius=4, num_levels=4)
    disp = c["f1"].new_zeros((B, 1, H, W))
    net = [t.clone() for t in c["net"]]
    hist = []
    for _ in range(ITERS):
        feat = corr_fn(disp, coords)
        net, delta = mod(net, c["inp"], feat, disp, iter16=True, iter08=True)
        disp = disp + delta
        hist.append(disp)
    save("loop_raft", disps=torch.stack(hist), net0=net[0], net1=net[1], net2=net[2], iters=ITERS)

    # ---- adjoints through reference autograd (training path, config 5) --------------------
    torch.set_grad_enabled(True)
    c = cases.igev_geo_case(seed=8, B=1, D=24, H=3, W=14, Dg=16)
    f1 = c["f1"].clone().requires_grad_(True)
    f2 = c["f2"].clone().requires_grad_(True)
    geo = c["geo"].clone().requires_grad_(True)
    blk = R.Combined_Geo_Encoding_Volume(f1, f2, geo, num_levels=2, radius=4)
    d = c["disps"]["uniform"]
    out = blk(d, coords_of(1, 3, 14))
    rng = np.random.RandomState(99)
    w = torch.from_numpy(rng.standard_normal(tuple(out.shape)).astype("float32"))
    (out * w).sum().backward()
    arrs = dict(cot=w, g_f1=f1.grad, g_f2=f2.grad, g_geo=geo.grad)
    g = cases.gwc_cases()["odd"]
    L = g["left"].clone().requires_grad_(True)
    Rr = g["right"].clone().requires_grad_(True)
    vol = R.build_gwc_volume(L, Rr, g["maxdisp"], g["groups"])
    wv = torch.from_numpy(rng.standard_normal(tuple(vol.shape)).astype("float32"))
    (vol * wv).sum().backward()
    arrs.update(gwc_cot=wv, gwc_gL=L.grad, gwc_gR=Rr.grad)
    save("adjoints", **arrs)


def slowfast():
    """args.slow_fast_gru loop (continuous_IGEVstereo.py:284-295 with :288-291 active), reference classes, 6 iterations."""
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    R = ref_loader.load()
    c = cases.loop_case("igev", seed=61, B=1, H=16, W=24)
    args = ref_loader.update_block_args("igev")
    mod = R.IGEVUpdateBlock(args, hidden_dims=[128, 128, 128]).eval()
    load_params_into(mod, O.make_update_block_params(162, seed=8))
    B, _, H, W = c["f1"].shape
    geo_fn = R.Combined_Geo_Encoding_Volume(c["f1"].float(), c["f2"].float(), c["geo"].float(), radius=4, num_levels=2)
    coords = coords_of(B, H, W)
    disp = c["init_disp"]
    net = [t.clone() for t in c["net"]]
    hist = []
    for _ in range(6):
        feat = geo_fn(disp, coords)
        net = mod(net, c["inp"], iter16=True, iter08=False, iter04=False, update=False)
        net = mod(net, c["inp"], iter16=True, iter08=True, iter04=False, update=False)
        net, delta = mod(net, c["inp"], feat, disp, iter16=True, iter08=True)
        disp = disp + delta
        hist.append(disp)
    save("loop_igev_slowfast", disps=torch.stack(hist), iters=6)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "slowfast":
        slowfast()
        sys.exit(0)
    main()
    slowfast()
