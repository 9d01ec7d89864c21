"""Model-level golden vector: run the REAL reference RAFT-family model graph (continuous_RaftStereo, unmodified,
CPU fp32, default-initialised backbone, portable seeded update-block weights) on a small synthetic stereo pair
and record (a) the tensors that cross into the hot path -- fmaps, net_list, inp_list -- and (b) the low-resolution
disparity the reference produces after `ITERS` iterations (prune_raft_stereo.py:267-288).  The GPU test replays (a)
through this library's operators and must land within 0.01 px of (b).

    python tests/golden/make_model_golden.py      # build container only (needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_loader  # noqa: E402
from oracle import hotpath_oracle as O  # noqa: E402

ITERS = 32
H, W = 64, 96          # image size -> 16 x 24 at 1/4


def main():
    ref_loader.load()
    from models.corePrune_RAFT import prune_raft_stereo as prs
    from models.corePrune_RAFT.liif import make_coord

    args = types.SimpleNamespace(
        hidden_dims=[128] * 3, n_gru_layers=3, n_downsample=2, corr_radius=4, corr_levels=4, slow_fast_gru=False,
        agg_type="type5", multi_training=True, multi_input_training=False, unfold_similarity="with_v2ISU",
        mlphidden_list=[128, 64, 64], pos_dim=0, pos_enconding=False, pos_enconding_new=False, local_ensemble=False,
        decode_cell=False, lsp_width=3, lsp_height=3, lsp_dilation=[1, 2, 4, 8], quater_nearest=None,
        require_grad=False, disparity_norm=False, disparity_norm2=False, mixed_precision=False, max_disp=192, Raw_Mask_dim=32, unfold=False,
        corr_implementation="reg", shared_backbone=False)
    torch.manual_seed(0)
    model = prs.continuous_RaftStereo(args).eval()
    params = O.make_update_block_params(36, seed=77)
    model.update_block.load_state_dict(params, strict=True)

    captured = {}
    orig_corr = prs.CorrBlock1D

    class SpyCorr(orig_corr):
        def __init__(self, f1, f2, **kw):
            captured["f1"], captured["f2"] = f1.detach().clone(), f2.detach().clone()
            super().__init__(f1, f2, **kw)

    prs.CorrBlock1D = SpyCorr
    orig_fwd = model.update_block.forward

    def spy_fwd(net, inp, *a, **k):
        if "net" not in captured:
            captured["net"] = [t.detach().clone() for t in net]
            captured["inp"] = [[t.detach().clone() for t in lst] for lst in inp]
        return orig_fwd(net, inp, *a, **k)

    model.update_block.forward = spy_fwd
    rng = np.random.RandomState(3)
    base = rng.uniform(0, 255, size=(1, 3, H, W + 16)).astype("float32")
    img1 = torch.from_numpy(base[..., 8:8 + W].copy())
    img2 = torch.from_numpy(base[..., 2:2 + W].copy()) + torch.from_numpy(rng.normal(0, 2, size=(1, 3, H, W)).astype("float32"))
    with torch.no_grad():
        disp_lowres, disp_up = model(img1, img2, iters=ITERS, test_mode=True, hr_coord=make_coord([H, W])[None],
                                     scale=torch.tensor([[1.0]]), output_raw=True)
    prs.CorrBlock1D = orig_corr
    out = {"f1": captured["f1"], "f2": captured["f2"], "disp_lowres": disp_lowres, "iters": ITERS}
    for i in range(3):
        out["net%d" % i] = captured["net"][i]
        for j in range(3):
            out["inp%d_%d" % (i, j)] = captured["inp"][i][j]
    path = os.path.join(HERE, "model_raft_boundary.npz")
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "disp range", float(disp_lowres.min()), float(disp_lowres.max()))


def timm_shim():
    """timm is absent: a shape-identical MobileNetV2 from torchvision regrouped into the attributes
    extractor.py:331-343 reads (conv_stem, bn1, act1, blocks[0..6]); SURVEY.md 8c.  Off the hot path, random weights."""
    import timm
    import torchvision

    def create_model(name, pretrained=True, features_only=True):
        f = torchvision.models.mobilenet_v2(weights=None).features
        m = types.SimpleNamespace()
        m.conv_stem, m.bn1, m.act1 = f[0][0], f[0][1], f[0][2]
        groups = [[1], [2, 3], [4, 5, 6], [7, 8, 9, 10], [11, 12, 13], [14, 15, 16], [17]]
        m.blocks = [torch.nn.Sequential(*[f[i] for i in g]) for g in groups]
        return m

    timm.create_model = create_model


def main_igev():
    ref_loader.load()
    timm_shim()
    from models.coreContinuous_IGEV import continuous_IGEVstereo as cis
    from models.coreContinuous_IGEV.liif import make_coord

    args = types.SimpleNamespace(
        hidden_dims=[128] * 3, n_gru_layers=3, n_downsample=2, corr_radius=4, corr_levels=2, slow_fast_gru=False,
        agg_type="type5", multi_training=True, multi_input_training=False, unfold_similarity="with_v2ISU",
        mlphidden_list=[128, 64, 64], pos_dim=0, pos_enconding=False, pos_enconding_new=False, local_ensemble=False,
        decode_cell=False, lsp_width=3, lsp_height=3, lsp_dilation=[1, 2, 4, 8], quater_nearest=None,
        require_grad=False, disparity_norm=False, disparity_norm2=False, mixed_precision=False, max_disp=192, Raw_Mask_dim=32, unfold=False,
        corr_implementation="reg", shared_backbone=False)
    torch.manual_seed(0)
    model = cis.continuous_IGEVStereo(args).eval()
    params = O.make_update_block_params(162, seed=78)
    model.update_block.load_state_dict(params, strict=True)
    captured = {}
    orig_geo = cis.Combined_Geo_Encoding_Volume

    class SpyGeo(orig_geo):
        def __init__(self, f1, f2, geo, **kw):
            captured["f1"], captured["f2"], captured["geo"] = f1.detach().clone(), f2.detach().clone(), geo.detach().clone()
            super().__init__(f1, f2, geo, **kw)

        def __call__(self, disp, coords):
            if "init_disp" not in captured:
                captured["init_disp"] = disp.detach().clone()
            return super().__call__(disp, coords)

    cis.Combined_Geo_Encoding_Volume = SpyGeo
    orig_gwc = cis.build_gwc_volume

    def spy_gwc(l, r, maxdisp, groups):
        out = orig_gwc(l, r, maxdisp, groups)
        captured["gwc"] = out.detach().clone()
        return out

    cis.build_gwc_volume = spy_gwc
    orig_fwd = model.update_block.forward
    state = {}

    def spy_fwd(net, inp, *a, **k):
        if "net" not in captured:
            captured["net"] = [t.detach().clone() for t in net]
            captured["inp"] = [[t.detach().clone() for t in lst] for lst in inp]
        out = orig_fwd(net, inp, *a, **k)
        if isinstance(out, tuple):
            state["delta"] = out[1].detach()
        return out

    model.update_block.forward = spy_fwd
    # the model only returns the upsampled map in test mode: track disp = disp + delta ourselves
    rng = np.random.RandomState(4)
    Hi, Wi = 64, 128
    base = rng.uniform(0, 255, size=(1, 3, Hi, Wi + 16)).astype("float32")
    img1 = torch.from_numpy(base[..., 8:8 + Wi].copy())
    img2 = torch.from_numpy(base[..., 2:2 + Wi].copy()) + torch.from_numpy(rng.normal(0, 2, size=(1, 3, Hi, Wi)).astype("float32"))
    disp_track = {"d": None}
    orig_call = SpyGeo.__call__

    def tracking_call(self, disp, coords):
        disp_track["d"] = disp.detach().clone()       # disparity entering iteration k == result of iteration k-1
        return orig_call(self, disp, coords)

    SpyGeo.__call__ = tracking_call
    with torch.no_grad():
        model(img1, img2, iters=ITERS, test_mode=True, hr_coord=make_coord([Hi, Wi])[None], scale=torch.tensor([[1.0]]))
    final = disp_track["d"] + state["delta"]           # continuous_IGEVstereo.py:295 for the last iteration
    # ---- the arbitrary-scale upsampler at the end of the same model graph, x2.5 query grid ----
    up_cap = {}
    orig_up = model.upsample_disp

    def spy_up(disp, hidden, stem_4x, stem_2x, stem_1x, hr_coord=None, scale=1):
        out_ = orig_up(disp, hidden, stem_4x, stem_2x, stem_1x, hr_coord=hr_coord.clone(), scale=scale)
        up_cap.update(disp=disp.detach().clone(), hidden=hidden.detach().clone(), stem4=stem_4x.detach().clone(),
                      stem2=stem_2x.detach().clone(), out=out_.detach().clone())
        return out_

    model.upsample_disp = spy_up
    Ho, Wo = int(Hi * 2.5), int(Wi * 2.5)
    with torch.no_grad():
        model(img1, img2, iters=4, test_mode=True, hr_coord=make_coord([Ho, Wo])[None], scale=torch.tensor([[2.5]]))
    up = {"disp": up_cap["disp"], "hidden": up_cap["hidden"], "stem4": up_cap["stem4"], "stem2": up_cap["stem2"],
          "up_disp": up_cap["out"], "out_hw": np.asarray([Ho, Wo]), "scale": np.asarray([2.5], dtype="float32")}
    for k, v in model.liif_up.state_dict().items():
        up["liif." + k] = v
    path = os.path.join(HERE, "model_igev_upsample.npz")
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in up.items()})
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "up_disp", tuple(up_cap["out"].shape))
    cis.Combined_Geo_Encoding_Volume = orig_geo
    cis.build_gwc_volume = orig_gwc
    out = {"classifier_weight": model.classifier.weight.detach().clone(),      # init-disparity head, SURVEY 8(f)-3
           "f1": captured["f1"], "f2": captured["f2"], "geo": captured["geo"], "gwc": captured["gwc"],
           "init_disp": captured["init_disp"], "disp_lowres": final, "iters": ITERS}
    for i in range(3):
        out["net%d" % i] = captured["net"][i]
        for j in range(3):
            out["inp%d_%d" % (i, j)] = captured["inp"][i][j]
    path = os.path.join(HERE, "model_igev_boundary.npz")
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "disp range", float(final.min()), float(final.max()))


if __name__ == "__main__":
    main()
    main_igev()
