"""Golden vectors for the LIIF arbitrary-scale upsampler, produced by the UNMODIFIED
reference classes imported from /root/reference (build container only):

    python tests/golden/make_liif_golden.py      ->  tests/golden/liif_upsample.npz

Reference objects exercised: liif.py AffinityFeature / StructureFeature("with_v2ISU") /
liif_feat_multiscale_train / liif_out_multi_scale_Training, submodule.py context_upsample_multiscale_train,
and the arithmetic of continuous_IGEVStereo.upsample_disp (multi_training branch, no disparity_norm).
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
from oracle import liif_oracle as LO  # noqa: E402  (only for the portable weight generator)


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    ref_loader.load()                      # installs the stub modules and the .cuda() no-op
    import models.coreContinuous_IGEV.liif as RL
    from models.coreContinuous_IGEV.submodule import context_upsample_multiscale_train
    aff = {"win_w": 3, "win_h": 3, "dilation": [1, 2, 4, 8]}
    arrs = {}
    for n_in in (2, 3):
        c = cases.liif_case(n_in)
        chanels = [f.shape[1] for f in c["feats"]]
        mod = RL.liif_out_multi_scale_Training(encoder_dim=sum(chanels), mlphidden_list=[128, 64, 64], pos_dim=0,
                                               pos_enconding=False, pos_enconding_new=False, local_ensemble=False,
                                               decode_cell=False, unfold="with_v2ISU", affinity_settings=aff,
                                               quater_nearest=None, require_grad=True, number_input=n_in,
                                               chanels=chanels).eval()
        params = LO.make_liif_params(c["in_dim"], seed=40 + n_in)
        sd = mod.state_dict()
        assert set(sd.keys()) == set(params.keys()), set(sd.keys()) ^ set(params.keys())
        mod.load_state_dict(params, strict=True)
        t = "n%d_" % n_in
        # stage outputs
        sf = [m(f) for m, f in zip(mod.to_sf_l2, c["feats"])]
        arrs[t + "affinity0"] = sf[0][:, chanels[0]:]
        rel, q, _ = RL.liif_feat_multiscale_train(sf[-1], c["coords"].clone(), c["scale"], False, False)
        arrs[t + "rel_last"] = rel
        arrs[t + "qfeat_last_head"] = q[:, :, :4]
        logits = mod([f.clone() for f in c["feats"]], c["coords"].clone(), c["scale"])
        arrs[t + "logits"] = logits
        mask = torch.softmax(logits, 1)
        d = c["disp"] * 4.0 * c["scale"].view(-1, 1, 1, 1)          # continuous_IGEVstereo.py:206
        arrs[t + "up_disp"] = context_upsample_multiscale_train(d, mask, c["coords"].clone()).unsqueeze(1)
    out = {k: v.numpy() for k, v in arrs.items()}
    path = os.path.join(HERE, "liif_upsample.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
