"""Seeded synthetic inputs shared by the golden-vector generator and the parity tests.

Everything is drawn from numpy RandomState (portable across machines/torch builds), so
the fixtures under tests/golden/ only need to store the REFERENCE OUTPUTS.
"""
import numpy as np
import torch


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def disparity_fields(rng, B, H, W, maxd):
    """The disparity distributions SURVEY.md section 4/8d asks for."""
    xs = np.arange(W)[None, None, None, :]
    ys = np.arange(H)[None, None, :, None]
    return {
        "uniform": rng.uniform(0, maxd, size=(B, 1, H, W)),
        "smooth": 0.4 * maxd + 0.2 * maxd * np.sin(2 * np.pi * xs / W) * np.cos(2 * np.pi * ys / H)
                  + rng.normal(0, 0.25, size=(B, 1, H, W)),
        "integer": np.floor(rng.uniform(0, maxd, size=(B, 1, H, W))),
        "negative": -rng.uniform(0, maxd, size=(B, 1, H, W)),
        "far_oob": rng.uniform(3 * maxd + 3 * W, 5 * maxd + 5 * W, size=(B, 1, H, W)),
        "edge": rng.choice([-4.5, -0.25, 0.0, 0.5, maxd - 1.0, maxd - 0.5, maxd + 3.75, W - 1.0, W - 0.5],
                           size=(B, 1, H, W)),
    }


def raft_corr_case(seed=1, B=2, D=16, H=5, W=23, L=4):
    rng = np.random.RandomState(seed)
    f1 = _t(rng.standard_normal((B, D, H, W)))
    f2 = _t(rng.standard_normal((B, D, H, W)))
    disps = {k: _t(v) for k, v in disparity_fields(rng, B, H, W, 12).items()}
    return dict(f1=f1, f2=f2, L=L, r=4, disps=disps)


def igev_geo_case(seed=2, B=2, D=24, H=4, W=20, G=8, Dg=16, L=2):
    rng = np.random.RandomState(seed)
    f1 = _t(rng.standard_normal((B, D, H, W)))
    f2 = _t(rng.standard_normal((B, D, H, W)))
    geo = _t(rng.standard_normal((B, G, Dg, H, W)))
    disps = {k: _t(v) for k, v in disparity_fields(rng, B, H, W, Dg).items()}
    return dict(f1=f1, f2=f2, geo=geo, L=L, r=4, disps=disps)


def gwc_cases():
    out = {}
    for name, (seed, B, C, H, W, maxd, G) in {
        "small_wide_disp": (3, 2, 24, 3, 11, 16, 8),     # maxdisp > W: empty slices (submodule.py:266)
        "igev_shape": (4, 1, 96, 4, 40, 48, 8),
        "odd": (5, 2, 32, 2, 17, 7, 4),
    }.items():
        rng = np.random.RandomState(seed)
        out[name] = dict(left=_t(rng.standard_normal((B, C, H, W))),
                         right=_t(rng.standard_normal((B, C, H, W))), maxdisp=maxd, groups=G)
    return out


def update_block_case(family, seed=6, B=1, H=8, W=12, hidden=128):
    """net/inp at 1/4, 1/8, 1/16 (sizes follow the reference encoder's stride-2 convs:
    ceil halves), corr features and disparity."""
    rng = np.random.RandomState(seed + (0 if family == "igev" else 100))
    cor_planes = 162 if family == "igev" else 36
    sizes = [(H, W), ((H + 1) // 2, (W + 1) // 2), (((H + 1) // 2 + 1) // 2, ((W + 1) // 2 + 1) // 2)]
    net = [_t(np.tanh(rng.standard_normal((B, hidden, h, w)))) for h, w in sizes]
    inp = [[_t(0.5 * rng.standard_normal((B, hidden, h, w))) for _ in range(3)] for h, w in sizes]
    corr = _t(rng.standard_normal((B, cor_planes, H, W)))
    disp = _t(rng.uniform(0, 10, size=(B, 1, H, W)))
    return dict(net=net, inp=inp, corr=corr, disp=disp, cor_planes=cor_planes)


def loop_case(family, seed=7, B=1, H=16, W=24, hidden=128):
    rng = np.random.RandomState(seed + (0 if family == "igev" else 100))
    D = 96 if family == "igev" else 256
    f1 = _t(rng.standard_normal((B, D, H, W)) / np.sqrt(D) * 2.0)
    f2 = _t(rng.standard_normal((B, D, H, W)) / np.sqrt(D) * 2.0)
    sizes = [(H, W), ((H + 1) // 2, (W + 1) // 2), (((H + 1) // 2 + 1) // 2, ((W + 1) // 2 + 1) // 2)]
    net = [_t(np.tanh(rng.standard_normal((B, hidden, h, w)))) for h, w in sizes]
    inp = [[_t(0.5 * rng.standard_normal((B, hidden, h, w))) for _ in range(3)] for h, w in sizes]
    case = dict(f1=f1, f2=f2, net=net, inp=inp, cor_planes=162 if family == "igev" else 36)
    if family == "igev":
        case["geo"] = _t(rng.standard_normal((B, 8, 12, H, W)))
        case["init_disp"] = _t(rng.uniform(0, 11, size=(B, 1, H, W)))
    return case


def liif_case(n_in=2, seed=21, B=2, h=6, w=10, scale=2.5, extra_q=77):
    """Inputs of continuous_IGEVStereo.upsample_disp (multi-scale branch): feature maps at 1/4 (stem_4x ++ hidden,
    176 ch), 1/2 (32 ch) and, for agg_type type2, 1/1 (8 ch); query coordinates = the full output grid at `scale`
    plus random points (incl. exact +-1 borders); disparity at 1/4."""
    rng = np.random.RandomState(seed + n_in)
    x4 = _t(rng.standard_normal((B, 176, h, w)))
    x2 = _t(rng.standard_normal((B, 32, 2 * h, 2 * w)))
    x1 = _t(rng.standard_normal((B, 8, 4 * h, 4 * w)))
    x4[0, :, 1, 2] = 0.0                                     # an all-zero feature vector: normalisation eps path
    feats = [x4, x2] if n_in == 2 else [x1, x2, x4]           # continuous_IGEVstereo.py:214-217 ordering
    H, W = int(h * 4 * scale), int(w * 4 * scale)
    ry, rx = 1.0 / H, 1.0 / W
    ys = -1 + ry + 2 * ry * np.arange(H)
    xs = -1 + rx + 2 * rx * np.arange(W)
    grid = np.stack(np.meshgrid(ys, xs, indexing="ij"), -1).reshape(-1, 2)
    rnd = rng.uniform(-1, 1, size=(extra_q, 2))
    rnd[:4] = [[-1, -1], [1, 1], [-1, 1], [0.0, 0.0]]
    coords = np.concatenate([grid, rnd], 0).astype("float32")
    coords = _t(np.broadcast_to(coords, (B,) + coords.shape).copy())
    coords[1] = coords[1].flip(0)                             # different query order per batch element
    disp = _t(rng.uniform(0, 12, size=(B, 1, h, w)))
    sc = _t(np.full((B,), scale))
    in_dim = sum(f.shape[1] + 8 + 2 for f in feats)
    return dict(feats=feats, coords=coords, disp=disp, scale=sc, in_dim=in_dim, n_in=n_in)


def init_disp_case(seed=31, B=2, G=8, D=12, H=5, W=37):
    rng = np.random.RandomState(seed)
    return dict(geo=_t(rng.standard_normal((B, G, D, H, W)) * 2.0), weight=_t(rng.standard_normal((1, G, 3, 3, 3)) * 0.2))
