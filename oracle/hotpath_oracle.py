"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Any-Stereo iterative cost-volume hot path.

This file is a from-scratch CPU restatement (torch-CPU fp32 + integer index
math) of what the reference computes on the path named by
``BASELINE.json:north_star``.  Nothing under ``any-stereo_b200/`` imports it; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  Every function cites the reference lines it
follows (paths relative to ``/root/reference``).

PARITY PIN: the reference ships no tests, golden vectors or fixtures for this
path, so this oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports the unmodified
reference modules (``oracle/ref_loader.py``) in the build container and commits
their outputs on seeded inputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against those fixtures.

Two lookup formulations are kept on purpose:
  * ``*_gridsample``: the reference's Python route (normalise to [-1,1], then
    ``F.grid_sample(align_corners=True)``) -- the value oracle;
  * ``*_exact``: the reference's CUDA route (``sampler/sampler_kernel.cu``) with
    integer tap indices -- the index oracle (bit-exact ints).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# a1  all-pairs row correlation
# --------------------------------------------------------------------------------------

def all_pairs_corr(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
    """corr[b,y,x1,0,x2] = sum_d f1[b,d,y,x1] * f2[b,d,y,x2]; no 1/sqrt(D) scaling.

    Reference: corePrune_RAFT/geometry.py:46-56 and coreContinuous_IGEV/geometry.py:63-72
    (einsum 'aijk,aijh->ajkh', then reshape to [B,H,W1,1,W2]).  The reference's
    ``mask_invalid`` branch (geometry.py:53-54) is a discarded comparison, i.e. a no-op.
    """
    B, D, H, W1 = fmap1.shape
    W2 = fmap2.shape[3]
    # per (b,y): [W1,D] @ [D,W2]
    a = fmap1.permute(0, 2, 3, 1).reshape(B * H, W1, D)
    b = fmap2.permute(0, 2, 1, 3).reshape(B * H, D, W2)
    return torch.bmm(a, b).reshape(B, H, W1, 1, W2).contiguous()


# --------------------------------------------------------------------------------------
# a2  pyramids
# --------------------------------------------------------------------------------------

def halve_last(x: torch.Tensor) -> torch.Tensor:
    """out[..., j] = (x[..., 2j] + x[..., 2j+1]) / 2, odd tail dropped.

    Reference: F.avg_pool2d(x, [1,2], stride=[1,2]) at corePrune_RAFT/geometry.py:18 and
    coreContinuous_IGEV/geometry.py:24,28.  (a+b)*0.5 is bit-identical to avg_pool2d on CPU.
    """
    n = x.shape[-1] // 2
    return (x[..., 0:2 * n:2] + x[..., 1:2 * n:2]) * 0.5


def corr_pyramid(corr: torch.Tensor, num_levels: int) -> List[torch.Tensor]:
    """Levels [N,1,1,W2/2^i] with N=B*H*W1 (corePrune_RAFT/geometry.py:13-19)."""
    B, H, W1, _, W2 = corr.shape
    lvl = corr.reshape(B * H * W1, 1, 1, W2)
    out = [lvl]
    for _ in range(num_levels - 1):
        lvl = halve_last(lvl)
        out.append(lvl)
    return out


def geo_pyramid(geo_volume: torch.Tensor, num_levels: int) -> List[torch.Tensor]:
    """[B,G,D,H,W] -> levels [N,G,1,D/2^i] (coreContinuous_IGEV/geometry.py:17-25)."""
    B, G, D, H, W = geo_volume.shape
    lvl = geo_volume.permute(0, 3, 4, 1, 2).reshape(B * H * W, G, 1, D)
    out = [lvl]
    for _ in range(num_levels - 1):
        lvl = halve_last(lvl)
        out.append(lvl)
    return out


# --------------------------------------------------------------------------------------
# a5/a6  1-D linear-interpolation radius lookup
# --------------------------------------------------------------------------------------

def tap_indices(x: torch.Tensor, radius: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Integer base tap and fractional weight of a sample position.

    Reference: sampler/sampler_kernel.cu:42 (dx = x0 - floor(x0)) and :47
    (x1 = int(floor(x0)) - r + i).  Returns (floor(x) - r as int32, x - floor(x) as fp32).
    """
    fl = torch.floor(x)
    return fl.to(torch.int32) - radius, (x - fl)


def lookup_rows_exact(rows: torch.Tensor, x: torch.Tensor, radius: int) -> torch.Tensor:
    """rows [N,C,W], x [N] -> [N,C,2r+1]:
    out[n,c,k] = (1-f)*V[n,c,t+k] + f*V[n,c,t+k+1], t = floor(x)-r, f = x-floor(x),
    taps outside [0,W) contribute 0.  Reference: sampler/sampler_kernel.cu:39-58.
    """
    N, C, W = rows.shape
    t0, f = tap_indices(x, radius)
    k = torch.arange(2 * radius + 2, dtype=torch.int64, device=rows.device)
    idx = t0.to(torch.int64)[:, None] + k[None, :]                 # [N,2r+2]
    ok = (idx >= 0) & (idx < W)
    g = torch.gather(rows, 2, idx.clamp(0, W - 1)[:, None, :].expand(N, C, -1))
    g = g * ok[:, None, :].to(rows.dtype)
    omf = (1.0 - f).to(rows.dtype)[:, None, None]      # scalar_t(1.0f - dx): fp32 subtraction, then cast (:56)
    f = f.to(rows.dtype)[:, None, None]
    return g[:, :, :-1] * omf + g[:, :, 1:] * f


def lookup_rows_gridsample(rows: torch.Tensor, x: torch.Tensor, radius: int) -> torch.Tensor:
    """Same quantity through the reference's Python route.

    Reference: models/*/utils/utils.py:59-72 (xgrid = 2x/(W-1)-1; grid_sample with
    align_corners=True, bilinear, zero padding) as driven by geometry.py:29-41.
    rows [N,C,W], x [N] -> [N,C,2r+1].
    """
    N, C, W = rows.shape
    taps = torch.linspace(-radius, radius, 2 * radius + 1, device=rows.device).view(1, 1, -1, 1)
    xs = taps + x.reshape(N, 1, 1, 1)
    ys = torch.zeros_like(xs)
    grid = torch.cat([2 * xs / (W - 1) - 1, ys], dim=-1)           # [N,1,2r+1,2]
    out = F.grid_sample(rows.reshape(N, C, 1, W), grid, align_corners=True)  # [N,C,1,2r+1]
    return out.reshape(N, C, 2 * radius + 1)


def sampler_forward(volume: torch.Tensor, coords: torch.Tensor, radius: int) -> torch.Tensor:
    """corr_sampler.forward: volume [B,H,W1,W2], coords [B,>=1,H,W1] -> [B,2r+1,H,W1].

    Reference: sampler/sampler_kernel.cu:19-60 (kernel), :107-136 (launcher),
    sampler/sampler.cpp:24-32 (boundary).
    """
    B, H, W1, W2 = volume.shape
    x = coords[:, 0].reshape(-1).to(torch.float32)
    out = lookup_rows_exact(volume.reshape(B * H * W1, 1, W2), x, radius)   # [N,1,2r+1]
    return out.reshape(B, H, W1, 2 * radius + 1).permute(0, 3, 1, 2).contiguous()


def sampler_backward(volume: torch.Tensor, coords: torch.Tensor, corr_grad: torch.Tensor,
                     radius: int) -> torch.Tensor:
    """corr_sampler.backward: adjoint w.r.t. the volume only.

    volume_grad[n,y,x,t+j] = f*g[j-1] + (1-f)*g[j]   (j in [0,2r+1], terms outside [0,2r] dropped)
    Reference: sampler/sampler_kernel.cu:63-105, :138-166; sampler.cpp:34-45.
    """
    B, H, W1, W2 = volume.shape
    N = B * H * W1
    x = coords[:, 0].reshape(-1).to(torch.float32)
    t0, f = tap_indices(x, radius)
    g = corr_grad.permute(0, 2, 3, 1).reshape(N, 2 * radius + 1)
    omf = (1.0 - f).to(g.dtype)[:, None]
    f = f.to(g.dtype)[:, None]
    z = torch.zeros(N, 1, dtype=g.dtype)
    contrib = torch.cat([z, g], 1) * f + torch.cat([g, z], 1) * omf        # [N,2r+2]
    idx = t0.to(torch.int64)[:, None] + torch.arange(2 * radius + 2)[None, :]
    ok = (idx >= 0) & (idx < W2)
    out = torch.zeros(N, W2, dtype=g.dtype)
    out.scatter_add_(1, idx.clamp(0, W2 - 1), contrib * ok.to(g.dtype))
    return out.reshape(B, H, W1, W2)


# --------------------------------------------------------------------------------------
# a3/a4  per-iteration pyramid lookups
# --------------------------------------------------------------------------------------

def _level_positions(disp: torch.Tensor, coords: torch.Tensor, level: int):
    """Sample positions at pyramid level i: geo x = disp/2^i; corr x = coords/2^i - disp/2^i.
    Reference: coreContinuous_IGEV/geometry.py:43,52; corePrune_RAFT/geometry.py:31,35."""
    d = disp.reshape(-1) / 2 ** level
    c = coords.reshape(-1) / 2 ** level
    return d, c - d


def corrblock1d_lookup(pyr: Sequence[torch.Tensor], disp: torch.Tensor, coords: torch.Tensor,
                       radius: int, exact: bool = False) -> torch.Tensor:
    """CorrBlock1D.__call__: [B,L*(2r+1),H,W], channel = level*(2r+1)+tap.
    Reference: corePrune_RAFT/geometry.py:24-43."""
    B, _, H, W = disp.shape
    fn = lookup_rows_exact if exact else lookup_rows_gridsample
    outs = []
    for i, lvl in enumerate(pyr):
        _, xc = _level_positions(disp, coords, i)
        rows = lvl.reshape(B * H * W, 1, lvl.shape[-1])
        outs.append(fn(rows, xc, radius).reshape(B, H, W, -1))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def geo_lookup(geo_pyr: Sequence[torch.Tensor], corr_pyr: Sequence[torch.Tensor],
               disp: torch.Tensor, coords: torch.Tensor, radius: int,
               exact: bool = False) -> torch.Tensor:
    """Combined_Geo_Encoding_Volume.__call__: [B, L*(2r+1)*(G+1), H, W].

    Channel map per level i (base = i*(G+1)*(2r+1)): geo g*(2r+1)+k, then corr G*(2r+1)+k.
    Reference: coreContinuous_IGEV/geometry.py:34-60.
    """
    B, _, H, W = disp.shape
    N = B * H * W
    fn = lookup_rows_exact if exact else lookup_rows_gridsample
    outs = []
    for i in range(len(geo_pyr)):
        xg, xc = _level_positions(disp, coords, i)
        g = geo_pyr[i]
        g_rows = g.reshape(N, g.shape[1], g.shape[-1])
        outs.append(fn(g_rows, xg, radius).reshape(B, H, W, -1))
        c = corr_pyr[i]
        c_rows = c.reshape(N, 1, c.shape[-1])
        outs.append(fn(c_rows, xc, radius).reshape(B, H, W, -1))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


# --------------------------------------------------------------------------------------
# a7  group-wise correlation volume
# --------------------------------------------------------------------------------------

def gwc_volume(left: torch.Tensor, right: torch.Tensor, maxdisp: int, num_groups: int) -> torch.Tensor:
    """vol[b,g,d,y,x] = mean_{c in group g} L[b,c,y,x] * R[b,c,y,x-d] for x >= d else 0.
    Reference: coreContinuous_IGEV/submodule.py:253-271."""
    B, C, H, W = left.shape
    assert C % num_groups == 0
    cpg = C // num_groups
    vol = left.new_zeros(B, num_groups, maxdisp, H, W)
    for d in range(min(maxdisp, W)):
        prod = left[:, :, :, d:] * right[:, :, :, :W - d]
        vol[:, :, d, :, d:] = prod.view(B, num_groups, cpg, H, W - d).mean(dim=2)
    return vol


# --------------------------------------------------------------------------------------
# a8-a11  per-iteration update block
# --------------------------------------------------------------------------------------

def pool2x(x):
    """F.avg_pool2d(x, 3, stride=2, padding=1) (update.py:94-95); divisor always 9."""
    return F.avg_pool2d(x, 3, stride=2, padding=1)


def interp_to(x, ref):
    """Bilinear, align_corners=True, to ref's HxW (update.py:100-102)."""
    return F.interpolate(x, ref.shape[2:], mode="bilinear", align_corners=True)


def _conv(x, p, name, pad):
    return F.conv2d(x, p[name + ".weight"], p[name + ".bias"], padding=pad)


def conv_gru(p, prefix, h, cz, cr, cq, *xs):
    """ConvGRU.forward (update.py:33-41)."""
    x = torch.cat(xs, dim=1)
    hx = torch.cat([h, x], dim=1)
    z = torch.sigmoid(_conv(hx, p, prefix + ".convz", 1) + cz)
    r = torch.sigmoid(_conv(hx, p, prefix + ".convr", 1) + cr)
    q = torch.tanh(_conv(torch.cat([r * h, x], dim=1), p, prefix + ".convq", 1) + cq)
    return (1 - z) * h + z * q


def motion_encoder(p, disp, corr):
    """BasicMotionEncoder.forward (update.py:84-92)."""
    c = F.relu(_conv(corr, p, "encoder.convc1", 0))
    c = F.relu(_conv(c, p, "encoder.convc2", 1))
    d = F.relu(_conv(disp, p, "encoder.convd1", 3))
    d = F.relu(_conv(d, p, "encoder.convd2", 1))
    out = F.relu(_conv(torch.cat([c, d], dim=1), p, "encoder.conv", 1))
    return torch.cat([out, disp], dim=1)


def disp_head(p, h):
    """DispHead.forward (update.py:23-24)."""
    return _conv(F.relu(_conv(h, p, "disp_head.conv1", 1)), p, "disp_head.conv2", 1)


def update_block(p, net, inp, corr=None, disp=None, iter04=True, iter08=True, iter16=True,
                 update=True, n_gru_layers=3):
    """BasicMultiUpdateBlock.forward (update.py:116-136).  ``p`` maps parameter names
    (relative to update_block) to tensors.  Returns a NEW net list (and delta_disp)."""
    net = list(net)
    if iter16:
        net[2] = conv_gru(p, "gru16", net[2], *inp[2], pool2x(net[1]))
    if iter08:
        if n_gru_layers > 2:
            net[1] = conv_gru(p, "gru08", net[1], *inp[1], pool2x(net[0]), interp_to(net[2], net[1]))
        else:
            net[1] = conv_gru(p, "gru08", net[1], *inp[1], pool2x(net[0]))
    if iter04:
        mf = motion_encoder(p, disp, corr)
        if n_gru_layers > 1:
            net[0] = conv_gru(p, "gru04", net[0], *inp[0], mf, interp_to(net[1], net[0]))
        else:
            net[0] = conv_gru(p, "gru04", net[0], *inp[0], mf)
    if not update:
        return net
    return net, disp_head(p, net[0])


# --------------------------------------------------------------------------------------
# a12  loop glue
# --------------------------------------------------------------------------------------

def pixel_coords(B, H, W, device=None):
    """coords[b,y,x,0] = x as float (continuous_IGEVstereo.py:280, prune_raft_stereo.py:272)."""
    return torch.arange(W, device=device).float().reshape(1, 1, W, 1).repeat(B, H, 1, 1)


def igev_iterations(p, fmap1, fmap2, geo_volume, net, inp, init_disp, iters,
                    radius=4, num_levels=2, exact=False, keep_all=False, slow_fast_gru=False):
    """continuous_IGEVstereo.py:275-295: build the combined volume, then ``iters`` x
    {lookup -> update block -> disp += delta}.  Returns (disp, net[, all disps])."""
    B, _, H, W = fmap1.shape
    cp = corr_pyramid(all_pairs_corr(fmap1.float(), fmap2.float()), num_levels)
    gp = geo_pyramid(geo_volume.float(), num_levels)
    coords = pixel_coords(B, H, W, fmap1.device)
    disp = init_disp
    hist = []
    for _ in range(iters):
        feat = geo_lookup(gp, cp, disp, coords, radius, exact=exact)
        if slow_fast_gru:   # continuous_IGEVstereo.py:288-291 (n_gru_layers == 3): extra low-resolution GRU passes
            net = update_block(p, net, inp, iter16=True, iter08=False, iter04=False, update=False)
            net = update_block(p, net, inp, iter16=True, iter08=True, iter04=False, update=False)
        net, delta = update_block(p, net, inp, feat, disp)
        disp = disp + delta
        if keep_all:
            hist.append(disp)
    return (disp, net, hist) if keep_all else (disp, net)


def raft_iterations(p, fmap1, fmap2, net, inp, iters, radius=4, num_levels=4, exact=False,
                    keep_all=False):
    """prune_raft_stereo.py:267-286 (disp starts at zero, :274)."""
    B, _, H, W = fmap1.shape
    cp = corr_pyramid(all_pairs_corr(fmap1.float(), fmap2.float()), num_levels)
    coords = pixel_coords(B, H, W, fmap1.device)
    disp = fmap1.new_zeros(B, 1, H, W)
    hist = []
    for _ in range(iters):
        feat = corrblock1d_lookup(cp, disp, coords, radius, exact=exact)
        net, delta = update_block(p, net, inp, feat, disp)
        disp = disp + delta
        if keep_all:
            hist.append(disp)
    return (disp, net, hist) if keep_all else (disp, net)


# --------------------------------------------------------------------------------------
# a13  adjoints of the memory-bound operators (reference = autograd over a1-a7)
# --------------------------------------------------------------------------------------

def halve_last_bwd(g_coarse: torch.Tensor, fine_width: int) -> torch.Tensor:
    """Adjoint of halve_last: g_fine[2j] = g_fine[2j+1] = g_coarse[j]/2, dropped odd tail = 0."""
    out = g_coarse.new_zeros(*g_coarse.shape[:-1], fine_width)
    n = g_coarse.shape[-1]
    out[..., 0:2 * n:2] = g_coarse * 0.5
    out[..., 1:2 * n:2] = g_coarse * 0.5
    return out


def lookup_rows_bwd(g: torch.Tensor, x: torch.Tensor, radius: int, W: int) -> torch.Tensor:
    """Adjoint of lookup_rows_exact w.r.t. rows: g [N,C,2r+1] -> [N,C,W]
    (sampler/sampler_kernel.cu:83-103 generalised to C channels)."""
    N, C, _ = g.shape
    t0, f = tap_indices(x, radius)
    omf = (1.0 - f).to(g.dtype)[:, None, None]
    f = f.to(g.dtype)[:, None, None]
    z = g.new_zeros(N, C, 1)
    contrib = torch.cat([z, g], 2) * f + torch.cat([g, z], 2) * omf
    idx = t0.to(torch.int64)[:, None] + torch.arange(2 * radius + 2)[None, :]
    ok = (idx >= 0) & (idx < W)
    out = g.new_zeros(N, C, W)
    out.scatter_add_(2, idx.clamp(0, W - 1)[:, None, :].expand(N, C, -1),
                     contrib * ok[:, None, :].to(g.dtype))
    return out


def all_pairs_corr_bwd(g_corr: torch.Tensor, fmap1: torch.Tensor, fmap2: torch.Tensor):
    """dF1[b,d,y,x1] = sum_x2 gC[b,y,x1,x2] F2[b,d,y,x2]; dF2[b,d,y,x2] = sum_x1 gC[..] F1[b,d,y,x1]."""
    B, D, H, W1 = fmap1.shape
    W2 = fmap2.shape[3]
    g = g_corr.reshape(B * H, W1, W2)
    f1 = fmap1.permute(0, 2, 3, 1).reshape(B * H, W1, D)
    f2 = fmap2.permute(0, 2, 3, 1).reshape(B * H, W2, D)
    d1 = torch.bmm(g, f2).reshape(B, H, W1, D).permute(0, 3, 1, 2).contiguous()
    d2 = torch.bmm(g.transpose(1, 2), f1).reshape(B, H, W2, D).permute(0, 3, 1, 2).contiguous()
    return d1, d2


def gwc_volume_bwd(g_vol: torch.Tensor, left: torch.Tensor, right: torch.Tensor, num_groups: int):
    """Adjoint of gwc_volume."""
    B, C, H, W = left.shape
    cpg = C // num_groups
    D = g_vol.shape[2]
    dL = torch.zeros_like(left)
    dR = torch.zeros_like(right)
    for d in range(min(D, W)):
        g = g_vol[:, :, d, :, d:].repeat_interleave(cpg, dim=1) / cpg      # [B,C,H,W-d]
        dL[:, :, :, d:] += g * right[:, :, :, :W - d]
        dR[:, :, :, :W - d] += g * left[:, :, :, d:]
    return dL, dR


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench (synthetic parameters, portable across machines)
# --------------------------------------------------------------------------------------

def update_block_param_shapes(cor_planes: int, hidden: int = 128, n_gru_layers: int = 3):
    """Parameter names/shapes of BasicMultiUpdateBlock (update.py:74-82,104-114,16-20) in
    registration order."""
    enc_out = 128
    shapes = [
        ("encoder.convc1", (64, cor_planes, 1, 1)),
        ("encoder.convc2", (64, 64, 3, 3)),
        ("encoder.convd1", (64, 1, 7, 7)),
        ("encoder.convd2", (64, 64, 3, 3)),
        ("encoder.conv", (127, 128, 3, 3)),
    ]
    g04_in = hidden + enc_out + hidden * (n_gru_layers > 1)
    g08_in = hidden + hidden * (n_gru_layers == 3) + hidden
    g16_in = hidden + hidden
    for name, cin in (("gru04", g04_in), ("gru08", g08_in), ("gru16", g16_in)):
        for gate in ("convz", "convr", "convq"):
            shapes.append((f"{name}.{gate}", (hidden, cin, 3, 3)))
    shapes.append(("disp_head.conv1", (256, hidden, 3, 3)))
    shapes.append(("disp_head.conv2", (1, 256, 3, 3)))
    return shapes


def make_update_block_params(cor_planes: int, seed: int = 0, hidden: int = 128,
                             gain: float = 1.0) -> dict:
    """Deterministic synthetic weights, portable across machines (numpy RandomState),
    with the fan-in scaling nn.Conv2d's default init has (U(-1/sqrt(fan_in), +))."""
    import numpy as np

    rng = np.random.RandomState(seed)
    p = {}
    for name, shp in update_block_param_shapes(cor_planes, hidden):
        fan_in = shp[1] * shp[2] * shp[3]
        bound = gain / math.sqrt(fan_in)
        p[name + ".weight"] = torch.from_numpy(rng.uniform(-bound, bound, size=shp).astype("float32"))
        p[name + ".bias"] = torch.from_numpy(rng.uniform(-bound, bound, size=(shp[0],)).astype("float32"))
    return p


# --------------------------------------------------------------------------------------
# SURVEY 8(f)-3 (first half)  initial-disparity head before the loop
# --------------------------------------------------------------------------------------
def disparity_regression(x: torch.Tensor, maxdisp: int) -> torch.Tensor:
    """submodule.py:321-325."""
    d = torch.arange(0, maxdisp, dtype=x.dtype, device=x.device).view(1, maxdisp, 1, 1)
    return torch.sum(x * d, 1, keepdim=True)


def init_disparity(geo_volume: torch.Tensor, classifier_weight: torch.Tensor):
    """continuous_IGEVstereo.py:267-268 with classifier = nn.Conv3d(G,1,3,1,1,bias=False) (:176), written as explicit
    shifted sums (no conv3d call).  -> (init_disp [B,1,H,W], prob [B,D,H,W])."""
    B, G, D, H, W = geo_volume.shape
    p = torch.zeros(B, G, D + 2, H + 2, W + 2, dtype=geo_volume.dtype, device=geo_volume.device)
    p[:, :, 1:D + 1, 1:H + 1, 1:W + 1] = geo_volume
    cost = torch.zeros(B, D, H, W, dtype=geo_volume.dtype, device=geo_volume.device)
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                wv = classifier_weight[0, :, kd, kh, kw].view(1, G, 1, 1, 1)
                cost = cost + (p[:, :, kd:kd + D, kh:kh + H, kw:kw + W] * wv).sum(1)
    prob = torch.softmax(cost, dim=1)
    return disparity_regression(prob, D), prob
