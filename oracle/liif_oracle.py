"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the arbitrary-scale (LIIF) disparity upsampler that follows
the iterative loop.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this file; the product (any-stereo_b200/) never does.

Restated configuration = the reference's training default for the continuous models
(train_continuous_IGEV.py:323-361): ``liif_out_multi_scale_Training`` with ``unfold_similarity="with_v2ISU"``
(or "with_ISU"), ``pos_dim=0`` (raw relative coordinates), no cell decoding, no local ensemble, no quarter
sampling; 2 inputs (agg_type type1/3/4/5) or 3 inputs (type2); followed by softmax and
``context_upsample_multiscale_train``.

Pinned against the unmodified reference by tests/golden/make_liif_golden.py -> tests/golden/liif_upsample.npz
(tests/test_oracle_golden.py::test_liif_*).  Index arithmetic is written out by hand (no grid_sample / unfold),
all in fp32 like the reference.

Reference line numbers: models/coreContinuous_IGEV/liif.py unless stated otherwise.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch


def make_coord_axis(n: int, device=None) -> torch.Tensor:
    """Pixel-centre coordinates of an axis with n samples in [-1, 1]  (make_coord :32-45: v0 + r + 2r*i, r = 1/n,
    evaluated exactly like the reference: python-double scalars applied to an fp32 arange)."""
    r = (1 - (-1)) / (2 * n)
    return -1 + r + (2 * r) * torch.arange(n, device=device).float()


def nearest_index(c: torch.Tensor, n: int) -> torch.Tensor:
    """Index picked by F.grid_sample(mode='nearest', align_corners=False) for normalised coordinate c after the
    reference's clamp to +-(1-1e-6) (:118): round-half-even of ((c+1)*n-1)/2."""
    c = c.float().clamp(-1 + 1e-6, 1 - 1e-6)
    x = ((c + 1) * n - 1) / 2
    return torch.round(x).long().clamp(0, n - 1)


def isu_affinity(x: torch.Tensor) -> torch.Tensor:
    """AffinityFeature.forward (:434-449), 3x3 window, dilation 1: cosine similarity of every pixel with its 8
    neighbours (zero outside the image), negatives clipped to 0.  [B,C,H,W] -> [B,8,H,W], neighbour order = unfold
    order (row-major 3x3) with the centre removed."""
    B, C, H, W = x.shape
    n = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)           # F.normalize(p=2, dim=1)
    p = torch.zeros(B, C, H + 2, W + 2, dtype=x.dtype, device=x.device)
    p[:, :, 1:H + 1, 1:W + 1] = n
    out = []
    for dy in range(3):
        for dx in range(3):
            if dy == 1 and dx == 1:
                continue
            out.append((p[:, :, dy:dy + H, dx:dx + W] * n).sum(1))
    a = torch.stack(out, dim=1)
    return torch.where(a < 0, torch.zeros_like(a), a)


def structure_feature(x: torch.Tensor) -> torch.Tensor:
    """StructureFeature.forward, "with_ISU" / "with_v2ISU" branches (:518-526): cat([x, affinity])."""
    return torch.cat([x, isu_affinity(x)], dim=1)


def liif_query(feat: torch.Tensor, coords: torch.Tensor):
    """liif_feat_multiscale_train (:108-137), local=False, cell=False.
    feat [B,C,h,w], coords [B,Q,2] as (y, x) in [-1,1]  ->  q_feat [B,Q,C], rel_coord [B,Q,2]."""
    B, C, h, w = feat.shape
    iy = nearest_index(coords[:, :, 0], h)
    ix = nearest_index(coords[:, :, 1], w)
    bidx = torch.arange(B, device=feat.device).view(B, 1).expand_as(iy)
    q_feat = feat[bidx, :, iy, ix]                                 # [B,Q,C]
    qy = make_coord_axis(h, feat.device)[iy]
    qx = make_coord_axis(w, feat.device)[ix]
    rel = torch.stack([(coords[:, :, 0].float() - qy) * h, (coords[:, :, 1].float() - qx) * w], dim=-1)
    return q_feat, rel


def mlp(params: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "imnet.layers.") -> torch.Tensor:
    """MLP.forward (:9-25): Linear/ReLU stack, last layer linear."""
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in params if k.startswith(prefix)})
    for n, i in enumerate(idx):
        x = x @ params["%s%d.weight" % (prefix, i)].t() + params["%s%d.bias" % (prefix, i)]
        if n + 1 < len(idx):
            x = torch.relu(x)
    return x


def liif_logits(params: Dict[str, torch.Tensor], feats: Sequence[torch.Tensor], coords: torch.Tensor) -> torch.Tensor:
    """liif_out_multi_scale_Training.forward (:652-678) -> [B, 9, Q] (pre-softmax)."""
    latent: List[torch.Tensor] = []
    for f in feats:
        q, rel = liif_query(structure_feature(f.float()), coords)
        latent.append(torch.cat([q, rel], dim=-1))
    z = torch.cat(latent, dim=-1)
    B, Q, _ = z.shape
    return mlp(params, z.reshape(B * Q, -1)).reshape(B, Q, -1).permute(0, 2, 1).contiguous()


def context_upsample_multiscale(disp_low: torch.Tensor, up_weights: torch.Tensor, hr_coord: torch.Tensor) -> torch.Tensor:
    """context_upsample_multiscale_train (submodule.py:357-372): 3x3 zero-padded neighbourhood of the nearest low-res
    pixel, weighted by up_weights [B,9,Q].  -> [B,Q]."""
    B, _, h, w = disp_low.shape
    iy = nearest_index(hr_coord[:, :, 0], h)
    ix = nearest_index(hr_coord[:, :, 1], w)
    p = torch.zeros(B, h + 2, w + 2, dtype=disp_low.dtype, device=disp_low.device)
    p[:, 1:h + 1, 1:w + 1] = disp_low[:, 0]
    bidx = torch.arange(B, device=disp_low.device).view(B, 1).expand_as(iy)
    out = torch.zeros(B, hr_coord.shape[1], dtype=disp_low.dtype, device=disp_low.device)
    k = 0
    for dy in range(3):
        for dx in range(3):
            out = out + p[bidx, iy + dy, ix + dx] * up_weights[:, k]
            k += 1
    return out


def upsample_disp_multiscale(params, disp, feats, hr_coord, scale):
    """continuous_IGEVStereo.upsample_disp, multi_training branch without disparity_norm
    (continuous_IGEVstereo.py:192-237): disp*4*scale, logits -> softmax -> context upsample.  -> [B,1,Q]."""
    d = disp.float() * 4.0 * scale.view(-1, 1, 1, 1).float()
    mask = torch.softmax(liif_logits(params, feats, hr_coord), dim=1)
    return context_upsample_multiscale(d, mask, hr_coord).unsqueeze(1)


def make_liif_params(in_dim: int, hidden=(128, 64, 64), out_dim: int = 9, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Portable seeded MLP weights with the reference's state_dict names (imnet.layers.{0,2,4,6}.{weight,bias})."""
    import numpy as np
    rng = np.random.RandomState(seed)
    p = {}
    last = in_dim
    dims = list(hidden) + [out_dim]
    for n, d in enumerate(dims):
        p["imnet.layers.%d.weight" % (2 * n)] = torch.from_numpy((rng.standard_normal((d, last)) / np.sqrt(last)).astype("float32"))
        p["imnet.layers.%d.bias" % (2 * n)] = torch.from_numpy((0.1 * rng.standard_normal(d)).astype("float32"))
        last = d
    return p
