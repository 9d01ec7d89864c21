#!/usr/bin/env python
"""The drop-in, executed: build the REAL reference model graphs (continuous_IGEVStereo / continuous_RaftStereo, the
unmodified code under baseline/_ref -- or /root/reference in the build container), run them as shipped, then rebind the
names SURVEY 8(b) lists to this library (install_into_reference + adopt_update_block + adopt_liif_up) and run the SAME
forward again: same weights, same images, same iteration count.

    python tools/dropin.py [--family igev|raft|both] [--size 384x1248] [--iters 32] [--json out.json]

Reports, per family / shape / engine: mean and max |final full-resolution disparity - reference| in pixels (the
BASELINE.json EPE gate is 0.01 px) and the wall time of one forward (CUDA events) for the reference torch path on the
same GPU (strict fp32 = the parity oracle, and PyTorch's default TF32 convolutions) and for the drop-in call pattern
(``geo_fn(disp, coords)`` materialising the lookup tensor, then ``update_block(...)`` on NCHW tensors).

Test / measurement infrastructure: imports oracle/ref_loader.py (never imported by the product package).
"""
import argparse
import contextlib
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402


def reference_available() -> bool:
    return ref_loader.available()


def make_pair(B, H, W, device, seed=3):
    """Synthetic stereo pair with structure: band-limited texture + fine noise, right view = left view warped by a
    smooth disparity field (20 + 10 sin cos px, SURVEY 8(d) 'smooth'), values in [0, 255)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    low = torch.rand(B, 3, H // 8 + 2, W // 8 + 2, generator=g)
    tex = F.interpolate(low, size=(H, W), mode="bicubic", align_corners=True).clamp(0, 1)
    img1 = (0.75 * tex + 0.25 * torch.rand(B, 3, H, W, generator=g)) * 254.0
    ys = torch.arange(H).float().view(1, H, 1)
    xs = torch.arange(W).float().view(1, 1, W)
    disp = 20.0 + 10.0 * torch.sin(2 * math.pi * xs / W) * torch.cos(2 * math.pi * ys / H)
    # right(x) = left(x + d): sample the left image at x + d
    gx = (xs + disp) / (W - 1) * 2 - 1
    gy = (ys / (H - 1) * 2 - 1).expand(1, H, W)
    grid = torch.stack([gx.expand(1, H, W), gy], -1).expand(B, H, W, 2)
    img2 = F.grid_sample(img1, grid, mode="bilinear", padding_mode="border", align_corners=True)
    img2 = (img2 + torch.randn(B, 3, H, W, generator=g)).clamp(0, 254.9)
    return img1.to(device), img2.to(device)


def build_model(family, device, seed=0, **over):
    R = ref_loader.load_models()
    torch.manual_seed(seed)
    args = ref_loader.model_args(family, **over)
    cls = R.IGEV if family == "igev" else R.RAFT
    with contextlib.redirect_stdout(open(os.devnull, "w")):       # the RAFT constructor prints a banner
        model = cls(args)
    return model.to(device).eval(), R


@contextlib.contextmanager
def installed(model, R, family, fuse_corr_stem=False, defer_lookup=False, replay=None, fold_cnet=False, fused_fnet=False, fold_bn=False):
    """Rebind the reference's names to this library for the duration of the block.
    fuse_corr_stem (IGEV): also adopt corr_stem / corr_feature_att (build_gwc_volume fused with them).
    replay: adopt_update_block(..., replay=...) -- every update-block call replayed from a CUDA graph.
    fold_cnet: model.cnet = adopt_context_encoder(model.cnet) (eval BatchNorm folded, channels-last).
    fused_fnet (RAFT): model.fnet = adopt_feature_encoder(model.fnet) (InstanceNorm + ReLU / residual kernels).
    fold_bn: fold_basic_convs(model) -- every reference BasicConv (hourglass, FeatureAtt, Conv2x) with its BatchNorm folded."""
    import anystereo_b200 as A
    mod = R.igev_module if family == "igev" else R.raft_module
    names = ["Combined_Geo_Encoding_Volume", "build_gwc_volume", "context_upsample_multiscale_train", "CorrBlock1D"]
    saved = {n: getattr(mod, n) for n in names if hasattr(mod, n)}
    ub, lu, cnet = model.update_block, model.liif_up, model.cnet
    fnet = getattr(model, "fnet", None)
    folded = []
    stem = (model.corr_stem, model.corr_feature_att) if family == "igev" else None
    try:
        if family == "igev":
            A.install_into_reference(ref_igev_module=mod, defer_lookup=defer_lookup)
            if fuse_corr_stem:
                A.adopt_corr_stem(model, mod)
        else:
            A.install_into_reference(ref_raft_module=mod, defer_lookup=defer_lookup)
        model.update_block = A.adopt_update_block(ub, family, replay=replay)
        if fold_cnet:
            model.cnet = A.adopt_context_encoder(cnet)
        if fused_fnet and family == "raft":
            model.fnet = A.adopt_feature_encoder(fnet)
        if fold_bn:
            folded = A.fold_basic_convs(model)
        hd = model.args.hidden_dims[2]
        model.liif_up = A.adopt_liif_up(lu, chanels=[48 + hd, 32])          # agg_type 'type5': [stem_4x|hidden, stem_2x]
        yield model
    finally:
        A.unfold_basic_convs(folded)
        for n, v in saved.items():
            setattr(mod, n, v)
        model.update_block, model.liif_up, model.cnet = ub, lu, cnet
        if fnet is not None:
            model.fnet = fnet
        if stem is not None:
            model.corr_stem, model.corr_feature_att = stem


def forward(model, R, img1, img2, iters, scale=1.0):
    B, _, H, W = img1.shape
    Ho, Wo = int(round(H * scale)), int(round(W * scale))
    hr = R.make_coord([Ho, Wo]).to(img1.device)[None].expand(B, -1, -1).contiguous()
    sc = torch.full((B, 1), float(scale), device=img1.device)
    with torch.no_grad():
        out = model(img1, img2, iters=iters, test_mode=True, hr_coord=hr, scale=sc)
    return out.reshape(B, Ho, Wo)


def _timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def run(family, H, W, iters=32, B=1, engines=("bf16x3", "fp16"), device="cuda", timing=True, scale=1.0):
    import anystereo_b200 as A
    model, R = build_model(family, device)
    img1, img2 = make_pair(B, H, W, device)
    res = {"family": family, "image": [H, W], "batch": B, "iters": iters, "scale": scale, "engines": {}}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        # the parity oracle: the reference as shipped, strict fp32 on the same GPU (SURVEY 8c "oracle precision caveat")
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        ref, ms_ref = _timed(lambda: forward(model, R, img1, img2, iters, scale), 1 if not timing else 2)
        res["reference_strict_fp32_ms"] = ms_ref
        if timing:
            torch.backends.cudnn.allow_tf32 = True
            ref_tf32, ms_tf32 = _timed(lambda: forward(model, R, img1, img2, iters, scale))
            res["reference_default_tf32_ms"] = ms_tf32
            res["reference_default_tf32_epe_px"] = float((ref_tf32 - ref).abs().mean())
            torch.backends.cudnn.allow_tf32 = False
        prev_engine, prev_corr = A.get_update_engine(), A.get_corr_mode()
        for eng in engines:
            # None = whatever the library defaults to (what a user who only rebinds the names gets)
            if eng is not None:
                A.set_update_engine(eng)
                A.set_corr_mode({"fp32": "fp32", "bf16": "bf16"}.get(eng, "bf16x3"))
            with installed(model, R, family) as m:
                ours, ms = _timed(lambda: forward(m, R, img1, img2, iters, scale), 1 if not timing else 2)
            d = (ours - ref).abs()
            res["engines"][eng or "default(%s)" % A.get_update_engine()] = {
                "epe_mean_px": float(d.mean()), "epe_max_px": float(d.max()), "dropin_forward_ms": ms,
                "disp_range_px": [float(ref.min()), float(ref.max())]}
        A.set_update_engine(prev_engine)
        A.set_corr_mode(prev_corr)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    return res


def profile(family="igev", H=384, W=1248, iters=32, B=1, device="cuda", **kw):
    """Where one forward of the REAL reference graph spends its time with this library installed: CUDA-event time per
    top-level child module (update_block = all iterations; 'other' = lookups, volume build, glue outside any child)."""
    import anystereo_b200 as A
    model, R = build_model(family, device)
    img1, img2 = make_pair(B, H, W, device)
    out = {"family": family, "image": [H, W], "batch": B, "iters": iters, "engine": A.get_update_engine(), "options": kw,
           "ms": {}}
    import gc
    with installed(model, R, family, **kw) as m:
        forward(m, R, img1, img2, iters)                       # warm-up: weight packing, cudnn autotune
        forward(m, R, img1, img2, iters)
        # a generation-2 collection in the middle of a forward starves the GPU for 50-150 ms and lands in whichever module
        # is running (seen: cnet 48 ms instead of 7): no collections inside the measured forward
        gc.collect()
        gc.disable()
        spans = {}
        handles = []
        for name, child in m.named_children():
            def pre(mod, inp, name=name):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                spans.setdefault(name, []).append([e, None])
            def post(mod, inp, outp, name=name):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                spans[name][-1][1] = e
            handles.append(child.register_forward_pre_hook(pre))
            handles.append(child.register_forward_hook(post))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        forward(m, R, img1, img2, iters)
        e1.record()
        torch.cuda.synchronize()
        gc.enable()
        for h in handles:
            h.remove()
        total = e0.elapsed_time(e1)
        acc = 0.0
        for name, evs in spans.items():
            t = sum(a.elapsed_time(b) for a, b in evs if b is not None)
            out["ms"][name] = {"ms": round(t, 3), "calls": len(evs)}
            acc += t
        out["ms"]["other (lookups, volume build, glue)"] = {"ms": round(total - acc, 3), "calls": 1}
        out["total_ms"] = round(total, 3)
    return out


def train_step_time(H=320, W=736, iters=16, B=4, steps=4, device="cuda"):
    """Config-5 structure through the REAL reference graph (train() mode, frozen BatchNorm2d, sequence loss over every
    iteration's upsampled disparity, backward): ms per forward+backward for the reference as shipped (its defaults: TF32
    convolutions allowed) and with this library installed, plus the upsampler's share of the drop-in forward."""
    import anystereo_b200 as A
    model, R = build_model("igev", device)
    img1, img2 = make_pair(B, H, W, device)
    hr = R.make_coord([H, W]).to(device)[None].expand(B, -1, -1).contiguous()
    sc = torch.ones(B, 1, device=device)
    gt = torch.rand(B, 1, H * W, device=device) * 40.0
    out = {"image": [H, W], "batch": B, "iters": iters, "engine": A.get_update_engine()}

    def step(m, spans=None):
        m.zero_grad(set_to_none=True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        init_disp, preds = m(img1, img2, iters=iters, test_mode=False, hr_coord=hr, scale=sc)
        loss = sum(0.9 ** (len(preds) - 1 - i) * (p - gt).abs().mean() for i, p in enumerate(preds)) + init_disp.abs().mean()
        e[1].record()
        loss.backward()
        e[2].record()
        torch.cuda.synchronize()
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), float(loss.detach())

    def run(m, tag):
        m.train()
        m.freeze_bn()
        res = [step(m) for _ in range(steps)]
        res = res[1:]                                    # first step: autotune / weight packing
        out[tag] = {"forward_ms": round(sorted(r[0] for r in res)[len(res) // 2], 2),
                    "backward_ms": round(sorted(r[1] for r in res)[len(res) // 2], 2), "loss": res[-1][2],
                    "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
        m.eval()

    run(model, "reference_as_shipped")
    torch.cuda.reset_peak_memory_stats()
    with installed(model, R, "igev") as m:
        run(m, "dropin")
        # share of the upsampler (differentiable ATen formulation in training) in the drop-in forward
        spans = []
        lu = m.liif_up

        def pre(mod, inp):
            a = torch.cuda.Event(enable_timing=True)
            a.record()
            spans.append([a, None])

        def post(mod, inp, outp):
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            spans[-1][1] = b
        h1, h2 = lu.register_forward_pre_hook(pre), lu.register_forward_hook(post)
        m.train()
        m.freeze_bn()
        step(m)
        h1.remove()
        h2.remove()
        m.eval()
        out["dropin"]["liif_up_forward_ms_total"] = round(sum(a.elapsed_time(b) for a, b in spans if b is not None), 2)
        out["dropin"]["liif_up_calls"] = len(spans)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train", action="store_true", help="time a training step through the real graph (reference vs drop-in)")
    ap.add_argument("--profile", action="store_true", help="per-module time of one drop-in forward instead of the EPE table")
    ap.add_argument("--family", default="both", choices=["igev", "raft", "both"])
    ap.add_argument("--size", default=None, help="HxW (default: 384x1248 for igev, 320x736 for raft; both for 'both')")
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--engines", default="default,bf16x3,fp16,fp32")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    if not reference_available():
        print(json.dumps({"unavailable": "reference tree not found (baseline/_ref; run oracle/install_ref.py in the build container)"}))
        return
    if a.train:
        r = train_step_time(B=a.batch if a.batch > 1 else 4, iters=min(a.iters, 16))
        print(json.dumps(r), flush=True)
        if a.json:
            with open(a.json, "w") as f:
                json.dump(r, f, indent=1)
        return
    if a.profile:
        fams = ["igev", "raft"] if a.family == "both" else [a.family]
        res = []
        for fam in fams:
            H, W = (tuple(int(v) for v in a.size.split("x")) if a.size else (384, 1248))
            for kw in ({}, dict(defer_lookup=True, replay=True, fold_cnet=True,
                                **({"fuse_corr_stem": True, "fold_bn": True} if fam == "igev" else {"fused_fnet": True}))):
                r = profile(fam, H, W, a.iters, a.batch, **kw)
                res.append(r)
                print(json.dumps(r), flush=True)
        if a.json:
            with open(a.json, "w") as f:
                json.dump(res, f, indent=1)
        return
    engines = [None if e == "default" else e for e in a.engines.split(",")]
    jobs = []
    fams = ["igev", "raft"] if a.family == "both" else [a.family]
    for fam in fams:
        if a.size:
            sizes = [tuple(int(v) for v in a.size.split("x"))]
        else:
            sizes = [(384, 1248), (320, 736)]
        for hw in sizes:
            jobs.append((fam, hw))
    out = []
    for fam, (H, W) in jobs:
        r = run(fam, H, W, a.iters, a.batch, engines)
        out.append(r)
        print(json.dumps(r), flush=True)
    if a.json:
        os.makedirs(os.path.dirname(os.path.abspath(a.json)), exist_ok=True)
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
