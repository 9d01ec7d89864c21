"""Measure the LIIF arbitrary-scale upsampler (BASELINE config 4: x2.5 / x3.7 queries on a
384x1248 IGEV pair) on one B200: our kernels (CUDA events, per stage and end to end), the same arithmetic in plain
torch on the same GPU, and the oracle on the host cores for a bounded sample.

    python tools/liif_bench.py [--B 1] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import anystereo_b200 as A  # noqa: E402
from oracle import liif_oracle as LO  # noqa: E402  (checker / baseline only)

AFF = {"win_w": 3, "win_h": 3, "dilation": [1, 2, 4, 8]}


def grid_coords(B, H, W, dev):
    ys = -1 + 1.0 / H + (2.0 / H) * torch.arange(H, device=dev).float()
    xs = -1 + 1.0 / W + (2.0 / W) * torch.arange(W, device=dev).float()
    g = torch.stack(torch.meshgrid(ys, xs, indexing="ij"), -1).reshape(1, -1, 2)
    return g.expand(B, -1, -1).contiguous()


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--json", default=None)
    ap.add_argument("--engine", default="bf16x3")
    ap.add_argument("--no-torch", action="store_true")
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    B, h, w = a.B, 96, 312
    stem4 = torch.randn(B, 48, h, w, device=dev)
    hid = torch.tanh(torch.randn(B, 128, h, w, device=dev))
    stem2 = torch.randn(B, 32, 2 * h, 2 * w, device=dev)
    disp = torch.rand(B, 1, h, w, device=dev) * 48
    m = A.liif_out_multi_scale_Training(encoder_dim=208, mlphidden_list=[128, 64, 64], pos_dim=0, unfold="with_v2ISU",
                                        affinity_settings=AFF, number_input=2, chanels=[176, 32])
    params = LO.make_liif_params(228, seed=1)
    m.load_state_dict(params, strict=True)
    m = m.cuda().eval()
    A.set_update_engine(a.engine)
    res = {"B": B, "engine": a.engine, "lowres": [h, w], "scales": {}}
    x = torch.cat([stem4, hid], 1)
    feats = [x, stem2]
    for scale in (2.5, 3.7):
        H, W = int(h * 4 * scale), int(w * 4 * scale)
        coords = grid_coords(B, H, W, dev)
        Q = coords.shape[1]
        sc = torch.full((B,), scale, device=dev)
        ms_all = ev_time(lambda: A.upsample_disp(m, disp, hid, stem4, stem2, None, hr_coord=coords, scale=sc))
        split = a.engine != "bf16"
        wts = m._weights(split)
        ms_pre = ev_time(lambda: m._first_layer_maps(feats, wts, split))
        r = {"out_hw": [H, W], "queries_per_pair": Q, "ms_total": ms_all, "ms_source_res_stage": ms_pre,
             "ms_query_kernel": ms_all - ms_pre, "Mqueries_per_s": B * Q / ms_all / 1e3,
             "pairs_per_s": B / (ms_all / 1e3)}
        flops = 2.0 * B * Q * (128 * 64 + 64 * 64 + 64 * 9)
        r["query_kernel_logical_TFLOPs"] = flops / ((ms_all - ms_pre) * 1e-3) / 1e12
        if not a.no_torch and B * Q * 228 * 4 < 40e9:
            pg = {k: v.to(dev) for k, v in params.items()}
            fn = lambda: LO.upsample_disp_multiscale(pg, disp, feats, coords, sc)   # noqa: E731
            ref = fn()
            got = A.upsample_disp(m, disp, hid, stem4, stem2, None, hr_coord=coords, scale=sc)
            r["max_rel_err_vs_torch_gpu"] = float((got - ref).abs().max() / ref.abs().max())
            r["torch_same_gpu_ms"] = ev_time(fn, reps=3, warm=1)
            del ref, got
        # host cores, bounded sample of the same workload
        nq = 200_000
        cs = coords[:1, :: max(1, Q // nq)][:, :nq].cpu().contiguous()
        cf = [f[:1].cpu() for f in feats]
        t0 = time.time()
        LO.upsample_disp_multiscale(params, disp[:1].cpu(), cf, cs, sc[:1].cpu())
        dt = time.time() - t0
        r["cpu_oracle"] = {"sample_queries": int(cs.shape[1]), "seconds": dt, "Mqueries_per_s": cs.shape[1] / dt / 1e6,
                           "cores": torch.get_num_threads(), "note": "includes the full-map affinity stage once"}
        res["scales"]["x%.1f" % scale] = r
        print("x%.1f" % scale, json.dumps(r), flush=True)
        del coords
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
