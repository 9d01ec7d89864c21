#!/usr/bin/env python
"""Experiment: one forward of the REAL graph at batch 1, 384x1248, 32 iterations, PyTorch defaults (TF32 convolutions
allowed): reference as shipped vs drop-in installed; per-repetition CUDA-event times with the Python garbage collector
enabled and disabled (a generation-2 collection in the middle of a forward starves the GPU for tens of ms)."""
import gc
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import dropin  # noqa: E402

fam = sys.argv[1] if len(sys.argv) > 1 else "igev"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
model, R = dropin.build_model(fam, "cuda")
H, W = (int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "384x1248").split("x"))
img1, img2 = dropin.make_pair(B, H, W, "cuda")
print("%s batch %d %dx%d, 32 iterations" % (fam, B, H, W))


def reps(fn, n=6):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(round(e0.elapsed_time(e1), 1))
    return out


for collect in (False,):
    (gc.enable if collect else gc.disable)()
    print("gc %s" % ("on" if collect else "off"))
    print("  reference as shipped (TF32 default):", reps(lambda: dropin.forward(model, R, img1, img2, 32), 4), flush=True)
    variants = [{}, {"defer_lookup": True}, {"defer_lookup": True, "replay": True},
                {"defer_lookup": True, "replay": True, "fold_cnet": True}]
    if fam == "raft":
        variants.append({"defer_lookup": True, "replay": True, "fold_cnet": True, "fused_fnet": True})
    if fam == "igev":
        variants.append({"fuse_corr_stem": True, "defer_lookup": True})
        variants.append({"fuse_corr_stem": True, "defer_lookup": True, "replay": True, "fold_cnet": True})
        variants.append({"fuse_corr_stem": True, "defer_lookup": True, "replay": True, "fold_cnet": True, "fold_bn": True})
    for kw in variants:
        with dropin.installed(model, R, fam, **kw) as m:
            print("  drop-in %s:" % kw, reps(lambda: dropin.forward(m, R, img1, img2, 32)), flush=True)
