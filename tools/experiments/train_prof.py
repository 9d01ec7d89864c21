"""Experiment: top CUDA kernels of one training step through the REAL reference graph with the drop-in installed
(torch.profiler; run on the GPU box):  python tools/experiments/train_prof.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import dropin as D  # noqa: E402

B, H, W, iters = 4, 320, 736, 16
model, R = D.build_model("igev", "cuda")
img1, img2 = D.make_pair(B, H, W, "cuda")
hr = R.make_coord([H, W]).cuda()[None].expand(B, -1, -1).contiguous()
sc = torch.ones(B, 1, device="cuda")
gt = torch.rand(B, 1, H * W, device="cuda") * 40.0


def step(m):
    m.zero_grad(set_to_none=True)
    init_disp, preds = m(img1, img2, iters=iters, test_mode=False, hr_coord=hr, scale=sc)
    loss = sum(0.9 ** (len(preds) - 1 - i) * (p - gt).abs().mean() for i, p in enumerate(preds)) + init_disp.abs().mean()
    loss.backward()


with D.installed(model, R, "igev") as m:
    m.train()
    m.freeze_bn()
    step(m)
    step(m)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(m)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=70))
