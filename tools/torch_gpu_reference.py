"""The reference's own torch path (einsum / avg_pool2d / grid_sample / cuDNN conv2d -- restated in
oracle/hotpath_oracle.py) timed ON THE SAME B200 at BASELINE config 2: the honest GPU "implementation to beat"
(SURVEY.md 8d).  Two arithmetic settings: PyTorch's defaults (cuDNN convolutions in TF32) and strict fp32.
MEASUREMENT TOOL ONLY: this runs the oracle, not the product."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hotpath_oracle as O  # noqa: E402


def main():
    dev = "cuda"
    B, H, W, D, Dg, iters = 8, 96, 312, 96, 48, 32
    torch.manual_seed(0)
    f1 = torch.randn(B, D, H, W, device=dev)
    f2 = torch.randn(B, D, H, W, device=dev)
    sizes = [(H, W), (H // 2, W // 2), (H // 4, W // 4)]
    net = [torch.tanh(torch.randn(B, 128, h, w, device=dev)) for h, w in sizes]
    inp = [[torch.relu(torch.randn(B, 128, h, w, device=dev)) for _ in range(3)] for h, w in sizes]
    disp0 = torch.rand(B, 1, H, W, device=dev) * 40
    p = {k: v.to(dev) for k, v in O.make_update_block_params(162, seed=0).items()}
    out = {}
    for name, tf32 in (("torch_default_tf32_convs", True), ("torch_strict_fp32", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        ts = []
        with torch.no_grad():
            for rep in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                geo = O.gwc_volume(f1, f2, Dg, 8)
                O.igev_iterations(p, f1, f2, geo, net, inp, disp0, iters)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
        best = min(ts[1:])
        out[name] = {"ms_per_step": best * 1e3, "pairs_per_s": B / best}
        print(name, out[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "torch_gpu_reference_r01.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
