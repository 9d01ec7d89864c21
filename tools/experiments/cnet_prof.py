#!/usr/bin/env python
"""Experiment: where does the reference's context encoder (IGEV `cnet` = MultiBasicEncoder) spend its time at
batch 1, 384x1248?  Times the module alone under the cuDNN settings a user could flip (TF32, channels_last, benchmark mode,
bf16 autocast) and prints the top kernels of the default setting.  Reads baseline/_ref through oracle/ref_loader.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import dropin  # noqa: E402


def timed(fn, reps=3):
    fn()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    fam = sys.argv[1] if len(sys.argv) > 1 else "igev"
    model, R = dropin.build_model(fam, "cuda")
    img1, img2 = dropin.make_pair(1, 384, 1248, "cuda")
    x = (2 * (img1 / 255.0) - 1.0).contiguous()
    cnet = model.cnet
    n = model.args.n_gru_layers

    def call(mod=cnet, inp=x):
        with torch.no_grad():
            return mod(inp, num_layers=n)

    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        for bench in (False, True):
            torch.backends.cudnn.benchmark = bench
            print("cnet NCHW tf32=%d benchmark=%d : %.2f ms" % (tf32, bench, timed(call)), flush=True)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = False
    import copy
    cl = copy.deepcopy(cnet).to(memory_format=torch.channels_last)
    xcl = x.contiguous(memory_format=torch.channels_last)
    print("cnet channels_last tf32 : %.2f ms" % timed(lambda: call(cl, xcl)), flush=True)

    def call_bf16():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return cl(xcl, num_layers=n)
    print("cnet channels_last bf16 autocast : %.2f ms" % timed(call_bf16), flush=True)

    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        call()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
    # per top-level child of the encoder
    spans = {}
    hs = []
    for name, child in cnet.named_children():
        def pre(m, i, name=name):
            e = torch.cuda.Event(enable_timing=True); e.record(); spans.setdefault(name, []).append([e, None])
        def post(m, i, o, name=name):
            e = torch.cuda.Event(enable_timing=True); e.record(); spans[name][-1][1] = e
        hs += [child.register_forward_pre_hook(pre), child.register_forward_hook(post)]
    call()
    torch.cuda.synchronize()
    for h in hs:
        h.remove()
    for k, v in spans.items():
        print("  %-12s %.3f ms (%d calls)" % (k, sum(a.elapsed_time(b) for a, b in v), len(v)))
    if fam == "igev":
        f = model.feature
        print("feature (both images) : %.2f ms" % timed(lambda: torch.no_grad()(lambda: f(torch.cat([x, x])))()), flush=True)


if __name__ == "__main__":
    main()
