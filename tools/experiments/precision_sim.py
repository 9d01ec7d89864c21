"""Experiment (CPU, not product code): which operand arithmetic keeps the 32-iteration result inside the EPE gate?

The REAL reference models (unmodified code) run on CPU at 192x640; only the update block's nn.Conv2d arithmetic is
emulated (SURVEY 7 'chaotic amplification' protocol):
   fp32      reference
   bf16x3    a = ah + al (bf16), w = wh + wl:  ah*wh + ah*wl + al*wh          (3 tensor-core passes, today's parity engine)
   fp16      round(a) * round(w) in IEEE half                                   (1 pass)
   f16f8     ah*wh in IEEE half (11-bit) + both cross terms in ONE fp8 pass: [al*2^s | ah*2^-t] . [wh*2^-s | wl*2^t] with
             e5m2 operands (K doubled, fp8 runs at twice the f16 rate) -> 2 pass-equivalents
all products accumulated in fp32 like the tensor core does.
    python tools/experiments/precision_sim.py [--family igev|raft] [--iters 32]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import dropin  # noqa: E402


def q(x, dt):
    return x.to(dt).to(torch.float32)


def split(x, dt):
    hi = q(x, dt)
    return hi, x - hi


def conv_mode(mode, s_act=6, t_act=8):
    def conv(x, w, b, pad):
        if mode == "fp32":
            return F.conv2d(x, w, b, padding=pad)
        if mode == "bf16":
            return F.conv2d(q(x, torch.bfloat16), q(w, torch.bfloat16), b, padding=pad)
        if mode == "fp16":
            return F.conv2d(q(x, torch.float16), q(w, torch.float16), b, padding=pad)
        if mode == "bf16x3":
            ah, al = split(x, torch.bfloat16)
            wh, wl = split(w, torch.bfloat16)
            al, wl = q(al, torch.bfloat16), q(wl, torch.bfloat16)
            return F.conv2d(ah, wh, b, padding=pad) + F.conv2d(ah, wl, None, padding=pad) + F.conv2d(al, wh, None, padding=pad)
        if mode in ("f16f8", "f16f8_e4m3hi", "f16_act_only"):
            ah, al = split(x, torch.float16)
            wh, wl = split(w, torch.float16)
            main = F.conv2d(ah, wh, b, padding=pad)
            if mode == "f16_act_only":      # 2 f16 passes: activations exact, weights rounded to half
                return main + F.conv2d(q(al, torch.float16), wh, None, padding=pad)
            f8 = torch.float8_e5m2
            hi8 = torch.float8_e4m3fn if mode == "f16f8_e4m3hi" else f8
            sa, ta = 2.0 ** s_act, 2.0 ** t_act
            c1 = F.conv2d(q(al * sa, f8), q(wh / sa, hi8 if mode != "f16f8_e4m3hi" else f8), None, padding=pad)
            c2 = F.conv2d(q(ah / ta, f8), q(wl * ta, f8), None, padding=pad)
            return main + c1 + c2
        if mode in ("f16f8_wl", "f16f8_al"):          # ONE cross term only (half of the e5m2 pass): weight residual / activation residual
            ah, al = split(x, torch.float16)
            wh, wl = split(w, torch.float16)
            main = F.conv2d(ah, wh, b, padding=pad)
            f8 = torch.float8_e5m2
            sa, ta = 2.0 ** s_act, 2.0 ** t_act
            if mode == "f16f8_wl":
                return main + F.conv2d(q(ah / ta, f8), q(wl * ta, f8), None, padding=pad)
            return main + F.conv2d(q(al * sa, f8), q(wh / sa, f8), None, padding=pad)
        raise ValueError(mode)
    return conv


def run(family, iters, H, W, modes):
    model, R = dropin.build_model(family, "cpu")
    img1, img2 = dropin.make_pair(1, H, W, "cpu")
    convs = [m for m in model.update_block.modules() if isinstance(m, torch.nn.Conv2d)]
    out = {}
    names = {id(m): n for n, m in model.update_block.named_modules()}
    for mode in modes:
        # mixed modes "mix_<pattern>[+<pattern>]": the convolutions whose name contains a pattern run single-pass fp16, the rest
        # f16f8 (e.g. mix_convz+convr: the GRU gates z, r in one pass; mix_gru16+gru08: the low-resolution GRUs)
        # "mixw_<patterns>": the named convolutions keep only the WEIGHT-residual cross term (half of the e5m2 pass)
        cheap = "f16f8_wl" if mode.startswith("mixw_") else "fp16"
        pats = mode[5:].split("+") if mode.startswith("mixw_") else (mode[4:].split("+") if mode.startswith("mix_") else None)
        for m in convs:
            mm = mode if pats is None else (cheap if any(p in names[id(m)] for p in pats) else "f16f8")
            fn = conv_mode(mm)
            m.forward = (lambda x, m=m, fn=fn: fn(x.float(), m.weight, m.bias, m.padding))
        res = dropin.forward(model, R, img1, img2, iters)
        out[mode] = res
        if mode != "fp32":
            d = (res - out["fp32"]).abs()
            print("%-14s mean |d| %.2e px   max %.2e px" % (mode, float(d.mean()), float(d.max())), flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="igev")
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--size", default="192x640")
    ap.add_argument("--modes", default="fp32,bf16x3,fp16,f16f8,f16_act_only,bf16")
    a = ap.parse_args()
    H, W = (int(v) for v in a.size.split("x"))
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        run(a.family, a.iters, H, W, a.modes.split(","))
