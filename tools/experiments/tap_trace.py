"""Timeline of the tap-major fused lookup kernel (experiment; needs a library built with `make EXTRA=-DAS_TAP_TRACE`):
CTA 0 records clock64 at the hand-over points of its first 16 tiles; prints per-tile intervals in cycles."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import anystereo_b200 as A  # noqa: E402

dev = "cuda"
B, D, H, W, Dg = 8, 96, 96, 312, 48
torch.manual_seed(0)
f1 = torch.randn(B, D, H, W, device=dev) * 0.2
f2 = torch.randn(B, D, H, W, device=dev) * 0.2
geo = torch.randn(B, 8, Dg, H, W, device=dev)
A.set_corr_mode("bf16x3")
A.set_update_engine(sys.argv[1] if len(sys.argv) > 1 else "bf16x3")
vol = A.Combined_Geo_Encoding_Volume(f1, f2, geo, num_levels=2, radius=4)
disp = (torch.rand(B, 1, H, W, device=dev) * Dg).contiguous()
bias = torch.randn(64, device=dev)
with A._lib.operand_format_scope(A.update_umma._sixteen_bit_format()):
    w_hi, w_lo = A.geometry.DeferredGeoLookup.pack_convc1_weight(torch.randn(64, 162, 1, 1, device=dev) * 0.1, True, True, bias)
o_hi = torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16)
o_lo = torch.empty_like(o_hi)
dl = vol.deferred(disp, None)
for _ in range(3):
    dl.convc1_planes(w_hi, w_lo, bias, o_hi, o_lo, True)
torch.cuda.synchronize()
tr = torch.zeros(28 * 16 * 8, device=dev, dtype=torch.int64)
lib = A._lib.lib()
lib.as_tap_trace_set.argtypes = [ctypes.c_void_p]
assert lib.as_tap_trace_set(tr.data_ptr()) == 0
dl.convc1_planes(w_hi, w_lo, bias, o_hi, o_lo, True)
torch.cuda.synchronize()
t = tr.view(28, 16, 8).cpu()
t0 = int(t[t > 0].min())
rel = lambda x: int(x) - t0 if int(x) > 0 else -1   # noqa: E731
print("producer warp w: step start, landed, next loads issued, stage free, arrive")
for w in (0, 5, 13, 19):
    for k in range(8):
        print("P%02d tile %2d:" % (w, k), [rel(t[w, k, e]) for e in range(5)])
print("MMA warp: a_full seen, acc_empty seen, committed")
for k in range(10):
    print("MMA tile %2d:" % k, [rel(t[23, k, e]) for e in range(3)])
print("epilogue warp 24: wait start, acc_full seen, tmem drained, stores issued")
for k in range(10):
    print("EPI tile %2d:" % k, [rel(t[20, k, e]) for e in range(4)])
