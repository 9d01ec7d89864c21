import torch, time
x = torch.randn(942080, 128, device="cuda")
W = torch.randn(64, 128, device="cuda") * 0.1
def split(t):
    h = t.to(torch.bfloat16); l = (t - h.float()).to(torch.bfloat16); return h, l
xh, xl = split(x); Wh, Wl = split(W)
try:
    y = torch.mm(xh, Wh.t(), out_dtype=torch.float32)
    y = y + torch.mm(xh, Wl.t(), out_dtype=torch.float32) + torch.mm(xl, Wh.t(), out_dtype=torch.float32)
    ref = (x.double() @ W.double().t())
    print("out_dtype ok; rel err", float((y.double()-ref).abs().max()/ref.abs().max()))
except Exception as e:
    print("out_dtype failed:", repr(e)[:200])
torch.backends.cuda.matmul.allow_tf32 = False
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("fp32 mm ms", t(lambda: x @ W.t()))
print("3x bf16 mm ms", t(lambda: torch.mm(xh, Wh.t(), out_dtype=torch.float32) + torch.mm(xh, Wl.t(), out_dtype=torch.float32) + torch.mm(xl, Wh.t(), out_dtype=torch.float32)))
print("split ms", t(lambda: split(x)))
g = torch.randn(942080, 64, device="cuda"); gh, gl = split(g)
print("fp32 wgrad ms", t(lambda: g.t() @ x))
print("3x bf16 wgrad ms", t(lambda: torch.mm(gh.t(), xh, out_dtype=torch.float32) + torch.mm(gh.t(), xl, out_dtype=torch.float32) + torch.mm(gl.t(), xh, out_dtype=torch.float32)))
