import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import torch, dropin
import anystereo_b200 as A
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
model, R = dropin.build_model("igev", "cuda")
H, W, iters = 64, 128, 3
img1, img2 = dropin.make_pair(2, H, W, "cuda")
hr = R.make_coord([H, W]).cuda()[None].expand(2, -1, -1).contiguous()
sc = torch.ones(2, 1, device="cuda")
gt = torch.rand(2, 1, H * W, device="cuda") * 40.0
names = ["update_block.gru04.convz.weight", "update_block.encoder.convc1.weight", "update_block.disp_head.conv2.weight",
         "liif_up.imnet.layers.0.weight", "liif_up.imnet.layers.6.bias", "stem_2.conv1.conv.weight", "classifier.weight", "desc.weight", "context_zqr_convs.0.weight"]
params = dict(model.named_parameters()); names = [n for n in names if n in params]
def run(m):
    m.train(); m.freeze_bn(); m.zero_grad(set_to_none=True)
    init_disp, preds = m(img1, img2, iters=iters, test_mode=False, hr_coord=hr, scale=sc)
    loss = 0.0
    for i, p in enumerate(preds): loss = loss + 0.9 ** (len(preds) - 1 - i) * (p - gt).abs().mean()
    loss = loss + init_disp.abs().mean(); loss.backward()
    g = {n: dict(m.named_parameters())[n].grad.detach().clone() for n in names}; m.eval()
    return float(loss.detach()), g
l0, g0 = run(model)
l0b, g0b = run(model)
print("ref vs ref:", abs(l0b - l0) / abs(l0), max(float((g0b[n] - g0[n]).abs().max() / g0[n].abs().max()) for n in names))
for rep in range(8):
    with dropin.installed(model, R, "igev") as m:
        l1, g1 = run(m)
    errs = {n.split('.')[-2] + '.' + n.split('.')[-1] if n.count('.') > 1 else n: float((g1[n] - g0[n]).abs().max() / g0[n].abs().max()) for n in names}
    print(rep, "loss rel", abs(l1 - l0) / abs(l0), "max grad err", max(errs.values()), max(errs, key=errs.get))
