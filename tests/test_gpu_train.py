"""GPU: the update block's explicit adjoints (a13-vi) against torch autograd over the oracle restatement, and one
data-parallel-style training step of the hot path (config 5 structure at a tiny size)."""
import types

import numpy as np
import pytest
import torch

import cases
from oracle import hotpath_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _block(A, family, seed):
    cls = A.BasicMultiUpdateBlock if family == "igev" else A.BasicMultiUpdateBlockRAFT
    args = types.SimpleNamespace(corr_levels=2 if family == "igev" else 4, corr_radius=4, n_gru_layers=3)
    m = cls(args, hidden_dims=[128, 128, 128])
    p = O.make_update_block_params(162 if family == "igev" else 36, seed=seed)
    m.load_state_dict(p, strict=True)
    return m.cuda().train(), p


@pytest.fixture(autouse=True)
def _restore_engine():
    import anystereo_b200 as A
    A.set_update_engine("fp32")
    yield
    A.set_update_engine("fp32")


# "bf16x3": forward convolutions and data gradients on the tcgen05 kernel (3-term split, fp32 accumulate), weight
# gradients on CUDA cores; same tolerances as the exact-fp32 engine
@pytest.mark.parametrize("engine", ["fp32", "bf16x3"])
@pytest.mark.parametrize("family", ["igev", "raft"])
def test_update_block_backward_vs_autograd(family, engine):
    import anystereo_b200 as A
    A.set_update_engine(engine)
    c = cases.update_block_case(family, B=2, H=9, W=13)
    m, p = _block(A, family, 31)
    rng = np.random.RandomState(5)
    # ---- reference: torch autograd over the oracle (CPU fp32)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    net_r = [t.clone().requires_grad_(True) for t in c["net"]]
    inp_r = [[t.clone().requires_grad_(True) for t in lst] for lst in c["inp"]]
    corr_r = c["corr"].clone().requires_grad_(True)
    out_net, out_delta = O.update_block(pr, net_r, inp_r, corr_r, c["disp"])
    cots = [torch.from_numpy(rng.standard_normal(tuple(t.shape)).astype("float32")) for t in out_net]
    cotd = torch.from_numpy(rng.standard_normal(tuple(out_delta.shape)).astype("float32"))
    loss = sum((o * w).sum() for o, w in zip(out_net, cots)) + (out_delta * cotd).sum()
    loss.backward()
    # ---- ours
    net_g = [t.cuda().requires_grad_(True) for t in c["net"]]
    inp_g = [[t.cuda().requires_grad_(True) for t in lst] for lst in c["inp"]]
    corr_g = c["corr"].cuda().requires_grad_(True)
    net_o, delta_o = m(list(net_g), inp_g, corr_g, c["disp"].cuda())
    for i in range(3):
        assert rel(net_o[i], out_net[i]) < 1e-4
    assert rel(delta_o, out_delta) < 1e-4
    loss_g = sum((o * w.cuda()).sum() for o, w in zip(net_o, cots)) + (delta_o * cotd.cuda()).sum()
    loss_g.backward()
    tol = 5e-4
    for i in range(3):
        assert rel(net_g[i].grad, net_r[i].grad) < tol, "d net[%d]" % i
        for j in range(3):
            assert rel(inp_g[i][j].grad, inp_r[i][j].grad) < tol, "d inp[%d][%d]" % (i, j)
    assert rel(corr_g.grad, corr_r.grad) < tol
    for name, prm in m.named_parameters():
        assert prm.grad is not None, name
        assert rel(prm.grad, pr[name].grad) < tol, name


@pytest.mark.parametrize("engine", ["fp32", "bf16x3"])
def test_lowres_only_backward(engine):
    """slow_fast_gru pre-pass (iter16/iter08 only, update=False): gradients still flow (update.py:116-133)."""
    import anystereo_b200 as A
    A.set_update_engine(engine)
    c = cases.update_block_case("igev", B=1, H=8, W=12)
    m, p = _block(A, "igev", 32)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    net_r = [t.clone().requires_grad_(True) for t in c["net"]]
    out = O.update_block(pr, net_r, c["inp"], iter16=True, iter08=True, iter04=False, update=False)
    sum(o.sum() for o in out).backward()
    net_g = [t.cuda().requires_grad_(True) for t in c["net"]]
    out_g = m(list(net_g), [[t.cuda() for t in l] for l in c["inp"]], iter16=True, iter08=True, iter04=False, update=False)
    sum(o.sum() for o in out_g).backward()
    for i in range(3):
        assert rel(net_g[i].grad, net_r[i].grad) < 5e-4
    assert rel(m.gru08.convz.weight.grad, pr["gru08.convz.weight"].grad) < 5e-4
    assert m.gru04.convz.weight.grad is None or float(m.gru04.convz.weight.grad.abs().max()) == 0.0


@pytest.mark.parametrize("engine", ["fp32", "bf16x3"])
def test_training_step_hot_path(engine):
    """Three unrolled iterations of the IGEV hot path with a sequence loss: gradients reach the matching features,
    the geometry volume, the context and every update-block parameter, and agree with autograd over the oracle."""
    import anystereo_b200 as A
    A.set_corr_mode("fp32")
    A.set_update_engine(engine)
    c = cases.loop_case("igev", seed=41, B=1, H=8, W=16)
    m, p = _block(A, "igev", 33)
    iters, gamma = 3, 0.9
    target = torch.from_numpy(np.random.RandomState(1).uniform(0, 10, size=(1, 1, 8, 16)).astype("float32"))

    def run(lookup_ctor, block, f1, f2, geo, net, inp, disp0, coords, tgt):
        fn = lookup_ctor(f1, f2, geo)
        disp, loss = disp0, 0.0
        for it in range(iters):
            disp = disp.detach()
            feat = fn(disp, coords)
            net, delta = block(net, inp, feat, disp)
            disp = disp + delta
            loss = loss + gamma ** (iters - 1 - it) * (disp - tgt).abs().mean()      # sequence_loss structure
        return loss

    # oracle / autograd
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    f1r, f2r, geor = (c[k].clone().requires_grad_(True) for k in ("f1", "f2", "geo"))

    def ctor_ref(f1, f2, geo):
        cp = O.corr_pyramid(O.all_pairs_corr(f1, f2), 2)
        gp = O.geo_pyramid(geo, 2)
        return lambda d, co: O.geo_lookup(gp, cp, d, co, 4, exact=True)

    def block_ref(net, inp, feat, disp):
        return O.update_block(pr, net, inp, feat, disp)

    lr = run(ctor_ref, block_ref, f1r, f2r, geor, [t.clone() for t in c["net"]], c["inp"], c["init_disp"],
             O.pixel_coords(1, 8, 16), target)
    lr.backward()
    # ours
    f1g, f2g, geog = (c[k].cuda().requires_grad_(True) for k in ("f1", "f2", "geo"))
    lg = run(lambda a, b, g: A.Combined_Geo_Encoding_Volume(a, b, g, num_levels=2, radius=4),
             lambda net, inp, feat, disp: m(net, inp, feat, disp),
             f1g, f2g, geog, [t.cuda() for t in c["net"]], [[t.cuda() for t in l] for l in c["inp"]],
             c["init_disp"].cuda(), O.pixel_coords(1, 8, 16).cuda(), target.cuda())
    lg.backward()
    assert abs(float(lg) - float(lr)) < 1e-4 * max(1.0, abs(float(lr)))
    assert rel(f1g.grad, f1r.grad) < 2e-3
    assert rel(f2g.grad, f2r.grad) < 2e-3
    assert rel(geog.grad, geor.grad) < 2e-3
    for name, prm in m.named_parameters():
        assert rel(prm.grad, pr[name].grad) < 2e-3, name


@pytest.mark.parametrize("shape", [
    # (B, H, W, source channels, Cout, kernel)
    (2, 9, 13, [128, 128, 128], 256, 3),      # gru04/gru08 z|r: three sources, two Cout tiles
    (1, 20, 46, [128, 128], 128, 3),          # gru16 q at config 5's 1/16 scale
    (2, 7, 70, [128], 127, 3),                # encoder.conv: 127 outputs (dY pitch 128), two 64-pixel chunks per row
    (2, 6, 10, [162], 64, 1),                 # convc1: 1x1 over the 162 lookup channels (second N tile is partial)
    (1, 5, 8, [64], 64, 3),                   # convc2 / convd2
])
def test_wgrad_tensor_cores_vs_cuda_cores(shape):
    """as_conv2d_wgrad_umma (tcgen05, K = pixels over channel-major planes) against as_conv2d_wgrad_fp32 and torch."""
    import anystereo_b200 as A
    from anystereo_b200 import update_train as T
    B, H, W, chans, Cout, k = shape
    g = torch.Generator(device="cpu").manual_seed(sum(shape[:3]) + Cout)
    conv = torch.nn.Conv2d(sum(chans), Cout, k, padding=k // 2).cuda()
    xs = [torch.randn(B, H, W, c, generator=g).cuda() for c in chans]
    pitch = (Cout + 63) // 64 * 64
    dy = torch.zeros(B, H, W, pitch).cuda()
    dy[..., :Cout] = torch.randn(B, H, W, Cout, generator=g).cuda()
    c = T._Conv([conv])
    srcs = [T._src(x) for x in xs]
    A.set_update_engine("fp32")
    dw0, db0 = c.wgrad(B, H, W, srcs, dy, pitch)
    A.set_update_engine("bf16x3")
    assert T._WGRAD_TC["on"]
    before = T.L.launch_count
    dw1, db1 = c.wgrad(B, H, W, srcs, dy, pitch)
    assert T.L.launch_count - before >= 3 + len(xs)          # transposes + tensor-core GEMM + bias reduction
    torch.cuda.synchronize()
    # torch autograd on the same GPU, strict fp32
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x = torch.cat(xs, 3).permute(0, 3, 1, 2).contiguous()
        y = conv(x)
        gw, gb = torch.autograd.grad(y, [conv.weight, conv.bias], dy[..., :Cout].permute(0, 3, 1, 2).contiguous())
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert rel(dw0, gw) < 1e-4
    assert rel(dw1, gw) < 1e-4, "tensor-core weight gradient"
    assert rel(dw1, dw0) < 1e-4
    assert rel(db1, gb) < 1e-4
