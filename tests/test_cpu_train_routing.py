"""CPU: host-side routing of the training path (update_train.py) per update engine.  The C-ABI entry points are
replaced by a recorder (no GPU here), so this checks which kernels each convolution of one update-block call and its
backward pass is sent to -- and that nothing is computed by torch on the way (every tensor op is an ABI call)."""
import collections
import types

import pytest
import torch

import cases


@pytest.fixture()
def recorder(monkeypatch):
    import anystereo_b200 as A
    from anystereo_b200 import _lib as L
    calls = []
    monkeypatch.setattr(L, "call", lambda name, *a: calls.append(name))
    monkeypatch.setattr(L, "stream_ptr", lambda: 0)
    monkeypatch.setattr(L, "require_cuda", lambda t, *a, **k: t)
    prev = A.get_update_engine()
    yield A, calls
    monkeypatch.setattr(L, "set_operand_format", lambda fmt: None)     # restoring the engine must not touch the library
    A.update._ENGINE["engine"] = prev


def _step(A, family, engine, small_tc=True, convd1=True):
    from anystereo_b200 import update_train as T
    prev = dict(T._KNOBS)
    T._KNOBS.update(small_tc=small_tc, convd1=convd1)
    try:
        return _step_impl(A, family, engine)
    finally:
        T._KNOBS.update(prev)


def _step_impl(A, family, engine):
    c = cases.update_block_case(family, B=1, H=8, W=12)
    cls = A.BasicMultiUpdateBlock if family == "igev" else A.BasicMultiUpdateBlockRAFT
    args = types.SimpleNamespace(corr_levels=2 if family == "igev" else 4, corr_radius=4, n_gru_layers=3)
    m = cls(args, hidden_dims=[128, 128, 128]).train()
    A.update._ENGINE["engine"] = engine                  # set_update_engine would call into the (GPU) library
    net = [t.clone().requires_grad_(True) for t in c["net"]]
    inp = [[t.clone().requires_grad_(True) for t in lst] for lst in c["inp"]]
    corr = c["corr"].clone().requires_grad_(True)
    out_net, delta = m(list(net), inp, corr, c["disp"])
    (sum(o.sum() for o in out_net) + delta.sum()).backward()
    return m


@pytest.mark.parametrize("family", ["igev", "raft"])
def test_fp32_engine_routes_to_cuda_core_kernels(recorder, family):
    A, calls = recorder
    _step(A, family, "fp32")
    n = collections.Counter(calls)
    assert n["as_conv2d_fp32"] == 13 + 12                 # 13 forward convolutions (z|r fused) + 12 data gradients
    assert n["as_conv2d_wgrad_fp32"] == 13
    assert not any(k in n for k in ("as_conv2d_umma", "as_conv2d_wgrad_umma", "as_transpose_split"))


@pytest.mark.parametrize("family", ["igev", "raft"])
def test_tensor_core_engine_routes_forward_dgrad_wgrad_to_tcgen05(recorder, family):
    A, calls = recorder
    _step(A, family, "bf16x3")
    n = collections.Counter(calls)
    # forward: 13 convolutions, all but convd1 (7x7, one input channel: dedicated CUDA-core kernel) on tcgen05
    assert n["as_conv_epilogue_fp32"] == 12 and n["as_convd1_fp32"] == 1
    # 12 data gradients; gru08/gru04 have 384 input channels = two N <= 256 launches each for z|r and for q
    assert n["as_conv2d_umma"] == 12 + 12 + 4
    assert n["as_conv2d_wgrad_umma"] == 12 and n["as_convd1_wgrad_fp32"] == 1 and n["as_bias_grad_fp32"] == 13
    assert "as_conv2d_wgrad_fp32" not in n and "as_conv2d_fp32" not in n      # nothing left on the generic CUDA-core kernels
    # channel-major planes are shared between the z|r and q weight gradients of a GRU (per-backward cache):
    # 12 dY + 4 single-source convs (convc1, convc2, convd2, conv) + dh1 + dh2 + per GRU (h, rh, each x once)
    assert n["as_transpose_split"] == 12 + 6 + (2 + 1) + (2 + 2) + (2 + 2)


def test_small_conv_knobs_fall_back_to_generic_kernels(recorder):
    A, calls = recorder
    from anystereo_b200 import update_train as T
    assert set(T._KNOBS) == {"small_tc", "convd1"}
    _step(A, "igev", "bf16x3", small_tc=False, convd1=False)
    n = collections.Counter(calls)
    assert n["as_conv2d_fp32"] == 3 and n["as_conv2d_wgrad_fp32"] == 2 and n["as_conv2d_wgrad_umma"] == 11


def test_wgrad_knob_falls_back_to_cuda_cores(recorder):
    A, calls = recorder
    from anystereo_b200 import update_train as T
    T.set_wgrad_tensor_cores(False)
    try:
        _step(A, "igev", "bf16x3")
    finally:
        T.set_wgrad_tensor_cores(True)
    n = collections.Counter(calls)
    assert n["as_conv2d_wgrad_fp32"] == 12 and n["as_convd1_wgrad_fp32"] == 1 and "as_conv2d_wgrad_umma" not in n


def test_upsampler_differentiable_path_matches_oracle_autograd():
    """ADVICE r1 (high/medium): whenever a gradient is requested through the forward-only upsampler pieces they switch to a
    differentiable formulation (device-agnostic ATen ops, so this runs on the CPU): values and gradients must equal the
    oracle restatement of the reference (oracle/liif_oracle.py, itself pinned to the reference's outputs)."""
    import anystereo_b200 as A
    from oracle import liif_oracle as LO
    torch.manual_seed(0)
    lc = cases.liif_case(2, h=5, w=7, scale=1.7, extra_q=5)
    chanels = [f.shape[1] for f in lc["feats"]]
    m = A.liif_out_multi_scale_Training(encoder_dim=sum(chanels), mlphidden_list=[128, 64, 64], pos_dim=0,
                                        unfold="with_v2ISU", affinity_settings={"win_w": 3, "win_h": 3, "dilation": [1, 2, 4, 8]},
                                        number_input=2, chanels=chanels)
    lp = LO.make_liif_params(lc["in_dim"], seed=2)
    m.load_state_dict(lp, strict=True)
    m.train()
    feats = [f.clone().requires_grad_(True) for f in lc["feats"]]
    disp = lc["disp"].clone().requires_grad_(True)
    out = A.upsample_disp(m, disp, feats[0][:, 48:], feats[0][:, :48], feats[1], None, hr_coord=lc["coords"], scale=lc["scale"])
    assert out.grad_fn is not None
    wgt = torch.linspace(0.5, 1.5, out.numel()).view_as(out)
    (out * wgt).sum().backward()
    p2 = {k: v.clone().requires_grad_(True) for k, v in lp.items()}
    f2 = [f.clone().requires_grad_(True) for f in lc["feats"]]
    d2 = lc["disp"].clone().requires_grad_(True)
    ref = LO.upsample_disp_multiscale(p2, d2, f2, lc["coords"], lc["scale"])
    (ref * wgt).sum().backward()
    assert torch.allclose(out, ref, atol=1e-5, rtol=1e-5)
    assert torch.allclose(disp.grad, d2.grad, atol=1e-6, rtol=1e-4)
    for a, b in zip(feats, f2):
        # the first Linear is applied at source resolution (a re-association in fp32): gated relative to the largest
        # gradient element (the affinity's normalisation of near-zero vectors makes single elements ~1e12)
        assert float((a.grad - b.grad).abs().max() / b.grad.abs().max()) < 1e-5
    for k, v in m.state_dict(keep_vars=True).items():
        assert float((v.grad - p2[k].grad).abs().max() / p2[k].grad.abs().max()) < 1e-5, k
    # the small pieces on their own
    x = torch.rand(1, 6, 4, 5, requires_grad=True)
    assert torch.allclose(A.liif.isu_affinity(x), LO.isu_affinity(x.detach()), atol=1e-6)
    pr = torch.softmax(torch.rand(1, 12, 3, 4, requires_grad=True), 1)
    dr = A.disparity_regression(pr, 12)
    assert dr.grad_fn is not None and torch.allclose(dr, (pr * torch.arange(12.0).view(1, 12, 1, 1)).sum(1, keepdim=True))
