"""The drop-in executed on a B200: the reference's REAL model graphs with this library installed against
the same graphs as shipped (strict fp32 on the same GPU), at the BASELINE shapes -- 384x1248 (config 2 / 4) and
320x736 (config 1 / 5) -- 32 iterations, final FULL-RESOLUTION disparity.  Gate: mean |diff| <= 0.01 px
(BASELINE.json north_star "final disparity EPE within 0.01 px of the reference after the same iteration count").

The reference tree travels to the GPU box as the git-ignored pristine copy ``baseline/_ref/`` (oracle/install_ref.py);
without it these tests are skipped with that reason.
"""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import dropin  # noqa: E402

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not dropin.reference_available(),
                               reason="reference tree absent (baseline/_ref is created by oracle/install_ref.py)")


@pytest.fixture(autouse=True)
def library_defaults():
    """These tests run on what the library selects by itself (no set_update_engine / set_corr_mode by the caller)."""
    import anystereo_b200 as A
    A.set_update_engine(A.update.DEFAULT_ENGINE)
    A.set_corr_mode(A.geometry.DEFAULT_CORR_MODE)
    yield
    A.set_update_engine(A.update.DEFAULT_ENGINE)
    A.set_corr_mode(A.geometry.DEFAULT_CORR_MODE)


def _record(res):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "dropin_epe.jsonl"), "a") as f:
            f.write(json.dumps(res) + "\n")


@needs_ref
@pytest.mark.parametrize("family,H,W", [("igev", 384, 1248), ("igev", 320, 736), ("raft", 320, 736), ("raft", 384, 1248)])
def test_dropin_epe_at_baseline_shapes(family, H, W):
    # engines: None = the library's defaults, i.e. what rebinding the names alone selects (must be a tensor-core engine of
    # the fp32-parity class); both parity engines are gated at 1e-3 px (10x inside the BASELINE gate); "fp16" is
    # recorded live, gated at 0.01 px, and reported separately
    res = dropin.run(family, H, W, iters=32, B=1, engines=(None, "bf16x3", "f16f8", "fp16"), timing=False)
    _record(res)
    default = [k for k in res["engines"] if k.startswith("default")]
    assert default in (["default(bf16x3)"], ["default(f16f8)"]), res["engines"].keys()
    for k in (default[0], "bf16x3", "f16f8"):
        assert res["engines"][k]["epe_mean_px"] <= 1e-3, (k, res)
    assert res["engines"]["fp16"]["epe_mean_px"] <= 0.01, res


@needs_ref
def test_dropin_arbitrary_scale_query_x2p5():
    """Config 4: the same graph queried at x2.5 (960x3120 = 3.0 M points) through the adopted LIIF upsampler."""
    res = dropin.run("igev", 384, 1248, iters=8, B=1, engines=(None,), timing=False, scale=2.5)
    _record(res)
    e = list(res["engines"].values())[0]
    assert e["epe_mean_px"] <= 0.01 * 2.5, res        # disparities scale with the query grid


@needs_ref
def test_dropin_batch_and_restores_reference_names():
    import anystereo_b200 as A
    model, R = dropin.build_model("igev", "cuda")
    orig = R.igev_module.Combined_Geo_Encoding_Volume
    img1, img2 = dropin.make_pair(2, 128, 256, "cuda")
    ref = dropin.forward(model, R, img1, img2, 6)
    ref_ub = model.update_block
    with dropin.installed(model, R, "igev") as m:
        assert R.igev_module.Combined_Geo_Encoding_Volume is A.Combined_Geo_Encoding_Volume
        assert isinstance(m.update_block, A.BasicMultiUpdateBlock)
        # the adopted block shares the reference's Parameter objects (checkpoints / optimizers keep working)
        assert m.update_block.gru04.convz.weight is ref_ub.gru04.convz.weight
        ours = dropin.forward(m, R, img1, img2, 6)
    assert R.igev_module.Combined_Geo_Encoding_Volume is orig
    assert float((ours - ref).abs().mean()) <= 0.01


@needs_ref
def test_dropin_training_step_matches_reference():
    """Config-5 structure through the REAL reference graph (ADVICE r1 high / VERDICT missing #1): train() mode, frozen BN,
    sequence loss over every iteration's upsampled disparity AND the supervised initial disparity, backward.  With this
    library installed the loss and the gradients of the update block, the LIIF MLP, the stems, the context network and
    the cost-aggregation classifier must equal the reference's own autograd."""
    model, R = dropin.build_model("igev", "cuda")
    H, W, iters = 64, 128, 3
    img1, img2 = dropin.make_pair(2, H, W, "cuda")
    hr = R.make_coord([H, W]).cuda()[None].expand(2, -1, -1).contiguous()
    sc = torch.ones(2, 1, device="cuda")
    gt = torch.rand(2, 1, H * W, device="cuda") * 40.0
    names = ["update_block.gru04.convz.weight", "update_block.encoder.convc1.weight", "update_block.disp_head.conv2.weight",
             "liif_up.imnet.layers.0.weight", "liif_up.imnet.layers.6.bias", "stem_2.conv1.conv.weight",
             "classifier.weight", "desc.weight", "context_zqr_convs.0.weight"]
    params = dict(model.named_parameters())
    names = [n for n in names if n in params]
    assert len(names) >= 7, names

    def run(m):
        m.train()
        m.freeze_bn()
        m.zero_grad(set_to_none=True)
        init_disp, preds = m(img1, img2, iters=iters, test_mode=False, hr_coord=hr, scale=sc)
        loss = 0.0
        for i, p in enumerate(preds):                                   # train_continuous_IGEV.py:37-122 structure
            loss = loss + 0.9 ** (len(preds) - 1 - i) * (p - gt).abs().mean()
        loss = loss + init_disp.abs().mean()                           # --supervise_init path: init_disp carries a graph
        loss.backward()
        g = {n: dict(m.named_parameters())[n].grad.detach().clone() for n in names}
        m.eval()
        return float(loss.detach()), g

    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        loss_ref, g_ref = run(model)
        _, g_ref2 = run(model)          # the reference's own run-to-run noise (atomics in its grid_sample / index adjoints)
        with dropin.installed(model, R, "igev") as m:
            loss_our, g_our = run(m)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert abs(loss_our - loss_ref) <= 1e-4 * abs(loss_ref), (loss_our, loss_ref)
    for n in names:
        scale = g_ref[n].abs().max().clamp_min(1e-20)
        err = float((g_our[n] - g_ref[n]).abs().max() / scale)
        noise = float((g_ref2[n] - g_ref[n]).abs().max() / scale)      # up to 3.5e-3 on `desc.weight` (measured)
        # both sides accumulate their gather adjoints with atomics (run-to-run noise on both): seen failing once in ~10
        # runs at max(2e-3, 3 x noise)
        # (`desc.weight`: the reference differs from ITSELF by up to 3.5e-3 between runs, and a two-run noise estimate can come
        #  out at 2e-4 on the same box: the floor has to cover the noise, not its estimate)
        assert err <= max(5e-3, 4.0 * noise), (n, err, noise)


def test_lookup_rejects_mismatched_disp():
    """a disparity map that does not match the pyramid raises instead of reading out of bounds."""
    import anystereo_b200 as A
    f1 = torch.randn(2, 32, 8, 24, device="cuda")
    f2 = torch.randn(2, 32, 8, 24, device="cuda")
    blk = A.CorrBlock1D(f1, f2, num_levels=2, radius=4)
    coords = A.hotpath.pixel_coords(2, 8, 24, "cuda")
    blk(torch.zeros(2, 1, 8, 24, device="cuda"), coords)
    for bad in ((1, 1, 8, 24), (2, 1, 4, 12), (2, 1, 8, 25)):
        with pytest.raises(RuntimeError, match="cost volume was built for"):
            blk(torch.zeros(bad, device="cuda"), None)
    geo = torch.randn(2, 8, 12, 8, 24, device="cuda")
    vol = A.Combined_Geo_Encoding_Volume(f1, f2, geo, num_levels=2, radius=4)
    with pytest.raises(RuntimeError, match="cost volume was built for"):
        vol(torch.zeros(4, 1, 8, 24, device="cuda"), None)
    with pytest.raises(RuntimeError, match="cost volume was built for"):
        vol.deferred(torch.zeros(2, 1, 16, 24, device="cuda"), None)


def test_inference_mode_and_invalidate_weights():
    """inference tensors have no version counter; stale packed weights after a .data write."""
    import types

    import anystereo_b200 as A
    from oracle import hotpath_oracle as O
    import cases
    c = cases.loop_case("igev", seed=5, B=1, H=16, W=24)
    args = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=3)
    m = A.BasicMultiUpdateBlock(args, hidden_dims=[128, 128, 128])
    m.load_state_dict(O.make_update_block_params(162, seed=5), strict=True)
    m = m.cuda().eval()
    cu = lambda t: t.cuda()                                                     # noqa: E731
    with torch.no_grad():
        d0, _ = A.igev_iterations(m, cu(c["f1"]), cu(c["f2"]), cu(c["geo"]), [cu(t) for t in c["net"]],
                                  [[cu(t) for t in l] for l in c["inp"]], cu(c["init_disp"]), 3)
    with torch.inference_mode():
        d1, _ = A.igev_iterations(m, cu(c["f1"]), cu(c["f2"]), cu(c["geo"]), [cu(t) for t in c["net"]],
                                  [[cu(t) for t in l] for l in c["inp"]], cu(c["init_disp"]), 3)
    assert torch.equal(d0, d1)
    m.disp_head.conv2.bias.data.add_(1.0)        # bypasses the version counter
    m.invalidate_weights()
    with torch.no_grad():
        d2, _ = A.igev_iterations(m, cu(c["f1"]), cu(c["f2"]), cu(c["geo"]), [cu(t) for t in c["net"]],
                                  [[cu(t) for t in l] for l in c["inp"]], cu(c["init_disp"]), 1)
        m.disp_head.conv2.bias.data.sub_(1.0)
        m.invalidate_weights()
        d3, _ = A.igev_iterations(m, cu(c["f1"]), cu(c["f2"]), cu(c["geo"]), [cu(t) for t in c["net"]],
                                  [[cu(t) for t in l] for l in c["inp"]], cu(c["init_disp"]), 1)
    assert abs(float((d2 - d3).mean()) - 1.0) < 1e-4


@needs_ref
def test_dropin_fused_corr_stem_matches_reference():
    """SURVEY 8(f)-3 inside the REAL graph: build_gwc_volume + corr_stem + corr_feature_att through adopt_corr_stem (one
    fused kernel, the GWC volume never in HBM) vs the unmodified model: same final disparity as the plain drop-in; the
    state_dict keys of the adopted modules are the reference's; train() mode falls back to the unfused operators."""
    import anystereo_b200 as A
    D = dropin
    model, R = D.build_model("igev", "cuda")
    img1, img2 = D.make_pair(1, 320, 736, "cuda")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        keys = set(model.state_dict().keys())
        ref = D.forward(model, R, img1, img2, 12)
        with D.installed(model, R, "igev") as m:
            plain = D.forward(m, R, img1, img2, 12)
        n0 = A._lib.launch_count
        with D.installed(model, R, "igev", fuse_corr_stem=True) as m:
            assert set(m.state_dict().keys()) == keys
            fused = D.forward(m, R, img1, img2, 12)
            m.train()
            m.freeze_bn()
            with torch.no_grad():
                pending = m.corr_stem(A.DeferredGwcVolume(torch.randn(1, 96, 8, 16, device="cuda"),
                                                          torch.randn(1, 96, 8, 16, device="cuda"), 48, 8))
            # the reference's freeze_bn() only reaches BatchNorm2d: corr_stem's BatchNorm3d is in train mode, so the adopted
            # module takes the reference's arithmetic (batch statistics) instead of the fused eval-mode kernel
            assert torch.is_tensor(pending) and pending.shape == (1, 8, 48, 8, 16)
            x = torch.randn(1, 96, 8, 16, device="cuda", requires_grad=True)
            v = R.igev_module.build_gwc_volume(x, x, 48, 8)
            assert torch.is_tensor(v) and v.requires_grad                 # a gradient is requested: the plain operator
            m.eval()
        assert R.igev_module.build_gwc_volume is not A.submodule.build_gwc_volume_deferred    # restored
        e_plain = float((plain - ref).abs().mean())
        e_fused = float((fused - ref).abs().mean())
        assert e_fused < 1e-3 and abs(e_fused - e_plain) < 5e-4, (e_plain, e_fused)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


@needs_ref
@pytest.mark.parametrize("family", ["igev", "raft"])
def test_dropin_deferred_lookup_matches_plain_dropin(family):
    """install_into_reference(defer_lookup=True): the reference's own loop hands an unevaluated lookup to the adopted update
    block, which runs it fused with convc1.  Same final disparity as the plain drop-in (K order of the fp32
    accumulation differs) and within the EPE gate of the unmodified model; fewer launches."""
    import anystereo_b200 as A
    D = dropin
    model, R = D.build_model(family, "cuda")
    img1, img2 = D.make_pair(1, 320, 736, "cuda")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = D.forward(model, R, img1, img2, 12)
        with D.installed(model, R, family) as m:
            D.forward(m, R, img1, img2, 12)
            n0 = A._lib.launch_count
            plain = D.forward(m, R, img1, img2, 12)
            n_plain = A._lib.launch_count - n0
        with D.installed(model, R, family, defer_lookup=True) as m:
            D.forward(m, R, img1, img2, 12)
            n0 = A._lib.launch_count
            fused = D.forward(m, R, img1, img2, 12)
            n_fused = A._lib.launch_count - n0
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert n_fused <= n_plain - 12, (n_plain, n_fused)            # at least one launch fewer per iteration
    assert float((fused - ref).abs().mean()) < 1e-3
    assert float((fused - plain).abs().mean()) < 5e-4


@needs_ref
@pytest.mark.parametrize("family,size", [("igev", (384, 1248)), ("raft", (320, 736))])
def test_dropin_lowres_single_pass_knob_epe(family, size):
    """Opt-in speed mode (set_lowres_single_pass): the 1/8- and 1/16-resolution GRUs in one tensor-core pass.  The final
    full-resolution disparity of the REAL graphs stays inside the 1e-3 px bar of the default engine (the simulation said
    1.3e-4 / 3.4e-4 px); recorded live here."""
    import anystereo_b200 as A
    D = dropin
    model, R = D.build_model(family, "cuda")
    img1, img2 = D.make_pair(1, size[0], size[1], "cuda")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    prev = A.set_lowres_single_pass(False)
    try:
        ref = D.forward(model, R, img1, img2, 32)
        with D.installed(model, R, family) as m:
            two = D.forward(m, R, img1, img2, 32)
            A.set_lowres_single_pass(True)
            one = D.forward(m, R, img1, img2, 32)
    finally:
        A.set_lowres_single_pass(prev)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    e2, e1 = float((two - ref).abs().mean()), float((one - ref).abs().mean())
    _record({"family": family, "image": list(size), "lowres_single_pass": {"epe_default_px": e2, "epe_knob_px": e1,
                                                                       "epe_knob_max_px": float((one - ref).abs().max())}})
    assert e2 < 1e-3 and e1 < 1e-3, (e2, e1)
    assert not torch.equal(one, two)


@needs_ref
def test_dropin_gate_weight_residual_only_epe():
    """IGEV default: the 1/4-resolution gates keep only the weight-residual cross term.  Final disparity of the REAL graph at
    384x1248 vs the unmodified model, with the mode on (default) and off: both inside the 1e-3 px bar (recorded live)."""
    import anystereo_b200 as A
    D = dropin
    model, R = D.build_model("igev", "cuda")
    img1, img2 = D.make_pair(1, 384, 1248, "cuda")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    prev = A.set_gate_weight_residual_only(None)
    try:
        ref = D.forward(model, R, img1, img2, 32)
        with D.installed(model, R, "igev") as m:
            on = D.forward(m, R, img1, img2, 32)
            A.set_gate_weight_residual_only(False)
            off = D.forward(m, R, img1, img2, 32)
    finally:
        A.set_gate_weight_residual_only(prev)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    e_on, e_off = float((on - ref).abs().mean()), float((off - ref).abs().mean())
    _record({"family": "igev", "image": [384, 1248], "gate_weight_residual_only": {"epe_on_px": e_on, "epe_off_px": e_off,
                                                                                 "epe_on_max_px": float((on - ref).abs().max())}})
    assert e_on < 1e-3 and e_off < 1e-3, (e_on, e_off)
    assert not torch.equal(on, off)


@needs_ref
@pytest.mark.parametrize("family,defer", [("igev", False), ("igev", True), ("raft", True)])
def test_dropin_call_replay_is_bit_identical(family, defer):
    """adopt_update_block(..., replay=True): every call of the block inside the reference's own loop is replayed from a CUDA
    graph (captured on the first call of a forward with new shapes / cost-volume buffers).  Same kernels on the same data:
    the final disparity must equal the eager drop-in bit for bit, on the capturing forward, on a replay-only forward, and
    on a second image pair (context refreshed, hidden states reloaded); the returned hidden states are private copies."""
    import anystereo_b200 as A
    D = dropin
    model, R = D.build_model(family, "cuda")
    pairs = [D.make_pair(1, 320, 736, "cuda", seed=s) for s in (3, 4)]
    with D.installed(model, R, family, defer_lookup=defer) as m:
        eager = [D.forward(m, R, a, b, 8) for a, b in pairs]
    with D.installed(model, R, family, defer_lookup=defer, replay=True) as m:
        first = D.forward(m, R, *pairs[0], 8)
        n0 = A._lib.launch_count
        again = D.forward(m, R, *pairs[0], 8)
        n_replayed = A._lib.launch_count - n0
        other = D.forward(m, R, *pairs[1], 8)
        back = D.forward(m, R, *pairs[0], 8)
        graphs = len(m.update_block.__dict__["_umma_state"]["calls"])
        assert 1 <= graphs <= 2
        # a different shape captures its own graph and still matches
        small = D.make_pair(1, 256, 512, "cuda", seed=5)
        got_small = D.forward(m, R, *small, 4)
    with D.installed(model, R, family, defer_lookup=defer) as m:
        want_small = D.forward(m, R, *small, 4)
        want_small2 = D.forward(m, R, *small, 4)
    assert n_replayed > 8 * 15                       # the replayed launches are counted (gpu_launches stays honest)
    assert torch.equal(first, eager[0]) and torch.equal(again, eager[0]) and torch.equal(back, eager[0])
    assert torch.equal(other, eager[1])
    if torch.equal(want_small, want_small2):
        assert torch.equal(got_small, want_small)
    else:       # the reference's own cuDNN layers (3-D deconvolutions of cost_agg) are not run-to-run deterministic at this shape
        assert float((got_small - want_small).abs().max()) < 1e-3


def test_call_replay_outputs_are_private_and_knobs_key_the_graph():
    """Operator level: a replayed call returns clones (a later call must not rewrite tensors the caller still holds), an
    engine / knob change captures a new graph instead of replaying a stale one, and autograd / fp32 calls stay eager."""
    import anystereo_b200 as A
    from anystereo_b200 import update_umma
    torch.manual_seed(0)
    import types
    args = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=3)
    ub = A.BasicMultiUpdateBlock(args, hidden_dims=[128, 128, 128]).cuda().eval()
    ub.call_replay = True
    B, H, W = 1, 40, 72
    net = [torch.tanh(torch.randn(B, 128, H >> i, W >> i, device="cuda")) for i in range(3)]
    inp = [[torch.randn(B, 128, H >> i, W >> i, device="cuda") for _ in range(3)] for i in range(3)]
    corr = torch.randn(B, 162, H, W, device="cuda")
    disp = torch.rand(B, 1, H, W, device="cuda") * 20

    def call(block, n):
        with torch.no_grad():
            return block(list(n), inp, corr, disp)
    ub.call_replay = False
    want1 = call(ub, net)
    want2 = call(ub, want1[0])
    ub.call_replay = True
    got1 = call(ub, net)
    keep = [t.clone() for t in got1[0]] + [got1[1].clone()]
    got2 = call(ub, got1[0])                          # replays; must not touch got1
    for a, b in zip(list(got1[0]) + [got1[1]], keep):
        assert torch.equal(a, b)
    for a, b in zip(list(got1[0]) + [got1[1]], list(want1[0]) + [want1[1]]):
        assert torch.equal(a, b)
    for a, b in zip(list(got2[0]) + [got2[1]], list(want2[0]) + [want2[1]]):
        assert torch.equal(a, b)
    calls = ub.__dict__["_umma_state"]["calls"]
    assert len(calls) == 1
    prev = A.set_lowres_single_pass(False)
    try:
        ub.call_replay = False
        want3 = call(ub, net)
        ub.call_replay = True
        got3 = call(ub, net)
        assert len(calls) == 2
        for a, b in zip(list(got3[0]) + [got3[1]], list(want3[0]) + [want3[1]]):
            assert torch.equal(a, b)
    finally:
        A.set_lowres_single_pass(prev)
    # a parameter update (version bump) may not replay the old weights
    with torch.no_grad():
        ub.disp_head.conv2.bias.add_(1.0)
    got4 = call(ub, net)
    assert torch.allclose(got4[1], got1[1] + 1.0, atol=1e-5)
    with torch.no_grad():
        ub.disp_head.conv2.bias.sub_(1.0)
    ub.invalidate_weights()
    assert len(calls) == 0
    # gradients requested: the differentiable path, no graph
    n_req = [t.clone().requires_grad_(True) for t in net]
    out = ub(n_req, inp, corr, disp)
    assert out[1].requires_grad and len(calls) == 0


@needs_ref
@pytest.mark.parametrize("family", ["igev", "raft"])
def test_dropin_everything_adopted_epe(family):
    """The full drop-in a user can install today -- deferred lookups, replayed update-block calls, folded context encoder
    (SURVEY 8(f)-4), fused corr stem (IGEV) -- on the REAL graph against the same graph as shipped in strict fp32: final
    full-resolution disparity far inside the BASELINE gate (0.01 px).  Measured 1.4e-4 px (IGEV) / 9.2e-4 px (RAFT; 7.9e-4
    without the folded encoder: folding reorders fp32 roundings of the context features by ~1e-5 relative and RAFT's
    iteration amplifies any such perturbation -- the reference's own TF32 default sits at 1.5e-2 px); bar 2e-3 px.
    The folded encoder alone agrees with the reference module to fp32 rounding."""
    import anystereo_b200 as A
    D = dropin
    model, R = D.build_model(family, "cuda")
    img1, img2 = D.make_pair(1, 384, 1248, "cuda")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = D.forward(model, R, img1, img2, 32)
        x = (2 * (img1 / 255.0) - 1.0).contiguous()
        with torch.no_grad():
            want = model.cnet(x, num_layers=3)
            got = A.adopt_context_encoder(model.cnet)(x, num_layers=3)
        for lw, lg in zip(want, got):
            for w, g in zip(lw, lg):
                assert float((g - w).abs().max()) <= 1e-4 * float(w.abs().max())
        kw = dict(defer_lookup=True, replay=True, fold_cnet=True)
        if family == "igev":
            kw["fuse_corr_stem"] = True
            kw["fold_bn"] = True
        with D.installed(model, R, family, **kw) as m:
            assert isinstance(m.cnet, A.ContextEncoder)
            assert family != "igev" or isinstance(m.cost_agg.conv1[0], A.FoldedBasicConv)
            D.forward(m, R, img1, img2, 32)
            ours = D.forward(m, R, img1, img2, 32)
        assert not isinstance(model.cnet, A.ContextEncoder)
        assert family != "igev" or type(model.cost_agg.conv1[0]).__name__ == "BasicConv"
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    epe = float((ours - ref).abs().mean())
    _record({"test": "everything_adopted", "family": family, "epe_mean_px": epe, "options": kw})
    assert epe < 2e-3, epe


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 37, 53), (1, 96, 48, 80), (3, 128, 9, 7), (1, 4, 1, 2), (2, 64, 384, 416)])
def test_instnorm_kernel_matches_torch(B, C, H, W):
    """csrc/instnorm.cu vs F.instance_norm in fp64 (biased variance, eps 1e-5): plain, + ReLU, + residual; in place; a
    channel mean 300x its standard deviation (the shifted sums must not cancel); ragged sizes, a single pixel."""
    import torch.nn.functional as F
    from anystereo_b200 import extractor
    torch.manual_seed(B * 1000 + C)
    x = torch.randn(B, C, H, W, device="cuda")
    x[:, 1] = x[:, 1] * 0.01 + 3.0                                   # mean >> std
    x[:, 2] = x[:, 2] * 50.0 - 20.0
    x = x.contiguous(memory_format=torch.channels_last)
    r = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    want = F.instance_norm(x.double(), eps=1e-5)
    for relu, resid in ((False, None), (True, None), (True, r), (False, r)):
        w = want.relu() if relu else want
        if resid is not None:
            w = (w + resid.double()).relu()
        got = extractor._instnorm_(x.clone(memory_format=torch.channels_last), 1e-5, relu=relu, resid=resid)
        assert got.is_contiguous(memory_format=torch.channels_last)
        err = float((got.double() - w).abs().max())
        # (the mean is held in fp32: a channel whose mean is 300x its standard deviation sees ulp(mean) * rstd ~ 3e-5)
        assert err <= 5e-5 * max(1.0, float(w.abs().max())), (relu, resid is not None, err)
    # out aliasing the residual (the block's running tensor) is allowed
    rr = r.clone(memory_format=torch.channels_last)
    import anystereo_b200 as A
    L = A._lib
    ws_bytes = L.lib().as_instnorm_workspace_bytes(B, C)
    ws = torch.empty(ws_bytes, device="cuda", dtype=torch.uint8)
    L.call("as_instnorm_nhwc", x.data_ptr(), rr.data_ptr(), rr.data_ptr(), ws.data_ptr(), ws_bytes, B, H * W, C, 1e-5, 1,
           L.stream_ptr())
    w = (want.relu() + r.double()).relu()
    assert float((rr.double() - w).abs().max()) <= 5e-5 * max(1.0, float(w.abs().max()))


@needs_ref
def test_adopted_feature_encoder_matches_reference_module_and_graph():
    """SURVEY 8(f)-4, RAFT: adopt_feature_encoder(model.fnet) -- channels-last convolutions + the fused InstanceNorm kernels --
    against the reference module in strict fp32 (list input as prune_raft_stereo.py:252 passes it, and a single tensor),
    then inside the real graph with everything else adopted: final disparity within the bar."""
    import anystereo_b200 as A
    D = dropin
    model, R = D.build_model("raft", "cuda")
    img1, img2 = D.make_pair(1, 320, 736, "cuda")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        a = (2 * (img1 / 255.0) - 1.0).contiguous()
        b = (2 * (img2 / 255.0) - 1.0).contiguous()
        ours = A.adopt_feature_encoder(model.fnet)
        with torch.no_grad():
            w1, w2 = model.fnet([a, b])
            n0 = A._lib.launch_count
            g1, g2 = ours([a, b])
            assert A._lib.launch_count - n0 == 3 * 15           # 15 normalisations, 3 kernels each
            ws, gs = model.fnet(a), ours(a)
            with torch.autocast("cuda", dtype=torch.float16):          # prune_raft_stereo.py:251 autocast(mixed_precision)
                ga = ours(a)
            assert ga.dtype == torch.float32 and torch.equal(ga, gs)
        for w, g in ((w1, g1), (w2, g2), (ws, gs)):
            assert g.shape == w.shape and g.is_contiguous()
            assert float((g - w).abs().max()) <= 1e-4 * float(w.abs().max())
        # gradients requested -> the reference's forward
        a_req = a.clone().requires_grad_(True)
        assert ours(a_req).requires_grad
        ref = D.forward(model, R, img1, img2, 32)
        with D.installed(model, R, "raft", defer_lookup=True, replay=True, fold_cnet=True, fused_fnet=True) as m:
            assert isinstance(m.fnet, A.FeatureEncoder)
            D.forward(m, R, img1, img2, 32)
            got = D.forward(m, R, img1, img2, 32)
        assert not isinstance(model.fnet, A.FeatureEncoder)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    epe = float((got - ref).abs().mean())
    _record({"test": "raft_everything_adopted_incl_fnet", "epe_mean_px": epe})
    assert epe < 2e-3, epe


@needs_ref
@pytest.mark.parametrize("family", ["igev", "raft"])
def test_adopt_model_one_call(family):
    """A.adopt_model(model, module, family, replay=True): the one-call installation gives the same final disparity as the
    step-by-step installation tools/dropin.py performs (same kernels; cuDNN's own layers are not bit-reproducible run to
    run) and stays inside the bar against the model as shipped."""
    import copy
    import anystereo_b200 as A
    D = dropin
    model, R = D.build_model(family, "cuda")
    mod = R.igev_module if family == "igev" else R.raft_module
    img1, img2 = D.make_pair(1, 320, 736, "cuda")
    ref = D.forward(model, R, img1, img2, 16)
    kw = dict(defer_lookup=True, replay=True, fold_cnet=True)
    kw.update({"fuse_corr_stem": True} if family == "igev" else {"fused_fnet": True})
    with D.installed(model, R, family, **kw) as m:
        D.forward(m, R, img1, img2, 16)
        stepwise = D.forward(m, R, img1, img2, 16)
    names = ["Combined_Geo_Encoding_Volume", "build_gwc_volume", "context_upsample_multiscale_train", "CorrBlock1D"]
    saved = {n: getattr(mod, n) for n in names if hasattr(mod, n)}
    try:
        twin = A.adopt_model(copy.deepcopy(model), mod, family, replay=True)
        D.forward(twin, R, img1, img2, 16)
        one_call = D.forward(twin, R, img1, img2, 16)
    finally:
        for n, v in saved.items():
            setattr(mod, n, v)
    assert float((one_call - stepwise).abs().mean()) < 5e-4
    assert float((one_call - ref).abs().mean()) < 1e-2            # vs the TF32 reference as shipped: its own error dominates
