"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header declares, host-side
logic, the C restatement of the sampler against the golden vectors, and error behaviour at the boundary."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

import cases
from oracle import hotpath_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def A():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "any-stereo_b200", "csrc", "libanystereo_b200.so")):
        os.environ["ANYSTEREO_SKIP_REF_BUILD"] = "1"
        g.build()
    import anystereo_b200 as a
    return a


def test_header_symbols_exported(A):
    hdr = open(os.path.join(ROOT, "include", "anystereo_b200.h")).read()
    declared = set(re.findall(r"\b(as_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"as_stream_t"}
    assert len(declared) >= 25
    lib = A._lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: " + name
    # and the ctypes table binds exactly the declared functions
    assert set(A._lib.SIGNATURES) == declared
    assert lib.as_abi_version() == 1
    assert lib.as_compiled_sm() == 100
    assert b"ok" == lib.as_error_string(0)
    assert b"alignment" in lib.as_error_string(-4)


def test_struct_layouts_match_header(A, tmp_path):
    """ctypes mirrors of the descriptor structs agree with what a C compiler makes of include/anystereo_b200.h."""
    import ctypes
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "anystereo_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(as_conv_src), sizeof(as_conv_desc), offsetof(as_conv_desc, src),
         offsetof(as_conv_desc, weight), offsetof(as_conv_desc, out), offsetof(as_conv_desc, save));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(as_umma_src), sizeof(as_conv_umma_desc), offsetof(as_conv_umma_desc, src),
         offsetof(as_conv_umma_desc, w_hi), offsetof(as_conv_umma_desc, out_hi), offsetof(as_conv_umma_desc, u));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(as_liif_query_desc), offsetof(as_liif_query_desc, P), offsetof(as_liif_query_desc, coords),
         offsetof(as_liif_query_desc, w2_hi), offsetof(as_liif_query_desc, disp), offsetof(as_liif_query_desc, out));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(as_wgrad_umma_desc), offsetof(as_wgrad_umma_desc, src), offsetof(as_wgrad_umma_desc, dy_hi),
         offsetof(as_wgrad_umma_desc, Wp), offsetof(as_wgrad_umma_desc, ws), offsetof(as_wgrad_umma_desc, dw_acc));
  return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    c = [int(x) for x in out]
    L = A._lib
    d, u = L.ConvDesc, L.ConvUmmaDesc
    assert c[:6] == [ctypes.sizeof(L.ConvSrc), ctypes.sizeof(d), d.src.offset, d.weight.offset, d.out.offset, d.save.offset]
    assert c[6:12] == [ctypes.sizeof(L.UmmaSrc), ctypes.sizeof(u), u.src.offset, u.w_hi.offset, u.out_hi.offset, u.u.offset]
    q = L.LiifQueryDesc
    assert c[12:18] == [ctypes.sizeof(q), q.P.offset, q.coords.offset, q.w2_hi.offset, q.disp.offset, q.out.offset]
    g = L.WgradUmmaDesc
    assert c[18:] == [ctypes.sizeof(g), g.src.offset, g.dy_hi.offset, g.Wp.offset, g.ws.offset, g.dw_acc.offset]


def test_argument_errors_without_gpu(A):
    lib = A._lib.lib()
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.as_sampler_fwd(None, None, 1, None, 1, 1, 1, 1, 4, 0, None) == -1
    assert lib.as_gwc_build_fwd(None, None, None, 1, 8, 1, 1, 4, 8, None) == -1
    assert lib.as_pool1d_halve(None, None, 1, 4, 4, 2, None) == -1
    assert lib.as_isu_affinity(None, 1, 8, 4, 4, None, None, None, 0, 0, None) == -1
    assert lib.as_liif_query(None, None) == -1
    assert lib.as_context_upsample_multiscale(None, None, None, None, 1, 2, 2, 4, None) == -1
    assert lib.as_corr1d_workspace_bytes(1, 8, 2, 4, 4, 0) == 0
    # training entry points (a13-vi): epilogue on a raw conv output, channel-major planes, tensor-core weight gradient
    assert lib.as_conv_epilogue_fp32(None, 64, 16, 64, 1, None, 0, None, None, None, None, 64, 0, None) == -1
    assert lib.as_transpose_split(None, 64, 0, 64, 1, 4, 4, None, None, 8, 1, None) == -1
    assert lib.as_conv2d_wgrad_umma(None, None) == -1
    wd = A._lib.WgradUmmaDesc()                      # a descriptor without planes is rejected before any CUDA call
    wd.B, wd.H, wd.W, wd.KH, wd.KW, wd.Cout, wd.num_src, wd.Wp, wd.nsplit = 1, 4, 4, 3, 3, 64, 1, 8, 3
    assert lib.as_conv2d_wgrad_umma(ctypes.byref(wd), None) == -1
    assert lib.as_bias_grad_fp32(None, 64, 64, 16, None, None) == -1
    assert lib.as_convd1_fp32(None, None, None, None, 1, 4, 4, 64, 0, None) == -1
    assert lib.as_convd1_wgrad_fp32(None, None, 64, 1, 4, 4, None, None) == -1
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        A.corr_sampler.forward(torch.zeros(1, 2, 3, 4), torch.zeros(1, 1, 2, 3), 4)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        A.CorrBlock1D(torch.zeros(1, 4, 2, 8), torch.zeros(1, 4, 2, 8))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        A.build_gwc_volume(torch.zeros(1, 8, 2, 8), torch.zeros(1, 8, 2, 8), 4, 8)
    with pytest.raises(AssertionError):
        A.build_gwc_volume(torch.zeros(1, 9, 2, 8), torch.zeros(1, 9, 2, 8), 4, 8)


def test_update_block_state_dict_matches_reference_names(A):
    for cls, planes in ((A.BasicMultiUpdateBlock, 162), (A.BasicMultiUpdateBlockRAFT, 36)):
        args = types.SimpleNamespace(corr_levels=2 if planes == 162 else 4, corr_radius=4, n_gru_layers=3)
        m = cls(args, hidden_dims=[128, 128, 128])
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        want = {}
        for name, shp in O.update_block_param_shapes(planes):
            want[name + ".weight"] = shp
            want[name + ".bias"] = (shp[0],)
        assert ours == want
        assert list(ours) == list(want)          # registration order too (RNG-stream compatible init)
        assert sum(int(np.prod(s)) for s in ours.values()) in (4071488, 4063424)


def test_lowres_single_pass_default(A):
    """The f16f8 engine's low-resolution GRUs run one pass by default; the switch returns the previous state."""
    assert A.update_umma._LOWRES_1PASS["on"] is True
    assert A.set_lowres_single_pass(False) is True
    assert A.set_lowres_single_pass(True) is False


def test_call_replay_switches(A):
    """Per-call graph replay is opt-in: module switch, per-block override (adopt_update_block(..., replay=...)), suspended
    inside the library's own loops (those are captured as a whole), and never taken without corr / disp."""
    import types
    um = A.update_umma
    assert um._CALL_REPLAY["on"] is False
    args = types.SimpleNamespace(corr_levels=2, corr_radius=4, n_gru_layers=3)
    ub = A.BasicMultiUpdateBlock(args, hidden_dims=[128, 128, 128])
    assert ub.call_replay is None and not um.call_replay_enabled(ub)
    assert A.set_call_replay(True) is False
    try:
        assert um.call_replay_enabled(ub)
        ub.call_replay = False
        assert not um.call_replay_enabled(ub)
        ub.call_replay = True
        with um.call_replay_suspended():
            assert not um.call_replay_enabled(ub)
        assert um.call_replay_enabled(ub)
    finally:
        assert A.set_call_replay(False) is True
    assert um.forward_replayed(ub, [], [], None, None) is None
    ours = A.hotpath.adopt_update_block(ub, "igev", replay=True)
    assert ours.call_replay is True and ours.gru04.convz.weight is ub.gru04.convz.weight


def test_default_engine_is_tensor_core_parity(A):
    """rebinding the names alone must select the tcgen05 parity engine, not the CUDA-core one."""
    import subprocess
    import sys as _sys
    out = subprocess.run([_sys.executable, "-c",
                          "import anystereo_b200 as A; print(A.get_update_engine(), A.get_corr_mode())"],
                         capture_output=True, text=True, cwd=ROOT)
    assert out.stdout.split() == ["f16f8", "bf16x3"], (out.stdout, out.stderr[-500:])
    assert A.update.DEFAULT_ENGINE in ("f16f8", "bf16x3") and A.geometry.DEFAULT_CORR_MODE == "bf16x3"


def test_reference_install_is_pristine():
    """baseline/_ref (git-ignored; travels to the GPU box) is a byte-identical copy of the reference tree."""
    import filecmp
    from oracle import install_ref
    if not os.path.isdir(os.path.join(install_ref.SRC, "models")):
        pytest.skip("no /root/reference in this container")
    assert install_ref.install(verbose=False)
    for rel in install_ref._files(install_ref.SRC):
        assert filecmp.cmp(os.path.join(install_ref.SRC, rel), os.path.join(install_ref.DST, rel), shallow=False), rel
    tracked = subprocess_out(["git", "ls-files", "baseline"])
    assert tracked.strip() == "", "baseline/_ref must never be committed"


def subprocess_out(cmd):
    import subprocess
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout


def test_shard_pairs(A):
    for n in (0, 1, 7, 8, 64):
        for ws in (1, 2, 3, 8):
            got = [A.shard_pairs(n, r, ws) for r in range(ws)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        A.shard_pairs(4, 2, 2)


def test_c_sampler_oracle_against_golden(golden):
    """The plain-C restatement of sampler_kernel.cu reproduces the reference's Python lookup (level 0)."""
    from oracle import sampler_c
    g = golden("raft_corrblock")
    c = cases.raft_corr_case()
    B, _, H, W = c["f1"].shape
    vol = g["corr"].reshape(B, H, W, W)
    for dist in ("uniform", "smooth", "integer", "negative", "far_oob", "edge"):
        x0 = (O.pixel_coords(B, H, W).reshape(B, 1, H, W) - c["disps"][dist]).numpy()
        out = sampler_c.forward(vol, x0, 4)
        ref = g["lookup_" + dist][:, :9]
        denom = max(np.abs(ref).max(), 1e-30)
        assert np.abs(out - ref).max() / denom < 1e-4, dist
        # and the torch restatement agrees with the C one to rounding
        t = O.sampler_forward(torch.from_numpy(vol), torch.from_numpy(x0), 4).numpy()
        assert np.abs(out - t).max() <= 2e-6 * max(1.0, np.abs(t).max())
    gr = np.random.RandomState(0).standard_normal((B, 9, H, W)).astype("float32")
    bw = sampler_c.backward(vol.shape, x0, gr, 4)
    tb = O.sampler_backward(torch.from_numpy(vol), torch.from_numpy(x0), torch.from_numpy(gr), 4).numpy()
    assert np.abs(bw - tb).max() <= 2e-6 * max(1.0, np.abs(tb).max())


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (no CPU fallback behind the API)."""
    pkg = os.path.join(ROOT, "any-stereo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f


def test_no_predicated_tensor_core_mma_in_sass():
    """Lint against a ptxas code-generation hazard hit in this repo: with an `if (split) {3 MMAs} else {1 MMA}` shape
    around inline-asm tcgen05.mma, the non-split path of one kernel was emitted as `@UPn UTCHMMA` under a stale uniform
    predicate and silently skipped a K-step.  Every tcgen05 MMA in the library must be unpredicated."""
    import re
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "any-stereo_b200", "csrc", "libanystereo_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    mma = [l for l in sass.splitlines() if "UTCHMMA" in l]
    assert len(mma) > 50
    bad = [l.strip() for l in mma if re.search(r"@!?UP\d+\s+UTCHMMA", l)]
    assert not bad, bad[:3]


def test_round2_entry_points_reject_bad_arguments(A):
    """The entry points added in round 2 validate before any CUDA call (no GPU here)."""
    lib = A._lib.lib()
    assert lib.as_geo_lookup_convc1_tap(None, 8, 48, None, None, None, 2, None, None, None, None, None, 3, None, None,
                                        1, 4, 4, 4, None) == -1
    assert lib.as_gwc_corr_stem_fwd(None, None, None, None, None, None, None, 1, 96, 4, 4, 48, 8, 0.01, None) == -1
    assert lib.as_nearest_gather_fwd(None, None, None, 1, 4, 4, 128, 16, None) == -1
    assert lib.as_nearest_gather_bwd(None, None, None, 1, 4, 4, 128, 16, None) == -1
    assert lib.as_context_upsample_multiscale_bwd(None, None, None, None, None, None, 1, 4, 4, 16, None) == -1
    assert lib.as_liif_layer1_fwd(2, None, None, None, None, None, None, None, 1, 128, 16, None) == -1
    assert lib.as_liif_layer1_bwd(2, None, None, None, None, None, None, None, None, None, None, 1, 128, 16, None) == -1
    # as_instnorm_nhwc: null / sizes / workspace, channel count, alignment
    ws = lib.as_instnorm_workspace_bytes(2, 64)
    assert ws == 2 * 64 * 24 and lib.as_instnorm_workspace_bytes(0, 64) == 0
    assert lib.as_instnorm_nhwc(None, None, 16, 16, ws, 2, 100, 64, 1e-5, 1, None) == -1
    assert lib.as_instnorm_nhwc(16, None, 16, 16, ws - 1, 2, 100, 64, 1e-5, 1, None) == -1
    assert lib.as_instnorm_nhwc(16, None, 16, 16, ws, 2, 100, 64, 0.0, 1, None) == -1
    assert lib.as_instnorm_nhwc(16, None, 16, 16, ws, 2, 100, 66, 1e-5, 1, None) == -2
    assert lib.as_instnorm_nhwc(16, 8, 16, 16, ws, 2, 100, 64, 1e-5, 1, None) == -4


def test_deferred_wrappers_host_logic(A):
    """Host side of the round-2 fusions: the tap-major K order of the fused lookup's weights, and what the deferring
    build_gwc_volume / corr_stem pair does when the fused kernel does not apply (CPU tensors, gradients, train-mode BN)."""
    import torch.nn as nn
    # K order: channel (level l, group g, tap k) -> l*96 + k*8 + g; correlation taps -> l*96 + 72 + k; bias -> column 81
    c = torch.arange(162)
    g, k = (c % 81) // 9, c % 9
    kidx = (c // 81) * 96 + torch.where(g < 8, k * 8 + g, 72 + k)
    assert len(set(kidx.tolist())) == 162 and 81 not in set(kidx.tolist()) and int(kidx.max()) < 192
    assert int(kidx[0]) == 0 and int(kidx[9]) == 1 and int(kidx[72]) == 72 and int(kidx[81]) == 96
    # build_gwc_volume_deferred: CPU tensors / gradients / unsupported shapes take the plain operator (which then raises
    # its own "must be a CUDA tensor" here), CUDA-eligible inputs would be deferred
    from anystereo_b200.submodule import build_gwc_volume_deferred
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        build_gwc_volume_deferred(torch.zeros(1, 96, 2, 8), torch.zeros(1, 96, 2, 8), 48, 8)
    # CorrStem: same submodules (state_dict keys), plain tensors take the reference's arithmetic on any device
    class Ref(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv3d(8, 8, 3, 1, 1, bias=False)
            self.bn = nn.BatchNorm3d(8)
            self.relu, self.use_bn = True, True
    ref = Ref().eval()
    ours = A.hotpath.CorrStem(ref)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(1, 8, 4, 3, 5)
    want = nn.functional.leaky_relu(ref.bn(ref.conv(x)))
    assert torch.allclose(ours(x), want)
    with torch.no_grad():
        assert ours._fusable()
        ours.train()
        assert not ours._fusable()                   # batch statistics: never the fused eval-mode kernel
        ours.eval()
    with torch.enable_grad():
        assert not ours._fusable()                   # parameters want gradients
    # the deferring cost-volume classes only change __call__
    assert A.geometry.Combined_Geo_Encoding_Volume_Deferred.__mro__[1] is A.Combined_Geo_Encoding_Volume
    assert A.geometry.CorrBlock1D_Deferred.__mro__[1] is A.CorrBlock1D


def test_adopted_context_encoder_matches_reference_module(A):
    """adopt_context_encoder folds the eval-mode BatchNorm2d layers of the reference's MultiBasicEncoder
    (models/*/extractor.py:200-300) into its convolutions.  Same outputs as the reference module (fp32, CPU), same
    state_dict keys, reference arithmetic when a gradient is requested or a BatchNorm is in training mode."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree absent")
    ref_loader.load_models()
    from models.coreContinuous_IGEV.extractor import MultiBasicEncoder
    torch.manual_seed(0)
    ref = MultiBasicEncoder(output_dim=[[128] * 3, [128] * 3], norm_fn="batch", downsample=2)
    for m in ref.modules():                       # non-trivial running statistics and affine parameters
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.3)
            m.running_var.uniform_(0.5, 2.0)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    ref.eval()
    x = torch.randn(2, 3, 32, 64)
    with torch.no_grad():
        want = ref(x, num_layers=3)
    ours = A.adopt_context_encoder(ref)
    assert A.adopt_context_encoder(ours) is ours
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    with torch.no_grad():
        got = ours(x, num_layers=3)
        assert ours.__dict__["_fold_cache"] is not None            # the folded path ran
        for lw, lg in zip(want, got):
            for w, g in zip(lw, lg):
                assert g.shape == w.shape and g.is_contiguous()
                assert float((g - w).abs().max()) <= 2e-5 * float(w.abs().max())
        for n in (1, 2):
            w_, g_ = ref(x, num_layers=n), ours(x, num_layers=n)
            assert len(w_) == len(g_) == n
            assert float((g_[-1][1] - w_[-1][1]).abs().max()) <= 2e-5 * float(w_[-1][1].abs().max())
        wd, gd = ref(x, dual_inp=True, num_layers=3), ours(x, dual_inp=True, num_layers=3)
        assert gd[0][0].shape[0] == 1 and float((gd[3] - wd[3]).abs().max()) <= 2e-5 * float(wd[3].abs().max())
        assert float((gd[0][0] - wd[0][0]).abs().max()) <= 2e-5 * float(wd[0][0].abs().max())
        # under autocast (the reference wraps the call in autocast(enabled=args.mixed_precision)) the folded path stays fp32
        with torch.autocast("cpu", dtype=torch.bfloat16):
            ga = ours(x, num_layers=1)
        assert ga[0][0].dtype == torch.float32 and torch.equal(ga[0][0], got[0][0])
        # a parameter update refolds
        ref.conv1.weight.mul_(1.5)
        w2, g2 = ref(x), ours(x)
        assert float((g2[0][0] - w2[0][0]).abs().max()) <= 2e-5 * float(w2[0][0].abs().max())
    # gradients requested -> the reference's own forward (autograd graph intact)
    out = ours(x, num_layers=1)
    assert out[0][0].requires_grad
    # a BatchNorm in training mode -> reference arithmetic (batch statistics), bit-identical to the reference module
    ours.train()
    torch.manual_seed(1)
    with torch.no_grad():
        a = ours(x, num_layers=1)[0][0]
    with pytest.raises(TypeError):
        A.adopt_context_encoder(torch.nn.Conv2d(3, 3, 1))
    assert a.shape == want[0][0].shape


def test_adopted_feature_encoder_host_logic(A):
    """SURVEY 8(f)-4, RAFT: adopt_feature_encoder wraps the reference's BasicEncoder(norm_fn='instance') around the same
    submodules; without CUDA tensors (or with gradients / other norms) it is the reference module's own forward."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree absent")
    ref_loader.load_models()
    from models.corePrune_RAFT.extractor import BasicEncoder, MultiBasicEncoder
    torch.manual_seed(0)
    ref = BasicEncoder(output_dim=256, norm_fn="instance", downsample=2).eval()
    ours = A.adopt_feature_encoder(ref)
    assert A.adopt_feature_encoder(ours) is ours
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(1, 3, 32, 64)
    assert not ours._fusable([x])                                   # CPU tensor: no kernel, no fallback arithmetic of ours
    with torch.no_grad():
        a, b = ours([x, x + 1])
        c, d = ref([x, x + 1])
    assert torch.equal(a, c) and torch.equal(b, d)
    f = ours._weights()
    assert len(f["layers"]) == 3 and f["layers"][0][0][2] is None and f["layers"][1][0][2] is not None
    assert ours._weights() is f
    with torch.no_grad():
        ref.conv2.weight.mul_(2.0)
    assert ours._weights() is not f                                 # parameter update: repacked
    with pytest.raises(TypeError):
        A.adopt_feature_encoder(MultiBasicEncoder(output_dim=[[128] * 3], norm_fn="batch", downsample=2))
    with pytest.raises(TypeError):
        A.adopt_context_encoder(ref)
    assert not A.adopt_feature_encoder(BasicEncoder(output_dim=64, norm_fn="batch", downsample=2))._fusable([x])


def test_fold_basic_convs_matches_reference_hourglass(A):
    """fold_basic_convs swaps every reference BasicConv (Conv / ConvTranspose 2-D or 3-D + BatchNorm +
    LeakyReLU, submodule.py:6-32) for a wrapper with the eval-mode BatchNorm folded into the convolution.  Checked on the
    reference's 3-D hourglass (continuous_IGEVstereo.py:22-89: strided Conv3d, ConvTranspose3d, 1x1x1, FeatureAtt's 2-D
    blocks, a block without BatchNorm / activation): same output, same state_dict keys, originals restored by unfold."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree absent")
    R = ref_loader.load_models()
    torch.manual_seed(0)
    hg = R.igev_module.hourglass(8)
    for m in hg.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.running_mean.normal_(0, 0.3)
            m.running_var.uniform_(0.5, 2.0)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    hg.eval()
    x = torch.randn(1, 8, 8, 16, 24)
    feats = [None, torch.randn(1, 64, 8, 12), torch.randn(1, 192, 4, 6), torch.randn(1, 160, 2, 3)]
    with torch.no_grad():
        want = hg(x, feats)
    keys = list(hg.state_dict().keys())
    n_blocks = sum(1 for m in hg.modules() if type(m).__name__ == "BasicConv")
    handles = A.fold_basic_convs(hg)
    assert len(handles) == n_blocks >= 20
    assert list(hg.state_dict().keys()) == keys
    assert not any(type(m).__name__ == "BasicConv" for m in hg.modules())
    with torch.no_grad():
        got = hg(x, feats)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 5e-5 * float(want.abs().max())
    # gradients requested / training-mode BatchNorm: the reference block's own forward
    blk = hg.conv1[0]
    assert isinstance(blk, A.FoldedBasicConv) and blk(torch.randn(1, 8, 8, 16, 24)).requires_grad
    A.unfold_basic_convs(handles)
    assert sum(1 for m in hg.modules() if type(m).__name__ == "BasicConv") == n_blocks
    with torch.no_grad():
        assert torch.equal(hg(x, feats), want)


@pytest.mark.parametrize("family", ["igev", "raft"])
def test_adopt_model_swaps_everything_in_one_call(A, family):
    """adopt_model = install_into_reference + adopt_update_block + adopt_liif_up + the encoders (+ adopt_corr_stem) on a real
    reference model instance (CPU here: only the wiring; tests/test_gpu_dropin.py runs it)."""
    import contextlib
    import io
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree absent")
    R = ref_loader.load_models()
    mod = R.igev_module if family == "igev" else R.raft_module
    names = ["Combined_Geo_Encoding_Volume", "build_gwc_volume", "context_upsample_multiscale_train", "CorrBlock1D"]
    saved = {n: getattr(mod, n) for n in names if hasattr(mod, n)}
    try:
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model = (R.IGEV if family == "igev" else R.RAFT)(ref_loader.model_args(family)).eval()
        keys = list(model.state_dict().keys())
        out = A.adopt_model(model, mod, family, replay=True)
        assert out is model
        assert isinstance(model.update_block, A.BasicMultiUpdateBlock) and model.update_block.call_replay is True
        assert isinstance(model.liif_up, A.liif_out_multi_scale_Training)
        assert isinstance(model.cnet, A.ContextEncoder)
        if family == "igev":
            assert mod.Combined_Geo_Encoding_Volume is A.geometry.Combined_Geo_Encoding_Volume_Deferred
            assert isinstance(model.corr_stem, A.hotpath.CorrStem) and isinstance(model.corr_feature_att, A.hotpath.CorrFeatureAtt)
            assert mod.build_gwc_volume is A.submodule.build_gwc_volume_deferred
        else:
            assert mod.CorrBlock1D is A.geometry.CorrBlock1D_Deferred
            assert isinstance(model.fnet, A.FeatureEncoder)
        assert mod.context_upsample_multiscale_train is A.liif.context_upsample_multiscale_train
        assert list(model.state_dict().keys()) == keys            # a reference checkpoint still loads unchanged
        with pytest.raises(ValueError):
            A.adopt_model(model, mod, "psm")
    finally:
        for n, v in saved.items():
            setattr(mod, n, v)
