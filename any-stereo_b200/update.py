"""Per-iteration update block -- drop-in for models/*/update.py (BasicMultiUpdateBlock :104-136).

The module tree and parameter names are the reference's (``encoder.convc1.weight`` ...
``gru04.convz.weight`` ... ``disp_head.conv2.bias``), so reference checkpoints load unchanged
(``strict=True``); the forward pass runs entirely on the library's kernels:

* activations are pixel-major ("NHWC") so every convolution is an implicit GEMM with M = pixels,
  N = Cout, K = taps x Cin, reading up to 4 sources concatenated along K -- ``torch.cat`` of
  update.py:35-36,39,90 never materialises;
* convz and convr share one N = 2*hidden GEMM; sigmoid, the ``r*h`` product, tanh and the
  ``(1-z)h + zq`` blend are the GEMM epilogues (update.py:37-40);
* hidden states returned in ``net`` are channels-last views ([B,128,h,w] shape, NHWC memory): the next
  call consumes them without any layout change.

``engine``: "fp32" = exact-fp32 CUDA-core kernels (parity baseline);
"bf16x3" / "bf16" = tcgen05 tensor-core kernels (split-bf16 fp32-parity mode / fast mode).
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from . import _lib as L

#: default engine = the 2-pass fp32-parity tensor-core path ("f16f8": IEEE-half hi*hi + one e5m2 pass for both cross terms,
#: fp32 accumulation in TMEM; training runs its kernels in the 3-pass "bf16x3" split).  "fp32" selects the exact CUDA-core
#: kernels (the on-GPU parity baseline, ~30x slower); "fp16" / "bf16" are the single-pass fast modes.
DEFAULT_ENGINE = "f16f8"
_ENGINE = {"engine": DEFAULT_ENGINE}


ENGINES = ("fp32", "bf16x3", "f16f8", "bf16", "fp16")


def set_update_engine(engine: str):
    """fp32   exact CUDA-core kernels (parity baseline on the GPU)
    bf16x3 tcgen05, 3 passes: x = hi + lo in bf16, hi*hi + hi*lo + lo*hi -- fp32 parity (default)
    f16f8  tcgen05, 2 pass-equivalents: hi*hi in IEEE half + BOTH cross terms in one e5m2 pass (csrc/common.cuh) -- fp32
           parity class (final-disparity drift 1e-4 px like bf16x3), 1.5x fewer tensor-core cycles
    fp16 / bf16  single pass -- fast modes, outside the 1e-4 operator tolerance (fp16 inside the 0.01 px EPE gate)"""
    if engine not in ENGINES:
        raise ValueError("engine must be one of %s" % (ENGINES,))
    _ENGINE["engine"] = engine
    sync_operand_format()


def sync_operand_format():
    """Make the library's process-wide operand format the one the current engine uses ("fp16": IEEE half, the analogue of
    the reference's autocast mixed precision, continuous_IGEVstereo.py:287; "f16f8": half + e5m2 pair planes; else bf16).
    Called by set_update_engine and at the top of every forward (the import-time default engine has not touched the
    CUDA library yet: importing works without a GPU)."""
    L.set_operand_format({"fp16": L.FMT_F16, "f16f8": L.FMT_F16F8}.get(_ENGINE["engine"], L.FMT_BF16))


def get_update_engine() -> str:
    return _ENGINE["engine"]


# ---- parameter containers with the reference's names/shapes ---------------------------------------

class DispHead(nn.Module):  # update.py:16-24
    def __init__(self, input_dim=128, hidden_dim=256, output_dim=1):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, output_dim, 3, padding=1)


class ConvGRU(nn.Module):  # update.py:26-41
    def __init__(self, hidden_dim, input_dim, kernel_size=3):
        super().__init__()
        self.convz = nn.Conv2d(hidden_dim + input_dim, hidden_dim, kernel_size, padding=kernel_size // 2)
        self.convr = nn.Conv2d(hidden_dim + input_dim, hidden_dim, kernel_size, padding=kernel_size // 2)
        self.convq = nn.Conv2d(hidden_dim + input_dim, hidden_dim, kernel_size, padding=kernel_size // 2)


class BasicMotionEncoder(nn.Module):  # update.py:73-92
    def __init__(self, args, geo_groups=8):
        super().__init__()
        self.args = args
        # IGEV: L*(2r+1)*(8+1) (coreContinuous_IGEV/update.py:77); RAFT: L*(2r+1) (corePrune_RAFT/update.py:77)
        cor_planes = args.corr_levels * (2 * args.corr_radius + 1) * (geo_groups + 1)
        self.convc1 = nn.Conv2d(cor_planes, 64, 1, padding=0)
        self.convc2 = nn.Conv2d(64, 64, 3, padding=1)
        self.convd1 = nn.Conv2d(1, 64, 7, padding=3)
        self.convd2 = nn.Conv2d(64, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 64, 128 - 1, 3, padding=1)


# ---- helpers --------------------------------------------------------------------------------------

def _nhwc_view(t: torch.Tensor):
    """[B,C,H,W] tensor -> (pixel-major [B,H,W,C] contiguous tensor, was_copy)."""
    p = t.permute(0, 2, 3, 1)
    if p.is_contiguous():
        return p, False
    B, C, H, W = t.shape
    t = t.contiguous()
    out = torch.empty((B, H, W, C), device=t.device, dtype=torch.float32)
    L.call("as_nchw_to_nhwc", t.data_ptr(), out.data_ptr(), B, C, H, W, C, 0, L.stream_ptr())
    return out, True


class _PackedConv:
    """GEMM-ready weights of one (possibly fused) convolution, cached per parameter version."""

    def __init__(self):
        self.key = None
        self.w = None
        self.b = None

    def get(self, convs):
        key = tuple((c.weight.data_ptr(), L.version_of(c.weight), c.bias.data_ptr(), L.version_of(c.bias)) for c in convs)
        if key != self.key:
            with torch.no_grad():
                w = torch.cat([c.weight.detach().float() for c in convs], dim=0).contiguous()
                b = torch.cat([c.bias.detach().float() for c in convs], dim=0).contiguous()
                Cout, Cin, KH, KW = w.shape
                packed = torch.empty((KH * KW * Cin, Cout), device=w.device, dtype=torch.float32)
                L.call("as_pack_conv_weight", w.data_ptr(), packed.data_ptr(), Cout, Cin, KH, KW, L.stream_ptr())
            self.key, self.w, self.b = key, packed, b
            self.shape = (Cout, Cin, KH, KW)
        return self.w, self.b, self.shape


def _conv(B, H, W, srcs, pk, epilogue, out, out_pitch, out_coff=0, out_layout=L.LAYOUT_NHWC,
          ctx=None, ctx_pitch=0, h=None, z=None):
    """srcs: list of (tensor, channels, pitch, layout)."""
    w, b, (Cout, Cin, KH, KW) = pk
    assert sum(s[1] for s in srcs) == Cin, (Cin, [s[1] for s in srcs])
    d = L.ConvDesc()
    d.B, d.H, d.W, d.KH, d.KW, d.Cout = B, H, W, KH, KW, Cout
    d.num_src = len(srcs)
    for i, (t, ch, pitch, layout) in enumerate(srcs):
        d.src[i].ptr = t.data_ptr()
        d.src[i].channels = ch
        d.src[i].pitch = pitch
        d.src[i].layout = layout
    d.weight = w.data_ptr()
    d.bias = b.data_ptr()
    d.epilogue = epilogue
    d.out = out.data_ptr()
    d.out_pitch, d.out_coff, d.out_layout = out_pitch, out_coff, out_layout
    d.ctx = L.ptr(ctx)
    d.ctx_pitch = ctx_pitch
    d.h = L.ptr(h)
    d.z = L.ptr(z)
    L.call("as_conv2d_fp32", d, L.stream_ptr())


class BasicMultiUpdateBlock(nn.Module):
    """reference: models/*/update.py:104-136; call sites continuous_IGEVstereo.py:287-293,
    prune_raft_stereo.py:279-284."""

    #: 8 geometry groups feed the motion encoder in the IGEV family, 0 in the RAFT family
    GEO_GROUPS = 8
    #: "f16f8" engine: the 1/4-resolution gate convolutions keep only the weight-residual cross term (update_umma._GATE_WL)
    gate_weight_residual_only = True
    #: replay every inference call from a CUDA graph: True / False / None = update_umma.set_call_replay's setting
    call_replay = None

    def __init__(self, args, hidden_dims=[]):
        super().__init__()
        self.args = args
        self.encoder = BasicMotionEncoder(args, self.GEO_GROUPS)
        encoder_output_dim = 128
        self.gru04 = ConvGRU(hidden_dims[2], encoder_output_dim + hidden_dims[1] * (args.n_gru_layers > 1))
        self.gru08 = ConvGRU(hidden_dims[1], hidden_dims[0] * (args.n_gru_layers == 3) + hidden_dims[2])
        self.gru16 = ConvGRU(hidden_dims[0], hidden_dims[1])
        self.disp_head = DispHead(hidden_dims[2], hidden_dim=256, output_dim=1)
        self._packed = {}
        self._ctx_cache = {}

    # -- caches ---------------------------------------------------------------------------------
    def reset_caches(self):
        """Forget the loop-invariant context conversions and cached hi/lo planes (weights stay packed)."""
        self._ctx_cache.clear()
        st = self.__dict__.get("_umma_state")
        if st is not None:
            st["ctx"].clear()
            st["planes"].clear()

    def invalidate_weights(self):
        """Drop every packed / split copy of the parameters.  The caches key on (data_ptr, version counter); writes
        that bypass the counter (``param.data.copy_``, EMA / clipping through ``.data``, in-place edits of inference
        tensors) must be followed by this call."""
        self._packed.clear()
        self.reset_caches()
        st = self.__dict__.get("_umma_state")
        if st is not None:
            st["w"].clear()
            st["calls"].clear()

    def _pk(self, name, convs):
        if name not in self._packed:
            self._packed[name] = _PackedConv()
        return self._packed[name].get(convs)

    def _context(self, idx, inp_i):
        """(cz|cr) and cq of one scale as pixel-major tensors; loop-invariant, cached per tensor version."""
        cz, cr, cq = inp_i
        # identity (weak refs) + version: a recycled allocation with new contents must not hit the cache
        key = tuple((t.data_ptr(), L.version_of(t), tuple(t.shape)) for t in (cz, cr, cq))
        hit = self._ctx_cache.get(idx)
        if hit is not None and hit[0] == key and all(r() is t for r, t in zip(hit[3], (cz, cr, cq))):
            return hit[1], hit[2]
        B, C, H, W = cz.shape
        st = L.stream_ptr()
        zr = torch.empty((B, H, W, 2 * C), device=cz.device, dtype=torch.float32)
        q = torch.empty((B, H, W, C), device=cz.device, dtype=torch.float32)
        for t, dst, pitch, off in ((cz, zr, 2 * C, 0), (cr, zr, 2 * C, C), (cq, q, C, 0)):
            t = t.detach().float().contiguous()
            L.call("as_nchw_to_nhwc", t.data_ptr(), dst.data_ptr(), B, C, H, W, pitch, off, st)
        self._ctx_cache[idx] = (key, zr, q, tuple(weakref.ref(t) for t in (cz, cr, cq)))
        return zr, q

    # -- pieces ---------------------------------------------------------------------------------
    def _gru(self, name, gru, idx, h, inp_i, xs):
        """ConvGRU.forward (update.py:33-41) on pixel-major tensors.  h [B,H,W,Hd]; xs list of [B,H,W,C]."""
        B, H, W, Hd = h.shape
        ctx_zr, ctx_q = self._context(idx, inp_i)
        srcs_x = [(x, x.shape[3], x.shape[3], L.LAYOUT_NHWC) for x in xs]
        z = torch.empty_like(h)
        rh = torch.empty_like(h)
        _conv(B, H, W, [(h, Hd, Hd, L.LAYOUT_NHWC)] + srcs_x, self._pk(name + ".zr", [gru.convz, gru.convr]),
              L.EPI_GRU_ZR, rh, Hd, ctx=ctx_zr, ctx_pitch=2 * Hd, h=h, z=z)
        hn = torch.empty_like(h)
        _conv(B, H, W, [(rh, Hd, Hd, L.LAYOUT_NHWC)] + srcs_x, self._pk(name + ".q", [gru.convq]),
              L.EPI_GRU_Q, hn, Hd, ctx=ctx_q, ctx_pitch=Hd, h=h, z=z)
        return hn

    def _pool2x(self, x):
        B, H, W, C = x.shape
        out = torch.empty((B, (H + 1) // 2, (W + 1) // 2, C), device=x.device, dtype=torch.float32)
        L.call("as_pool2x_nhwc", x.data_ptr(), out.data_ptr(), B, H, W, C, L.stream_ptr())
        return out

    def _interp(self, x, ref):
        B, H, W, C = x.shape
        Ho, Wo = ref.shape[1], ref.shape[2]
        out = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.float32)
        L.call("as_interp_bilinear_nhwc", x.data_ptr(), out.data_ptr(), B, H, W, Ho, Wo, C, L.stream_ptr())
        return out

    def _encoder(self, disp, corr):
        """BasicMotionEncoder.forward (update.py:84-92) -> pixel-major [B,H,W,128]."""
        e = self.encoder
        B, Cc, H, W = corr.shape
        dev = corr.device
        st = L.stream_ptr()
        c1 = torch.empty((B, H, W, 64), device=dev, dtype=torch.float32)
        _conv(B, H, W, [(corr, Cc, 0, L.LAYOUT_NCHW)], self._pk("convc1", [e.convc1]), L.EPI_BIAS_RELU, c1, 64)
        cd = torch.empty((B, H, W, 128), device=dev, dtype=torch.float32)     # cat(cor, disp_) never copied
        _conv(B, H, W, [(c1, 64, 64, L.LAYOUT_NHWC)], self._pk("convc2", [e.convc2]), L.EPI_BIAS_RELU, cd, 128, 0)
        d1 = torch.empty((B, H, W, 64), device=dev, dtype=torch.float32)
        _conv(B, H, W, [(disp, 1, 1, L.LAYOUT_NHWC)], self._pk("convd1", [e.convd1]), L.EPI_BIAS_RELU, d1, 64)
        _conv(B, H, W, [(d1, 64, 64, L.LAYOUT_NHWC)], self._pk("convd2", [e.convd2]), L.EPI_BIAS_RELU, cd, 128, 64)
        mo = torch.empty((B, H, W, 128), device=dev, dtype=torch.float32)     # cat(out, disp)
        _conv(B, H, W, [(cd, 128, 128, L.LAYOUT_NHWC)], self._pk("conv", [e.conv]), L.EPI_BIAS_RELU, mo, 128, 0)
        L.call("as_nchw_to_nhwc", disp.data_ptr(), mo.data_ptr(), B, 1, H, W, 128, 127, st)
        return mo

    def _disp_head(self, h):
        B, H, W, Hd = h.shape
        dh = self.disp_head
        t = torch.empty((B, H, W, 256), device=h.device, dtype=torch.float32)
        _conv(B, H, W, [(h, Hd, Hd, L.LAYOUT_NHWC)], self._pk("dh1", [dh.conv1]), L.EPI_BIAS_RELU, t, 256)
        delta = torch.empty((B, 1, H, W), device=h.device, dtype=torch.float32)
        _conv(B, H, W, [(t, 256, 256, L.LAYOUT_NHWC)], self._pk("dh2", [dh.conv2]), L.EPI_BIAS, delta, 1, 0,
              out_layout=L.LAYOUT_NCHW)
        return delta

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, net, inp, corr=None, disp=None, iter04=True, iter08=True, iter16=True, update=True):
        deferred = corr is not None and not torch.is_tensor(corr)     # geometry.DeferredGeoLookup
        if deferred and (get_update_engine() == "fp32" or torch.is_grad_enabled()):
            corr, deferred = corr.materialize(), False
        needs_grad = torch.is_grad_enabled() and (
            any(p.requires_grad for p in self.parameters()) or any(t.requires_grad for t in net)
            or any(t.requires_grad for lst in inp for t in lst) or (corr is not None and corr.requires_grad))
        if needs_grad:
            # training (config 5): differentiable path on the exact-fp32 kernels, explicit adjoints
            from . import update_train
            return update_train.forward(self, net, inp, corr, disp, iter04, iter08, iter16, update)
        if get_update_engine() != "fp32":
            from . import update_umma
            sync_operand_format()
            if iter04 and update and update_umma.call_replay_enabled(self):
                # the reference's own loop, one call per iteration: replay the call from a CUDA graph (opt-in)
                out = update_umma.forward_replayed(self, net, inp, corr, disp, iter08, iter16)
                if out is not None:
                    return out
            return update_umma.forward(self, net, inp, corr, disp, iter04, iter08, iter16, update)
        for t in net:
            L.require_cuda(t, "net[i]", contiguous=False)
        n_layers = self.args.n_gru_layers
        with torch.cuda.device(net[0].device):
            # half inputs (the reference runs this block under autocast when mixed_precision is on) are widened
            hs = [None if t is None else _nhwc_view(t.detach().float())[0] for t in net]
            if iter16:
                hs[2] = self._gru("gru16", self.gru16, 2, hs[2], inp[2], [self._pool2x(hs[1])])
            if iter08:
                xs = [self._pool2x(hs[0])]
                if n_layers > 2:
                    xs.append(self._interp(hs[2], hs[1]))
                hs[1] = self._gru("gru08", self.gru08, 1, hs[1], inp[1], xs)
            if iter04:
                L.require_cuda(corr, "corr", contiguous=False)
                L.require_cuda(disp, "disp", contiguous=False)
                corr = corr.detach().float().contiguous()
                disp = disp.detach().float().contiguous()
                xs = [self._encoder(disp, corr)]
                if n_layers > 1:
                    xs.append(self._interp(hs[1], hs[0]))
                hs[0] = self._gru("gru04", self.gru04, 0, hs[0], inp[0], xs)
            for i in range(len(net)):
                if hs[i] is not None:
                    net[i] = hs[i].permute(0, 3, 1, 2)     # [B,C,H,W] shape, channels-last memory
            if not update:
                return net
            delta_disp = self._disp_head(hs[0])
        return net, delta_disp


class BasicMultiUpdateBlockRAFT(BasicMultiUpdateBlock):
    """The corePrune_RAFT flavour (models/corePrune_RAFT/update.py:77: cor_planes = L*(2r+1))."""
    GEO_GROUPS = 0
    gate_weight_residual_only = False         # RAFT's EPE sits closer to the 1e-3 px bar: both cross terms everywhere
