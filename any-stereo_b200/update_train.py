"""Differentiable execution of BasicMultiUpdateBlock.forward (training, BASELINE.json config 5).

The reference obtains these gradients from autograd over models/*/update.py:16-136; here the adjoint is written
out explicitly as one ``torch.autograd.Function`` per update-block call:

  conv data gradient   = the forward implicit-GEMM kernel on dY with flipped/transposed weights
  conv weight gradient = as_conv2d_wgrad_fp32 (K = pixels, fp32 atomics)
  GRU gates            = as_gru_bwd_gates1/2 (update.py:37-40), ReLU masks, pool2x / bilinear adjoints

Engines (``set_update_engine``): "fp32" runs every convolution on the exact-fp32 CUDA-core kernel
(include/anystereo_b200.h, section a13-vi).  "bf16x3" / "bf16" / "fp16" run the forward convolutions and the data
gradients on the tcgen05 kernel (as_conv2d_umma with AS_UEPI_LINEAR_F32: raw fp32 output, then
as_conv_epilogue_fp32 applies the gate / ReLU epilogue and keeps r and q for the backward pass); convd1 (7x7, one
input channel), DispHead.conv2 (one output channel) and every weight gradient stay on the CUDA cores.

Gradients flow to: every parameter, the hidden states ``net``, the context terms ``inp``, and the lookup
features ``corr``.  ``disp`` receives none: both model forwards detach it every iteration
(continuous_IGEVstereo.py:285, prune_raft_stereo.py:277).
"""
from __future__ import annotations

import os

import torch

from . import _lib as L

NHWC, NCHW = L.LAYOUT_NHWC, L.LAYOUT_NCHW


def _s():
    return L.stream_ptr()


def _to_nhwc(t):
    """[B,C,H,W] (any strides) -> contiguous [B,H,W,C] fp32."""
    p = t.permute(0, 2, 3, 1)
    if p.is_contiguous():
        return p
    B, C, H, W = t.shape
    t = t.contiguous()
    out = torch.empty((B, H, W, C), device=t.device, dtype=torch.float32)
    L.call("as_nchw_to_nhwc", t.data_ptr(), out.data_ptr(), B, C, H, W, C, 0, _s())
    return out


def _to_nchw(x, C=None, pitch=None, coff=0):
    """pixel-major [B,H,W,pitch] slice -> [B,C,H,W] contiguous."""
    B, H, W, P = x.shape
    C = C or P
    out = torch.empty((B, C, H, W), device=x.device, dtype=torch.float32)
    L.call("as_nhwc_to_nchw", x.data_ptr(), out.data_ptr(), B, C, H, W, pitch or P, coff, _s())
    return out


_WGRAD_TC = {"on": os.environ.get("AS_WGRAD_TC", "1") != "0"}


def set_wgrad_tensor_cores(on: bool):
    """Weight gradients of the 1x1 / 3x3 convolutions on the tcgen05 kernel (as_conv2d_wgrad_umma) when the update engine
    is a tensor-core engine; off = CUDA-core as_conv2d_wgrad_fp32 (A/B knob, also AS_WGRAD_TC=0)."""
    _WGRAD_TC["on"] = bool(on)


# A/B knobs of the tensor-core training path (both on by default; AS_TRAIN_SMALL_TC=0 / AS_TRAIN_CONVD1=0 turn them off;
# measured on B200: config-5 step 293 -> 207 ms with both on, 153 GPU tests green):
#  small_tc: DispHead.conv2 (one output channel) forward / data gradient / weight gradient on the tcgen05 kernels
#            (rows / channels zero-padded to the kernels' granules) instead of the 64x64-tile CUDA-core kernels
#  convd1:   dedicated CUDA-core kernels for the 7x7 single-input-channel convd1 (forward and weight gradient)
_KNOBS = {"small_tc": os.environ.get("AS_TRAIN_SMALL_TC", "1") != "0",
          "convd1": os.environ.get("AS_TRAIN_CONVD1", "1") != "0"}


def _engine():
    """(tensor cores?, MMAs per K-step) of the current update engine."""
    from .update import get_update_engine
    e = get_update_engine()
    return e != "fp32", (3 if e == "bf16x3" else 1)


def _split(x, split):
    """hi (+lo) operand planes of a contiguous pixel-major fp32 tensor whose channel count is a multiple of 64."""
    from .update_umma import _Planes
    assert x.is_contiguous() and x.shape[3] % 64 == 0, tuple(x.shape)
    pl = _Planes(x.shape, x.device, split)
    L.call("as_split_f32", x.data_ptr(), pl.hi.data_ptr(), L.ptr(pl.lo), x.numel(), _s())
    return pl


def _dgrad_weights(ub, name, convs, split):
    """Data-gradient GEMM weights of a convolution: the [Cin][Cout][KH][KW] transpose with flipped taps, packed like
    forward weights (rows padded to 32, K channels padded to 64), cached per parameter version."""
    from .update_umma import _state
    st = _state(ub)["w"]
    key = (split, L.operand_format()) + tuple((c.weight.data_ptr(), L.version_of(c.weight)) for c in convs)
    hit = st.get("train.dgrad." + name)
    if hit is not None and hit["key"] == key:
        return hit
    with torch.no_grad():
        w = torch.cat([c.weight.detach().float() for c in convs], dim=0)
        wt = w.transpose(0, 1).flip(2, 3).contiguous()                    # [Cin][Cout][KH][KW]
        Cin, Cout, KH, KW = wt.shape
        n_pad, cin_pad = (Cin + 31) // 32 * 32, (Cout + 63) // 64 * 64
        hi = torch.empty((n_pad, KH * KW * cin_pad), device=w.device, dtype=torch.bfloat16)
        lo = torch.empty_like(hi) if split else None
        L.call("as_pack_conv_weight_bf16", wt.data_ptr(), hi.data_ptr(), L.ptr(lo), Cin, Cout, KH, KW, n_pad, cin_pad, _s())
    hit = dict(key=key, hi=hi, lo=lo, n=n_pad, cin=cin_pad, k=KH)
    st["train.dgrad." + name] = hit
    return hit


class _Conv:
    """One convolution (or z|r pair) with its packed forward / data-gradient weights."""

    def __init__(self, convs, ub=None, name=None):
        self.ub, self.name = ub, name
        self.tcache = None                     # per-backward cache of channel-major operand planes
        self.convs = convs
        self.Cout = sum(c.weight.shape[0] for c in convs)
        _, self.Cin, self.KH, self.KW = convs[0].weight.shape
        self.dev = convs[0].weight.device
        self._w = self._b = self._wf = self._wd = None     # fp32 copies / CUDA-core packings, made on first use only

    @property
    def w(self):
        if self._w is None:
            with torch.no_grad():
                self._w = torch.cat([c.weight.detach().float() for c in self.convs], 0).contiguous()
        return self._w

    @property
    def b(self):
        if self._b is None:
            with torch.no_grad():
                self._b = torch.cat([c.bias.detach().float() for c in self.convs], 0).contiguous()
        return self._b

    @property
    def wf(self):
        if self._wf is None:
            self._wf = torch.empty((self.KH * self.KW * self.Cin, self.Cout), device=self.dev, dtype=torch.float32)
            L.call("as_pack_conv_weight", self.w.data_ptr(), self._wf.data_ptr(), self.Cout, self.Cin, self.KH, self.KW, _s())
        return self._wf

    @property
    def wd(self):
        if self._wd is None:
            self._wd = torch.empty((self.KH * self.KW * self.Cout, self.Cin), device=self.dev, dtype=torch.float32)
            L.call("as_pack_conv_weight_dgrad", self.w.data_ptr(), self._wd.data_ptr(), self.Cout, self.Cin, self.KH, self.KW,
                   _s())
        return self._wd

    def _desc(self, B, H, W, srcs):
        d = L.ConvDesc()
        d.B, d.H, d.W, d.KH, d.KW = B, H, W, self.KH, self.KW
        d.num_src = len(srcs)
        for i, (t, ch, pitch, layout) in enumerate(srcs):
            d.src[i].ptr = t.data_ptr()
            d.src[i].channels = ch
            d.src[i].pitch = pitch
            d.src[i].layout = layout
        return d

    def fwd(self, B, H, W, srcs, epi, out, out_pitch, out_coff=0, out_layout=NHWC, ctx=None, ctx_pitch=0, h=None, z=None,
            save=None, bias=True):
        d = self._desc(B, H, W, srcs)
        d.Cout = self.Cout
        d.weight = self.wf.data_ptr()
        d.bias = self.b.data_ptr() if bias else None
        d.epilogue = epi
        d.out = out.data_ptr()
        d.out_pitch, d.out_coff, d.out_layout = out_pitch, out_coff, out_layout
        d.ctx, d.ctx_pitch, d.h, d.z, d.save = L.ptr(ctx), ctx_pitch, L.ptr(h), L.ptr(z), L.ptr(save)
        L.call("as_conv2d_fp32", d, _s())

    # ---- tensor-core variants (engines other than "fp32")
    def tc_ok(self):
        """The tcgen05 kernel covers 1x1 / 3x3 with at least 32 input channels; fewer than 32 output channels are
        zero-padded rows of the packed weights (DispHead.conv2)."""
        return (self.KH in (1, 3) and self.KH == self.KW and self.Cin >= 32 and self.Cout <= 256
                and (self.Cout >= 32 or _KNOBS["small_tc"]))

    def fwd_tc(self, B, H, W, planes, nsplit, epi, out, out_pitch, out_coff=0, ctx=None, ctx_pitch=0, h=None, z=None,
               save=None):
        """Same contract as fwd() with NHWC output; `planes` are the hi/lo operand planes of the sources."""
        from . import update_umma as U
        cin_pad = sum(pl.shape[3] for pl in planes)
        # fewer than 32 outputs (DispHead.conv2): zero rows up to N = 64, the narrowest GEMM the encoder convs exercise
        wt = U._weights(self.ub, "train." + self.name, self.convs, n_pad=None if self.Cout >= 32 else 64, cin_pad=cin_pad,
                        split=nsplit == 3)
        raw = torch.empty((B, H, W, wt["n"]), device=out.device, dtype=torch.float32)
        U._conv(B, H, W, planes, wt, nsplit, L.UEPI_LINEAR_F32, out_f32=raw)
        L.call("as_conv_epilogue_fp32", raw.data_ptr(), wt["n"], B * H * W, self.Cout, epi, L.ptr(ctx), ctx_pitch, L.ptr(h),
               L.ptr(z), L.ptr(save), out.data_ptr(), out_pitch, out_coff, _s())

    def dgrad_tc(self, B, H, W, dy, nsplit):
        """dX [B,H,W,roundup32(Cin)] from a contiguous dY [B,H,W,roundup64(Cout)] (padding channels zero)."""
        from . import update_umma as U
        wt = _dgrad_weights(self.ub, self.name, self.convs, nsplit == 3)
        assert dy.shape[3] == wt["cin"], (tuple(dy.shape), wt["cin"])
        dyS = _split(dy, nsplit == 3)
        n_pad = wt["n"]
        dx = torch.empty((B, H, W, n_pad), device=dy.device, dtype=torch.float32)
        r0 = 0
        while r0 < n_pad:                                   # GEMM N <= 256 per launch: chunk the Cin rows
            rows = min(256, n_pad - r0)
            chunk = dict(hi=wt["hi"][r0:r0 + rows], lo=None if wt["lo"] is None else wt["lo"][r0:r0 + rows], n=rows,
                         cin=wt["cin"], k=wt["k"])
            U._conv(B, H, W, [dyS], chunk, nsplit, L.UEPI_LINEAR_F32, bias=False, out_f32=dx, out_coff=r0,
                    f32_pitch=n_pad if rows != n_pad else 0)
            r0 += rows
        return dx

    def dgrad(self, B, H, W, dy, dy_pitch):
        """dX [B,H,W,Cin] from dY [.., Cout] (pixel-major, pitch dy_pitch)."""
        tc, nsplit = _engine()
        if tc and self.ub is not None and self.tc_ok() and dy.shape[3] == dy_pitch:
            cpad = (self.Cout + 63) // 64 * 64
            if dy_pitch == cpad:                                          # note: Cin comes back padded to 32
                return self.dgrad_tc(B, H, W, dy, nsplit)
            if dy_pitch == self.Cout and _KNOBS["small_tc"]:              # DispHead.conv2: pad dY to the K granule
                dyp = torch.zeros((B, H, W, cpad), device=dy.device, dtype=torch.float32)
                L.call("as_add_slice", dy.data_ptr(), dy_pitch, 0, dyp.data_ptr(), cpad, 0, B * H * W, self.Cout, _s())
                return self.dgrad_tc(B, H, W, dyp, nsplit)
        dx = torch.empty((B, H, W, self.Cin), device=dy.device, dtype=torch.float32)
        d = self._desc(B, H, W, [(dy, self.Cout, dy_pitch, NHWC)])
        d.Cout = self.Cin
        d.weight = self.wd.data_ptr()
        d.bias = None
        d.epilogue = L.EPI_BIAS
        d.out = dx.data_ptr()
        d.out_pitch, d.out_coff, d.out_layout = self.Cin, 0, NHWC
        L.call("as_conv2d_fp32", d, _s())
        return dx

    def _tplanes(self, t, ch, pitch, B, H, W, Wp, split, nshift=1):
        """Channel-major hi/lo planes [nshift][ch][B][H][Wp] of a pixel-major fp32 tensor, copy j shifted by j - nshift/2
        pixels along x (shared by the z|r and q weight gradients of a GRU through the per-backward cache)."""
        cache = self.tcache if self.tcache is not None else {}
        hit = cache.get((id(t), ch, nshift))
        if hit is None:
            hi = torch.empty((nshift, ch, B, H, Wp), device=t.device, dtype=torch.bfloat16)
            lo = torch.empty_like(hi) if split else None
            L.call("as_transpose_split", t.data_ptr(), pitch, 0, ch, B, H, W, hi.data_ptr(), L.ptr(lo), Wp, nshift, _s())
            hit = cache[(id(t), ch, nshift)] = (t, hi, lo)
        return hit[1], hit[2]

    def wgrad(self, B, H, W, srcs, dy, dy_pitch):
        """(dW [Cout,Cin,KH,KW], db [Cout]) for this call."""
        dw = torch.zeros((self.Cout, self.Cin, self.KH, self.KW), device=self.dev, dtype=torch.float32)
        db = torch.zeros((self.Cout,), device=self.dev, dtype=torch.float32)
        tc, nsplit = _engine()
        if tc and _KNOBS["convd1"] and self.KH == 7 and self.KW == 7 and self.Cin == 1 and self.Cout == 64 \
                and len(srcs) == 1 and srcs[0][2] == 1:
            L.call("as_convd1_wgrad_fp32", srcs[0][0].data_ptr(), dy.data_ptr(), dy_pitch, B, H, W, dw.data_ptr(), _s())
            L.call("as_bias_grad_fp32", dy.data_ptr(), dy_pitch, self.Cout, B * H * W, db.data_ptr(), _s())
            return dw, db
        if (tc and _WGRAD_TC["on"] and self.KH in (1, 3) and self.KH == self.KW and self.Cin >= 32
                and (self.Cout >= 32 or _KNOBS["small_tc"]) and len(srcs) <= 3 and all(s[3] == NHWC for s in srcs) and all(s[1] % 128 == 0 for s in srcs[:-1])
                and B * H <= 65535):
            # tensor cores: K = pixels GEMM over channel-major operand planes (as_conv2d_wgrad_umma)
            split = nsplit == 3
            Wp = (W + 7) // 8 * 8
            d = L.WgradUmmaDesc()
            d.B, d.H, d.W, d.KH, d.KW, d.Cout, d.num_src = B, H, W, self.KH, self.KW, self.Cout, len(srcs)
            keep = []
            for i, (t, ch, pitch, _) in enumerate(srcs):
                hi, lo = self._tplanes(t, ch, pitch, B, H, W, Wp, split, self.KW)
                d.src[i].hi, d.src[i].lo, d.src[i].channels = hi.data_ptr(), L.ptr(lo), ch
                keep += [hi, lo]
            dyh, dyl = self._tplanes(dy, self.Cout, dy_pitch, B, H, W, Wp, split)
            d.dy_hi, d.dy_lo, d.Wp, d.nsplit = dyh.data_ptr(), L.ptr(dyl), Wp, nsplit
            ws = torch.empty((self.KH * self.KW, self.Cout, self.Cin), device=dy.device, dtype=torch.float32)
            d.ws, d.dw_acc = ws.data_ptr(), dw.data_ptr()
            L.call("as_conv2d_wgrad_umma", d, _s())
            L.call("as_bias_grad_fp32", dy.data_ptr(), dy_pitch, self.Cout, B * H * W, db.data_ptr(), _s())
            return dw, db
        d = self._desc(B, H, W, srcs)
        L.call("as_conv2d_wgrad_fp32", d, dy.data_ptr(), dy_pitch, self.Cout, dw.data_ptr(), db.data_ptr(), _s())
        L.launch_count += 1
        return dw, db


def _src(t):
    return (t, t.shape[3], t.shape[3], NHWC)


class _train_engine:
    """The training path has no 2-pass kernels: under the "f16f8" inference engine its tensor-core kernels run in the
    3-pass split-bf16 mode (same fp32-parity class), for the forward and -- whenever autograd calls it -- the backward."""

    def __enter__(self):
        from .update import get_update_engine, set_update_engine
        self.prev = get_update_engine()
        if self.prev == "f16f8":
            set_update_engine("bf16x3")

    def __exit__(self, *exc):
        if self.prev == "f16f8":
            from .update import set_update_engine
            set_update_engine("f16f8")


class UpdateBlockFn(torch.autograd.Function):
    """forward(ub, flags, *tensors) with tensors = net[0..n) + flat(inp) + [corr, disp] + parameters."""

    @staticmethod
    def forward(ctx, ub, flags, n_net, has_corr, *tensors):
        with _train_engine():
            return UpdateBlockFn._forward(ctx, ub, flags, n_net, has_corr, *tensors)

    @staticmethod
    def backward(ctx, *gouts):
        with _train_engine():
            return UpdateBlockFn._backward(ctx, *gouts)

    @staticmethod
    def _forward(ctx, ub, flags, n_net, has_corr, *tensors):
        iter04, iter08, iter16, update = flags
        n_layers = ub.args.n_gru_layers
        net = list(tensors[:n_net])
        inp = [list(tensors[n_net + 3 * i:n_net + 3 * i + 3]) for i in range(n_net)]
        k = n_net + 3 * n_net
        corr = disp = None
        if has_corr:
            corr, disp = tensors[k], tensors[k + 1]
            k += 2
        dev = net[0].device
        e, dh = ub.encoder, ub.disp_head
        C = {k: _Conv([m], ub, k) for k, m in (("convc1", e.convc1), ("convc2", e.convc2), ("convd1", e.convd1),
                                               ("convd2", e.convd2), ("conv", e.conv), ("dh1", dh.conv1), ("dh2", dh.conv2))}
        for name in ("gru04", "gru08", "gru16"):
            g = getattr(ub, name)
            C[name + ".zr"] = _Conv([g.convz, g.convr], ub, name + ".zr")
            C[name + ".q"] = _Conv([g.convq], ub, name + ".q")
        tc, nsplit = _engine()
        planes = {}                                          # id(fp32 tensor) -> (tensor, hi/lo planes), this call only

        def pl(x):
            hit = planes.get(id(x))
            if hit is None:
                hit = planes[id(x)] = (x, _split(x, nsplit == 3))
            return hit[1]

        def conv(name, B, H, W, xs, epi, out, out_pitch, out_coff=0, **kw):
            """One forward convolution over the pixel-major fp32 sources xs, on the engine's kernel."""
            c = C[name]
            if tc and c.tc_ok() and all(x.shape[3] % 64 == 0 for x in xs):
                c.fwd_tc(B, H, W, [pl(x) for x in xs], nsplit, epi, out, out_pitch, out_coff, **kw)
            else:
                c.fwd(B, H, W, [_src(x) for x in xs], epi, out, out_pitch, out_coff, **kw)
        tape = {"C": C, "flags": flags, "n_net": n_net, "has_corr": has_corr, "gru": {}}
        hs = [_to_nhwc(t.detach().float()) for t in net]

        def context(i):
            cz, cr, cq = inp[i]
            B, Cc, H, W = cz.shape
            zr = torch.empty((B, H, W, 2 * Cc), device=dev, dtype=torch.float32)
            q = torch.empty((B, H, W, Cc), device=dev, dtype=torch.float32)
            for t, dst, pitch, off in ((cz, zr, 2 * Cc, 0), (cr, zr, 2 * Cc, Cc), (cq, q, Cc, 0)):
                t = t.detach().float().contiguous()
                L.call("as_nchw_to_nhwc", t.data_ptr(), dst.data_ptr(), B, Cc, H, W, pitch, off, _s())
            return zr, q

        def pool2x(x):
            B, H, W, Cc = x.shape
            out = torch.empty((B, (H + 1) // 2, (W + 1) // 2, Cc), device=dev, dtype=torch.float32)
            L.call("as_pool2x_nhwc", x.data_ptr(), out.data_ptr(), B, H, W, Cc, _s())
            return out

        def interp(x, ref):
            B, H, W, Cc = x.shape
            out = torch.empty((B, ref.shape[1], ref.shape[2], Cc), device=dev, dtype=torch.float32)
            L.call("as_interp_bilinear_nhwc", x.data_ptr(), out.data_ptr(), B, H, W, ref.shape[1], ref.shape[2], Cc, _s())
            return out

        def gru(name, i, h, xs):
            B, H, W, Hd = h.shape
            czr, cq = context(i)
            z, rh, r, q, hn = (torch.empty_like(h) for _ in range(5))
            conv(name + ".zr", B, H, W, [h] + xs, L.EPI_GRU_ZR, rh, Hd, ctx=czr, ctx_pitch=2 * Hd, h=h, z=z, save=r)
            conv(name + ".q", B, H, W, [rh] + xs, L.EPI_GRU_Q, hn, Hd, ctx=cq, ctx_pitch=Hd, h=h, z=z, save=q)
            tape["gru"][name] = dict(h=h, xs=xs, z=z, r=r, q=q, rh=rh)
            return hn

        if iter16:
            tape["p08"] = hs[1]
            hs[2] = gru("gru16", 2, hs[2], [pool2x(hs[1])])
        if iter08:
            xs = [pool2x(hs[0])]
            if n_layers > 2:
                tape["i16_src"] = hs[2]
                xs.append(interp(hs[2], hs[1]))
            hs[1] = gru("gru08", 1, hs[1], xs)
        if iter04:
            corr_c = corr.detach().float().contiguous()
            disp_c = disp.detach().float().contiguous()
            B, Cc, H, W = corr_c.shape
            corr_n = _to_nhwc(corr_c)                                   # wgrad reads pixel-major
            c1 = torch.empty((B, H, W, 64), device=dev, dtype=torch.float32)
            if tc and C["convc1"].tc_ok() and C["convc1"].Cout == 64:
                from .update_umma import _Planes
                cpad = (Cc + 63) // 64 * 64                             # lookup channels zero-padded to the K granule
                corrS = _Planes((B, H, W, cpad), dev, nsplit == 3)
                L.call("as_nchw_to_nhwc_split", corr_c.data_ptr(), corrS.hi.data_ptr(), L.ptr(corrS.lo), B, Cc, H, W, cpad,
                       _s())
                C["convc1"].fwd_tc(B, H, W, [corrS], nsplit, L.EPI_BIAS_RELU, c1, 64)
            else:
                C["convc1"].fwd(B, H, W, [_src(corr_n)], L.EPI_BIAS_RELU, c1, 64)
            cd = torch.empty((B, H, W, 128), device=dev, dtype=torch.float32)
            conv("convc2", B, H, W, [c1], L.EPI_BIAS_RELU, cd, 128, 0)
            d1 = torch.empty((B, H, W, 64), device=dev, dtype=torch.float32)
            disp_n = disp_c.view(B, H, W, 1)
            cd1 = C["convd1"]
            if tc and _KNOBS["convd1"] and (cd1.Cout, cd1.Cin, cd1.KH, cd1.KW) == (64, 1, 7, 7):
                L.call("as_convd1_fp32", disp_c.data_ptr(), C["convd1"].w.data_ptr(), C["convd1"].b.data_ptr(), d1.data_ptr(),
                       B, H, W, 64, 0, _s())
            else:
                C["convd1"].fwd(B, H, W, [_src(disp_n)], L.EPI_BIAS_RELU, d1, 64)
            conv("convd2", B, H, W, [d1], L.EPI_BIAS_RELU, cd, 128, 64)
            mo = torch.empty((B, H, W, 128), device=dev, dtype=torch.float32)
            conv("conv", B, H, W, [cd], L.EPI_BIAS_RELU, mo, 128, 0)
            L.call("as_nchw_to_nhwc", disp_c.data_ptr(), mo.data_ptr(), B, 1, H, W, 128, 127, _s())
            tape["enc"] = dict(corr=corr_n, c1=c1, cd=cd, d1=d1, disp=disp_n, mo=mo)
            xs = [mo]
            if n_layers > 1:
                tape["i08_src"] = hs[1]
                xs.append(interp(hs[1], hs[0]))
            hs[0] = gru("gru04", 0, hs[0], xs)
        outs = [h.permute(0, 3, 1, 2) for h in hs]
        if update:
            B, H, W, Hd = hs[0].shape
            t = torch.empty((B, H, W, 256), device=dev, dtype=torch.float32)
            conv("dh1", B, H, W, [hs[0]], L.EPI_BIAS_RELU, t, 256)
            delta = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
            if tc and C["dh2"].tc_ok() and C["dh2"].Cout == 1:          # [B,1,H,W] is [B,H,W,1]
                C["dh2"].fwd_tc(B, H, W, [pl(t)], nsplit, L.EPI_BIAS, delta, 1, 0)
            else:
                C["dh2"].fwd(B, H, W, [_src(t)], L.EPI_BIAS, delta, 1, 0, out_layout=NCHW)
            tape["head"] = dict(h=hs[0], t=t)
            outs.append(delta)
        tape["shapes"] = [tuple(h.shape) for h in hs]
        ctx.tape = tape
        ctx.ub = ub
        ctx.n_params = len(tensors) - k
        return tuple(outs)

    @staticmethod
    def _backward(ctx, *gouts):
        tape, ub = ctx.tape, ctx.ub
        C = tape["C"]
        iter04, iter08, iter16, update = tape["flags"]
        n_net, has_corr = tape["n_net"], tape["has_corr"]
        n_layers = ub.args.n_gru_layers
        dev = gouts[0].device if gouts[0] is not None else next(g.device for g in gouts if g is not None)
        pgrads = {}                                          # id(parameter) -> grad
        tcache = {}
        for c in C.values():
            c.tcache = tcache

        def acc_param(conv, dw, db):
            off = 0
            for c in conv.convs:
                n = c.weight.shape[0]
                pgrads[id(c.weight)] = dw[off:off + n]
                pgrads[id(c.bias)] = db[off:off + n]
                off += n

        # incoming gradients as pixel-major fp32 (zeros where autograd passed None)
        dh = []
        for i, shp in enumerate(tape["shapes"]):
            g = gouts[i]
            dh.append(torch.zeros(shp, device=dev, dtype=torch.float32) if g is None
                      else _to_nhwc(g.float()).clone())
        d_inp = [[None, None, None] for _ in range(n_net)]
        d_corr = None

        def relu_bwd(dy, dy_coff, y, y_coff, Cc):
            """dy * (y > 0) over Cc channels; the row is zero-padded to a multiple of 64 channels when the tensor-core
            data gradient consumes it (only `conv`, 127 channels, is affected)."""
            B, H, W, _ = y.shape
            Cp = (Cc + 63) // 64 * 64 if _engine()[0] else Cc
            dx = (torch.empty if Cp == Cc else torch.zeros)((B, H, W, Cp), device=dev, dtype=torch.float32)
            L.call("as_relu_bwd", dy.data_ptr(), dy.shape[3], dy_coff, y.data_ptr(), y.shape[3], y_coff, dx.data_ptr(), Cp, 0,
                   B * H * W, Cc, _s())
            return dx

        def gru_bwd(name, i, dhn):
            """returns (dh_in, [dx per source])"""
            sv = tape["gru"][name]
            h, xs, z, r, q, rh = sv["h"], sv["xs"], sv["z"], sv["r"], sv["q"], sv["rh"]
            B, H, W, Hd = h.shape
            N = B * H * W
            dq = torch.empty_like(h)
            dzr = torch.empty((B, H, W, 2 * Hd), device=dev, dtype=torch.float32)
            dhin = torch.zeros_like(h)
            L.call("as_gru_bwd_gates1", dhn.data_ptr(), z.data_ptr(), q.data_ptr(), h.data_ptr(), dq.data_ptr(), dzr.data_ptr(),
                   dhin.data_ptr(), N, Hd, _s())
            cq_, czr_ = C[name + ".q"], C[name + ".zr"]
            srcs_x = [_src(x) for x in xs]
            acc_param(cq_, *cq_.wgrad(B, H, W, [_src(rh)] + srcs_x, dq, Hd))
            din_q = cq_.dgrad(B, H, W, dq, Hd)                                  # [.., Hd + Cx]
            cin = din_q.shape[3]
            L.call("as_gru_bwd_gates2", din_q.data_ptr(), cin, h.data_ptr(), r.data_ptr(), dzr.data_ptr(), dhin.data_ptr(),
                   N, Hd, _s())
            acc_param(czr_, *czr_.wgrad(B, H, W, [_src(h)] + srcs_x, dzr, 2 * Hd))
            din_zr = czr_.dgrad(B, H, W, dzr, 2 * Hd)
            L.call("as_add_slice", din_zr.data_ptr(), cin, 0, dhin.data_ptr(), Hd, 0, N, Hd, _s())
            # x gradients: sum of both paths, then slice per source
            L.call("as_add_slice", din_zr.data_ptr(), cin, Hd, din_q.data_ptr(), cin, Hd, N, cin - Hd, _s())
            dxs, off = [], Hd
            for x in xs:
                cx = x.shape[3]
                dx = torch.zeros((B, H, W, cx), device=dev, dtype=torch.float32)
                L.call("as_add_slice", din_q.data_ptr(), cin, off, dx.data_ptr(), cx, 0, N, cx, _s())
                dxs.append(dx)
                off += cx
            # context gradients (cz, cr, cq), back to [B,C,H,W]
            d_inp[i] = [_to_nchw(dzr, Hd, 2 * Hd, 0), _to_nchw(dzr, Hd, 2 * Hd, Hd), _to_nchw(dq, Hd, Hd, 0)]
            return dhin, dxs

        def pool_bwd(dy, like):
            B, H, W, Cc = like.shape
            L.call("as_pool2x_nhwc_bwd", dy.data_ptr(), like.data_ptr(), B, H, W, Cc, _s())

        def interp_bwd(dy, acc):
            B, H, W, Cc = acc.shape
            L.call("as_interp_bilinear_nhwc_bwd", dy.data_ptr(), acc.data_ptr(), B, H, W, dy.shape[1], dy.shape[2], Cc, _s())

        # ---- disparity head
        if update:
            hd = tape["head"]
            h, t = hd["h"], hd["t"]
            B, H, W, Hd = h.shape
            gd = gouts[n_net]
            if gd is not None:
                gd = gd.float().contiguous().view(B, H, W, 1)
                acc_param(C["dh2"], *C["dh2"].wgrad(B, H, W, [_src(t)], gd, 1))
                dt = C["dh2"].dgrad(B, H, W, gd, 1)
                dpre = relu_bwd(dt, 0, t, 0, 256)
                acc_param(C["dh1"], *C["dh1"].wgrad(B, H, W, [_src(h)], dpre, 256))
                dhh = C["dh1"].dgrad(B, H, W, dpre, 256)
                L.call("as_add_slice", dhh.data_ptr(), Hd, 0, dh[0].data_ptr(), Hd, 0, B * H * W, Hd, _s())

        # ---- gru04 + motion encoder
        g0 = dh[0]
        if iter04:
            dhin, dxs = gru_bwd("gru04", 0, dh[0])
            g0 = dhin
            en = tape["enc"]
            B, H, W, _ = en["mo"].shape
            N = B * H * W
            dmo = dxs[0]
            if n_layers > 1:
                interp_bwd(dxs[1], dh[1])                       # into the NEW h08 gradient
            dpre = relu_bwd(dmo, 0, en["mo"], 0, 127)
            acc_param(C["conv"], *C["conv"].wgrad(B, H, W, [_src(en["cd"])], dpre, dpre.shape[3]))
            dcd = C["conv"].dgrad(B, H, W, dpre, dpre.shape[3])           # [..,128]
            dp_c2 = relu_bwd(dcd, 0, en["cd"], 0, 64)
            acc_param(C["convc2"], *C["convc2"].wgrad(B, H, W, [_src(en["c1"])], dp_c2, 64))
            dc1 = C["convc2"].dgrad(B, H, W, dp_c2, 64)
            dp_d2 = relu_bwd(dcd, 64, en["cd"], 64, 64)
            acc_param(C["convd2"], *C["convd2"].wgrad(B, H, W, [_src(en["d1"])], dp_d2, 64))
            dd1 = C["convd2"].dgrad(B, H, W, dp_d2, 64)
            dp_d1 = relu_bwd(dd1, 0, en["d1"], 0, 64)
            acc_param(C["convd1"], *C["convd1"].wgrad(B, H, W, [_src(en["disp"])], dp_d1, 64))
            dp_c1 = relu_bwd(dc1, 0, en["c1"], 0, 64)
            acc_param(C["convc1"], *C["convc1"].wgrad(B, H, W, [_src(en["corr"])], dp_c1, 64))
            dcn = C["convc1"].dgrad(B, H, W, dp_c1, 64)          # Cin channels (padded to 32 on the tensor cores)
            d_corr = _to_nchw(dcn, C["convc1"].Cin, dcn.shape[3])
        # ---- gru08
        g1 = dh[1]
        if iter08:
            dhin, dxs = gru_bwd("gru08", 1, dh[1])
            g1 = dhin
            pool_bwd(dxs[0], g0)                                # pool2x(old h04)
            if n_layers > 2:
                interp_bwd(dxs[1], dh[2])                       # interp(new h16)
        # ---- gru16
        g2 = dh[2] if n_net > 2 else None
        if iter16:
            dhin, dxs = gru_bwd("gru16", 2, dh[2])
            g2 = dhin
            pool_bwd(dxs[0], g1)                                # pool2x(old h08)
        gnet = [g0, g1, g2][:n_net]
        grads = [None, None, None, None]                       # ub, flags, n_net, has_corr
        grads += [g.permute(0, 3, 1, 2) for g in gnet]
        for i in range(n_net):
            grads += d_inp[i]
        if has_corr:
            grads += [d_corr, None]
        for p in ub.parameters():
            grads.append(pgrads.get(id(p)))
        tcache.clear()
        return tuple(grads)


def forward(ub, net, inp, corr, disp, iter04, iter08, iter16, update):
    n_net = len(net)
    has_corr = corr is not None
    flat = list(net)
    for i in range(n_net):
        flat += list(inp[i])
    if has_corr:
        flat += [corr, disp]
    flat += list(ub.parameters())
    outs = UpdateBlockFn.apply(ub, (iter04, iter08, iter16, update), n_net, has_corr, *flat)
    for i in range(n_net):
        net[i] = outs[i]
    if not update:
        return net
    return net, outs[n_net]
