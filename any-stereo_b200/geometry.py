"""Drop-in replacements for the reference's cost-volume objects, backed by the sm_100a kernels.

  CorrBlock1D                  <- models/corePrune_RAFT/geometry.py:6-56
  Combined_Geo_Encoding_Volume <- models/coreContinuous_IGEV/geometry.py:6-72

Same constructor/``__call__``/``corr`` signatures, same public attributes (``num_levels``, ``radius``,
``init_corr_pyramid``, ``geo_volume_pyramid``), same output tensors.  Internals differ:

* the pyramid levels live in row-pitched buffers (pitch = width rounded up to 4 floats, so every window
  load is a 16-byte aligned vector); ``init_corr_pyramid[i]`` are ``[N,1,1,w_i]`` views of them;
* the geometry pyramid is stored ``[pixel][disparity][group]``; ``geo_volume_pyramid[i]`` exposes the
  reference's ``[N,G,1,D_i]`` shape as a (strided) view of the same memory;
* ``__call__`` is ONE kernel launch with no host synchronisation (the reference's ``bilinear_sampler``
  syncs through ``torch.unique`` once per level per sampler, utils/utils.py:64).
"""
from __future__ import annotations

import os

import torch

from . import _lib as L

#: arithmetic of the all-pairs correlation: "bf16x3" (default: tcgen05, fp32-parity split, <= 1e-4 of max|ref|),
#: "fp32" (CUDA cores, exact), "bf16" (tcgen05, fast).  Set by anystereo_b200.set_corr_mode().
DEFAULT_CORR_MODE = "bf16x3"
_CORR_MODE = {"mode": DEFAULT_CORR_MODE}
_TAP_MAJOR = os.environ.get("AS_LOOKUP_C1_TAP", "1") != "0"     # second-generation fused IGEV lookup + convc1 kernel
_MODE_ID = {"fp32": L.CORR_FP32_SIMT, "bf16x3": L.CORR_BF16X3, "bf16": L.CORR_BF16}


def set_corr_mode(mode: str):
    if mode not in _MODE_ID:
        raise ValueError("corr mode must be one of %s" % sorted(_MODE_ID))
    _CORR_MODE["mode"] = mode


def get_corr_mode() -> str:
    return _CORR_MODE["mode"]


def _pitch(w: int) -> int:
    return max(4, (w + 3) // 4 * 4)


def _check_pair(fmap1, fmap2):
    L.require_cuda(fmap1, "fmap1", torch.float32, contiguous=False)
    L.require_cuda(fmap2, "fmap2", torch.float32, contiguous=False)
    if fmap1.dim() != 4 or fmap2.dim() != 4 or fmap1.shape[:3] != fmap2.shape[:3]:
        raise RuntimeError("fmap1/fmap2 must be [B,D,H,W1] and [B,D,H,W2]")
    return fmap1.contiguous(), fmap2.contiguous()


def _build_corr_levels(fmap1, fmap2, num_levels, mode=None):
    """all-pairs correlation + pooled levels -> (buffers [N,pitch_l], widths, pitches)."""
    fmap1, fmap2 = _check_pair(fmap1, fmap2)
    B, D, H, W1 = fmap1.shape
    W2 = fmap2.shape[3]
    N = B * H * W1
    widths = [W2 >> l for l in range(num_levels)]
    pitches = [_pitch(w) for w in widths]
    with torch.cuda.device(fmap1.device):
        bufs = [torch.empty((N, p), device=fmap1.device, dtype=torch.float32) for p in pitches]
        mode_id = _MODE_ID[mode or _CORR_MODE["mode"]]
        ws_bytes = L.lib().as_corr1d_workspace_bytes(B, D, H, W1, W2, mode_id)
        ws = torch.empty((max(ws_bytes, 1),), device=fmap1.device, dtype=torch.uint8) if ws_bytes else None
        L.call("as_corr1d_build", fmap1.data_ptr(), fmap2.data_ptr(), B, D, H, W1, W2, num_levels,
               L.ptr_array(bufs), L.int_array(pitches), mode_id, L.ptr(ws), ws_bytes, L.stream_ptr())
        if mode_id == L.CORR_FP32_SIMT:
            L.launch_count += num_levels - 1      # one pooling kernel per extra level
    return bufs, widths, pitches


class _CorrBuildFn(torch.autograd.Function):
    """Differentiable pyramid build (training, config 5): grads flow to the feature maps."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, num_levels, mode):
        bufs, widths, pitches = _build_corr_levels(fmap1, fmap2, num_levels, mode)
        ctx.save_for_backward(fmap1, fmap2)
        ctx.meta = (widths, pitches)
        return tuple(bufs)

    @staticmethod
    def backward(ctx, *g_levels):
        fmap1, fmap2 = ctx.saved_tensors
        widths, pitches = ctx.meta
        B, D, H, W1 = fmap1.shape
        W2 = fmap2.shape[3]
        N = B * H * W1
        st = L.stream_ptr()
        gl = [None if g is None else g.contiguous() for g in g_levels]
        # pool adjoint, coarse -> fine
        acc = None
        for l in range(len(gl) - 1, -1, -1):
            cur = gl[l].clone() if gl[l] is not None else torch.zeros((N, pitches[l]), device=fmap1.device)
            if acc is not None:
                L.call("as_pool1d_halve_bwd_acc", acc.data_ptr(), cur.data_ptr(), N, widths[l], pitches[l + 1],
                       pitches[l], st)
            acc = cur
        g1 = torch.empty_like(fmap1)
        g2 = torch.empty_like(fmap2)
        L.call("as_corr1d_bwd", acc.data_ptr(), pitches[0], fmap1.data_ptr(), fmap2.data_ptr(), g1.data_ptr(),
               g2.data_ptr(), B, D, H, W1, W2, st)
        return g1, g2, None, None


class _CorrLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, coords, meta, *bufs):
        widths, pitches, radius = meta
        out = _corr_lookup(bufs, widths, pitches, radius, disp, coords)
        ctx.save_for_backward(disp, coords)
        ctx.meta = (widths, pitches, radius, [tuple(b.shape) for b in bufs])
        return out

    @staticmethod
    def backward(ctx, g_out):
        disp, coords = ctx.saved_tensors
        widths, pitches, radius, shapes = ctx.meta
        B, _, H, W = disp.shape
        g_out = g_out.contiguous()
        gb = [torch.zeros(s, device=disp.device, dtype=torch.float32) for s in shapes]
        L.call("as_corr_lookup_bwd", L.ptr_array(gb), L.int_array(widths), L.int_array(pitches), len(gb),
               disp.data_ptr(), L.ptr(coords), g_out.data_ptr(), B, H, W, radius, L.stream_ptr())
        return (None, None, None) + tuple(gb)


def _check_disp_coords(disp, coords, extent=None):
    """``extent`` = (B, H, W1, device) of the feature maps the pyramid was built from: the lookup kernels index volume
    rows as (b*H + y)*W1 + x, so a disparity map of any other size (wrong scale, a different batch after sharding) or
    on another device would read out of bounds instead of failing like the reference's reshape / grid_sample does."""
    L.require_cuda(disp, "disp", torch.float32, contiguous=False)
    if disp.dim() != 4 or disp.shape[1] != 1:
        raise RuntimeError("disp must be [B,1,H,W]")
    B, _, H, W = disp.shape
    if extent is not None:
        if (B, H, W) != tuple(extent[:3]):
            raise RuntimeError("disp is [%d,1,%d,%d] but the cost volume was built for [%d,1,%d,%d]"
                               % ((B, H, W) + tuple(extent[:3])))
        if disp.device != extent[3]:
            raise RuntimeError("disp is on %s but the cost volume lives on %s" % (disp.device, extent[3]))
    disp = disp.contiguous()
    if coords is not None:
        L.require_cuda(coords, "coords", torch.float32, contiguous=False)
        if coords.numel() != B * H * W:
            raise RuntimeError("coords must be [B,H,W,1]")
        coords = coords.contiguous()
    return disp, coords


def _corr_lookup(bufs, widths, pitches, radius, disp, coords):
    B, _, H, W = disp.shape
    Lv = len(bufs)
    with torch.cuda.device(disp.device):
        out = torch.empty((B, Lv * (2 * radius + 1), H, W), device=disp.device, dtype=torch.float32)
        L.call("as_corr_lookup_fwd", L.ptr_array(bufs), L.int_array(widths), L.int_array(pitches), Lv,
               disp.data_ptr(), L.ptr(coords), out.data_ptr(), B, H, W, radius, L.stream_ptr())
    return out


class CorrBlock1D:
    """RAFT-style 1-D all-pairs correlation pyramid + radius lookup
    (reference: models/corePrune_RAFT/geometry.py:6-56; call site prune_raft_stereo.py:267-278)."""

    def __init__(self, init_fmap1, init_fmap2, num_levels=2, radius=4, mask_invalid=False):
        self.num_levels = num_levels
        self.radius = radius
        # mask_invalid is accepted and ignored: in the reference it is a discarded comparison
        # (geometry.py:53-54), i.e. a no-op.
        needs_grad = torch.is_grad_enabled() and (init_fmap1.requires_grad or init_fmap2.requires_grad)
        if needs_grad:
            f1, f2 = _check_pair(init_fmap1, init_fmap2)
            bufs = list(_CorrBuildFn.apply(f1, f2, num_levels, None))
            W2 = f2.shape[3]
            self._widths = [W2 >> l for l in range(num_levels)]
            self._pitches = [_pitch(w) for w in self._widths]
        else:
            bufs, self._widths, self._pitches = _build_corr_levels(init_fmap1.detach(), init_fmap2.detach(), num_levels)
        self._bufs = bufs
        self._extent = (init_fmap1.shape[0], init_fmap1.shape[2], init_fmap1.shape[3], init_fmap1.device)
        N = bufs[0].shape[0]
        self.init_corr_pyramid = [b[:, :w].unflatten(1, (1, 1, w)) if w > 0 else b[:, :0].reshape(N, 1, 1, 0)
                                  for b, w in zip(bufs, self._widths)]

    def __call__(self, disp, coords):
        disp, coords = _check_disp_coords(disp, coords, self._extent)
        if torch.is_grad_enabled() and any(b.requires_grad for b in self._bufs):
            return _CorrLookupFn.apply(disp, coords, (self._widths, self._pitches, self.radius), *self._bufs)
        return _corr_lookup(self._bufs, self._widths, self._pitches, self.radius, disp, coords)

    def deferred(self, disp, coords):
        """The same lookup, not yet run (see Combined_Geo_Encoding_Volume.deferred): the tensor-core engines fuse it
        with BasicMotionEncoder.convc1."""
        disp, coords = _check_disp_coords(disp, coords, self._extent)
        return DeferredCorrLookup(self, disp, coords)

    @staticmethod
    def corr(fmap1, fmap2, mask_invalid=False):
        """[B,D,H,W1] x [B,D,H,W2] -> [B,H,W1,1,W2] (geometry.py:46-56); no 1/sqrt(D) scaling."""
        f1, f2 = _check_pair(fmap1, fmap2)
        B, D, H, W1 = f1.shape
        W2 = f2.shape[3]
        if torch.is_grad_enabled() and (f1.requires_grad or f2.requires_grad):
            lvl = _CorrBuildFn.apply(f1, f2, 1, None)[0]      # differentiable like the constructor's pyramid build
            pitches = [_pitch(W2)]
        else:
            bufs, widths, pitches = _build_corr_levels(f1.detach(), f2.detach(), 1)
            lvl = bufs[0]
        if pitches[0] != W2:
            lvl = lvl[:, :W2].contiguous()
        return lvl.view(B, H, W1, 1, W2)


# ------------------------------------------------------------------------------------------------
# IGEV
# ------------------------------------------------------------------------------------------------

def _build_geo_levels(geo_volume, num_levels):
    L.require_cuda(geo_volume, "geo_volume", torch.float32, contiguous=False)
    if geo_volume.dim() != 5:
        raise RuntimeError("geo_volume must be [B,G,D,H,W]")
    geo_volume = geo_volume.contiguous()
    B, G, Dg, H, W = geo_volume.shape
    N = B * H * W
    with torch.cuda.device(geo_volume.device):
        bufs = [torch.empty((N, max(Dg >> l, 0), G), device=geo_volume.device, dtype=torch.float32)
                for l in range(num_levels)]
        L.call("as_geo_pyramid_build", geo_volume.data_ptr(), B, G, Dg, H, W, num_levels, L.ptr_array(bufs),
               L.stream_ptr())
    return bufs


class _GeoBuildFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, geo_volume, num_levels):
        ctx.shape = tuple(geo_volume.shape)
        return tuple(_build_geo_levels(geo_volume, num_levels))

    @staticmethod
    def backward(ctx, *g_levels):
        B, G, Dg, H, W = ctx.shape
        N = B * H * W
        dev = g_levels[0].device if g_levels[0] is not None else None
        gl = [g.contiguous() if g is not None else torch.zeros((N, Dg >> l, G), device=dev)
              for l, g in enumerate(g_levels)]
        g_geo = torch.empty(ctx.shape, device=gl[0].device, dtype=torch.float32)
        L.call("as_geo_pyramid_bwd", L.ptr_array(gl), B, G, Dg, H, W, len(gl), g_geo.data_ptr(), L.stream_ptr())
        return g_geo, None


def _geo_lookup(geo_bufs, G, Dg, corr_bufs, widths, pitches, radius, disp, coords):
    B, _, H, W = disp.shape
    Lv = len(geo_bufs)
    with torch.cuda.device(disp.device):
        out = torch.empty((B, Lv * (G + 1) * (2 * radius + 1), H, W), device=disp.device, dtype=torch.float32)
        L.call("as_geo_lookup_fwd", L.ptr_array(geo_bufs), G, Dg, L.ptr_array(corr_bufs), L.int_array(widths),
               L.int_array(pitches), Lv, disp.data_ptr(), L.ptr(coords), out.data_ptr(), B, H, W, radius,
               L.stream_ptr())
    return out


class _GeoLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, coords, meta, *bufs):
        G, Dg, widths, pitches, radius, Lv = meta
        out = _geo_lookup(bufs[:Lv], G, Dg, bufs[Lv:], widths, pitches, radius, disp, coords)
        ctx.save_for_backward(disp, coords)
        ctx.meta = meta
        ctx.shapes = [tuple(b.shape) for b in bufs]
        return out

    @staticmethod
    def backward(ctx, g_out):
        disp, coords = ctx.saved_tensors
        G, Dg, widths, pitches, radius, Lv = ctx.meta
        B, _, H, W = disp.shape
        g_out = g_out.contiguous()
        gb = [torch.zeros(s, device=disp.device, dtype=torch.float32) for s in ctx.shapes]
        L.call("as_geo_lookup_bwd", L.ptr_array(gb[:Lv]), G, Dg, L.ptr_array(gb[Lv:]), L.int_array(widths),
               L.int_array(pitches), Lv, disp.data_ptr(), L.ptr(coords), g_out.data_ptr(), B, H, W, radius,
               L.stream_ptr())
        return (None, None, None) + tuple(gb)


class Combined_Geo_Encoding_Volume:
    """IGEV combined geometry-encoding volume: all-pairs correlation pyramid + pooled GWC/geometry volume,
    looked up together every iteration
    (reference: models/coreContinuous_IGEV/geometry.py:6-72; call site continuous_IGEVstereo.py:275-286)."""

    def __init__(self, init_fmap1, init_fmap2, geo_volume, num_levels=2, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        grad_on = torch.is_grad_enabled()
        if grad_on and (init_fmap1.requires_grad or init_fmap2.requires_grad):
            f1, f2 = _check_pair(init_fmap1, init_fmap2)
            corr_bufs = list(_CorrBuildFn.apply(f1, f2, num_levels, None))
            W2 = f2.shape[3]
            self._widths = [W2 >> l for l in range(num_levels)]
            self._pitches = [_pitch(w) for w in self._widths]
        else:
            corr_bufs, self._widths, self._pitches = _build_corr_levels(init_fmap1.detach(), init_fmap2.detach(),
                                                                        num_levels)
        if grad_on and geo_volume.requires_grad:
            L.require_cuda(geo_volume, "geo_volume", torch.float32, contiguous=False)
            geo_bufs = list(_GeoBuildFn.apply(geo_volume.contiguous(), num_levels))
        else:
            geo_bufs = _build_geo_levels(geo_volume.detach(), num_levels)
        self._corr_bufs = corr_bufs
        self._geo_bufs = geo_bufs
        B, G, Dg, H, W = geo_volume.shape
        if (B, H, W) != (init_fmap1.shape[0], init_fmap1.shape[2], init_fmap1.shape[3]):
            raise RuntimeError("geo_volume [B,G,D,H,W] and init_fmap1 [B,C,H,W] disagree on B/H/W")
        self._G, self._Dg = G, Dg
        self._extent = (B, H, W, geo_volume.device)
        N = B * H * W
        # reference-shaped views: [N,1,1,w_i] and [N,G,1,D_i]
        self.init_corr_pyramid = [b[:, :w].unflatten(1, (1, 1, w)) for b, w in zip(corr_bufs, self._widths)]
        self.geo_volume_pyramid = [b.permute(0, 2, 1).unsqueeze(2) for b in geo_bufs]

    def __call__(self, disp, coords):
        disp, coords = _check_disp_coords(disp, coords, self._extent)
        bufs = list(self._geo_bufs) + list(self._corr_bufs)
        if torch.is_grad_enabled() and any(b.requires_grad for b in bufs):
            meta = (self._G, self._Dg, self._widths, self._pitches, self.radius, self.num_levels)
            return _GeoLookupFn.apply(disp, coords, meta, *bufs)
        return _geo_lookup(self._geo_bufs, self._G, self._Dg, self._corr_bufs, self._widths, self._pitches,
                           self.radius, disp, coords)

    def deferred(self, disp, coords):
        """The same lookup, not yet run: hand the result to BasicMultiUpdateBlock.forward as `corr` and the
        tensor-core engines fuse it with BasicMotionEncoder.convc1, so the 162-channel tensor never
        reaches HBM.  Engines / shapes without the fused kernel call .materialize() and behave as before."""
        disp, coords = _check_disp_coords(disp, coords, self._extent)
        return DeferredGeoLookup(self, disp, coords)

    @staticmethod
    def corr(fmap1, fmap2):
        """coreContinuous_IGEV/geometry.py:63-72."""
        return CorrBlock1D.corr(fmap1, fmap2)


class DeferredGeoLookup:
    """(volume, disp, coords) of one Combined_Geo_Encoding_Volume.__call__ (geometry.py:34-60), evaluated by its consumer."""

    def __init__(self, vol, disp, coords):
        self.vol, self.disp, self.coords = vol, disp, coords
        B, _, H, W = disp.shape
        self.shape = (B, vol.num_levels * (vol._G + 1) * (2 * vol.radius + 1), H, W)
        self.device = disp.device
        self.events = None          # bench.py: list collecting (start, end) CUDA events around the fused kernel

    @property
    def fusable(self):
        v = self.vol
        needs_grad = torch.is_grad_enabled() and any(b.requires_grad for b in list(v._geo_bufs) + list(v._corr_bufs))
        return v._G == 8 and v.radius == 4 and v.num_levels in (1, 2) and not needs_grad

    def materialize(self):
        return self.vol(self.disp, self.coords)

    @property
    def tap_major(self):
        """True when the second-generation kernel (as_geo_lookup_convc1_tap) takes this lookup: 2 levels, 16-byte aligned
        level buffers, correlation pitches in multiples of 4 floats (AS_LOOKUP_C1_TAP=0 keeps the first kernel)."""
        v = self.vol
        return (_TAP_MAJOR and v.num_levels == 2 and v._Dg >= 2 and all(p % 4 == 0 for p in v._pitches)
                and all(b.data_ptr() % 16 == 0 for b in list(v._geo_bufs) + list(v._corr_bufs)))

    @staticmethod
    def pack_convc1_weight(weight, split=True, tap_major=False, bias=None):
        """convc1.weight [64, L*81, 1, 1] -> bf16 hi/lo [64][192] in the K order of the fused kernel:
        channel (level l, group g, tap k) at K = l*96 + g*10 + k (g == 8: correlation taps); pads are zero.
        tap_major: K = l*96 + k*8 + g, correlation taps at l*96 + 72 + k, and `bias` (required) in column 81
        (as_geo_lookup_convc1_tap feeds a constant 1 there, so the bias goes through the GEMM)."""
        w = weight.detach().float().reshape(weight.shape[0], -1)
        Cout, Cin = w.shape
        if Cout != 64 or Cin not in (81, 162):
            raise RuntimeError("fused lookup+convc1 needs convc1: 81|162 -> 64 channels")
        c = torch.arange(Cin, device=w.device)
        g, k = (c % 81) // 9, c % 9
        if tap_major:
            kidx = (c // 81) * 96 + torch.where(g < 8, k * 8 + g, 72 + k)
        else:
            kidx = (c // 81) * 96 + g * 10 + k
        wp = torch.zeros((Cout, 192), device=w.device, dtype=torch.float32)
        wp[:, kidx] = w
        if tap_major:
            if bias is None:
                raise RuntimeError("the tap-major packing carries the bias in column 81")
            wp[:, 81] = bias.detach().float().to(w.device)
        hi = torch.empty((Cout, 192), device=w.device, dtype=torch.bfloat16)
        lo = torch.empty_like(hi) if split else None
        with torch.cuda.device(w.device):
            L.call("as_pack_conv_weight_bf16", wp.data_ptr(), hi.data_ptr(), L.ptr(lo), Cout, 192, 1, 1, Cout, 192,
                   L.stream_ptr())
        return hi, lo

    def convc1_planes(self, w_hi, w_lo, bias, out_hi, out_lo, tap_major=False):
        """relu(convc1(lookup)) as bf16 planes [B,H,W,64] (update.py:78,85 applied to geometry.py:34-60).
        tap_major must say how pack_convc1_weight laid the weights out."""
        v = self.vol
        B, _, H, W = self.disp.shape
        with torch.cuda.device(self.device):
            if self.events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            L.call("as_geo_lookup_convc1_tap" if tap_major else "as_geo_lookup_convc1",
                   L.ptr_array(v._geo_bufs), v._G, v._Dg, L.ptr_array(v._corr_bufs),
                   L.int_array(v._widths), L.int_array(v._pitches), v.num_levels, self.disp.data_ptr(),
                   L.ptr(self.coords), w_hi.data_ptr(), L.ptr(w_lo), bias.data_ptr(), 3 if w_lo is not None else 1,
                   out_hi.data_ptr(), L.ptr(out_lo), B, H, W, v.radius, L.stream_ptr())
            if self.events is not None:
                e1.record()
                self.events.append((e0, e1))


class DeferredCorrLookup:
    """(pyramid, disp, coords) of one CorrBlock1D.__call__ (corePrune_RAFT/geometry.py:24-43), evaluated by its consumer."""

    def __init__(self, blk, disp, coords):
        self.blk, self.disp, self.coords = blk, disp, coords
        B, _, H, W = disp.shape
        self.shape = (B, blk.num_levels * (2 * blk.radius + 1), H, W)
        self.device = disp.device
        self.events = None

    @property
    def fusable(self):
        b = self.blk
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in b._bufs)
        return b.radius == 4 and b.num_levels in (2, 4) and min(b._widths) > 0 and not needs_grad

    def materialize(self):
        return self.blk(self.disp, self.coords)

    @staticmethod
    def pack_convc1_weight(weight, split=True):
        """convc1.weight [64, L*9, 1, 1] -> bf16 hi/lo [64][64], channel (level l, tap k) at K = l*10 + k."""
        w = weight.detach().float().reshape(weight.shape[0], -1)
        Cout, Cin = w.shape
        if Cout != 64 or Cin not in (18, 36):
            raise RuntimeError("fused lookup+convc1 needs convc1: 18|36 -> 64 channels")
        c = torch.arange(Cin, device=w.device)
        wp = torch.zeros((Cout, 64), device=w.device, dtype=torch.float32)
        wp[:, (c // 9) * 10 + (c % 9)] = w
        hi = torch.empty((Cout, 64), device=w.device, dtype=torch.bfloat16)
        lo = torch.empty_like(hi) if split else None
        with torch.cuda.device(w.device):
            L.call("as_pack_conv_weight_bf16", wp.data_ptr(), hi.data_ptr(), L.ptr(lo), Cout, 64, 1, 1, Cout, 64, L.stream_ptr())
        return hi, lo

    def convc1_planes(self, w_hi, w_lo, bias, out_hi, out_lo):
        b = self.blk
        B, _, H, W = self.disp.shape
        with torch.cuda.device(self.device):
            L.call("as_corr_lookup_convc1", L.ptr_array(b._bufs), L.int_array(b._widths), L.int_array(b._pitches),
                   b.num_levels, self.disp.data_ptr(), L.ptr(self.coords), w_hi.data_ptr(), L.ptr(w_lo), bias.data_ptr(),
                   3 if w_lo is not None else 1, out_hi.data_ptr(), L.ptr(out_lo), B, H, W, b.radius, L.stream_ptr())


class Combined_Geo_Encoding_Volume_Deferred(Combined_Geo_Encoding_Volume):
    """Drop-in variant for models whose update block has been adopted (hotpath.adopt_update_block): ``geo_fn(disp, coords)``
    returns the lookup UNEVALUATED (a DeferredGeoLookup) and the adopted update block fuses it with convc1
    (continuous_IGEVstereo.py:275-295 hands the result straight to ``self.update_block``).  Gradients requested, the
    exact-fp32 engine or unsupported shapes materialise it inside the update block, as ``deferred`` documents."""

    def __call__(self, disp, coords):
        return self.deferred(disp, coords)


class CorrBlock1D_Deferred(CorrBlock1D):
    """RAFT-family twin of Combined_Geo_Encoding_Volume_Deferred (prune_raft_stereo.py:267-290)."""

    def __call__(self, disp, coords):
        return self.deferred(disp, coords)
