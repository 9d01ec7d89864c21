"""Loop glue of the hot path (a12) and the hook that installs the operators into the reference models.

  igev_iterations  <- models/coreContinuous_IGEV/continuous_IGEVstereo.py:275-297
  raft_iterations  <- models/corePrune_RAFT/prune_raft_stereo.py:267-288
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib as L
from .geometry import CorrBlock1D, Combined_Geo_Encoding_Volume
from .submodule import build_gwc_volume
from .update import BasicMultiUpdateBlock, get_update_engine


def pixel_coords(B, H, W, device):
    """coords[b,y,x,0] = x (continuous_IGEVstereo.py:280 / prune_raft_stereo.py:272)."""
    return torch.arange(W, device=device, dtype=torch.float32).reshape(1, 1, W, 1).repeat(B, H, 1, 1)


def _add(a, b):
    a = a.contiguous()
    b = b.contiguous()
    y = torch.empty_like(a)
    L.call("as_add_f32", a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), L.stream_ptr())
    return y


_FUSION = {"on": True}


def set_lookup_fusion(on: bool) -> bool:
    """Fuse the IGEV lookup with BasicMotionEncoder.convc1 inside igev_iterations (tensor-core engines only)."""
    prev = _FUSION["on"]
    _FUSION["on"] = bool(on)
    return prev


def _iterate(lookup, update_block, net_list, inp_list, disp, coords, iters, slow_fast_gru=False, keep_all=False,
             lookup_events=None, update_events=None, fused_lookup=None):
    from .update_umma import call_replay_suspended
    with call_replay_suspended():          # this loop is captured as a whole (HotLoopGraph), not call by call
        return _iterate_body(lookup, update_block, net_list, inp_list, disp, coords, iters, slow_fast_gru, keep_all,
                             lookup_events, update_events, fused_lookup)


def _iterate_body(lookup, update_block, net_list, inp_list, disp, coords, iters, slow_fast_gru, keep_all, lookup_events,
                  update_events, fused_lookup):
    n_layers = update_block.args.n_gru_layers
    hist = []
    net_list = list(net_list)
    for _ in range(iters):
        disp = disp.detach()
        if fused_lookup is not None:       # lookup evaluated inside the update block, fused with convc1 (8(f)-1)
            feat = fused_lookup(disp, coords)
            feat.events = lookup_events
        elif lookup_events is not None:    # bench.py: CUDA events around the lookup launch, on its stream
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            feat = lookup(disp, coords)
            e1.record()
            lookup_events.append((e0, e1))
        else:
            feat = lookup(disp, coords)
        if n_layers == 3 and slow_fast_gru:
            net_list = update_block(net_list, inp_list, iter16=True, iter08=False, iter04=False, update=False)
        if n_layers >= 2 and slow_fast_gru:
            net_list = update_block(net_list, inp_list, iter16=n_layers == 3, iter08=True, iter04=False, update=False)
        if update_events is not None:
            u0 = torch.cuda.Event(enable_timing=True)
            u1 = torch.cuda.Event(enable_timing=True)
            u0.record()
        net_list, delta = update_block(net_list, inp_list, feat, disp, iter16=n_layers == 3, iter08=n_layers >= 2)
        if update_events is not None:
            u1.record()
            update_events.append((u0, u1))
        disp = _add(disp, delta)
        if keep_all:
            hist.append(disp)
    return (disp, net_list, hist) if keep_all else (disp, net_list)


# ---- CUDA-graph replay by default ----------------------------------------------------------------------------------------
# One hot-path step is ~20 launches per iteration; at small shapes (config 1: one 320x736 pair) the loop is bound by
# launch latency (18.2 ms eager vs 12.0 ms replayed).  igev_iterations / raft_iterations therefore capture the step once
# per (update block, shapes, engine, parameter versions) and replay it: inputs are copied into the graph's static buffers,
# outputs are cloned out.  Nothing is traced or compiled: the graph records exactly this library's kernel launches.
# Eager execution remains for everything a graph cannot serve: keep_all / timing events, autograd, an ongoing capture,
# the exact-fp32 engine, or AS_HOTLOOP_GRAPH=0 / set_graph_replay(False).
_GRAPH = {"on": os.environ.get("AS_HOTLOOP_GRAPH", "1") != "0", "cache": {}, "max": 2}


def set_graph_replay(on: bool) -> bool:
    prev = _GRAPH["on"]
    _GRAPH["on"] = bool(on)
    if not on:
        _GRAPH["cache"].clear()
    return prev


def _graph_key(kind, ub, tensors, scalars):
    from .update_umma import _OVERLAP, _LOWRES_1PASS, _GATE_WL
    shapes = tuple((tuple(t.shape), t.dtype, t.device.index) for t in tensors)
    params = tuple((p.data_ptr(), L.version_of(p)) for p in ub.parameters())
    from .geometry import get_corr_mode
    return (kind, id(ub), shapes, params, scalars, get_update_engine(), get_corr_mode(), _FUSION["on"], _OVERLAP["on"],
            _LOWRES_1PASS["on"], _GATE_WL["on"], torch.cuda.current_device())


def _graph_eligible(ub, tensors, keep_all, events):
    if not _GRAPH["on"] or keep_all or events or get_update_engine() == "fp32":
        return False
    if torch.cuda.is_current_stream_capturing():
        return False
    return all(t.is_cuda and not t.requires_grad for t in tensors)


def _graph_run(key, build, load):
    cache = _GRAPH["cache"]
    g = cache.get(key)
    if g is None:
        while len(cache) >= _GRAPH["max"]:
            cache.pop(next(iter(cache)))                 # oldest first: a graph pins its private memory pool
        g = build()
        cache[key] = g
    load(g)
    disp, net = g.replay()
    return disp.clone(), [t.clone() for t in net]


def graph_cache_clear():
    """Release the cached step graphs (each pins the memory pool of one captured step)."""
    _GRAPH["cache"].clear()


@torch.no_grad()
def igev_iterations(update_block: BasicMultiUpdateBlock, match_left, match_right, geo_encoding_volume, net_list,
                    inp_list, init_disp, iters, radius=4, num_levels=2, slow_fast_gru=False, keep_all=False,
                    lookup_events=None, update_events=None):
    """Build the combined volume, then ``iters`` x {lookup -> update -> disp += delta} (replayed from a CUDA graph when
    possible, see above)."""
    tensors = [match_left, match_right, geo_encoding_volume, init_disp] + list(net_list) + [t for l in inp_list for t in l]
    if not slow_fast_gru and _graph_eligible(update_block, tensors, keep_all, lookup_events is not None or update_events is not None):
        key = _graph_key("igev", update_block, tensors, (iters, radius, num_levels))
        return _graph_run(
            key,
            lambda: HotLoopGraph(update_block, match_left, match_right, geo_encoding_volume, net_list, inp_list, init_disp,
                                 iters=iters, radius=radius, num_levels=num_levels),
            lambda g: g.load(match_left, match_right, geo_encoding_volume, net_list, inp_list, init_disp))
    return _igev_iterations_eager(update_block, match_left, match_right, geo_encoding_volume, net_list, inp_list, init_disp,
                                  iters, radius, num_levels, slow_fast_gru, keep_all, lookup_events, update_events)


def _igev_iterations_eager(update_block, match_left, match_right, geo_encoding_volume, net_list, inp_list, init_disp, iters,
                           radius=4, num_levels=2, slow_fast_gru=False, keep_all=False, lookup_events=None,
                           update_events=None):
    geo_fn = Combined_Geo_Encoding_Volume(match_left.float(), match_right.float(), geo_encoding_volume.float(),
                                          radius=radius, num_levels=num_levels)
    B, _, H, W = match_left.shape
    coords = pixel_coords(B, H, W, match_left.device)
    fused = geo_fn.deferred if _FUSION["on"] and get_update_engine() != "fp32" else None
    return _iterate(geo_fn, update_block, net_list, inp_list, init_disp, coords, iters, slow_fast_gru, keep_all,
                    lookup_events, update_events, fused_lookup=fused)


@torch.no_grad()
def raft_iterations(update_block: BasicMultiUpdateBlock, match_left, match_right, net_list, inp_list, iters,
                    radius=4, num_levels=4, slow_fast_gru=False, keep_all=False):
    tensors = [match_left, match_right] + list(net_list) + [t for l in inp_list for t in l]
    if not slow_fast_gru and _graph_eligible(update_block, tensors, keep_all, False):
        key = _graph_key("raft", update_block, tensors, (iters, radius, num_levels))
        return _graph_run(
            key,
            lambda: HotLoopGraph(update_block, match_left, match_right, net_list=net_list, inp_list=inp_list, iters=iters,
                                 radius=radius, num_levels=num_levels),
            lambda g: g.load(match_left, match_right, net_list=net_list, inp_list=inp_list))
    return _raft_iterations_eager(update_block, match_left, match_right, net_list, inp_list, iters, radius, num_levels,
                                  slow_fast_gru, keep_all)


def _raft_iterations_eager(update_block, match_left, match_right, net_list, inp_list, iters, radius=4, num_levels=4,
                           slow_fast_gru=False, keep_all=False):
    corr_fn = CorrBlock1D(match_left.float(), match_right.float(), radius=radius, num_levels=num_levels)
    B, _, H, W = match_left.shape
    coords = pixel_coords(B, H, W, match_left.device)
    disp = match_left.new_zeros((B, 1, H, W))       # prune_raft_stereo.py:274
    fused = corr_fn.deferred if _FUSION["on"] and get_update_engine() != "fp32" else None
    return _iterate(corr_fn, update_block, net_list, inp_list, disp, coords, iters, slow_fast_gru, keep_all,
                    fused_lookup=fused)


class HotLoopGraph:
    """CUDA-graph replay of one hot-path step (volume build + all iterations) on static buffers.

    The per-iteration work is ~20 launches; at small shapes (config 1: one 320x736 pair) the loop is launch-bound, so
    the whole step is captured once and replayed (no tracing compiler involved: the graph records exactly the
    library's kernels).  ``geo_volume`` / ``init_disp`` given -> IGEV family, else RAFT family."""

    def __init__(self, update_block, match_left, match_right, geo_volume=None, net_list=None, inp_list=None,
                 init_disp=None, iters=32, radius=4, num_levels=None):
        self.igev = geo_volume is not None
        if num_levels is None:
            num_levels = 2 if self.igev else 4
        self.static_in = dict(ml=match_left.clone(), mr=match_right.clone(),
                              geo=geo_volume.clone() if self.igev else None,
                              net=[t.clone() for t in net_list],
                              inp=[[t.clone() for t in lst] for lst in inp_list],
                              disp=init_disp.clone() if init_disp is not None else None)
        self.args = (update_block, iters, radius, num_levels)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):   # warm-up: packs the weights outside capture
                self._run()
        torch.cuda.current_stream().wait_stream(s)
        # the loop-invariant context / hi-lo plane caches were filled by the warm-up: drop them so the conversion
        # kernels are part of the captured graph and follow the static buffers on every replay
        update_block.reset_caches()
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count
        # thread_local: CUDA calls of OTHER host threads (a DataLoader's pin-memory thread, say) do not invalidate the capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out_disp, self.out_net = self._run()
        self.launches = L.launch_count - n0          # kernel-launching ABI calls recorded in the graph (per replay)

    def _run(self):
        ub, iters, radius, num_levels = self.args
        si = self.static_in
        with torch.no_grad():
            if self.igev:
                return _igev_iterations_eager(ub, si["ml"], si["mr"], si["geo"], si["net"], si["inp"], si["disp"], iters,
                                              radius=radius, num_levels=num_levels)
            return _raft_iterations_eager(ub, si["ml"], si["mr"], si["net"], si["inp"], iters, radius=radius,
                                          num_levels=num_levels)

    def load(self, match_left, match_right, geo_volume=None, net_list=None, inp_list=None, init_disp=None):
        si = self.static_in
        si["ml"].copy_(match_left, non_blocking=True)
        si["mr"].copy_(match_right, non_blocking=True)
        if self.igev:
            si["geo"].copy_(geo_volume, non_blocking=True)
            si["disp"].copy_(init_disp, non_blocking=True)
        for d, s in zip(si["net"], net_list):
            d.copy_(s, non_blocking=True)
        for dl, sl in zip(si["inp"], inp_list):
            for d, s in zip(dl, sl):
                d.copy_(s, non_blocking=True)

    def replay(self):
        self.graph.replay()
        L.launch_count += self.launches
        return self.out_disp, self.out_net


def install_into_reference(ref_igev_module=None, ref_raft_module=None, defer_lookup=False):
    """Rebind the names the reference's model graphs resolve at call time:

        models.coreContinuous_IGEV.continuous_IGEVstereo.{Combined_Geo_Encoding_Volume, build_gwc_volume}
        models.coreContinuous_IGEV.continuous_IGEVstereo.context_upsample_multiscale_train
        models.corePrune_RAFT.prune_raft_stereo.{CorrBlock1D, context_upsample_multiscale_train}

    ``model.update_block`` / ``model.liif_up`` are swapped per model instance with ``adopt_update_block`` /
    ``adopt_liif_up``.  Everything installed is differentiable (training, config 5): the cost-volume objects,
    ``build_gwc_volume`` and the update block through hand-written adjoint kernels; the upsampler pieces (forward-only
    fused kernels) switch to the same arithmetic in differentiable ATen ops whenever a gradient is requested through
    them, so no ``disp_preds`` loss term ever loses its graph.

    defer_lookup=True binds the *_Deferred cost-volume classes instead: ``geo_fn(disp, coords)`` / ``corr_fn(disp, coords)``
    return the lookup unevaluated and the ADOPTED update block runs it fused with convc1 (SURVEY 8(f)-1 inside the
    reference's own loop).  Requires ``adopt_update_block`` on every model built from that module."""
    if defer_lookup:
        from .geometry import Combined_Geo_Encoding_Volume_Deferred, CorrBlock1D_Deferred
    if ref_igev_module is not None:
        ref_igev_module.Combined_Geo_Encoding_Volume = Combined_Geo_Encoding_Volume_Deferred if defer_lookup else Combined_Geo_Encoding_Volume
        ref_igev_module.build_gwc_volume = build_gwc_volume
        from .liif import context_upsample_multiscale_train      # SURVEY 8(f)-2, continuous_IGEVstereo.py:219
        ref_igev_module.context_upsample_multiscale_train = context_upsample_multiscale_train
    if ref_raft_module is not None:
        ref_raft_module.CorrBlock1D = CorrBlock1D_Deferred if defer_lookup else CorrBlock1D
        from .liif import context_upsample_multiscale_train      # prune_raft_stereo.py:227
        ref_raft_module.context_upsample_multiscale_train = context_upsample_multiscale_train


class _PendingStem:
    """corr_stem applied to a DeferredGwcVolume, still not run: corr_feature_att launches the fused kernel."""

    def __init__(self, vol, stem):
        self.vol, self.stem = vol, stem

    def run(self, att=None):
        from .submodule import gwc_corr_stem
        v, bn = self.vol, self.stem.bn
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        shift = bn.bias - bn.running_mean * scale
        return gwc_corr_stem(v.left, v.right, v.maxdisp, v.num_groups, self.stem.conv.weight, scale, shift, 0.01, att)


class CorrStem(nn.Module):
    """The reference's ``corr_stem`` (submodule.BasicConv: Conv3d + BatchNorm3d + LeakyReLU, continuous_IGEVstereo.py:172)
    around the SAME submodules (state_dict keys unchanged).  A DeferredGwcVolume input is handed on unevaluated when the
    fused kernel applies (eval-mode BatchNorm, no gradient requested); anything else takes the reference's arithmetic."""

    def __init__(self, ref):
        super().__init__()
        self.conv, self.bn, self.relu, self.use_bn = ref.conv, ref.bn, ref.relu, ref.use_bn

    def _fusable(self):
        c = self.conv
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        return (isinstance(c, nn.Conv3d) and c.bias is None and tuple(c.weight.shape) == (8, 8, 3, 3, 3)
                and tuple(c.stride) == (1, 1, 1) and tuple(c.padding) == (1, 1, 1) and tuple(c.dilation) == (1, 1, 1)
                and self.use_bn and self.relu and not self.bn.training and self.bn.affine
                and self.bn.running_mean is not None and not grad)

    def forward(self, x):
        from .submodule import DeferredGwcVolume
        if isinstance(x, DeferredGwcVolume):
            if self._fusable():
                return _PendingStem(x, self)
            x = x.materialize()
        x = self.conv(x)                                  # submodule.py:26-32
        if self.use_bn:
            x = self.bn(x)
        if self.relu:
            x = nn.functional.leaky_relu(x)
        return x


class CorrFeatureAtt(nn.Module):
    """The reference's ``corr_feature_att`` (submodule.FeatureAtt, :328-341) around the same ``feat_att`` stack; multiplies a
    pending corr_stem inside the fused kernel, a tensor like the reference does."""

    def __init__(self, ref):
        super().__init__()
        self.feat_att = ref.feat_att

    def forward(self, cv, feat):
        feat_att = self.feat_att(feat)
        if isinstance(cv, _PendingStem):
            return cv.run(torch.sigmoid(feat_att))
        return torch.sigmoid(feat_att.unsqueeze(2)) * cv


def adopt_corr_stem(model, ref_igev_module):
    """fuse build_gwc_volume with corr_stem's 3-D convolution (+ BatchNorm + LeakyReLU + the FeatureAtt
    multiply) for ``model`` (a continuous_IGEVStereo): swaps ``model.corr_stem`` / ``model.corr_feature_att`` for modules
    around the same parameters and rebinds the module-level ``build_gwc_volume`` to the deferring variant.  Every model
    built from ``ref_igev_module`` must be adopted (a plain corr_stem cannot take a DeferredGwcVolume).  Training
    (gradients requested, or BatchNorm in train mode) falls back to the unfused operators automatically."""
    from .submodule import build_gwc_volume_deferred
    model.corr_stem = CorrStem(model.corr_stem)
    model.corr_feature_att = CorrFeatureAtt(model.corr_feature_att)
    ref_igev_module.build_gwc_volume = build_gwc_volume_deferred
    return model


def adopt_update_block(ref_update_block, family="igev", replay=None):
    """Build our update block around the SAME nn.Parameter objects as a reference BasicMultiUpdateBlock.
    replay=True: every inference call of the block is replayed from a CUDA graph (update_umma.set_call_replay; for loops
    that run at the host's pace, i.e. small batches); None = the module-wide setting."""
    from .update import BasicMultiUpdateBlock, BasicMultiUpdateBlockRAFT
    cls = BasicMultiUpdateBlock if family == "igev" else BasicMultiUpdateBlockRAFT
    hd = [ref_update_block.gru16.convz.out_channels, ref_update_block.gru08.convz.out_channels,
          ref_update_block.gru04.convz.out_channels]
    ours = cls(ref_update_block.args, hidden_dims=hd)
    ours.load_state_dict(ref_update_block.state_dict(), strict=True)
    for (n1, p1), (n2, p2) in zip(ours.named_parameters(), ref_update_block.named_parameters()):
        assert n1 == n2
    # share storage: point our modules' parameters at the reference's Parameter objects
    ref_params = dict(ref_update_block.named_parameters())
    for name, _ in list(ours.named_parameters()):
        mod = ours
        parts = name.split(".")
        for p in parts[:-1]:
            mod = getattr(mod, p)
        mod._parameters[parts[-1]] = ref_params[name]
    ours.call_replay = replay
    return ours.to(next(ref_update_block.parameters()).device)


def adopt_liif_up(ref_liif_up, chanels):
    """Our liif_out_multi_scale_Training around the parameters of a reference one (model.liif_up);
    ``chanels`` = channel counts of the feature maps in call order (continuous_IGEVstereo.py:120-163)."""
    from .liif import liif_out_multi_scale_Training
    aff = {"win_w": 3, "win_h": 3, "dilation": [1, 2, 4, 8]}
    ours = liif_out_multi_scale_Training(encoder_dim=sum(chanels), mlphidden_list=[128, 64, 64], pos_dim=0,
                                         unfold=ref_liif_up.unfold, affinity_settings=aff, number_input=len(chanels),
                                         chanels=list(chanels))
    ours.load_state_dict(ref_liif_up.state_dict(), strict=True)
    ref_params = dict(ref_liif_up.named_parameters())
    for name, _ in list(ours.named_parameters()):
        mod = ours
        parts = name.split(".")
        for p in parts[:-1]:
            mod = getattr(mod, p)
        mod._parameters[parts[-1]] = ref_params[name]
    return ours.to(next(ref_liif_up.parameters()).device)


def adopt_model(model, ref_module, family="igev", defer_lookup=True, replay=None, encoders=True, corr_stem=True,
                liif_chanels=None):
    """Everything this library can take over in one reference model, in one call (INTEGRATION.md section 1 spelled out):

        import models.coreContinuous_IGEV.continuous_IGEVstereo as igev_mod
        A.adopt_model(model.module, igev_mod, "igev", replay=True)

    * rebinds the module-level names (``install_into_reference``; ``defer_lookup``: lookups fused with convc1),
    * ``model.update_block`` -> adopt_update_block(..., replay=replay),
    * ``model.liif_up`` -> adopt_liif_up when it is the multi-scale upsampler this library builds (``liif_chanels`` =
      channel counts of its feature maps in call order; default: from ``model.args`` as continuous_IGEVstereo.py:104-168 /
      prune_raft_stereo.py:100-189 derive them); variants outside the built set keep the reference module,
    * ``encoders``: ``model.cnet`` -> adopt_context_encoder, RAFT's ``model.fnet`` -> adopt_feature_encoder,
    * ``corr_stem`` (IGEV): adopt_corr_stem.
    Every model built from ``ref_module`` must be adopted (the rebound names are module-wide).  Returns ``model``."""
    from .extractor import adopt_context_encoder, adopt_feature_encoder
    igev = family == "igev"
    if family not in ("igev", "raft"):
        raise ValueError("family must be 'igev' or 'raft'")
    install_into_reference(ref_igev_module=ref_module if igev else None, ref_raft_module=None if igev else ref_module,
                           defer_lookup=defer_lookup)
    model.update_block = adopt_update_block(model.update_block, family, replay=replay)
    lu = getattr(model, "liif_up", None)
    if lu is not None and type(lu).__name__ == "liif_out_multi_scale_Training":
        if liif_chanels is None:
            hd = model.args.hidden_dims[2]
            agg = str(getattr(model.args, "agg_type", "type5"))
            if "type2" in agg:
                liif_chanels = [8, 32, 48 + hd]
            elif any(t in agg for t in ("type1", "type3", "type4", "type5")):
                liif_chanels = [48 + hd, 32]
        if liif_chanels is not None:
            try:
                model.liif_up = adopt_liif_up(lu, chanels=list(liif_chanels))
            except (NotImplementedError, RuntimeError):
                pass                                     # an upsampler variant outside the built set: the reference's stays
    if encoders:
        try:
            model.cnet = adopt_context_encoder(model.cnet)
        except (TypeError, AttributeError):
            pass
        if not igev and hasattr(model, "fnet"):
            try:
                model.fnet = adopt_feature_encoder(model.fnet)
            except TypeError:
                pass
    if corr_stem and igev and hasattr(model, "corr_stem") and hasattr(model, "corr_feature_att"):
        adopt_corr_stem(model, ref_module)
    return model
