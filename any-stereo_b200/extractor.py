"""SURVEY 8(f)-4, first step: the context encoder ``cnet`` (reference: models/*/extractor.py:200-300 MultiBasicEncoder,
ResidualBlock :9-58; call sites continuous_IGEVstereo.py:270, prune_raft_stereo.py:253).

With the iteration loop on the tensor cores, one 384x1248 pair spends 7 ms of a 36 ms forward in this module -- and only
2.4 ms of those in convolutions (profiles/cnet_prof_r02.txt): 2.8 ms are 33 eval-mode BatchNorm2d kernels, 1.1 ms NCHW <->
NHWC conversions around every cuDNN call, 1.1 ms separate ReLU / add kernels.  None of that is arithmetic the model needs
at inference: an eval-mode BatchNorm is a per-channel affine map of the convolution in front of it.

``adopt_context_encoder(model.cnet)`` wraps the SAME submodules (state_dict keys unchanged) and, when no gradient is
requested and every BatchNorm2d is in eval mode, runs

    conv(x, w * s) + ((b - mean) * s + beta),   s = gamma / sqrt(var + eps)

per Conv2d + BatchNorm2d pair, channels-last end to end, ReLU and the residual add in place.  The convolutions themselves
stay cuDNN library calls (SURVEY names "channels-last cuDNN or custom kernels" for this row; the tensor-core kernels of
this library cover stride-1 3x3 layers with 64-multiple channels, which is layer1 only -- DESIGN 9).  Anything else
(training, gradients, GroupNorm / InstanceNorm variants) takes the reference module's own forward.
"""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L


class _Folded:
    """One Conv2d with the eval-mode BatchNorm2d behind it folded in (channels-last weights)."""
    __slots__ = ("w", "b", "stride", "padding")

    def __init__(self, conv, bn=None):
        w = conv.weight.detach().float()
        b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
        if isinstance(bn, nn.BatchNorm2d):
            s = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
            if bn.weight is not None:
                s = s * bn.weight.detach().float()
            w = w * s.view(-1, 1, 1, 1)
            b = (b - bn.running_mean.detach().float()) * s
            if bn.bias is not None:
                b = b + bn.bias.detach().float()
        self.w = w.contiguous(memory_format=torch.channels_last)
        self.b = b.contiguous()
        self.stride, self.padding = conv.stride, conv.padding

    def __call__(self, x, relu=False):
        y = F.conv2d(x, self.w, self.b, self.stride, self.padding)
        return y.relu_() if relu else y


class _FoldedBlock:
    """ResidualBlock (extractor.py:9-58): relu(x' + relu(bn2(conv2(relu(bn1(conv1(x))))))), x' = bn3(conv1x1(x)) or x."""
    __slots__ = ("c1", "c2", "down")

    def __init__(self, blk):
        self.c1 = _Folded(blk.conv1, blk.norm1)
        self.c2 = _Folded(blk.conv2, blk.norm2)
        self.down = None if blk.downsample is None else _Folded(blk.downsample[0], blk.downsample[1])

    def __call__(self, x):
        y = self.c2(self.c1(x, True), True)
        if self.down is not None:
            x = self.down(x)
        return y.add_(x).relu_()


def _is_block(m):
    return all(hasattr(m, n) for n in ("conv1", "conv2", "norm1", "norm2", "downsample"))


class ContextEncoder(nn.Module):
    """The reference's MultiBasicEncoder around the same submodules; see the module docstring."""

    def __init__(self, ref):
        super().__init__()
        for name, child in ref.named_children():
            self.add_module(name, child)
        self.norm_fn = ref.norm_fn
        self.downsample = ref.downsample
        self.__dict__["_ref"] = ref            # not registered: its parameters are already ours through the children
        self.__dict__["_fold_cache"] = None

    # -- when the folded path applies --------------------------------------------------------------------------------
    def _fusable(self, x):
        if self.norm_fn != "batch" or not torch.is_floating_point(x):
            return False
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                if m.training or m.running_mean is None:
                    return False
            elif isinstance(m, (nn.GroupNorm, nn.InstanceNorm2d)):
                return False
        return all(_is_block(b) for layer in self._layers() for b in layer)

    def _layers(self):
        return [getattr(self, "layer%d" % i) for i in range(1, 6)]

    def _folded(self):
        key = tuple((t.data_ptr(), L.version_of(t)) for t in itertools.chain(self.parameters(), self.buffers()))
        hit = self.__dict__["_fold_cache"]
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            f = {"stem": _Folded(self.conv1, self.norm1),
                 "layers": [[_FoldedBlock(b) for b in layer] for layer in self._layers()],
                 "out04": [(_FoldedBlock(h[0]), _Folded(h[1])) for h in self.outputs04],
                 "out08": [(_FoldedBlock(h[0]), _Folded(h[1])) for h in self.outputs08],
                 "out16": [_Folded(h) for h in self.outputs16]}
        self.__dict__["_fold_cache"] = (key, f)
        return f

    def invalidate_weights(self):
        """After writes that bypass the version counter (``param.data``, EMA through ``.data``)."""
        self.__dict__["_fold_cache"] = None

    # -- forward (extractor.py:280-304) --------------------------------------------------------------------------------
    def forward(self, x, dual_inp=False, num_layers=3):
        if not self._fusable(x):
            return self.__dict__["_ref"](x, dual_inp=dual_inp, num_layers=num_layers)
        f = self._folded()
        # (the reference calls this module under autocast(enabled=args.mixed_precision): the folded path keeps fp32 / TF32)
        with torch.no_grad(), torch.autocast(device_type=x.device.type, enabled=False):
            x = f["stem"](x.float().contiguous(memory_format=torch.channels_last), True)
            for layer in f["layers"][:3]:
                for blk in layer:
                    x = blk(x)
            v = None
            if dual_inp:
                v = x.contiguous()
                x = x[:(x.shape[0] // 2)]
            # the heads leave in the reference's (NCHW-contiguous) layout: whatever consumes them may .view() them
            outputs04 = [conv(blk(x)).contiguous() for blk, conv in f["out04"]]
            if num_layers == 1:
                return (outputs04, v) if dual_inp else (outputs04,)
            y = x
            for blk in f["layers"][3]:
                y = blk(y)
            outputs08 = [conv(blk(y)).contiguous() for blk, conv in f["out08"]]
            if num_layers == 2:
                return (outputs04, outputs08, v) if dual_inp else (outputs04, outputs08)
            z = y
            for blk in f["layers"][4]:
                z = blk(z)
            outputs16 = [conv(z).contiguous() for conv in f["out16"]]
            return (outputs04, outputs08, outputs16, v) if dual_inp else (outputs04, outputs08, outputs16)


def _instnorm_(y, eps, relu=True, resid=None):
    """In place on a channels-last fp32 tensor [B,C,H,W]: InstanceNorm2d (no affine, no running stats) + ReLU, or the
    block's final relu(resid + relu(norm(y))) -- csrc/instnorm.cu."""
    B, C, H, W = y.shape
    ws_bytes = L.lib().as_instnorm_workspace_bytes(B, C)
    ws = torch.empty((ws_bytes,), device=y.device, dtype=torch.uint8)
    L.call("as_instnorm_nhwc", y.data_ptr(), L.ptr(resid), y.data_ptr(), ws.data_ptr(), ws_bytes, B, H * W, C, float(eps),
           1 if relu else 0, L.stream_ptr())
    L.launch_count += 2                              # stats + finalize + apply
    return y


def _cl(t):
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


class FeatureEncoder(nn.Module):
    """The reference's BasicEncoder(norm_fn='instance') -- RAFT's ``fnet`` (corePrune_RAFT/extractor.py:126-201; call site
    prune_raft_stereo.py:108,252) -- around the same submodules.  Inference on CUDA: channels-last cuDNN convolutions, every
    InstanceNorm2d + ReLU (+ the block's residual add + ReLU) in ONE elementwise pass over the tensor after a one-read
    statistics kernel (csrc/instnorm.cu) instead of ATen's reshaped batch norm + separate ReLU / add kernels.  Gradients,
    CPU tensors, other norm layers: the reference module's own forward."""

    def __init__(self, ref):
        super().__init__()
        for name, child in ref.named_children():
            self.add_module(name, child)
        self.norm_fn = ref.norm_fn
        self.downsample = ref.downsample
        self.__dict__["_ref"] = ref
        self.__dict__["_fold_cache"] = None

    def _fusable(self, xs):
        if self.norm_fn != "instance":
            return False
        if not all(t.is_cuda and t.dtype == torch.float32 for t in xs):
            return False
        if torch.is_grad_enabled() and (any(t.requires_grad for t in xs) or any(p.requires_grad for p in self.parameters())):
            return False
        if self.training and getattr(self, "dropout", None) is not None:
            return False
        for m in self.modules():
            if isinstance(m, nn.InstanceNorm2d):
                if m.affine or m.track_running_stats:
                    return False
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                return False
        return all(_is_block(b) for layer in (self.layer1, self.layer2, self.layer3) for b in layer)

    def _weights(self):
        key = tuple((t.data_ptr(), L.version_of(t)) for t in self.parameters())
        hit = self.__dict__["_fold_cache"]
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            f = {"stem": _Folded(self.conv1), "out": _Folded(self.conv2),
                 "layers": [[(_Folded(b.conv1), _Folded(b.conv2), None if b.downsample is None else _Folded(b.downsample[0]),
                              b.norm1.eps, b.norm2.eps, None if b.downsample is None else b.downsample[1].eps)
                             for b in layer] for layer in (self.layer1, self.layer2, self.layer3)]}
        self.__dict__["_fold_cache"] = (key, f)
        return f

    def invalidate_weights(self):
        self.__dict__["_fold_cache"] = None

    def forward(self, x, dual_inp=False):
        is_list = isinstance(x, (tuple, list))
        if not self._fusable(list(x) if is_list else [x]):
            return self.__dict__["_ref"](x, dual_inp=dual_inp)
        f = self._weights()
        # autocast off: the normalisation kernels read fp32 (the reference calls this module under autocast(mixed_precision))
        with torch.no_grad(), torch.cuda.device(x[0].device if is_list else x.device), torch.autocast("cuda", enabled=False):
            if is_list:
                batch_dim = x[0].shape[0]
                x = torch.cat(x, dim=0)
            x = _instnorm_(_cl(f["stem"](_cl(x))), self.norm1.eps)
            for layer in f["layers"]:
                for c1, c2, down, e1, e2, e3 in layer:
                    y = _instnorm_(_cl(c1(x)), e1)
                    y = _cl(c2(y))
                    if down is not None:
                        x = _instnorm_(_cl(down(x)), e3, relu=False)
                    x = _instnorm_(y, e2, relu=True, resid=x)
            x = f["out"](x).contiguous()
            if is_list:
                x = x.split(split_size=batch_dim, dim=0)
        return x


def adopt_feature_encoder(ref_fnet):
    """``model.fnet = adopt_feature_encoder(model.fnet)`` (RAFT family): same parameters and state_dict keys; see FeatureEncoder."""
    if isinstance(ref_fnet, FeatureEncoder):
        return ref_fnet
    need = ("conv1", "norm1", "layer1", "layer2", "layer3", "conv2")
    if not all(hasattr(ref_fnet, n) for n in need) or hasattr(ref_fnet, "outputs04"):
        raise TypeError("adopt_feature_encoder expects the reference's BasicEncoder (models/corePrune_RAFT/extractor.py:126)")
    return FeatureEncoder(ref_fnet)


def adopt_context_encoder(ref_cnet):
    """``model.cnet = adopt_context_encoder(model.cnet)``: eval-mode BatchNorm folded into the convolutions, channels-last,
    in-place ReLU / residual adds; same parameters, same state_dict keys, reference arithmetic whenever a gradient is
    requested or a norm layer is not an eval-mode BatchNorm2d."""
    if isinstance(ref_cnet, ContextEncoder):
        return ref_cnet
    need = ("conv1", "norm1", "layer1", "layer2", "layer3", "layer4", "layer5", "outputs04", "outputs08", "outputs16")
    if not all(hasattr(ref_cnet, n) for n in need):
        raise TypeError("adopt_context_encoder expects the reference's MultiBasicEncoder (models/*/extractor.py:200)")
    return ContextEncoder(ref_cnet)


# ---- BasicConv (submodule.py:6-32): Conv / ConvTranspose (2-D or 3-D, bias=False) + BatchNorm + LeakyReLU -------------------
# The reference builds its 3-D hourglass (continuous_IGEVstereo.py:22-89), FeatureAtt (submodule.py:328-341), the Conv2x
# up-blocks of `Feature` (extractor.py) and corr_stem out of this one block.  At inference each is three launches (conv,
# batch norm, out-of-place LeakyReLU) over small tensors; with the BatchNorm folded it is one convolution and an in-place
# LeakyReLU.  fold_basic_convs(model) swaps every such block for a wrapper around the SAME conv / bn submodules.
class FoldedBasicConv(nn.Module):
    _OPS = {nn.Conv2d: F.conv2d, nn.Conv3d: F.conv3d, nn.ConvTranspose2d: F.conv_transpose2d,
            nn.ConvTranspose3d: F.conv_transpose3d}

    def __init__(self, ref):
        super().__init__()
        self.conv, self.bn = ref.conv, ref.bn
        self.relu, self.use_bn = ref.relu, ref.use_bn
        self.__dict__["_ref"] = ref
        self.__dict__["_fold_cache"] = None

    def _fusable(self, x):
        c = self.conv
        if type(c) not in self._OPS or c.groups != 1 or c.bias is not None or c.padding_mode != "zeros":
            return False
        if self.use_bn and (self.bn.training or self.bn.running_mean is None):
            return False
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False
        return torch.is_floating_point(x)

    def _folded(self):
        ts = list(self.conv.parameters()) + (list(self.bn.parameters()) + list(self.bn.buffers()) if self.use_bn else [])
        key = tuple((t.data_ptr(), L.version_of(t)) for t in ts)
        hit = self.__dict__["_fold_cache"]
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        with torch.no_grad():
            w = self.conv.weight.detach().float()
            b = None
            if self.use_bn:
                bn = self.bn
                s = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
                if bn.weight is not None:
                    s = s * bn.weight.detach().float()
                transposed = isinstance(self.conv, (nn.ConvTranspose2d, nn.ConvTranspose3d))
                shape = [1] * w.dim()
                shape[1 if transposed else 0] = -1          # output channels: dim 1 of a transposed convolution's weight
                w = (w * s.view(shape)).contiguous()
                b = -bn.running_mean.detach().float() * s
                if bn.bias is not None:
                    b = b + bn.bias.detach().float()
        self.__dict__["_fold_cache"] = (key, w, b)
        return w, b

    def forward(self, x):
        if not self._fusable(x):
            return self.__dict__["_ref"](x)
        w, b = self._folded()
        c = self.conv
        with torch.no_grad(), torch.autocast(device_type=x.device.type, enabled=False):
            x = x.float()
            if isinstance(c, (nn.ConvTranspose2d, nn.ConvTranspose3d)):
                y = self._OPS[type(c)](x, w, b, c.stride, c.padding, c.output_padding, 1, c.dilation)
            else:
                y = self._OPS[type(c)](x, w, b, c.stride, c.padding, c.dilation, 1)
            return F.leaky_relu_(y, 0.01) if self.relu else y          # nn.LeakyReLU() default slope, submodule.py:30


def _is_basic_conv(m):
    return (type(m).__name__ == "BasicConv" and all(hasattr(m, n) for n in ("conv", "bn", "use_bn", "relu"))
            and isinstance(getattr(m, "bn"), (nn.BatchNorm2d, nn.BatchNorm3d)))


def fold_basic_convs(model):
    """Swap every reference ``BasicConv`` under ``model`` for a FoldedBasicConv (same conv / bn submodules, same state_dict
    keys).  Returns the handles ``unfold_basic_convs`` takes to put the originals back."""
    handles = []
    for parent in list(model.modules()):
        for name, child in list(parent._modules.items()):
            if child is not None and _is_basic_conv(child):
                parent._modules[name] = FoldedBasicConv(child)
                handles.append((parent, name, child))
    return handles


def unfold_basic_convs(handles):
    for parent, name, child in handles:
        parent._modules[name] = child
