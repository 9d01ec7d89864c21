"""``build_gwc_volume`` -- drop-in for models/coreContinuous_IGEV/submodule.py:253-271."""
from __future__ import annotations

import torch

from . import _lib as L


def _gwc_fwd(left, right, maxdisp, num_groups):
    B, C, H, W = left.shape
    with torch.cuda.device(left.device):
        out = torch.empty((B, num_groups, maxdisp, H, W), device=left.device, dtype=torch.float32)
        L.call("as_gwc_build_fwd", left.data_ptr(), right.data_ptr(), out.data_ptr(), B, C, H, W, maxdisp,
               num_groups, L.stream_ptr())
    return out


class _GwcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, left, right, maxdisp, num_groups):
        ctx.save_for_backward(left, right)
        ctx.meta = (maxdisp, num_groups)
        return _gwc_fwd(left, right, maxdisp, num_groups)

    @staticmethod
    def backward(ctx, g):
        left, right = ctx.saved_tensors
        maxdisp, G = ctx.meta
        B, C, H, W = left.shape
        g = g.contiguous().float()
        gl = torch.empty_like(left)
        gr = torch.empty_like(right)
        L.call("as_gwc_build_bwd", g.data_ptr(), left.data_ptr(), right.data_ptr(), gl.data_ptr(), gr.data_ptr(),
               B, C, H, W, maxdisp, G, L.stream_ptr())
        return gl, gr, None, None


def build_gwc_volume(refimg_fea, targetimg_fea, maxdisp, num_groups):
    """[B,C,H,W] x2 -> [B,num_groups,maxdisp,H,W], same dtype/device as the inputs.

    vol[b,g,d,y,x] = mean_{c in group g} ref[b,c,y,x] * tgt[b,c,y,x-d] for x >= d, else 0.
    The arithmetic is fp32; half inputs (the reference calls this under autocast,
    continuous_IGEVstereo.py:244,262) are widened on the way in and narrowed on the way out.
    """
    if refimg_fea.shape != targetimg_fea.shape or refimg_fea.dim() != 4:
        raise RuntimeError("refimg_fea/targetimg_fea must both be [B,C,H,W]")
    C = refimg_fea.shape[1]
    assert C % num_groups == 0  # submodule.py:255
    L.require_cuda(refimg_fea, "refimg_fea", contiguous=False)
    L.require_cuda(targetimg_fea, "targetimg_fea", contiguous=False)
    dt = refimg_fea.dtype
    left = refimg_fea.float().contiguous()
    right = targetimg_fea.float().contiguous()
    if torch.is_grad_enabled() and (left.requires_grad or right.requires_grad):
        vol = _GwcFn.apply(left, right, int(maxdisp), int(num_groups))
    else:
        vol = _gwc_fwd(left, right, int(maxdisp), int(num_groups))
    return vol if dt == torch.float32 else vol.to(dt)


class DeferredGwcVolume:
    """(left, right, maxdisp, num_groups) of one build_gwc_volume call that has not run yet: its consumer (the adopted
    ``corr_stem`` / ``corr_feature_att`` pair, hotpath.adopt_corr_stem) evaluates it fused with the first 3-D convolution
    (SURVEY 8(f)-3), so the correlation volume never reaches HBM.  ``materialize()`` is the plain volume."""

    def __init__(self, left, right, maxdisp, num_groups):
        self.left, self.right, self.maxdisp, self.num_groups = left, right, int(maxdisp), int(num_groups)
        B, _, H, W = left.shape
        self.shape = (B, self.num_groups, self.maxdisp, H, W)
        self.dtype, self.device = left.dtype, left.device

    def materialize(self):
        return build_gwc_volume(self.left, self.right, self.maxdisp, self.num_groups)


def build_gwc_volume_deferred(refimg_fea, targetimg_fea, maxdisp, num_groups):
    """build_gwc_volume for models whose corr_stem has been adopted: returns a DeferredGwcVolume whenever the fused kernel
    can take it (8 groups, maxdisp <= 48, no gradient requested), the volume itself otherwise."""
    needs_grad = torch.is_grad_enabled() and (refimg_fea.requires_grad or targetimg_fea.requires_grad)
    if (needs_grad or num_groups != 8 or maxdisp > 48 or refimg_fea.dim() != 4 or not refimg_fea.is_cuda
            or refimg_fea.shape != targetimg_fea.shape or refimg_fea.shape[1] % 8):
        return build_gwc_volume(refimg_fea, targetimg_fea, maxdisp, num_groups)
    return DeferredGwcVolume(refimg_fea, targetimg_fea, maxdisp, num_groups)


def gwc_corr_stem(refimg_fea, targetimg_fea, maxdisp, num_groups, conv_weight, scale, shift, negative_slope=0.01, att=None):
    """build_gwc_volume -> Conv3d(G, G, 3, 1, 1, bias=False) -> per-channel affine (eval BatchNorm3d) -> LeakyReLU
    [-> * att[:, :, None]] in ONE kernel (continuous_IGEVstereo.py:262-264; submodule.py:6-32, 253-271, 328-341).
    Forward only.  refimg_fea/targetimg_fea [B,C,H,W]; conv_weight [G,G,3,3,3]; scale/shift [G]; att [B,G,H,W] or None."""
    if refimg_fea.shape != targetimg_fea.shape or refimg_fea.dim() != 4:
        raise RuntimeError("refimg_fea/targetimg_fea must both be [B,C,H,W]")
    B, C, H, W = refimg_fea.shape
    assert C % num_groups == 0  # submodule.py:255
    L.require_cuda(refimg_fea, "refimg_fea", contiguous=False)
    L.require_cuda(targetimg_fea, "targetimg_fea", contiguous=False)
    if tuple(conv_weight.shape) != (num_groups, num_groups, 3, 3, 3):
        raise RuntimeError("conv_weight must be [G,G,3,3,3]")
    dt = refimg_fea.dtype
    dev = refimg_fea.device
    left = refimg_fea.detach().float().contiguous()
    right = targetimg_fea.detach().float().contiguous()
    w = conv_weight.detach().float().contiguous()
    sc = scale.detach().float().contiguous()
    sh = shift.detach().float().contiguous()
    a = None
    if att is not None:
        if tuple(att.shape) != (B, num_groups, H, W):
            raise RuntimeError("att must be [B,G,H,W]")
        a = att.detach().float().contiguous()
    out = torch.empty((B, num_groups, int(maxdisp), H, W), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        L.call("as_gwc_corr_stem_fwd", left.data_ptr(), right.data_ptr(), w.data_ptr(), sc.data_ptr(), sh.data_ptr(), L.ptr(a),
               out.data_ptr(), B, C, H, W, int(maxdisp), int(num_groups), float(negative_slope), L.stream_ptr())
    return out if dt == torch.float32 else out.to(dt)


def disparity_regression(x, maxdisp):
    """submodule.py:321-325: sum_d x[:, d] * d  -> [B,1,H,W]."""
    assert len(x.shape) == 4
    if x.shape[1] != maxdisp:
        raise RuntimeError("x must have maxdisp channels")
    if torch.is_grad_enabled() and x.requires_grad:      # training (`--supervise_init`): differentiable formulation
        return (x * torch.arange(maxdisp, device=x.device, dtype=x.dtype).view(1, maxdisp, 1, 1)).sum(1, keepdim=True)
    L.require_cuda(x, "x", torch.float32, contiguous=False)
    x = x.detach().contiguous()
    B, D, H, W = x.shape
    out = torch.empty((B, 1, H, W), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        L.call("as_disparity_regression", x.data_ptr(), out.data_ptr(), B, D, H, W, L.stream_ptr())
    return out


def init_disparity(geo_encoding_volume, classifier_weight, maxdisp=None, return_prob=False):
    """The initial-disparity head in one kernel (continuous_IGEVstereo.py:267-268):
        prob = F.softmax(classifier(geo_encoding_volume).squeeze(1), dim=1); init_disp = disparity_regression(prob, D)
    with ``classifier = nn.Conv3d(G, 1, 3, 1, 1, bias=False)``.  -> init_disp [B,1,H,W] (and prob [B,D,H,W])."""
    if torch.is_grad_enabled() and (geo_encoding_volume.requires_grad or classifier_weight.requires_grad):
        # training (`--supervise_init`, train_continuous_IGEV.py:106): the fused kernel is forward-only, so the same
        # arithmetic runs in differentiable ATen ops and autograd provides the adjoint
        D = geo_encoding_volume.shape[2]
        prob = torch.softmax(torch.nn.functional.conv3d(geo_encoding_volume, classifier_weight, padding=1).squeeze(1), dim=1)
        disp = (prob * torch.arange(D, device=prob.device, dtype=prob.dtype).view(1, D, 1, 1)).sum(1, keepdim=True)
        return (disp, prob) if return_prob else disp
    L.require_cuda(geo_encoding_volume, "geo_encoding_volume", torch.float32, contiguous=False)
    L.require_cuda(classifier_weight, "classifier_weight", torch.float32, contiguous=False)
    g = geo_encoding_volume.detach().contiguous()
    w = classifier_weight.detach().contiguous()
    B, G, D, H, W = g.shape
    if tuple(w.shape) != (1, G, 3, 3, 3):
        raise RuntimeError("classifier weight must be [1,G,3,3,3]")
    if maxdisp is not None and maxdisp != D:
        raise RuntimeError("geo_encoding_volume must have maxdisp disparity planes")
    disp = torch.empty((B, 1, H, W), device=g.device, dtype=torch.float32)
    prob = torch.empty((B, D, H, W), device=g.device, dtype=torch.float32) if return_prob else None
    with torch.cuda.device(g.device):
        L.call("as_init_disparity", g.data_ptr(), w.data_ptr(), disp.data_ptr(), L.ptr(prob), B, G, D, H, W, L.stream_ptr())
    return (disp, prob) if return_prob else disp
