"""``corr_sampler`` -- drop-in for the reference's pybind extension module (sampler/sampler.cpp:48-51).

    corr_sampler.forward(volume, coords, radius)              -> [corr]
    corr_sampler.backward(volume, coords, corr_grad, radius)  -> [volume_grad]

Same argument order, same list-of-one-tensor returns, same error behaviour
("x must be a CUDA tensor" / "x must be contiguous", sampler.cpp:20-22).  Unlike the reference kernels these
run on the caller's current stream, write their outputs exactly once and check the launch.
"""
from __future__ import annotations

import torch

from . import _lib as L

_DT = {torch.float32: L.DTYPE_F32, torch.float16: L.DTYPE_F16, torch.float64: L.DTYPE_F64}


def _check(volume, coords):
    L.require_cuda(volume, "volume")
    L.require_cuda(coords, "coords")
    if volume.dtype not in _DT:
        raise RuntimeError("volume must be float32, float16 or float64 (AT_DISPATCH_FLOATING_TYPES_AND_HALF)")
    if coords.dtype != torch.float32:
        raise RuntimeError("coords must be float32")
    if volume.dim() != 4 or coords.dim() != 4:
        raise RuntimeError("volume must be [B,H,W1,W2] and coords [B,C,H,W1]")
    B, H, W1, W2 = volume.shape
    if coords.shape[0] != B or coords.shape[2] != H or coords.shape[3] != W1 or coords.shape[1] < 1:
        raise RuntimeError("coords shape %s does not match volume %s" % (tuple(coords.shape), tuple(volume.shape)))
    if volume.numel() >= 2 ** 31:
        raise RuntimeError("volume exceeds 32-bit indexing (PackedTensorAccessor32 in the reference)")
    return B, H, W1, W2


def forward(volume, coords, radius):
    """sampler/sampler.cpp:24-32 -> sampler_kernel.cu:107-136."""
    B, H, W1, W2 = _check(volume, coords)
    radius = int(radius)
    with torch.cuda.device(volume.device):
        out = torch.empty((B, 2 * radius + 1, H, W1), device=volume.device, dtype=volume.dtype)
        L.call("as_sampler_fwd", volume.data_ptr(), coords.data_ptr(), coords.shape[1], out.data_ptr(), B, H, W1,
               W2, radius, _DT[volume.dtype], L.stream_ptr())
    return [out]


def backward(volume, coords, corr_grad, radius):
    """sampler/sampler.cpp:34-45 -> sampler_kernel.cu:138-166 (adjoint w.r.t. the volume only)."""
    B, H, W1, W2 = _check(volume, coords)
    radius = int(radius)
    L.require_cuda(corr_grad, "corr_grad")
    if corr_grad.dtype != volume.dtype or tuple(corr_grad.shape) != (B, 2 * radius + 1, H, W1):
        raise RuntimeError("corr_grad must be [B,2r+1,H,W1] with the volume's dtype")
    with torch.cuda.device(volume.device):
        g = torch.empty_like(volume)
        L.call("as_sampler_bwd", coords.data_ptr(), coords.shape[1], corr_grad.data_ptr(), g.data_ptr(), B, H, W1,
               W2, radius, _DT[volume.dtype], L.stream_ptr())
    return [g]


class CorrSampler(torch.autograd.Function):
    """The autograd wrapper RAFT-Stereo pairs with this extension (absent from the reference)."""

    @staticmethod
    def forward(ctx, volume, coords, radius):
        ctx.save_for_backward(volume, coords)
        ctx.radius = radius
        corr, = forward(volume, coords, radius)
        return corr

    @staticmethod
    def backward(ctx, grad_output):
        volume, coords = ctx.saved_tensors
        grad_volume, = backward(volume, coords, grad_output.contiguous(), ctx.radius)
        return grad_volume, None, None
