"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/sampler_oracle.c (C restatement of
sampler/sampler_kernel.cu).  Build with `make -C oracle`."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libsampler_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = C.CDLL(_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def forward(volume: np.ndarray, coords: np.ndarray, r: int) -> np.ndarray:
    volume = np.ascontiguousarray(volume, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    B, H, W1, W2 = volume.shape
    out = np.empty((B, 2 * r + 1, H, W1), np.float32)
    lib().sampler_forward_f32(_p(volume), _p(coords), C.c_int(coords.shape[1]), _p(out), B, H, W1, W2, r)
    return out


def backward(volume_shape, coords: np.ndarray, grad: np.ndarray, r: int) -> np.ndarray:
    coords = np.ascontiguousarray(coords, np.float32)
    grad = np.ascontiguousarray(grad, np.float32)
    B, H, W1, W2 = volume_shape
    out = np.empty((B, H, W1, W2), np.float32)
    lib().sampler_backward_f32(_p(coords), C.c_int(coords.shape[1]), _p(grad), _p(out), B, H, W1, W2, r)
    return out
