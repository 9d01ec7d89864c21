"""TEST INFRASTRUCTURE ONLY -- loader for oracle/_ref/corr_sampler_ref.so, the reference's own CUDA sampler
compiled for sm_100a by oracle/build_ref_sampler.py.  Returns None when the binary is absent."""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "corr_sampler_ref.so")
_mod = None


def load():
    global _mod
    if _mod is None and os.path.exists(_PATH):
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        spec = importlib.util.spec_from_file_location("corr_sampler_ref", _PATH)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mod = m
    return _mod
