"""TEST INFRASTRUCTURE ONLY -- loader for the *real* Any-Stereo reference.

Imports the unmodified reference modules from ``/root/reference`` (read-only)
so that ``tests/golden/make_golden.py`` can generate golden vectors and so the
oracle restatement in ``oracle/hotpath_oracle.py`` can be pinned against the
reference's own code.  ``/root/reference`` exists only in the build container;
on the GPU box ``available()`` is False and nothing here is used.

The reference cannot be imported as shipped:
  * ``models/__init__.py:2-3`` eagerly imports both networks,
  * ``update.py:4`` imports the absent ``opt_einsum`` (never called),
  * ``extractor.py:5`` / ``liif.py:6`` import the absent ``timm``,
  * ``corePrune_RAFT/liif.py:5`` imports a package that does not exist,
  * ``liif.py`` hard-codes ``.cuda()``.
All of that is routed around here with stub modules; no reference semantics on
the hot path are altered.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("ANYSTEREO_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models", "coreContinuous_IGEV"))


_loaded = {}


def load():
    """Return a namespace with the reference's hot-path callables."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # third-party stubs that are imported but never used on the hot path
    if "opt_einsum" not in sys.modules:
        oe = types.ModuleType("opt_einsum")
        oe.contract = None
        sys.modules["opt_einsum"] = oe
    if "timm" not in sys.modules:
        sys.modules["timm"] = types.ModuleType("timm")
    # bare packages so models/__init__.py is never executed
    for pkg in ("models", "models.corePrune_RAFT", "models.coreContinuous_IGEV",
                "models.corePrune_RAFT.utils", "models.coreContinuous_IGEV.utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF_ROOT, pkg.replace(".", "/"))]
            sys.modules[pkg] = m
    import models.coreContinuous_IGEV.submodule as igev_sub
    if "models.coreContinuous_A2A4IGEV" not in sys.modules:
        sys.modules["models.coreContinuous_A2A4IGEV"] = types.ModuleType("models.coreContinuous_A2A4IGEV")
        sys.modules["models.coreContinuous_A2A4IGEV.submodule"] = igev_sub
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    from models.corePrune_RAFT.geometry import CorrBlock1D
    from models.coreContinuous_IGEV.geometry import Combined_Geo_Encoding_Volume
    from models.coreContinuous_IGEV.submodule import build_gwc_volume
    from models.coreContinuous_IGEV.update import BasicMultiUpdateBlock as IGEVUpdateBlock
    from models.corePrune_RAFT.update import BasicMultiUpdateBlock as RAFTUpdateBlock
    from models.coreContinuous_IGEV.utils.utils import bilinear_sampler

    _loaded.update(
        CorrBlock1D=CorrBlock1D,
        Combined_Geo_Encoding_Volume=Combined_Geo_Encoding_Volume,
        build_gwc_volume=build_gwc_volume,
        IGEVUpdateBlock=IGEVUpdateBlock,
        RAFTUpdateBlock=RAFTUpdateBlock,
        bilinear_sampler=bilinear_sampler,
    )
    return types.SimpleNamespace(**_loaded)


def update_block_args(family: str, corr_levels=None, corr_radius=4, n_gru_layers=3):
    """argparse.Namespace-alike with the fields update.py reads
    (update.py:77,111-113,121,127)."""
    if corr_levels is None:
        corr_levels = 2 if family == "igev" else 4
    return types.SimpleNamespace(corr_levels=corr_levels, corr_radius=corr_radius,
                                 n_gru_layers=n_gru_layers)
