"""TEST INFRASTRUCTURE ONLY -- loader for the *real* Any-Stereo reference.

Imports the unmodified reference modules from ``/root/reference`` (read-only)
so that ``tests/golden/make_golden.py`` can generate golden vectors and so the
oracle restatement in ``oracle/hotpath_oracle.py`` can be pinned against the
reference's own code.  ``/root/reference`` exists only in the build container;
on the GPU box the pristine copy under the git-ignored ``baseline/_ref/`` (oracle/install_ref.py) is used instead.

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

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """$ANYSTEREO_REFERENCE, else /root/reference (build container), else the pristine copy that
    oracle/install_ref.py placed under the git-ignored baseline/_ref/ (the only one present on the GPU box)."""
    for cand in (os.environ.get("ANYSTEREO_REFERENCE"), "/root/reference",
                 os.path.join(os.path.dirname(_HERE), "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "models", "coreContinuous_IGEV")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


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


def model_args(family: str, **over):
    """The argparse fields the reference model constructors / forwards read, at the defaults SURVEY.md 8(d) lists
    (train_continuous_IGEV.py:284-369, train_continuous_Raft.py, evaluation.py:556-674)."""
    a = dict(
        hidden_dims=[128] * 3, n_gru_layers=3, n_downsample=2, corr_radius=4, corr_levels=2 if family == "igev" else 4,
        slow_fast_gru=False, agg_type="type5", multi_training=True, multi_input_training=False,
        unfold_similarity="with_v2ISU", mlphidden_list=[128, 64, 64], pos_dim=0, pos_enconding=False,
        pos_enconding_new=False, local_ensemble=False, decode_cell=False, lsp_width=3, lsp_height=3,
        lsp_dilation=[1, 2, 4, 8], quater_nearest=None, require_grad=False, disparity_norm=False, disparity_norm2=False,
        mixed_precision=False, max_disp=192, Raw_Mask_dim=32, unfold=False, corr_implementation="reg",
        shared_backbone=False)
    a.update(over)
    return types.SimpleNamespace(**a)


def timm_shim():
    """timm is absent: a shape-identical MobileNetV2 from torchvision regrouped into the attributes
    extractor.py:331-343 reads (conv_stem, bn1, act1, blocks[0..6]); SURVEY.md 8c.  Off the hot path, random weights."""
    import timm
    import torch
    import torchvision

    def create_model(name, pretrained=True, features_only=True):
        f = torchvision.models.mobilenet_v2(weights=None).features
        m = types.SimpleNamespace()
        m.conv_stem, m.bn1, m.act1 = f[0][0], f[0][1], f[0][2]
        groups = [[1], [2, 3], [4, 5, 6], [7, 8, 9, 10], [11, 12, 13], [14, 15, 16], [17]]
        m.blocks = [torch.nn.Sequential(*[f[i] for i in g]) for g in groups]
        return m

    timm.create_model = create_model


def load_models():
    """The two unmodified reference model graphs + their modules (for rebinding names) + make_coord."""
    load()
    timm_shim()
    from models.coreContinuous_IGEV import continuous_IGEVstereo as cis
    from models.corePrune_RAFT import prune_raft_stereo as prs
    from models.coreContinuous_IGEV.liif import make_coord
    return types.SimpleNamespace(igev_module=cis, raft_module=prs, IGEV=cis.continuous_IGEVStereo,
                                 RAFT=prs.continuous_RaftStereo, make_coord=make_coord)
