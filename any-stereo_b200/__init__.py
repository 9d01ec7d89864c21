"""anystereo_b200 -- B200-native (sm_100a) implementation of Any-Stereo's iterative cost-volume hot path.

The public names are the reference's operator surface:

    CorrBlock1D                    models/corePrune_RAFT/geometry.py
    Combined_Geo_Encoding_Volume   models/coreContinuous_IGEV/geometry.py
    build_gwc_volume               models/coreContinuous_IGEV/submodule.py
    corr_sampler.forward/backward  sampler/ (pybind module `corr_sampler`)
    BasicMultiUpdateBlock          models/*/update.py

Everything computes in hand-written CUDA kernels behind the C ABI of include/anystereo_b200.h
(csrc/libanystereo_b200.so).  There is no CPU path: importing works anywhere, calling needs a B200.
"""
from . import _lib
from . import corr_sampler
from . import submodule
from . import hotpath
from . import update_umma
from .update_umma import set_lowres_single_pass, set_gate_weight_residual_only, set_call_replay
from .geometry import CorrBlock1D, Combined_Geo_Encoding_Volume, set_corr_mode, get_corr_mode
from .submodule import build_gwc_volume, disparity_regression, init_disparity, gwc_corr_stem, DeferredGwcVolume
from .update import (BasicMultiUpdateBlock, BasicMultiUpdateBlockRAFT, BasicMotionEncoder, ConvGRU, DispHead,
                     set_update_engine, get_update_engine)
from .hotpath import (igev_iterations, raft_iterations, install_into_reference, HotLoopGraph, set_lookup_fusion,
                      adopt_update_block, adopt_liif_up, adopt_corr_stem, adopt_model, set_graph_replay)
from .parallel import shard_pairs, allreduce_gradients, GradientAllReducer
from . import liif
from . import extractor
from .extractor import (adopt_context_encoder, adopt_feature_encoder, ContextEncoder, FeatureEncoder, fold_basic_convs,
                        unfold_basic_convs, FoldedBasicConv)
from .liif import liif_out_multi_scale_Training, context_upsample_multiscale_train, upsample_disp

__all__ = [
    "CorrBlock1D", "Combined_Geo_Encoding_Volume", "build_gwc_volume", "corr_sampler",
    "BasicMultiUpdateBlock", "BasicMultiUpdateBlockRAFT", "BasicMotionEncoder", "ConvGRU", "DispHead",
    "set_corr_mode", "get_corr_mode", "set_update_engine", "get_update_engine",
    "igev_iterations", "raft_iterations", "install_into_reference", "HotLoopGraph",
    "shard_pairs", "allreduce_gradients", "GradientAllReducer",
    "gwc_corr_stem", "adopt_corr_stem", "adopt_update_block", "adopt_liif_up", "set_graph_replay",
    "adopt_model", "adopt_context_encoder", "adopt_feature_encoder", "fold_basic_convs", "unfold_basic_convs",
    "set_call_replay", "set_lowres_single_pass", "set_gate_weight_residual_only",
]
