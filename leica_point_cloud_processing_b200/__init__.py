"""B200-native GICP registration engine: the hot path of catec/leica_point_cloud_processing
(GICPAlignment + Filter::removeFromCloud) behind the reference's own class surface.

CUDA (sm_100a) does all the work through libgicp_b200.so; there is no CPU implementation in this package.
"""
from ._capi import Engine, EngineGroup, GicpError, Params, load_library, EXPORTED_SYMBOLS, LIB_PATH  # noqa: F401
from .gicp_alignment import GICPAlignment, remove_from_cloud, is_valid_transform  # noqa: F401
from .fod_detector import FODDetector, downsample_cloud  # noqa: F401

__all__ = ["Engine", "EngineGroup", "GicpError", "Params", "GICPAlignment", "remove_from_cloud", "is_valid_transform", "FODDetector",
           "downsample_cloud",
           "load_library", "EXPORTED_SYMBOLS", "LIB_PATH"]
