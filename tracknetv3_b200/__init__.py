"""tracknetv3_b200: B200-native (sm_100a) implementation of the TrackNetV3 data-parallel hot path."""
from .model import TrackNet, InpaintNet  # noqa: F401
from .metric import WBCELoss, get_metric  # noqa: F401
from .decode import decode_heatmaps, predict_location, bbox_to_center  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .data import DevicePrefetcher, ScalarReader, label_discs  # noqa: F401
from .ensemble import TemporalEnsemble, get_ensemble_weight  # noqa: F401
from .frames import FramePreprocessor, resample_table  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["TrackNet", "InpaintNet", "WBCELoss", "get_metric", "decode_heatmaps", "predict_location",
           "bbox_to_center", "FusedAdam", "DevicePrefetcher", "ScalarReader", "label_discs", "TemporalEnsemble", "FramePreprocessor"]
