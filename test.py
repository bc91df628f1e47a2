"""Hot-path subset of the reference's test.py: get_ensemble_weight (:25-50) and predict_location (:52-79).
The evaluation drivers of the reference (dataset loops, COCO export) are out of scope."""
import math

import torch

from tracknetv3_b200.decode import predict_location, decode_heatmaps  # noqa: F401


def get_ensemble_weight(seq_len, eval_mode):
    """ Weights for the temporal ensemble: uniform ('average') or triangular ('weight'), sum 1. """
    if eval_mode == 'average':
        return torch.full((seq_len,), 1.0 / seq_len)
    if eval_mode == 'weight':
        half = torch.arange(1, seq_len + 1, dtype=torch.float32)
        weight = torch.minimum(half, half.flip(0))
        return weight / weight.sum()
    raise ValueError('Invalid mode')
