"""GPU temporal ensemble of sliding-window predictions (reference predict.py:163-209 and :245-301, test.py:635-692):
the reference copies every heatmap to the host and combines frames in a per-frame Python loop over a growing
torch.cat buffer; here the predictions stay on the device and one kernel per batch produces the ensembled frames,
which can feed `decode_heatmaps` directly."""
import ctypes as C
import math

import torch

from . import _lib


def get_ensemble_weight(seq_len, eval_mode):
    """ Weights for the temporal ensemble (reference test.py:25-50): uniform ('average') or positional ('weight'). """
    if eval_mode == 'average':
        return torch.ones(seq_len) / seq_len
    if eval_mode == 'weight':
        weight = torch.ones(seq_len)
        for i in range(math.ceil(seq_len / 2)):
            weight[i] = (i + 1)
            weight[seq_len - i - 1] = (i + 1)
        return weight / weight.sum()
    raise ValueError('Invalid mode')


class TemporalEnsemble:
    """Streaming ensemble over samples taken with sliding_step 1: sample s predicts frames s .. s+seq_len-1.

    ``push(pred)`` takes the next batch of predictions, CUDA ``(B, seq_len, ...)`` (heatmaps ``(B, L, H, W)`` or
    coordinates ``(B, L, 2)``), and returns the ensembled frames that batch completes, ``(B, ...)`` - plus the last
    ``seq_len-1`` frames once ``num_sample`` samples have been pushed (reference predict.py:192-201).
    """

    def __init__(self, seq_len, eval_mode, num_sample):
        if not 1 <= seq_len <= 16:
            raise ValueError('seq_len must be in 1..16')
        self.seq_len, self.num_sample, self.sample_count = seq_len, num_sample, 0
        self.weight = get_ensemble_weight(seq_len, eval_mode).float().contiguous()
        self._w = (C.c_float * seq_len)(*self.weight.tolist())
        self.state = None

    def push(self, pred):
        _lib.require_cuda(pred)
        lib = _lib.load()
        L, S = self.seq_len, self.seq_len - 1
        if pred.dim() < 3 or pred.shape[1] != L:
            raise RuntimeError(f"TemporalEnsemble expects (B, {L}, ...), got {tuple(pred.shape)}")
        pred = pred.contiguous().float()
        b = pred.shape[0]
        if self.sample_count + b > self.num_sample:
            raise RuntimeError("TemporalEnsemble: more samples pushed than num_sample")
        frame_shape = tuple(pred.shape[2:])
        elems = int(math.prod(frame_shape))
        if self.state is None:
            self.state = torch.zeros((S, L) + frame_shape, dtype=torch.float32, device=pred.device)
        last = self.sample_count + b == self.num_sample
        n_tail = S if last else 0
        out = torch.empty((b + n_tail,) + frame_shape, dtype=torch.float32, device=pred.device)
        _lib.check(lib.tnb_temporal_ensemble(self.state.data_ptr(), pred.data_ptr(), out.data_ptr(), self._w, L, elems, b,
                                             self.sample_count, b - 1, n_tail, _lib.stream_ptr()))
        self.sample_count += b
        if S > 0:  # keep the newest seq_len-1 samples for the next batch
            self.state = pred[b - S:].clone() if b >= S else torch.cat((self.state[b:], pred), dim=0)
        return out
