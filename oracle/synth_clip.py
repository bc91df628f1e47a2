"""Synthetic video clips and "ball detector" checkpoints for pinning the COMPOSITION of the predict path.

TEST INFRASTRUCTURE (see oracle/__init__.py). The reference ships no checkpoints and no videos, and a randomly
initialised TrackNet in eval() mode produces a constant heatmap (default running statistics let the activations decay
by ~0.4 per layer), so a composition test on random weights would compare nothing. This module builds

  * `make_clip`: a seeded uint8 RGB clip - dark noisy court, one bright ball on a curved path, a few frames where the
    ball is missing (so that `generate_inpaint_mask`, reference test.py:223-258, selects a gap for InpaintNet);
  * `detector_tracknet_state`: a TrackNet state_dict (reference layout, model.py:44-55) whose full-resolution skip path
    (down_block_1 -> up_block_3 -> predictor, model.py:58,69-72) is rewired into "brightness of frame f minus the
    background" -> heatmap f, with logits at least 0.1 away from 0 at every pixel (the input is uint8 / 255, the
    weights are +-1 and K): the thresholded maps of a 1e-5-accurate implementation are then bit-identical to the
    reference's, and the decoded coordinates can be compared exactly. All other weights keep torch's default
    initialisation, so the whole network still executes.

oracle/gen_predict_flow.py runs the REAL reference predict.py `__main__` on these in the build container and stores its
prediction dictionaries in tests/golden/predict_flow.npz; tests/test_gpu_predict_flow.py runs this repo's predict path
on the same clip and weights and compares bit for bit.
"""
import numpy as np
import torch

from . import tracknet_oracle as O


def make_clip(n_frames, hs=360, ws=640, seed=0, gap=(5, 6)):
    """(n_frames, hs, ws, 3) uint8 RGB. np.random.RandomState is frozen by numpy's compatibility policy, so the clip is
    reproducible from its seed on any box."""
    rs = np.random.RandomState(seed)
    video = rs.randint(0, 40, size=(n_frames, hs, ws, 3)).astype(np.uint8)
    for i in range(n_frames):
        if i in gap:
            continue
        y0 = 60 + 11 * i + (i * i) // 3
        x0 = 40 + 29 * i
        video[i, y0:y0 + 9, x0:x0 + 9] = 250
        video[i, y0 + 2:y0 + 7, x0 - 2:x0 + 11] = 250          # a blob that is not a square
    return video


def detector_tracknet_state(seq_len, bg_mode, seed=0, gain=8.0, level=(306 + 0.5) / 255):
    """TrackNet state_dict: heatmap f = sigmoid(gain * (sum_rgb(frame f) - sum_rgb(background) - level)). The sums are
    integers / 255, the level sits half way between two of them: |logit| >= gain * 0.5 / 255 on the detector path."""
    if bg_mode not in ('', 'concat'):
        raise ValueError("detector checkpoints exist for bg_mode '' and 'concat'")
    in_dim = 3 * seq_len + (3 if bg_mode == 'concat' else 0)
    sd = O.init_tracknet_state(seed, in_dim, seq_len)
    first = 3 if bg_mode == 'concat' else 0

    def identity(key, in_off):
        w = sd[key]
        w[:seq_len] = 0
        for f in range(seq_len):
            w[f, in_off + f, 1, 1] = 1.0

    w = sd['down_block_1.conv_1.conv.weight']
    w[:seq_len] = 0
    for f in range(seq_len):
        w[f, first + 3 * f:first + 3 * f + 3, 1, 1] = 1.0
        if bg_mode == 'concat':
            w[f, 0:3, 1, 1] = -1.0
    identity('down_block_1.conv_2.conv.weight', 0)
    identity('up_block_3.conv_1.conv.weight', 128)             # the skip half of cat([up(x), x1]) (model.py:69)
    identity('up_block_3.conv_2.conv.weight', 0)
    pw = sd['predictor.weight']
    pw *= 0.05                                                 # the deep path still contributes, far below the margin
    for f in range(seq_len):
        pw[f, :seq_len] = 0
        pw[f, f, 0, 0] = gain
    sd['predictor.bias'][:] = -gain * level
    return sd


def inpaintnet_state(seed=0):
    return O.init_inpaintnet_state(seed)


CASES = [
    # name, frames, (hs, ws), TrackNet seq_len, bg_mode, InpaintNet seq_len, batch_size, gap frames
    ("odd13", 13, (360, 640), 8, 'concat', 6, 4, (5, 6)),
    ("even14", 14, (360, 640), 8, 'concat', 6, 4, (5, 6)),
    ("nobg11", 11, (288, 512), 4, '', 5, 3, (4,)),
]
EVAL_MODES = ('nonoverlap', 'average', 'weight')
