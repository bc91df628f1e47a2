"""Hot-path subset of the reference's utils/general.py: constants (:15-19), get_model (:46-80), to_img
(:110-122), to_img_format (:124-150). File/video/csv IO of the reference is out of scope."""
import math

import numpy as np

from model import TrackNet, InpaintNet

HEIGHT = 288
WIDTH = 512
SIGMA = 2.5
DELTA_T = 1 / math.sqrt(HEIGHT ** 2 + WIDTH ** 2)
COOR_TH = DELTA_T * 50
IMG_FORMAT = 'png'


# input channels of TrackNet per bg_mode: (channels per frame, extra channels for the median image)
_FRAME_CHANNELS = {'subtract': (1, 0), 'subtract_concat': (4, 0), 'concat': (3, 3)}


def get_model(model_name, seq_len=None, bg_mode=None):
    """ Create a model by name (same choices and input-channel table as reference :66-78): 'TrackNet' with L x 3 input
        channels for RGB frames, L x 1 for 'subtract', L x 4 for 'subtract_concat', (L + 1) x 3 for 'concat', and L
        heatmaps out; 'InpaintNet'; anything else raises ValueError('Invalid model name.'). """
    if model_name == 'InpaintNet':
        return InpaintNet()
    if model_name != 'TrackNet':
        raise ValueError('Invalid model name.')
    per_frame, extra = _FRAME_CHANNELS.get(bg_mode, (3, 0))
    return TrackNet(in_dim=seq_len * per_frame + extra, out_dim=seq_len)


def to_img(image):
    """ [0, 1] -> uint8 [0, 255] (reference :110-122). """
    return (image * 255).astype('uint8')


def to_img_format(input, num_ch=1):
    """ (N, L*C, H, W) model-input layout -> (N, L, H, W) when num_ch == 1, else (N, L, H, W, 3) taking the
        first three channels of every num_ch-channel frame (behaviour of reference :124-150, vectorised). """
    assert len(input.shape) == 4, 'Input must be 4D tensor.'
    if num_ch == 1:
        return input
    n, lc, h, w = input.shape
    frames = input.reshape(n, lc // num_ch, num_ch, h, w)[:, :, :3]
    return np.ascontiguousarray(np.moveaxis(frames, 2, -1)).astype(np.float64)
