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


def get_model(model_name, seq_len=None, bg_mode=None):
    """ Create model by name and the configuration parameter (same table as reference :66-78). """
    if model_name == 'TrackNet':
        if bg_mode == 'subtract':
            model = TrackNet(in_dim=seq_len, out_dim=seq_len)
        elif bg_mode == 'subtract_concat':
            model = TrackNet(in_dim=seq_len * 4, out_dim=seq_len)
        elif bg_mode == 'concat':
            model = TrackNet(in_dim=(seq_len + 1) * 3, out_dim=seq_len)
        else:
            model = TrackNet(in_dim=seq_len * 3, out_dim=seq_len)
    elif model_name == 'InpaintNet':
        model = InpaintNet()
    else:
        raise ValueError('Invalid model name.')
    return model


def to_img(image):
    """ [0, 1] -> uint8 [0, 255] (reference :110-122). """
    image = image * 255
    image = image.astype('uint8')
    return image


def to_img_format(input, num_ch=1):
    """ (N, L*C, H, W) model-input layout -> (N, L, H, W) when num_ch == 1, else (N, L, H, W, 3) taking the
        first three channels of every num_ch-channel frame (behaviour of reference :124-150, vectorised). """
    assert len(input.shape) == 4, 'Input must be 4D tensor.'
    if num_ch == 1:
        return input
    n, lc, h, w = input.shape
    frames = input.reshape(n, lc // num_ch, num_ch, h, w)[:, :, :3]
    return np.ascontiguousarray(np.moveaxis(frames, 2, -1)).astype(np.float64)
