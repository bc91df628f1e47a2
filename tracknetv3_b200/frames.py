"""GPU frame preprocessing (reference dataset.py:435-461 / 611-647 / 783-812): the per-frame PIL resize, HWC -> CHW,
/ 255 and channel stacking that the reference runs in DataLoader workers, as two integer kernels on the device.

The resize is Pillow's `Image.resize((W, H))` for uint8 RGB images - antialiased BICUBIC, 22-bit fixed point,
horizontal pass then vertical pass (Pillow 10.0.0 `src/libImaging/Resample.c`, the reference's pin) - reproduced
bit for bit: the coefficient tables are built here exactly as `precompute_coeffs` / `normalize_coeffs_8bpc` build them
(double arithmetic, same truncations), the kernels accumulate in int32 and clip like `clip8`.
"""
import math

import numpy as np
import torch

from . import _lib

_PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = -x if x < 0.0 else x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_table(in_size, out_size):
    """(bounds int32 [out][2], coefficients int32 [out][ksize]) of one resampling pass; identity when the sizes match
    (Pillow skips such a pass)."""
    if in_size == out_size:
        bounds = np.stack([np.arange(out_size), np.ones(out_size)], 1).astype(np.int32)
        return bounds, np.full((out_size, 1), 1 << _PRECISION_BITS, dtype=np.int32)
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


class FramePreprocessor:
    """Resize + normalise + stack frames of one source size on the GPU.

    ``process(imgs, median)``: ``imgs`` uint8 CUDA ``(N, L, Hs, Ws, 3)`` RGB frames (what the reference hands to
    ``Image.fromarray``), ``median`` optional uint8 ``(3, H, W)`` background already resized as at dataset.py:104-107
    (``prepare_median`` does that) -> float32 ``(N, 3 * L [+ 3], H, W)`` in [0, 1], median channels first (bg_mode
    'concat', dataset.py:636-640), exactly the reference's ``frames / 255.`` cast to float32.
    """

    def __init__(self, src_h, src_w, height=288, width=512, device=None):
        self.src_h, self.src_w, self.h, self.w = src_h, src_w, height, width
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        hb, hk = resample_table(src_w, width)
        vb, vk = resample_table(src_h, height)
        self.hb, self.hk = (dev(hb), dev(hk)) if src_w != width else (None, None)
        self.vb, self.vk = dev(vb), dev(vk)

    def _run(self, imgs, out, per_sample, chan_off, frame_stride=None):
        lib = _lib.load()
        _lib.require_cuda(imgs, out)
        if imgs.dtype != torch.uint8 or imgs.dim() != 4 or imgs.shape[1] != self.src_h or imgs.shape[2] != self.src_w:
            raise RuntimeError(f"FramePreprocessor expects uint8 (n, {self.src_h}, {self.src_w}, C), got {tuple(imgs.shape)}")
        imgs = imgs.contiguous()
        n, c = imgs.shape[0], imgs.shape[3]
        tmp = torch.empty((n, self.src_h, self.w, c), dtype=torch.uint8, device=imgs.device)
        _lib.check(lib.tnb_resize_frames(
            imgs.data_ptr(), n, self.src_h, self.src_w, c,
            self.hb.data_ptr() if self.hb is not None else None, self.hk.data_ptr() if self.hk is not None else None,
            self.hk.shape[1] if self.hk is not None else 0, self.vb.data_ptr(), self.vk.data_ptr(), self.vk.shape[1],
            self.h, self.w, tmp.data_ptr(), out.data_ptr(), per_sample, out.stride(0), chan_off,
            c if frame_stride is None else frame_stride, _lib.stream_ptr()))

    def median(self, frames_u8, as_float=False):
        """`np.median(frame_arr, 0)` of a whole clip (dataset.py:102-103) on the GPU: ``frames_u8`` uint8 CUDA
        ``(T, Hs, Ws, 3)``. Returns the uint8 image ``median.astype('uint8')`` that bg_mode 'concat' resizes (:105; pass it
        to ``prepare_median``), or with ``as_float`` numpy's float64 median, x.5 values included, that the subtract modes
        keep (:108-109; pass it to ``process`` as it is)."""
        lib = _lib.load()
        _lib.require_cuda(frames_u8)
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or tuple(frames_u8.shape[1:3]) != (self.src_h, self.src_w):
            raise RuntimeError(f"median expects uint8 (T, {self.src_h}, {self.src_w}, C), got {tuple(frames_u8.shape)}")
        frames_u8 = frames_u8.contiguous()
        t, per = frames_u8.shape[0], frames_u8[0].numel()
        out = torch.empty(frames_u8.shape[1:], dtype=torch.float64 if as_float else torch.uint8, device=frames_u8.device)
        _lib.check(lib.tnb_median_u8(frames_u8.data_ptr(), t, per, out.data_ptr() if as_float else None,
                                     None if as_float else out.data_ptr(), _lib.stream_ptr()))
        return out

    def prepare_median(self, median_hwc):
        """dataset.py:104-107 / :776-779: the median frame (any float or uint8 (Hs, Ws, 3)) as uint8, resized, CHW."""
        m = median_hwc.to(self.device) if torch.is_tensor(median_hwc) else torch.as_tensor(np.asarray(median_hwc)).to(self.device)
        m = m.to(torch.uint8) if m.dtype != torch.uint8 else m   # .astype('uint8') of values in [0, 255]: truncation
        out = torch.empty((1, 3, self.h, self.w), dtype=torch.float32, device=self.device)
        self._run(m.unsqueeze(0), out, 1, 0)
        return torch.round(out[0] * 255).to(torch.uint8)

    def _difference(self, flat, median_src):
        """np.sum(np.absolute(img - median), 2).astype('uint8') for every frame (dataset.py:438, 442): (n, Hs, Ws, 1)."""
        lib = _lib.load()
        if torch.is_tensor(median_src):
            med = median_src.to(self.device, torch.float64).contiguous()
        else:
            med = torch.as_tensor(np.asarray(median_src, dtype=np.float64)).to(self.device).contiguous()
        if tuple(med.shape) != (self.src_h, self.src_w, 3):
            raise RuntimeError(f"median must be ({self.src_h}, {self.src_w}, 3) at the source resolution, got {tuple(med.shape)}")
        diff = torch.empty((flat.shape[0], self.src_h, self.src_w, 1), dtype=torch.uint8, device=self.device)
        _lib.check(lib.tnb_bg_subtract_u8(flat.data_ptr(), med.data_ptr(), flat.shape[0], self.src_h, self.src_w,
                                          diff.data_ptr(), _lib.stream_ptr()))
        return diff

    def process(self, imgs, median=None, bg_mode=None, out=None):
        """bg_mode: '' | 'concat' (median = prepare_median(...) output) | 'subtract' | 'subtract_concat' (median = the
        float64 (Hs, Ws, 3) median at the source resolution, dataset.py:108-109). Default: 'concat' if a median is
        given, else ''. Output channels: 3L, 3L + 3, L, 4L (utils/general.py:66-74 get_model's in_dim). ``out``: optional
        preallocated result tensor (a step loop that cycles two of them allocates nothing)."""
        if bg_mode is None:
            bg_mode = 'concat' if median is not None else ''
        if bg_mode not in ('', 'concat', 'subtract', 'subtract_concat'):
            raise ValueError(f"unknown bg_mode {bg_mode!r}")
        if bg_mode != '' and median is None:
            raise ValueError(f"bg_mode {bg_mode!r} needs the median image")
        _lib.require_cuda(imgs)
        n, l = imgs.shape[0], imgs.shape[1]
        flat = imgs.reshape(n * l, *imgs.shape[2:]).contiguous()
        per_frame = {'': 3, 'concat': 3, 'subtract': 1, 'subtract_concat': 4}[bg_mode]
        extra = 3 if bg_mode == 'concat' else 0
        shape = (n, per_frame * l + extra, self.h, self.w)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.device:
            raise RuntimeError(f"FramePreprocessor.process: out must be a contiguous float32 {shape} tensor on {self.device}")
        if bg_mode in ('subtract', 'subtract_concat'):
            self._run(self._difference(flat, median), out, l, per_frame - 1, per_frame)
        if bg_mode != 'subtract':
            self._run(flat, out, l, extra, per_frame)
        if bg_mode == 'concat':
            out[:, :3] = (median.to(self.device).double() / 255.0).float()
        return out
