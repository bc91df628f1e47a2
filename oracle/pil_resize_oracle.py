"""numpy restatement of `PIL.Image.resize` for uint8 images with the BICUBIC filter. TEST INFRASTRUCTURE.

The reference resizes every frame with `img.resize(size=(WIDTH, HEIGHT))` (dataset.py:440-451, 617-628, 795-806): for
RGB / L images Pillow's default filter is BICUBIC, with antialiasing (the filter support is stretched by the
downscaling factor). The arithmetic lives in a third-party dependency that is not part of /root/reference: Pillow,
pinned `Pillow==10.0.0` (requirements.txt:5), file src/libImaging/Resample.c. This module restates that published
algorithm (unchanged between Pillow 3.4 and 12.x):

  * precompute_coeffs: for output sample xx, center = (xx + 0.5) * scale, support = 2.0 * max(scale, 1), taps
    xmin = max(0, int(center - support + 0.5)) .. min(in_size, int(center + support + 0.5)), weights
    bicubic((x + xmin - center + 0.5) / max(scale, 1)) normalised to sum 1 (doubles);
  * normalize_coeffs_8bpc: weights -> integers with 22 fractional bits, round half away from zero;
  * ImagingResampleHorizontal_8bpc / Vertical_8bpc: int32 accumulation started at 1 << 21, result >> 22 clipped to
    [0, 255]; the horizontal pass runs first and its output is rounded to uint8 before the vertical pass.

Pinned against the real Pillow of this image by tests/test_oracle.py (random sizes) and by tests/golden/pil_resize.npz.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """-> (bounds int32 [out_size][2] = (xmin, count), coefficients int32 [out_size][ksize]) as Resample.c builds them
    for the full-image box (in0 = 0, in1 = in_size)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """one resampling pass along `axis` (0 = vertical, 1 = horizontal) of an (H, W, C) uint8 array."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img, out_w, out_h):
    """`np.array(Image.fromarray(img).resize((out_w, out_h)))` for a uint8 (H, W) or (H, W, 3) array."""
    a = img[:, :, None] if img.ndim == 2 else img
    h, w = a.shape[:2]
    if w != out_w:
        a = _pass(a, *precompute_coeffs(w, out_w), axis=1)
    if h != out_h:
        a = _pass(a, *precompute_coeffs(h, out_h), axis=0)
    return a[:, :, 0] if img.ndim == 2 else a


def process_frames(imgs, median_chw=None, width=512, height=288, bg_mode=None, median_src=None):
    """`Video_IterableDataset.__process__` / the frame branch of `Shuttlecock_Trajectory_Dataset.__getitem__`
    (reference dataset.py:435-461, 783-812): every frame resized, HWC -> CHW, stacked along the channel axis, divided by
    255 in float64. bg_mode (default: 'concat' when median_chw is given, else ''):
      ''                plain RGB frames;
      'concat'          the (already resized, CHW, uint8) median image first, then the frames;
      'subtract'        one channel per frame: sum_c |frame - median_src| cast to uint8 (numpy's wrap-around cast of the
                        float64 sum, as the reference does it), then resized as an 8-bit grey image;
      'subtract_concat' per frame R, G, B, difference.
    median_src: float64 (Hs, Ws, 3) median at the source resolution (dataset.py:108-109) for the subtract modes."""
    if bg_mode is None:
        bg_mode = 'concat' if median_chw is not None else ''
    frames = np.array([]).reshape(0, height, width)
    for img in imgs:
        if bg_mode in ('subtract', 'subtract_concat'):
            with np.errstate(invalid='ignore'):
                diff = np.sum(np.absolute(img - median_src), 2).astype('uint8')
            diff = resize_bicubic_u8(diff, width, height).reshape(1, height, width)
        if bg_mode == 'subtract':
            img = diff
        else:
            img = np.moveaxis(resize_bicubic_u8(img, width, height), -1, 0)
            if bg_mode == 'subtract_concat':
                img = np.concatenate((img, diff), axis=0)
        frames = np.concatenate((frames, img), axis=0)
    if bg_mode == 'concat':
        frames = np.concatenate((median_chw, frames), axis=0)
    frames /= 255.
    return frames
