"""numpy restatement of the reference's heatmap decode. TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates `predict_location` (reference test.py:52-79) without OpenCV so that it can travel to the GPU
box: cv2.findContours(RETR_EXTERNAL) + boundingRect + "largest w*h, first wins" becomes
  8-connected components -> bounding boxes -> maximum area, ties to the component whose raster-first
  pixel is LAST (OpenCV lists external contours in reverse raster order of their start pixel).
Pinned against real OpenCV by tests/test_oracle.py (hypothesis masks, when cv2 is importable) and by the
cv2-generated fixtures in tests/golden/decode_golden.npz.
"""
import numpy as np


def to_img(image):
    """reference utils/general.py:110-122"""
    return (image * 255).astype("uint8")


def components(mask):
    """8-connected components of a boolean (H, W) array -> list of (first_pixel_index, x, y, w, h) in
    raster order of the first pixel. Iterative flood fill; meant for test-sized inputs."""
    h, w = mask.shape
    seen = np.zeros_like(mask, dtype=bool)
    out = []
    ys, xs = np.nonzero(mask)
    for y0, x0 in zip(ys.tolist(), xs.tolist()):
        if seen[y0, x0]:
            continue
        stack = [(y0, x0)]
        seen[y0, x0] = True
        minx = maxx = x0
        miny = maxy = y0
        while stack:
            y, x = stack.pop()
            minx, maxx, miny, maxy = min(minx, x), max(maxx, x), min(miny, y), max(maxy, y)
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    yy, xx = y + dy, x + dx
                    if 0 <= yy < h and 0 <= xx < w and mask[yy, xx] and not seen[yy, xx]:
                        seen[yy, xx] = True
                        stack.append((yy, xx))
        out.append((y0 * w + x0, minx, miny, maxx - minx + 1, maxy - miny + 1))
    return out


def predict_location(heatmap):
    """(x, y, w, h) of the selected bounding box; (0, 0, 0, 0) for an empty map (reference test.py:52-79)."""
    heatmap = np.asarray(heatmap)
    if heatmap.size == 0 or np.amax(heatmap) == 0:
        return 0, 0, 0, 0
    comps = components(heatmap != 0)
    # OpenCV order = reverse raster order of the first pixel; the reference keeps the first strict maximum
    best = None
    for first, x, y, w, h in sorted(comps, key=lambda c: -c[0]):
        if best is None or w * h > best[2] * best[3]:
            best = (x, y, w, h)
    return best


def center(bbox):
    """cx, cy = int(x + w/2), int(y + h/2) (reference predict.py:56)."""
    return int(bbox[0] + bbox[2] / 2), int(bbox[1] + bbox[3] / 2)


def decode_batch(y_pred, threshold=0.5):
    """predict.py:35 threshold + per-map predict_location for an array (..., H, W) -> (..., 4) int32."""
    y_pred = np.asarray(y_pred)
    lead = y_pred.shape[:-2]
    maps = y_pred.reshape((-1,) + y_pred.shape[-2:])
    out = np.zeros((maps.shape[0], 4), dtype=np.int32)
    for i, m in enumerate(maps):
        out[i] = predict_location(to_img(m > threshold))
    return out.reshape(lead + (4,))
