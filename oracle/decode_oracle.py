"""numpy restatement of the reference's heatmap decode. TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates `predict_location` (reference test.py:52-79) without OpenCV so that it can travel to the GPU
box: cv2.findContours(RETR_EXTERNAL) + boundingRect + "largest w*h, first wins" becomes
  8-connected components -> bounding boxes -> maximum area, ties to the component whose raster-first
  pixel is LAST (OpenCV lists external contours in reverse raster order of their start pixel).
Pinned against real OpenCV by tests/test_oracle.py (hypothesis masks, when cv2 is importable) and by the
cv2-generated fixtures in tests/golden/decode_golden.npz.
"""
import math

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


PRED_TYPES = ['TP', 'TN', 'FP1', 'FP2', 'FN']           # reference test.py:20-21
PRED_TYPES_MAP = {t: i for i, t in enumerate(PRED_TYPES)}


def classify(pred_any, true_any, cx_pred, cy_pred, cx_true, cy_true, tolerance):
    """Outcome of one frame (reference test.py:135-157 / :171-192): TN / FP2 / FN by presence, else FP1 when the
    predicted centre is farther than `tolerance` from the true one, else TP."""
    if not pred_any and not true_any:
        return PRED_TYPES_MAP['TN']
    if pred_any and not true_any:
        return PRED_TYPES_MAP['FP2']
    if not pred_any and true_any:
        return PRED_TYPES_MAP['FN']
    dist = math.sqrt(pow(cx_pred - cx_true, 2) + pow(cy_pred - cy_true, 2))
    return PRED_TYPES_MAP['FP1'] if dist > tolerance else PRED_TYPES_MAP['TP']


def evaluate(indices, y_true=None, y_pred=None, c_true=None, c_pred=None, tolerance=4., img_scaler=(1, 1),
             output_bbox=False, output_gt=False, height=288, width=512):
    """Per-frame evaluation (reference test.py:81-221) on numpy arrays: heatmap mode decodes prediction (> 0.5) and
    ground truth (to_img != 0) with predict_location, confidence = max of the heatmap inside the predicted box;
    coordinate mode scales normalised coordinates by (width, height). Frames repeat at the end of a padded sample:
    the first repeated index ends that sample (:130-133, :211-212)."""
    d = {'Frame': [], 'X': [], 'Y': [], 'Visibility': [], 'Type': [], 'BBox': [], 'Confidence': [], 'X_GT': [],
         'Y_GT': [], 'Visibility_GT': []}
    indices = np.asarray(indices).tolist()
    heat = y_true is not None and y_pred is not None
    if heat:
        y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
        h_pred = y_pred > 0.5
    else:
        c_true = np.asarray(c_true, dtype=np.float32) * np.array([width, height], dtype=np.float32)
        c_pred = np.asarray(c_pred, dtype=np.float32) * np.array([width, height], dtype=np.float32)
    for n in range(len(indices)):
        prev = [-1, -1]
        for f in range(len(indices[n])):
            d_i = indices[n][f]
            if d_i == prev:
                break
            if heat:
                bt = predict_location(to_img(y_true[n][f]))
                cx_true, cy_true = int(bt[0] + bt[2] / 2), int(bt[1] + bt[3] / 2)
                bp = predict_location(to_img(h_pred[n][f]))
                cx_pred, cy_pred = int(bp[0] + bp[2] / 2), int(bp[1] + bp[3] / 2)
                conf = float(np.amax(y_pred[n][f][bp[1]:bp[1] + bp[3], bp[0]:bp[0] + bp[2]])) if max(bp) > 0 else 0.
                typ = classify(np.amax(h_pred[n][f]) > 0, np.amax(y_true[n][f]) > 0, cx_pred, cy_pred, cx_true, cy_true,
                               tolerance)
            else:
                c_t, c_p = c_true[n][f], c_pred[n][f]
                cx_true, cy_true = int(c_t[0]), int(c_t[1])
                cx_pred, cy_pred = int(c_p[0]), int(c_p[1])
                typ = classify(np.amax(c_p) > 0, np.amax(c_t) > 0, cx_pred, cy_pred, cx_true, cy_true, tolerance)
            d['Type'].append(typ)
            d['Frame'].append(int(d_i[1]))
            d['X'].append(int(cx_pred * img_scaler[0]))
            d['Y'].append(int(cy_pred * img_scaler[1]))
            d['Visibility'].append(0 if cx_pred == 0 and cy_pred == 0 else 1)
            if output_bbox:
                d['BBox'].append([int(bp[0] * img_scaler[0]), int(bp[1] * img_scaler[1]), int(bp[2] * img_scaler[0]),
                                  int(bp[3] * img_scaler[1])])
                d['Confidence'].append(conf)
            if output_gt:
                d['X_GT'].append(int(cx_true * img_scaler[0]))
                d['Y_GT'].append(int(cy_true * img_scaler[1]))
                d['Visibility_GT'].append(0 if cx_true == 0 and cy_true == 0 else 1)
            prev = d_i
    if not output_bbox:
        del d['BBox'], d['Confidence']
    if not output_gt:
        del d['X_GT'], d['Y_GT'], d['Visibility_GT']
    return d
