"""Hot-path subset of the reference's test.py: get_ensemble_weight (:25-50) and predict_location (:52-79).
The evaluation drivers of the reference (dataset loops, COCO export) are out of scope."""
from tracknetv3_b200.decode import predict_location, decode_heatmaps  # noqa: F401
from tracknetv3_b200.ensemble import get_ensemble_weight, TemporalEnsemble  # noqa: F401

import math  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tracknetv3_b200 import _lib  # noqa: E402
from utils.general import HEIGHT, WIDTH  # noqa: E402

pred_types = ['TP', 'TN', 'FP1', 'FP2', 'FN']
pred_types_map = {pred_type: i for i, pred_type in enumerate(pred_types)}


def _frame_stats(y_true, y_pred):
    """Everything `evaluate` needs per frame, computed on the GPU: predicted / true bounding boxes (the OpenCV-rule
    decode), confidence, ground-truth presence. 44 bytes per frame come back instead of two full heatmaps."""
    lib = _lib.load()
    y_pred = y_pred.cuda().contiguous().float()
    y_true = y_true.cuda().contiguous().float()
    boxes_pred = decode_heatmaps(y_pred, threshold=0.5)                       # predict_location(to_img(y_pred > 0.5))
    boxes_true = decode_heatmaps((y_true * 255).to(torch.uint8))              # predict_location(to_img(y_true))
    nmaps = y_pred.numel() // (y_pred.shape[-2] * y_pred.shape[-1])
    conf = torch.empty(nmaps, dtype=torch.float32, device=y_pred.device)
    true_any = torch.empty(nmaps, dtype=torch.int32, device=y_pred.device)
    _lib.check(lib.tnb_eval_stats(y_pred.data_ptr(), y_true.data_ptr(), boxes_pred.data_ptr(), nmaps, y_pred.shape[-2],
                                  y_pred.shape[-1], conf.data_ptr(), true_any.data_ptr(), _lib.stream_ptr()))
    lead = y_pred.shape[:-2]
    return (boxes_pred.cpu().numpy(), boxes_true.cpu().numpy(), conf.reshape(lead).cpu().numpy(),
            true_any.reshape(lead).cpu().numpy())


def evaluate(indices, y_true=None, y_pred=None, c_true=None, c_pred=None, tolerance=4., img_scaler=(1, 1),
             output_bbox=False, output_gt=False):
    """ Predict and output the result of each frame (drop-in for reference test.py:81-221; same arguments and the
        same dictionary). Heatmap inputs are decoded on the GPU; the outcome logic (TP / TN / FP1 / FP2 / FN) runs on
        the few integers per frame that come back. Unlike the reference, the caller's tensors are not modified.
    """
    pred_dict = {'Frame': [], 'X': [], 'Y': [], 'Visibility': [], 'Type': [], 'BBox': [], 'Confidence': [],
                 'X_GT': [], 'Y_GT': [], 'Visibility_GT': []}
    batch_size, seq_len = indices.shape[0], indices.shape[1]
    indices = indices.detach().cpu().numpy().tolist() if torch.is_tensor(indices) else np.asarray(indices).tolist()

    heat = y_true is not None and y_pred is not None
    if heat:
        assert c_true is None and c_pred is None, 'Invalid input'
        y_true = y_true if torch.is_tensor(y_true) else torch.as_tensor(np.asarray(y_true))
        y_pred = y_pred if torch.is_tensor(y_pred) else torch.as_tensor(np.asarray(y_pred))
        boxes_pred, boxes_true, conf_all, true_any = _frame_stats(y_true, y_pred)
    elif c_true is not None and c_pred is not None:
        assert y_true is None and y_pred is None, 'Invalid input'
        assert output_bbox == False, 'Coordinate prediction cannot output detection'  # noqa: E712
        scale = np.array([WIDTH, HEIGHT], dtype=np.float32)
        c_true = (c_true.detach().cpu().numpy() if torch.is_tensor(c_true) else np.asarray(c_true)).astype(np.float32) * scale
        c_pred = (c_pred.detach().cpu().numpy() if torch.is_tensor(c_pred) else np.asarray(c_pred)).astype(np.float32) * scale

    for n in range(batch_size):
        prev_d_i = [-1, -1]  # for ignoring the same frame in sequence
        for f in range(seq_len):
            d_i = indices[n][f]
            if d_i == prev_d_i:
                break
            if heat:
                bbox_true, bbox_pred = boxes_true[n][f], boxes_pred[n][f]
                cx_true, cy_true = int(bbox_true[0] + bbox_true[2] / 2), int(bbox_true[1] + bbox_true[3] / 2)
                cx_pred, cy_pred = int(bbox_pred[0] + bbox_pred[2] / 2), int(bbox_pred[1] + bbox_pred[3] / 2)
                conf = float(conf_all[n][f]) if np.amax(bbox_pred) > 0 else 0.
                p_any, t_any = bool(np.amax(bbox_pred) > 0), bool(true_any[n][f])
            elif c_true is not None and c_pred is not None:
                c_t, c_p = c_true[n][f], c_pred[n][f]
                cx_true, cy_true = int(c_t[0]), int(c_t[1])
                cx_pred, cy_pred = int(c_p[0]), int(c_p[1])
                p_any, t_any = bool(np.amax(c_p) > 0), bool(np.amax(c_t) > 0)
            else:
                raise ValueError('Invalid input')
            vis_pred = 0 if cx_pred == 0 and cy_pred == 0 else 1
            if not p_any and not t_any:
                pred_dict['Type'].append(pred_types_map['TN'])
            elif p_any and not t_any:
                pred_dict['Type'].append(pred_types_map['FP2'])
            elif not p_any and t_any:
                pred_dict['Type'].append(pred_types_map['FN'])
            else:
                dist = math.sqrt(pow(cx_pred - cx_true, 2) + pow(cy_pred - cy_true, 2))
                pred_dict['Type'].append(pred_types_map['FP1'] if dist > tolerance else pred_types_map['TP'])
            pred_dict['Frame'].append(int(d_i[1]))
            pred_dict['X'].append(int(cx_pred * img_scaler[0]))
            pred_dict['Y'].append(int(cy_pred * img_scaler[1]))
            pred_dict['Visibility'].append(vis_pred)
            if output_bbox:
                pred_dict['BBox'].append([int(bbox_pred[0] * img_scaler[0]), int(bbox_pred[1] * img_scaler[1]),
                                          int(bbox_pred[2] * img_scaler[0]), int(bbox_pred[3] * img_scaler[1])])
                pred_dict['Confidence'].append(float(conf))
            if output_gt:
                vis_gt = 0 if cx_true == 0 and cy_true == 0 else 1
                pred_dict['X_GT'].append(int(cx_true * img_scaler[0]))
                pred_dict['Y_GT'].append(int(cy_true * img_scaler[1]))
                pred_dict['Visibility_GT'].append(vis_gt)
            prev_d_i = d_i

    if not output_bbox:
        del pred_dict['BBox']
        del pred_dict['Confidence']
    if not output_gt:
        del pred_dict['X_GT']
        del pred_dict['Y_GT']
        del pred_dict['Visibility_GT']
    return pred_dict


def generate_inpaint_mask(pred_dict, th_h=30):
    """ Generate inpaint mask from a predicted trajectory (drop-in for reference test.py:223-258): a run of invisible
        frames is marked for inpainting when the ball was inside the court on both sides of it - y above `th_h` at the
        last visible frame before the run (the run must start after index 1) and at the first visible frame after it - or,
        for a run at the very start, at the first visible frame after it. A run that reaches the last frame is never
        marked, like in the reference, whose scan treats the last index as the end of every run.

        Args:
            pred_dict (Dict): {'Frame': [], 'X': [], 'Y': [], 'Visibility': []}
            th_h (float): height threshold (pixels) for the y coordinate
        Returns:
            inpaint_mask (List[int])
    """
    y = np.asarray(pred_dict['Y'])
    vis = np.asarray(pred_dict['Visibility'])
    n = len(vis)
    mask = np.zeros_like(y)
    start = 0
    while n > 0:
        gone = np.flatnonzero(vis[start:n - 1] != 1)          # first invisible frame; the last index stands in for "none"
        i = start + int(gone[0]) if len(gone) else n - 1
        back = np.flatnonzero(vis[i:n - 1] != 0)              # first visible frame after it; ditto
        j = i + int(back[0]) if len(back) else n - 1
        if j == i:
            break
        if i == 0:
            if y[j] > th_h:
                mask[:j] = 1
        elif i > 1 and y[i - 1] > th_h and y[j] > th_h:
            mask[i:j] = 1
        start = j
    return mask.tolist()


inpaintnet_eval_types = ['inpaint', 'reconstruct', 'baseline']


def get_eval_res(pred_dict):
    """ np.array([TP, TN, FP1, FP2, FN]) counted from pred_dict['Type'] (reference test.py:289-306). """
    type_res = np.asarray(pred_dict['Type'])
    return np.array([float((type_res == pred_types_map[t]).sum()) for t in pred_types])


def _res_dict(confusion):
    from utils.metric import get_metric
    TP, TN, FP1, FP2, FN = confusion
    accuracy, precision, recall, f1, miss_rate = get_metric(TP, TN, FP1, FP2, FN)
    return {'TP': TP, 'TN': TN, 'FP1': FP1, 'FP2': FP2, 'FN': FN, 'accuracy': accuracy, 'precision': precision,
            'recall': recall, 'f1': f1, 'miss_rate': miss_rate}


def eval_tracknet(model, data_loader, param_dict):
    """ Evaluate TrackNet (validation loop of reference test.py:308-368, called from train.py:276): mean WBCE loss and
        the TP/TN/FP1/FP2/FN counts with derived metrics. Forward, loss and the heatmap decode inside `evaluate` run on the
        GPU; per batch one loss scalar and 44 B per frame come back to the host. """
    from utils.metric import WBCELoss
    model.eval()
    losses = []
    confusion = np.zeros(5)
    for step, (i, x, y, _, _) in enumerate(data_loader):
        x, y = x.float().cuda(), y.float().cuda()
        with torch.no_grad():
            y_pred = model(x)
            losses.append(WBCELoss(y_pred, y).item())
        confusion += get_eval_res(evaluate(i, y_true=y, y_pred=y_pred, tolerance=param_dict['tolerance']))
    return float(np.mean(losses)), _res_dict(confusion)


def eval_inpaintnet(model, data_loader, param_dict):
    """ Evaluate InpaintNet (reference test.py:370-443): masked MSE loss and, for the three comparisons of the reference
        ('inpaint': refined vs truth, 'reconstruct': refined vs TrackNet's prediction, 'baseline': prediction vs truth),
        the TP/TN/FP1/FP2/FN counts with derived metrics. """
    from utils.general import COOR_TH
    model.eval()
    losses = []
    confusion = {t: np.zeros(5) for t in inpaintnet_eval_types}
    for step, (i, coor_pred, coor, _, _, inpaint_mask) in enumerate(data_loader):
        coor_pred, coor, inpaint_mask = coor_pred.float().cuda(), coor.float().cuda(), inpaint_mask.float().cuda()
        with torch.no_grad():
            coor_inpaint = model(coor_pred, inpaint_mask)
            coor_inpaint = coor_inpaint * inpaint_mask + coor_pred * (1 - inpaint_mask)
            losses.append(torch.nn.MSELoss()(coor_inpaint * inpaint_mask, coor * inpaint_mask).item())
            th_mask = (coor_inpaint[:, :, 0] < COOR_TH) & (coor_inpaint[:, :, 1] < COOR_TH)
            coor_inpaint = coor_inpaint.masked_fill(th_mask[:, :, None], 0.)
        pairs = {'inpaint': (coor, coor_inpaint), 'reconstruct': (coor_pred, coor_inpaint), 'baseline': (coor, coor_pred)}
        for t, (c_true, c_pred) in pairs.items():
            confusion[t] += get_eval_res(evaluate(i, c_true=c_true, c_pred=c_pred, tolerance=param_dict['tolerance']))
    return float(np.mean(losses)), {t: _res_dict(confusion[t]) for t in inpaintnet_eval_types}
