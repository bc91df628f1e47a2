"""Hot-path subset of the reference's predict.py: `predict` (:14-69), GPU decode instead of per-frame cv2.

`predict()` keeps the reference's signature and output dictionary. The heatmap branch thresholds and
decodes ALL (n, f) maps in one kernel launch on the device (the reference copies every heatmap to the
host and loops in Python), then applies the reference's de-duplication of padded frames on the host.
"""
import numpy as np
import torch

from tracknetv3_b200.decode import decode_heatmaps
from tracknetv3_b200.ensemble import TemporalEnsemble
from utils.general import HEIGHT, WIDTH


def predict(indices, y_pred=None, c_pred=None, img_scaler=(1, 1)):
    """ Predict coordinates from heatmap or inpainted coordinates.

        Args:
            indices (torch.Tensor): indices of input sequence with shape (N, L, 2)
            y_pred (torch.Tensor, optional): predicted heatmap sequence with shape (N, L, H, W)
            c_pred (torch.Tensor, optional): predicted inpainted coordinates sequence with shape (N, L, 2)
            img_scaler (Tuple): image scaler (w_scaler, h_scaler)

        Returns:
            pred_dict (Dict): {'Frame':[], 'X':[], 'Y':[], 'Visibility':[]}
    """
    pred_dict = {'Frame': [], 'X': [], 'Y': [], 'Visibility': []}
    batch_size, seq_len = indices.shape[0], indices.shape[1]
    indices = indices.detach().cpu().numpy() if torch.is_tensor(indices) else np.asarray(indices)

    boxes = None
    if c_pred is not None:
        c_pred = c_pred.detach().cpu().numpy() if torch.is_tensor(c_pred) else np.asarray(c_pred)
    elif y_pred is not None:
        if not torch.is_tensor(y_pred):
            y_pred = torch.as_tensor(np.asarray(y_pred))
        boxes = decode_heatmaps(y_pred.cuda(), threshold=0.5).cpu().numpy()  # (N, L, 4) ints, 16 B per frame
    else:
        raise ValueError('Invalid input')

    prev_f_i = -1
    for n in range(batch_size):
        for f in range(seq_len):
            f_i = indices[n][f][1]
            if f_i == prev_f_i:
                break
            if c_pred is not None:
                c_p = c_pred[n][f]
                cx_pred, cy_pred = int(c_p[0] * WIDTH * img_scaler[0]), int(c_p[1] * HEIGHT * img_scaler[1])
            else:
                x, y, w, h = (int(v) for v in boxes[n][f])
                cx_pred, cy_pred = int(x + w / 2), int(y + h / 2)
                cx_pred, cy_pred = int(cx_pred * img_scaler[0]), int(cy_pred * img_scaler[1])
            vis_pred = 0 if cx_pred == 0 and cy_pred == 0 else 1
            pred_dict['Frame'].append(int(f_i))
            pred_dict['X'].append(cx_pred)
            pred_dict['Y'].append(cy_pred)
            pred_dict['Visibility'].append(vis_pred)
            prev_f_i = f_i
    return pred_dict


def predict_ensemble(ensemble, indices, y_pred=None, c_pred=None, img_scaler=(1, 1)):
    """ One iteration of the reference's temporal-ensemble loops (predict.py:168-209 with ``y_pred``, :252-301 with
        ``c_pred``) with the ensemble and the heatmap decode on the GPU: no heatmap leaves the device.

        Args:
            ensemble (TemporalEnsemble): created once per video with (seq_len, eval_mode, num_sample)
            indices (torch.Tensor): indices of this batch's input sequences, (N, L, 2)
            y_pred (torch.Tensor, optional): CUDA heatmaps of this batch, (N, L, H, W)
            c_pred (torch.Tensor, optional): CUDA inpainted + blended + thresholded coordinates, (N, L, 2)

        Returns:
            pred_dict (Dict): what ``predict()`` returns for the frames this batch completes
    """
    from utils.general import COOR_TH
    seq_len = indices.shape[1]
    b_size = indices.shape[0]
    last = ensemble.sample_count + b_size == ensemble.num_sample
    ens_i = [indices[b][0].reshape(1, 1, 2) for b in range(b_size)]
    if last:
        ens_i += [indices[-1][f].reshape(1, 1, 2) for f in range(1, seq_len)]
    ens_i = torch.cat([torch.as_tensor(t) for t in ens_i], dim=0)
    if y_pred is not None:
        ens = ensemble.push(y_pred)                       # (n_out, H, W)
        return predict(ens_i, y_pred=ens.unsqueeze(1), img_scaler=img_scaler)
    if c_pred is not None:
        ens = ensemble.push(c_pred)                       # (n_out, 2)
        th_mask = (ens[:, 0] < COOR_TH) & (ens[:, 1] < COOR_TH)
        ens = ens.masked_fill(th_mask[:, None], 0.)
        return predict(ens_i, c_pred=ens.unsqueeze(1), img_scaler=img_scaler)
    raise ValueError('Invalid input')
