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
                # fp32 products like the reference's 0-dim tensor arithmetic (python scalars do not promote a tensor)
                c_p = c_pred[n][f].astype(np.float32)
                cx_pred = int(c_p[0] * np.float32(WIDTH) * np.float32(img_scaler[0]))
                cy_pred = int(c_p[1] * np.float32(HEIGHT) * np.float32(img_scaler[1]))
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


def _windows(n_frames, seq_len, sliding_step, padding):
    """ Frame indices of every input sequence the reference's dataset builds over ``n_frames`` frames
        (dataset.py:329-355 `_gen_input_from_frame_arr`, :357-395 `_gen_input_from_pred_dict`): windows start every
        ``sliding_step`` frames; a window running past the end is dropped, or - with ``padding``, which the dataset only
        honours for non-overlapping sampling (:92) - completed by repeating the last frame seen so far. (N, L) int64. """
    padding = padding and sliding_step == seq_len
    out, last = [], -1
    for i in range(0, n_frames, sliding_step):
        idx = []
        for f in range(seq_len):
            if i + f < n_frames:
                idx.append(i + f)
                last = i + f
            elif padding:
                idx.append(last)
            else:
                break
        if len(idx) == seq_len:
            out.append(idx)
    return torch.tensor(out, dtype=torch.int64).reshape(-1, seq_len)


def _as_indices(win):
    """ (N, L) frame numbers -> the (N, L, 2) `(rally, frame)` index tensor the reference's datasets return. """
    return torch.stack([torch.zeros_like(win), win], dim=-1)


def _extend(dst, src):
    for k in src:
        dst[k].extend(src[k])


def run_video(frames_u8, tracknet, inpaintnet=None, seq_len=8, bg_mode='concat', batch_size=16, eval_mode='weight',
              img_scaler=None, inpaintnet_seq_len=16):
    """ The data path of the reference's `predict.py` __main__ (:110-301, small-video branch) for one video held in
        memory, entirely on the GPU: median background over all frames (dataset.py:102-107) -> Pillow-exact resize /
        stack (FramePreprocessor) -> TrackNet -> [temporal ensemble ->] heatmap decode -> generate_inpaint_mask ->
        InpaintNet + blend + threshold on the decoded trajectory [-> temporal ensemble]. All three ``eval_mode``s:
        'nonoverlap' (windows every seq_len frames, the last one padded, :123-145, :224-238), 'average' / 'weight'
        (windows every frame + temporal ensemble, :146-209, :239-301). Video file decoding and the csv / video writers
        stay host code around this function. Pinned bit for bit to the reference's own __main__ by
        tests/golden/predict_flow.npz (oracle/gen_predict_flow.py).

        Args:
            frames_u8 (torch.Tensor): (T, Hs, Ws, 3) uint8 RGB frames of the video (CUDA or host)
            tracknet, inpaintnet: models from utils.general.get_model, already .cuda().eval()
            seq_len, bg_mode: the TrackNet checkpoint's param_dict values (predict.py:100-101)
            inpaintnet_seq_len: the InpaintNet checkpoint's param_dict['seq_len'] (predict.py:106)
            img_scaler: (w_scaler, h_scaler); default (Ws / WIDTH, Hs / HEIGHT) as predict.py:114
        Returns:
            (tracknet_pred_dict incl. 'Inpaint_Mask' when inpaintnet is given, inpaint_pred_dict or None)
    """
    import tracknetv3_b200 as T
    from utils.general import COOR_TH
    if eval_mode not in ('nonoverlap', 'average', 'weight'):
        raise ValueError(f'invalid eval_mode {eval_mode!r}')
    frames_u8 = torch.as_tensor(frames_u8).cuda()
    t, hs, ws = frames_u8.shape[0], frames_u8.shape[1], frames_u8.shape[2]
    if img_scaler is None:
        img_scaler = (ws / WIDTH, hs / HEIGHT)
    fp = T.FramePreprocessor(hs, ws, HEIGHT, WIDTH)
    median = None
    if bg_mode == 'concat':
        median = fp.prepare_median(fp.median(frames_u8))
    elif bg_mode:
        median = fp.median(frames_u8, as_float=True)
    ensemble = eval_mode != 'nonoverlap'
    win = _windows(t, seq_len, 1 if ensemble else seq_len, padding=True)
    pred = {'Frame': [], 'X': [], 'Y': [], 'Visibility': []}
    ens = TemporalEnsemble(seq_len, eval_mode, len(win)) if ensemble and len(win) else None
    dev_win = win.cuda()
    with torch.no_grad():
        for s0 in range(0, len(win), batch_size):
            w_b = win[s0:s0 + batch_size]
            imgs = frames_u8[dev_win[s0:s0 + batch_size]]                      # (B, L, Hs, Ws, 3) gathered on the device
            y = tracknet(fp.process(imgs, median, bg_mode=bg_mode))
            if ensemble:
                _extend(pred, predict_ensemble(ens, _as_indices(w_b), y_pred=y, img_scaler=img_scaler))
            else:
                _extend(pred, predict(_as_indices(w_b), y_pred=y, img_scaler=img_scaler))
    if inpaintnet is None:
        return pred, None

    # InpaintNet pass over the decoded trajectory (predict.py:211-301): gaps selected by generate_inpaint_mask with the
    # reference's threshold of 5 % of the image height (:216); coordinates normalised by the image shape
    # (dataset.py:469-470)
    from test import generate_inpaint_mask
    L = inpaintnet_seq_len
    pred['Inpaint_Mask'] = generate_inpaint_mask(pred, th_h=hs * 0.05)
    n_pred = len(pred['Frame'])
    # float64 quotient, then .float(): the dataset's coordinate array is float64 (np.concatenate of a float32 array with
    # python ints promotes, dataset.py:391) and predict.py:253 casts the batch to fp32
    coor = torch.stack([torch.tensor(pred['X'], dtype=torch.float64) / ws,
                        torch.tensor(pred['Y'], dtype=torch.float64) / hs], dim=1).reshape(-1, 2).float().cuda()
    mask = torch.tensor(pred['Inpaint_Mask'], dtype=torch.float32).reshape(-1, 1).cuda()
    win = _windows(n_pred, L, 1 if ensemble else L, padding=True)
    out = {'Frame': [], 'X': [], 'Y': [], 'Visibility': []}
    ens_c = TemporalEnsemble(L, eval_mode, len(win)) if ensemble and len(win) else None
    dev_win = win.cuda()
    for s0 in range(0, len(win), batch_size):
        w_b = win[s0:s0 + batch_size]
        gather = dev_win[s0:s0 + batch_size]
        ci = inpaintnet.rectify(coor[gather], mask[gather], COOR_TH)           # forward + blend + threshold, one launch
        if ensemble:
            _extend(out, predict_ensemble(ens_c, _as_indices(w_b), c_pred=ci, img_scaler=img_scaler))
        else:
            _extend(out, predict(_as_indices(w_b), c_pred=ci, img_scaler=img_scaler))
    return pred, out


def generate_frames(video_file):
    """ All frames of an .mp4 as a list of BGR arrays (reference utils/general.py:202-226; cv2.VideoCapture on the host). """
    import cv2
    assert video_file[-4:] == '.mp4', 'Invalid video file format.'
    cap = cv2.VideoCapture(video_file)
    frame_list = []
    while True:
        success, frame = cap.read()
        if not success:
            break
        frame_list.append(frame)
    cap.release()
    return frame_list


def write_pred_csv(pred_dict, save_file):
    """ Frame, Visibility, X, Y - the csv the reference's `write_pred_csv` (utils/general.py:322-354) writes for a
        prediction (same header, column order and integer formatting as its pandas `to_csv(index=False)`). """
    import csv
    with open(save_file, 'w', newline='') as f:
        wr = csv.writer(f, lineterminator='\n')
        wr.writerow(['Frame', 'Visibility', 'X', 'Y'])
        for row in zip(pred_dict['Frame'], pred_dict['Visibility'], pred_dict['X'], pred_dict['Y']):
            wr.writerow([int(v) for v in row])


def load_models(tracknet_file, inpaintnet_file=''):
    """ Models from checkpoints in the reference's layout (predict.py:99-108): ckpt['model'] and
        ckpt['param_dict']['seq_len' / 'bg_mode']. Returns (tracknet, inpaintnet or None, seq_len, bg_mode,
        inpaintnet_seq_len or None). """
    from utils.general import get_model
    ckpt = torch.load(tracknet_file, map_location='cpu', weights_only=False)
    seq_len, bg_mode = ckpt['param_dict']['seq_len'], ckpt['param_dict']['bg_mode']
    tracknet = get_model('TrackNet', seq_len, bg_mode)
    tracknet.load_state_dict(ckpt['model'])
    inpaintnet, inpaintnet_seq_len = None, None
    if inpaintnet_file:
        ickpt = torch.load(inpaintnet_file, map_location='cpu', weights_only=False)
        inpaintnet_seq_len = ickpt['param_dict']['seq_len']
        inpaintnet = get_model('InpaintNet')
        inpaintnet.load_state_dict(ickpt['model'])
    return tracknet, inpaintnet, seq_len, bg_mode, inpaintnet_seq_len


if __name__ == '__main__':
    import argparse
    import os
    from utils.general import get_model
    # the reference's command line (predict.py:72-83) for the in-memory path; without --video_file a synthetic clip
    # and random weights exercise the same GPU data path
    parser = argparse.ArgumentParser()
    parser.add_argument('--video_file', type=str, default='', help='file path of the video (.mp4)')
    parser.add_argument('--tracknet_file', type=str, default='', help='file path of the TrackNet model checkpoint')
    parser.add_argument('--inpaintnet_file', type=str, default='', help='file path of the InpaintNet model checkpoint')
    parser.add_argument('--batch_size', type=int, default=16, help='batch size for inference')
    parser.add_argument('--eval_mode', type=str, default='weight', choices=['nonoverlap', 'average', 'weight'], help='evaluation mode')
    parser.add_argument('--save_dir', type=str, default='pred_result', help='directory to save the prediction result')
    parser.add_argument('--frames', type=int, default=40, help='length of the synthetic clip (no --video_file)')
    args = parser.parse_args()
    # models: from checkpoints in the reference's layout (predict.py:98-108), else randomly initialised
    if args.tracknet_file:
        tracknet, inpaintnet, seq_len, bg_mode, inpaint_seq_len = load_models(args.tracknet_file, args.inpaintnet_file)
    else:
        assert not args.video_file, 'a TrackNet checkpoint is required with --video_file'
        torch.manual_seed(0)
        seq_len, bg_mode, inpaint_seq_len = 8, 'concat', 16
        tracknet, inpaintnet = get_model('TrackNet', seq_len, bg_mode), get_model('InpaintNet')
    # frames: the video file (all frames in memory, BGR -> RGB as predict.py:128), else a synthetic clip
    if args.video_file:
        frames = np.array(generate_frames(args.video_file))[:, :, :, ::-1]
        video = torch.from_numpy(np.ascontiguousarray(frames))
        name = os.path.basename(args.video_file)[:-4]
    else:
        name = 'synthetic'
        video = torch.randint(0, 40, (args.frames, 360, 640, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
        for i in range(args.frames):                                   # a bright ball crossing a dark noisy court
            video[i, 100 + 3 * i:108 + 3 * i, 50 + 10 * i:58 + 10 * i] = 250
    tracknet = tracknet.cuda().eval()
    inpaintnet = inpaintnet.cuda().eval() if inpaintnet is not None else None
    h, w = video.shape[1], video.shape[2]
    p1, p2 = run_video(video, tracknet, inpaintnet, seq_len=seq_len, bg_mode=bg_mode, batch_size=args.batch_size,
                       eval_mode=args.eval_mode, img_scaler=(w / WIDTH, h / HEIGHT),
                       inpaintnet_seq_len=inpaint_seq_len or 16)
    os.makedirs(args.save_dir, exist_ok=True)
    out_csv = os.path.join(args.save_dir, f'{name}_ball.csv')
    write_pred_csv(p2 if p2 is not None else p1, out_csv)
    print(f"TrackNet: {len(p1['Frame'])} frames, {sum(p1['Visibility'])} visible"
          + (f"; InpaintNet: {len(p2['Frame'])} frames" if p2 is not None else '') + f"; wrote {out_csv}")
