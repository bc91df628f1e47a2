"""tests/golden/predict_flow.npz: the REAL reference predict.py `__main__` (predict.py:70-311) executed end to end.

Run in the build container only:  python oracle/gen_predict_flow.py
`runpy` executes /root/reference/predict.py as `__main__` with its own dataset.py / test.py / model.py / utils imported
from /root/reference - nothing is lifted or restated. What is replaced around it is only what does not exist here:
  * `parse`, `pycocotools` (not installed; used for file-name parsing / COCO evaluation, never on this path): empty stubs;
  * the checkpoint files: `torch.load` returns {'model': state_dict, 'param_dict': {seq_len, bg_mode}} built by
    oracle/synth_clip.py (the reference publishes its checkpoints behind a Google-Drive link only);
  * the video file: `generate_frames` / `cv2.VideoCapture` serve the seeded synthetic clip of oracle/synth_clip.py;
  * the GPU: `.cuda()` is the identity (CPU fp32 arithmetic of the reference's own modules); DataLoader workers = 0;
  * `write_pred_csv`: records the dictionary it is handed.
Stored per case and eval_mode: the TrackNet prediction dictionary incl. Inpaint_Mask and the InpaintNet one.
"""
import os
import runpy
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "tests", "golden")


def _stub_missing_modules():
    for name in ("parse", "pycocotools", "pycocotools.coco", "pycocotools.cocoeval"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools.cocoeval"].COCOeval = object


def run_reference_main(clip_rgb, tracknet_ckpt, inpaintnet_ckpt, eval_mode, batch_size):
    """predict.py `__main__` on one clip. Returns (tracknet_pred_dict, inpaint_pred_dict)."""
    import cv2
    import torch.utils.data as tud
    import utils.general as G                      # the reference's (REF is first on sys.path)
    assert G.__file__.startswith(REF)

    class FakeCapture:
        def __init__(self, path):
            pass

        def get(self, prop):
            return {cv2.CAP_PROP_FRAME_WIDTH: clip_rgb.shape[2], cv2.CAP_PROP_FRAME_HEIGHT: clip_rgb.shape[1]}[prop]

    class LoaderNoWorkers(tud.DataLoader):
        def __init__(self, *a, **k):
            k["num_workers"] = 0
            super().__init__(*a, **k)

    written = {}
    saved = (cv2.VideoCapture, tud.DataLoader, torch.load, torch.Tensor.cuda, torch.nn.Module.cuda, G.generate_frames,
             G.write_pred_csv, sys.argv)
    try:
        cv2.VideoCapture = FakeCapture
        tud.DataLoader = LoaderNoWorkers
        torch.load = lambda f, *a, **k: {"tracknet.pt": tracknet_ckpt, "inpaintnet.pt": inpaintnet_ckpt}[f]
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        G.generate_frames = lambda video_file: [f[:, :, ::-1] for f in clip_rgb]         # BGR, as cv2 would return
        G.write_pred_csv = lambda pred_dict, save_file, save_inpaint_mask=False: written.update(pred=pred_dict)
        sys.argv = ["predict.py", "--video_file", "clip.mp4", "--tracknet_file", "tracknet.pt", "--inpaintnet_file",
                    "inpaintnet.pt", "--batch_size", str(batch_size), "--eval_mode", eval_mode, "--save_dir",
                    "/tmp/tnb_predict_flow"]
        g = runpy.run_path(f"{REF}/predict.py", run_name="__main__")
    finally:
        (cv2.VideoCapture, tud.DataLoader, torch.load, torch.Tensor.cuda, torch.nn.Module.cuda, G.generate_frames,
         G.write_pred_csv, sys.argv) = saved
    assert written["pred"] is g["inpaint_pred_dict"]
    return g["tracknet_pred_dict"], g["inpaint_pred_dict"]


WINDOW_CASES = [(n, L, step, pad) for n in (1, 3, 8, 9, 16, 21) for L in (4, 8) for step in (1, L) for pad in (False, True)]


def gen_windows():
    """The input-sequence index table of the reference's dataset (dataset.py:329-355) for a range of clip lengths,
    window lengths, sliding steps and padding flags -> tests/golden/predict_windows.npz (pins predict._windows)."""
    from dataset import Shuttlecock_Trajectory_Dataset
    out = {}
    for n, L, step, pad in WINDOW_CASES:
        ds = Shuttlecock_Trajectory_Dataset(seq_len=L, sliding_step=step, data_mode='heatmap', bg_mode='',
                                            frame_arr=np.zeros((n, 4, 4, 3), np.uint8), padding=pad)
        out[f"{n}_{L}_{step}_{int(pad)}"] = np.asarray(ds.data_dict['id'], dtype=np.int64).reshape(-1, L, 2)
    np.savez_compressed(f"{OUT}/predict_windows.npz", **out)


def main():
    sys.path.insert(0, REF)
    sys.path.append(ROOT)
    _stub_missing_modules()
    from oracle import synth_clip as S
    gen_windows()
    if "--windows-only" in sys.argv:
        return
    out = {}
    for name, t, (hs, ws), L, bg_mode, L_inp, bs, gap in S.CASES:
        clip = S.make_clip(t, hs, ws, seed=len(name) + t, gap=gap)
        tck = {"model": S.detector_tracknet_state(L, bg_mode, seed=1), "param_dict": {"seq_len": L, "bg_mode": bg_mode}}
        ick = {"model": S.inpaintnet_state(seed=2), "param_dict": {"seq_len": L_inp}}
        for mode in S.EVAL_MODES:
            p1, p2 = run_reference_main(clip, tck, ick, mode, bs)
            for k in ("Frame", "X", "Y", "Visibility", "Inpaint_Mask"):
                out[f"{name}/{mode}/tracknet/{k}"] = np.asarray(p1[k], dtype=np.int64)
            for k in ("Frame", "X", "Y", "Visibility"):
                out[f"{name}/{mode}/inpaint/{k}"] = np.asarray(p2[k], dtype=np.int64)
            print(name, mode, "tracknet", list(zip(p1["Frame"], p1["X"], p1["Y"], p1["Inpaint_Mask"])), flush=True)
            print(name, mode, "inpaint ", list(zip(p2["Frame"], p2["X"], p2["Y"], p2["Visibility"])), flush=True)
    np.savez_compressed(f"{OUT}/predict_flow.npz", **out)


if __name__ == "__main__":
    main()
