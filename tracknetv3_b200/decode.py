"""Heatmap -> shuttlecock location decode on the GPU (reference test.py:52-79, predict.py:35,54-57)."""
import numpy as np
import torch

from . import _lib


def decode_heatmaps(y_pred, threshold=0.5):
    """Batched predict_location: (..., H, W) float CUDA tensor -> (..., 4) int32 CUDA tensor (x, y, w, h).

    Foreground = ``y_pred > threshold`` exactly as ``predict.py:35``; an empty map yields (0, 0, 0, 0).
    """
    lib = _lib.load()
    _lib.require_cuda(y_pred)
    h, w = y_pred.shape[-2], y_pred.shape[-1]
    lead = y_pred.shape[:-2]
    maps = y_pred.contiguous()
    is_u8 = maps.dtype == torch.uint8
    if not is_u8:
        maps = maps.float()
    nmaps = maps.numel() // (h * w) if h * w else 0
    out = torch.zeros((nmaps, 4), dtype=torch.int32, device=maps.device)
    if nmaps:
        ws = torch.empty(lib.tnb_heatmap_decode_workspace_bytes(nmaps, h, w), dtype=torch.uint8, device=maps.device)
        _lib.check(lib.tnb_heatmap_decode(maps.data_ptr(), int(is_u8), float(threshold), nmaps, h, w,
                                          ws.data_ptr(), out.data_ptr(), _lib.stream_ptr()))
    return out.reshape(*lead, 4)


def predict_location(heatmap):
    """ Get coordinates from the heatmap (drop-in for reference test.py:52).

        Args:
            heatmap (numpy.ndarray | torch.Tensor): a single uint8 map (H, W), non-zero = response
        Returns:
            x, y, w, h (Tuple[int, int, int, int])
    """
    if isinstance(heatmap, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(heatmap)).cuda()
    else:
        t = heatmap.cuda()
    if t.dtype != torch.uint8:
        t = (t != 0).to(torch.uint8)
    x, y, w, h = decode_heatmaps(t)[...].reshape(4).tolist()
    return x, y, w, h


def bbox_to_center(bbox):
    """cx = int(x + w/2), cy = int(y + h/2) (reference predict.py:56, test.py:162); integer tensor in/out."""
    x, y, w, h = bbox.unbind(-1)
    return torch.stack([(2 * x + w) // 2, (2 * y + h) // 2], dim=-1)
