"""WBCELoss (reference utils/metric.py:3-20) as one fused forward and one fused backward kernel."""
import torch

from . import _lib


class _WBCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y, reduce):
        lib = _lib.load()
        _lib.require_cuda(y_pred, y)
        if y_pred.shape != y.shape:
            raise RuntimeError(f"WBCELoss: shape mismatch {tuple(y_pred.shape)} vs {tuple(y.shape)}")
        p = y_pred.contiguous().float()
        t = y.contiguous().float()
        n = p.shape[0]
        per_sample = p.numel() // n
        part = torch.empty(lib.tnb_wbce_workspace_bytes(n), dtype=torch.uint8, device=p.device)
        out = torch.empty(1 if reduce else n, dtype=torch.float32, device=p.device)
        _lib.check(lib.tnb_wbce_fwd(p.data_ptr(), t.data_ptr(), n, per_sample, int(reduce), part.data_ptr(),
                                    out.data_ptr(), _lib.stream_ptr()))
        ctx.save_for_backward(p, t)
        ctx.reduce = bool(reduce)
        return out.reshape(()) if reduce else out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        p, t = ctx.saved_tensors
        n = p.shape[0]
        per_sample = p.numel() // n
        g = gout.contiguous().float().reshape(-1)
        dp = torch.empty_like(p)
        _lib.check(lib.tnb_wbce_bwd(p.data_ptr(), t.data_ptr(), g.data_ptr(), n, per_sample, int(ctx.reduce),
                                    dp.data_ptr(), _lib.stream_ptr()))
        return dp, None, None


def WBCELoss(y_pred, y, reduce=True):
    """ Weighted Binary Cross Entropy loss function defined in TrackNetV2 paper.

        Same signature and semantics as the reference (utils/metric.py:3): mean over all elements when
        ``reduce`` else per-sample mean of shape (N,). Gradient flows to ``y_pred`` only.
    """
    return _WBCEFunction.apply(y_pred, y, reduce)


def get_metric(TP, TN, FP1, FP2, FN):
    """ accuracy, precision, recall, f1, miss_rate (host arithmetic; reference utils/metric.py:22-46). """
    total = TP + TN + FP1 + FP2 + FN
    accuracy = (TP + TN) / total if total > 0 else 0
    precision = TP / (TP + FP1 + FP2) if (TP + FP1 + FP2) > 0 else 0
    recall = TP / (TP + FN) if (TP + FN) > 0 else 0
    f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0
    miss_rate = FN / (TP + FN) if (TP + FN) > 0 else 0
    return accuracy, precision, recall, f1, miss_rate
