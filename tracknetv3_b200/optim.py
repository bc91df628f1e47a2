"""Fused multi-tensor Adam (reference train.py:242 ``torch.optim.Adam(model.parameters(), lr)``)."""
import ctypes as C

import torch

from . import _lib


class _Entry(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_longlong)]


class FusedAdam(torch.optim.Optimizer):
    """Same update rule and defaults as ``torch.optim.Adam`` (amsgrad=False), one launch per step.

    The per-parameter state uses torch's keys (``step``, ``exp_avg``, ``exp_avg_sq``), so ``state_dict()`` /
    ``load_state_dict()`` interchange with ``torch.optim.Adam`` checkpoints (reference train.py:258-259, 286-301):
    ``step`` may arrive as a python int (torch 1.10, the reference's pin) or as a tensor (current torch).
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._table_key = None
        self._table = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                _lib.require_cuda(p)
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] = st["step"] + 1
            key = (gi,) + tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                                 self.state[p]["exp_avg_sq"].data_ptr()) for p in ps)
            if key != self._table_key:
                arr = (_Entry * len(ps))()
                for i, p in enumerate(ps):
                    st = self.state[p]
                    arr[i] = _Entry(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                    st["exp_avg_sq"].data_ptr(), p.numel())
                host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
                self._table = host.to(ps[0].device)
                self._table_key = key
            b1, b2 = group["betas"]
            _lib.check(lib.tnb_adam_multi(self._table.data_ptr(), len(ps), max(p.numel() for p in ps),
                                          group["lr"], b1, b2, group["eps"], group["weight_decay"],
                                          int(self.state[ps[0]]["step"]), _lib.stream_ptr()))
        return loss
