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
        self._tables = {}

    def _table(self, ps):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in ps)
        tab = self._tables.get(key)
        if tab is None:
            if len(self._tables) > 16:      # addresses keep changing (no stable gradient buffer): do not hoard tables
                self._tables.clear()
            arr = (_Entry * len(ps))()
            for i, p in enumerate(ps):
                st = self.state[p]
                arr[i] = _Entry(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                st["exp_avg_sq"].data_ptr(), p.numel())
            tab = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(ps[0].device)
            self._tables[key] = tab
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            # options a loaded torch.optim.Adam state_dict may carry (load_state_dict copies its param_groups)
            for flag in ("amsgrad", "maximize"):
                if group.get(flag, False):
                    raise RuntimeError(f"FusedAdam does not implement {flag}=True (found in a loaded optimizer state)")
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            by_step = {}
            for p in ps:
                _lib.require_cuda(p, p.grad)
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() \
                        or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous fp32 parameters and gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] = st["step"] + 1
                by_step.setdefault(int(st["step"]), []).append(p)
            b1, b2 = group["betas"]
            # the bias correction depends on the step count: parameters that skipped steps (no gradient) form their own
            # launch - one launch in the train loop of the reference, where every parameter gets a gradient every step
            for step, bucket in by_step.items():
                _lib.check(lib.tnb_adam_multi(self._table(bucket).data_ptr(), len(bucket),
                                              sum(p.numel() for p in bucket), group["lr"], b1, b2, group["eps"],
                                              group["weight_decay"], step, _lib.stream_ptr()))
        return loss
