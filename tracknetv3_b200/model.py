"""TrackNet / InpaintNet with the reference's class names, constructor arguments, forward signatures
and state_dict layout (reference model.py:4-129), executing on hand-written sm_100a kernels.

The sub-modules (``conv``, ``bn``, ``relu`` ...) exist only as parameter/buffer containers so that
``state_dict()`` / ``load_state_dict()`` / ``parameters()`` order are byte-compatible with reference
checkpoints (104 entries for TrackNet, 20 for InpaintNet). ``forward`` never calls
them: the whole network runs through ``tnb_tracknet_forward`` / ``tnb_tracknet_backward`` (C ABI).
"""
import ctypes as C
import weakref

import torch
import torch.nn as nn

from . import _lib


class Conv2DBlock(nn.Module):
    """ Conv2D + BN + ReLU (parameter container; reference model.py:4-16) """

    def __init__(self, in_dim, out_dim, **kwargs):
        super(Conv2DBlock, self).__init__(**kwargs)
        self.conv = nn.Conv2d(in_dim, out_dim, kernel_size=3, padding='same', bias=False)
        self.bn = nn.BatchNorm2d(out_dim)
        self.relu = nn.ReLU()


class Double2DConv(nn.Module):
    """ Conv2DBlock x 2 (reference model.py:18-28) """

    def __init__(self, in_dim, out_dim):
        super(Double2DConv, self).__init__()
        self.conv_1 = Conv2DBlock(in_dim, out_dim)
        self.conv_2 = Conv2DBlock(out_dim, out_dim)


class Triple2DConv(nn.Module):
    """ Conv2DBlock x 3 (reference model.py:30-42) """

    def __init__(self, in_dim, out_dim):
        super(Triple2DConv, self).__init__()
        self.conv_1 = Conv2DBlock(in_dim, out_dim)
        self.conv_2 = Conv2DBlock(out_dim, out_dim)
        self.conv_3 = Conv2DBlock(out_dim, out_dim)


# (forward terms, backward terms) of the 3x3 convolutions' operand split
_PRECISIONS = {"fp32x3": (3, 3), "tf32like": (1, 1), "fp32x3_bwd1": (3, 1)}


def _cfg(n, h, w, in_dim, out_dim, training, precision, variant=0):
    terms = _PRECISIONS[precision]
    return _lib.TrackNetCfg(n=n, h=h, w=w, in_dim=in_dim, out_dim=out_dim, training=int(training),
                            fwd_terms=terms[0], bwd_terms=terms[1], variant=variant, bn_eps=1e-5, bn_momentum=0.1)


class _TrackNetFunction(torch.autograd.Function):
    """autograd bridge: forward = tnb_tracknet_forward, backward = tnb_tracknet_backward."""

    @staticmethod
    def forward(ctx, x, module, saves_state, *params):
        lib = _lib.load()
        n, _, h, w = x.shape
        # training: 1 = batch statistics; 2 = eval() with a backward to follow (running statistics, backward state kept)
        mode = 1 if module.training else (2 if saves_state else 0)
        cfg = _cfg(n, h, w, module.in_dim, module.out_dim, mode, module.precision, module._variant)
        nbytes = lib.tnb_tracknet_workspace_bytes(C.byref(cfg))
        if nbytes == 0:
            # mirror the reference's failure for sizes its pooling / concat cannot handle (model.py:59-69)
            raise RuntimeError(lib.tnb_last_error().decode())
        ws, ctx.token = module._workspace(nbytes, x.device, saves_state)
        y = torch.empty((n, module.out_dim, h, w), dtype=torch.float32, device=x.device)
        tensors = module._state_tensors()
        _lib.check(lib.tnb_tracknet_forward(C.byref(cfg), x.data_ptr(), _lib.ptr_array(tensors), y.data_ptr(),
                                            ws.data_ptr(), nbytes, _lib.stream_ptr()))
        ctx.module, ctx.cfg, ctx.ws, ctx.nbytes = module, cfg, ws, nbytes
        ctx.save_for_backward(y)
        ctx.set_materialize_grads(True)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (y,) = ctx.saved_tensors
        module = ctx.module
        if ctx.needs_input_grad[0]:
            raise RuntimeError("tracknet_b200: the gradient w.r.t. the input frames is not computed (the reference's "
                               "train step never asks for it, train.py:86-95); pass x without requires_grad")
        if ctx.token is None or ctx.token.done:
            # the workspace holds this forward's activations and BatchNorm statistics; the backward kernels consume them
            raise RuntimeError("tracknet_b200: backward called twice on the same forward (retain_graph is not supported)")
        ctx.token.done = True
        params = list(module.parameters())
        # one block for all 53 gradients: a single allocation whose address repeats from step to step, which is what
        # lets the library replay its captured launch sequence (the pointers are part of the CUDA-graph key)
        flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dy.device)
        grads, off = [], 0
        for p in params:
            grads.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        dy = dy.contiguous()
        state, gptr = _lib.ptr_array(module._state_tensors()), _lib.ptr_array(grads)
        split = module._grad_split
        if split is None:
            _lib.check(lib.tnb_tracknet_backward(C.byref(ctx.cfg), dy.data_ptr(), y.data_ptr(), state, gptr,
                                                 ctx.ws.data_ptr(), ctx.nbytes, _lib.stream_ptr()))
        else:
            # data parallel (parallel.GradBucket.overlap): two ranges with an event in between - the gradients of the
            # bottleneck / decoder / predictor are final at the event and get all-reduced on the bucket's side stream
            # while the encoder's backward still runs
            s = lib.tnb_tracknet_grad_split_layer()
            for hi, lo in ((16, s), (s - 1, 0)):
                _lib.check(lib.tnb_tracknet_backward_range(C.byref(ctx.cfg), dy.data_ptr(), y.data_ptr(), state, gptr,
                                                           ctx.ws.data_ptr(), ctx.nbytes, hi, lo, _lib.stream_ptr()))
                if lo == s:
                    split.first_param = 3 * s  # parameters are (conv.weight, bn.weight, bn.bias) per block, predictor last
                    split.event.record()
        return (None, None, None) + tuple(grads)


class _WorkspaceToken:
    """Alive and not ``done`` while a forward's saved-for-backward state sits in the workspace it was handed."""
    __slots__ = ("done", "__weakref__")

    def __init__(self):
        self.done = False


class TrackNet(nn.Module):
    """Drop-in for reference ``model.TrackNet`` (model.py:44-73).

    ``precision``: "fp32x3" (default) computes every 3x3 convolution as a 3-term (hi, lo) split product with fp32
    accumulation - fp16 pairs in the forward pass (weights pre-scaled by 2^10 so that their lo halves stay normal
    numbers), bf16 pairs in the backward pass - within ~4e-5 of the reference's fp32 heatmaps;
    "tf32like" is a single 16-bit pass (what the reference's own cuDNN TF32 path amounts to; ~7e-3);
    "fp32x3_bwd1" keeps the fp32-faithful forward (the heatmap bound) and runs dgrad / wgrad as a single fp16 pass
    (11-bit operands, gradients stored multiplied by a per-layer power of two; TF32, what the reference's own GPU
    backward runs in, has 10) - measured next to the default by ``bench.py --precision fp32x3_bwd1``, never the headline.

    ``train()``: batch statistics, running statistics and counters advanced. ``eval()``: running statistics; under
    ``torch.no_grad()`` nothing is kept, with gradients enabled the forward keeps what ``backward()`` needs and the
    BatchNorm layers act as frozen affine maps in it (``tnb_tracknet_cfg_t.training`` = 1 / 0 / 2).
    """

    def __init__(self, in_dim, out_dim, precision="fp32x3"):
        super(TrackNet, self).__init__()
        self.down_block_1 = Double2DConv(in_dim, 64)
        self.down_block_2 = Double2DConv(64, 128)
        self.down_block_3 = Triple2DConv(128, 256)
        self.bottleneck = Triple2DConv(256, 512)
        self.up_block_1 = Triple2DConv(768, 256)
        self.up_block_2 = Double2DConv(384, 128)
        self.up_block_3 = Double2DConv(192, 64)
        self.predictor = nn.Conv2d(64, out_dim, (1, 1))
        self.sigmoid = nn.Sigmoid()
        self.in_dim, self.out_dim = in_dim, out_dim
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.precision = precision
        self._variant = 0
        self._grad_split = None   # set by parallel.GradBucket(overlap=True): run the backward as two ranges (see backward)
        self._ws_scratch = None   # forwards that keep nothing for a backward (eval(), no_grad)
        self._ws_saved = None     # the newest training forward's activations until its backward has run
        self._ws_saved_token = None

    def _blocks(self):
        for blk in (self.down_block_1, self.down_block_2, self.down_block_3, self.bottleneck, self.up_block_1,
                    self.up_block_2, self.up_block_3):
            for name in ("conv_1", "conv_2", "conv_3"):
                if hasattr(blk, name):
                    yield getattr(blk, name)

    def _state_tensors(self):
        """104 tensors in state_dict order (what tnb_tracknet_forward expects as ``params``)."""
        out = []
        for b in self._blocks():
            out += [b.conv.weight, b.bn.weight, b.bn.bias, b.bn.running_mean, b.bn.running_var,
                    b.bn.num_batches_tracked]
        out += [self.predictor.weight, self.predictor.bias]
        return out

    def _workspace(self, nbytes, device, saves_state):
        """(buffer, token). A forward whose backward may follow (``saves_state``) leaves its activations, BatchNorm
        statistics and packed weights in the workspace, so that buffer must not be handed out again before the backward
        has consumed it. The common train loop (forward, backward, forward, ...) reuses ONE cached buffer - a stable
        address is also what lets the library replay its CUDA graphs; a second forward before the first one's backward
        (gradient accumulation over micro-batches, an evaluation in between) gets a private buffer instead."""
        fits = lambda t: t is not None and t.numel() >= nbytes and t.device == device
        if not saves_state:
            if not fits(self._ws_scratch):
                self._ws_scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
            return self._ws_scratch, None
        token = _WorkspaceToken()
        held = self._ws_saved_token() if self._ws_saved_token is not None else None
        if held is not None and not held.done:
            return torch.empty(nbytes, dtype=torch.uint8, device=device), token
        if not fits(self._ws_saved):
            self._ws_saved = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self._ws_saved_token = weakref.ref(token)
        return self._ws_saved, token

    def forward(self, x):
        _lib.require_cuda(x)
        if x.dim() != 4 or x.shape[1] != self.in_dim:
            raise RuntimeError(f"TrackNet expects input (N, {self.in_dim}, H, W), got {tuple(x.shape)}")
        dev = x.device
        for t in self._state_tensors():
            _lib.require_cuda(t)
            if t.device != dev or not t.is_contiguous() or t.dtype != (torch.int64 if t.dim() == 0 else torch.float32):
                raise RuntimeError("tracknet_b200: parameters and buffers must be contiguous fp32 tensors on the input's "
                                   f"device (found {t.dtype} on {t.device}, input on {dev}); .half() / .double() models "
                                   "are not supported")
        x = x.contiguous().float()
        # decided HERE: inside autograd.Function.forward grad mode is always off
        saves_state = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        return _TrackNetFunction.apply(x, self, saves_state, *self.parameters())


class Conv1DBlock(nn.Module):
    """ Conv1D + LeakyReLU (parameter container; reference model.py:76-86) """

    def __init__(self, in_dim, out_dim, **kwargs):
        super(Conv1DBlock, self).__init__(**kwargs)
        self.conv = nn.Conv1d(in_dim, out_dim, kernel_size=3, padding='same', bias=True)
        self.relu = nn.LeakyReLU()


class Double1DConv(nn.Module):
    """ Conv1DBlock x 2 (reference model.py:88-98) """

    def __init__(self, in_dim, out_dim):
        super(Double1DConv, self).__init__()
        self.conv_1 = Conv1DBlock(in_dim, out_dim)
        self.conv_2 = Conv1DBlock(out_dim, out_dim)


class _InpaintNetFunction(torch.autograd.Function):
    """autograd bridge: forward = tnb_inpaintnet_fwd, backward = tnb_inpaintnet_bwd (recomputes the forward)."""

    @staticmethod
    def forward(ctx, x, m, module, *params):
        lib = _lib.load()
        n, l = x.shape[0], x.shape[1]
        out = torch.empty((n, l, 2), dtype=torch.float32, device=x.device)
        _lib.check(lib.tnb_inpaintnet_fwd(x.data_ptr(), m.data_ptr(), _lib.ptr_array(params), n, l,
                                          out.data_ptr(), _lib.stream_ptr()))
        ctx.save_for_backward(x, m, *params)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x, m, *params = ctx.saved_tensors
        n, l = x.shape[0], x.shape[1]
        flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=x.device)
        grads, off = [], 0
        for p in params:
            grads.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dout = dout.contiguous().float()
        _lib.check(lib.tnb_inpaintnet_bwd(x.data_ptr(), m.data_ptr(), _lib.ptr_array(params), dout.data_ptr(),
                                          _lib.ptr_array(grads), n, l, dx.data_ptr() if dx is not None else None,
                                          _lib.stream_ptr()))
        return (dx, None, None) + tuple(grads)


class InpaintNet(nn.Module):
    """Drop-in for reference ``model.InpaintNet`` (model.py:100-129): forward is one fused kernel, and so is its
    autograd backward (what ``train.py:147-166`` needs)."""

    def __init__(self):
        super(InpaintNet, self).__init__()
        self.down_1 = Conv1DBlock(3, 32)
        self.down_2 = Conv1DBlock(32, 64)
        self.down_3 = Conv1DBlock(64, 128)
        self.buttleneck = Double1DConv(128, 256)
        self.up_1 = Conv1DBlock(384, 128)
        self.up_2 = Conv1DBlock(192, 64)
        self.up_3 = Conv1DBlock(96, 32)
        self.predictor = nn.Conv1d(32, 2, 3, padding='same')
        self.sigmoid = nn.Sigmoid()

    def _param_tensors(self):
        convs = [self.down_1.conv, self.down_2.conv, self.down_3.conv, self.buttleneck.conv_1.conv,
                 self.buttleneck.conv_2.conv, self.up_1.conv, self.up_2.conv, self.up_3.conv, self.predictor]
        out = []
        for c in convs:
            out += [c.weight, c.bias]
        return out

    def _check(self, x, m):
        _lib.require_cuda(x, m)
        if x.dim() != 3 or x.shape[2] != 2 or m.dim() != 3 or m.shape[2] != 1 or m.shape[:2] != x.shape[:2]:
            raise RuntimeError(f"InpaintNet expects x (N, L, 2) and m (N, L, 1), got {tuple(x.shape)} {tuple(m.shape)}")
        tensors = self._param_tensors()
        for t in tensors:
            _lib.require_cuda(t)
            if t.device != x.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("tracknet_b200: InpaintNet parameters must be contiguous fp32 tensors on the input's "
                                   f"device (found {t.dtype} on {t.device}, input on {x.device})")
        return x.contiguous().float(), m.contiguous().float(), tensors

    def forward(self, x, m):
        x, m, tensors = self._check(x, m)
        return _InpaintNetFunction.apply(x, m, self, *tensors)

    @torch.no_grad()
    def rectify(self, coor_pred, inpaint_mask, coor_th):
        """The inference step of reference predict.py:256-261 / test.py:400-408 in ONE launch:
        ``c = self(coor_pred, mask); c = c * mask + coor_pred * (1 - mask); c[(c[..., 0] < th) & (c[..., 1] < th)] = 0``."""
        x, m, tensors = self._check(coor_pred, inpaint_mask)
        lib = _lib.load()
        out = torch.empty_like(x)
        _lib.check(lib.tnb_inpaintnet_rectify(x.data_ptr(), m.data_ptr(), _lib.ptr_array(tensors), x.shape[0], x.shape[1],
                                              float(coor_th), out.data_ptr(), _lib.stream_ptr()))
        return out
