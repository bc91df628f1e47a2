"""Helpers for the -m gpu parity tests: everything below goes through the C ABI (ctypes)."""
import ctypes as C

import numpy as np
import torch

from tracknetv3_b200 import _lib
from tracknetv3_b200._lib import Src, View, GradSrc, BnBwd

DEV = "cuda"


def lib():
    return _lib.load()


def st():
    return _lib.stream_ptr()


def nhwc(t):
    """NCHW torch tensor -> contiguous NHWC CUDA fp32."""
    return t.permute(0, 2, 3, 1).contiguous().float().to(DEV)


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def make_src(t_nhwc, mode=_lib.SRC_IDENTITY, scale=None, shift=None):
    n, h, w, c = t_nhwc.shape
    return Src(ptr=t_nhwc.data_ptr(), scale=scale.data_ptr() if scale is not None else None,
               shift=shift.data_ptr() if shift is not None else None, C=c, Hs=h, Ws=w, mode=mode)


def make_view(srcs, n, h, w):
    v = View()
    v.s[0] = srcs[0]
    v.s[1] = srcs[1] if len(srcs) > 1 else srcs[0]
    v.C0 = srcs[0].C
    v.C = sum(s.C for s in srcs)
    v.N, v.H, v.W = n, h, w
    return v


def pack_weights(w_oihw, mode, fmt):
    L = lib()
    co, ci = w_oihw.shape[:2]
    kside, nside = (ci, co) if mode == 0 else (co, ci)
    out = torch.zeros(L.tnb_conv3x3_wpack_elems(kside, nside), dtype=torch.int16, device=DEV)
    _lib.check(L.tnb_conv3x3_pack_weights(w_oihw.contiguous().data_ptr(), out.data_ptr(), co, ci, mode, fmt, st()))
    return out


def conv3x3(view, w_oihw, cout, terms=3, fmt=0, variant=0, stats=False, mode=0):
    """Run tnb_conv3x3_fwd; returns (out NHWC, stat partials or None)."""
    L = lib()
    wp = pack_weights(w_oihw, mode, fmt)
    out = torch.full((view.N, view.H, view.W, cout), float("nan"), device=DEV)
    part = None
    if stats:
        rows = L.tnb_conv3x3_stat_rows(view.N, view.H, view.W, view.C, cout, terms)
        part = torch.full((rows, 2, cout), float("nan"), device=DEV)
    _lib.check(L.tnb_conv3x3_fwd(C.byref(view), wp.data_ptr(), out.data_ptr(), part.data_ptr() if stats else None,
                                 cout, terms, fmt, variant, st()))
    torch.cuda.synchronize()
    return out, part


BWD_FMT = 1  # 16-bit format of the backward pass's pre-split operands: 1 = bf16 pairs (the 3-term default), 0 = fp16 pairs of x * 2^k
             # backward uses), 1 = bf16 pairs; the `bwd_fmt` fixture of the conv tests runs both


def pow2_mul(t):
    """the power of two that brings max |t| into [2^7, 2^8) (the rule of bn_bwd_kernel's dz_format 2)"""
    m = float(t.abs().max())
    if not (m > 0.0) or not np.isfinite(m):
        return 1.0
    return float(2.0 ** (8 - np.frexp(m)[1]))


def presplit(t_nhwc, fmt=None, mul=1.0):
    """fp32 NHWC -> the pre-split format (same byte size): bf16 (hi, lo) pairs, or fp16 pairs of x * mul; returned as an
    opaque uint8 tensor."""
    L = lib()
    fmt = BWD_FMT if fmt is None else fmt
    n, h, w, c = t_nhwc.shape
    out = torch.empty(t_nhwc.numel() * 4, dtype=torch.uint8, device=DEV)
    if fmt == 1:
        assert mul == 1.0
        _lib.check(L.tnb_presplit_bf16(t_nhwc.data_ptr(), out.data_ptr(), n * h * w, c, st()))
    else:
        _lib.check(L.tnb_presplit_fp16(t_nhwc.data_ptr(), out.data_ptr(), n * h * w, c, mul, st()))
    return out


def unsplit(buf, shape_nhwc, fmt=None, mul=1.0):
    """inverse of presplit ((hi + lo) / mul) for checking: uint8 buffer -> fp32 NHWC."""
    fmt = BWD_FMT if fmt is None else fmt
    n, h, w, c = shape_nhwc
    v = buf.view(torch.bfloat16 if fmt == 1 else torch.float16).reshape(n, h, w, 2, c).double()  # [pixel][2 (hi, lo)][C]
    return ((v[..., 0, :] + v[..., 1, :]) / mul).float()


def split_tol(fmt=None):
    """relative size of what hi + lo drops: 2^-17 for bf16 pairs, 2^-23 for fp16 pairs (fp32 itself)"""
    fmt = BWD_FMT if fmt is None else fmt
    return 2e-5 if fmt == 1 else 3e-7


def wgrad3x3(view, dz_nhwc, cout, cin_real, terms=3, variant=0, scratch=False, fmt=None):
    """production configuration: dz pre-split (fp16 pairs of dz * 2^k with the multiplier in a device scalar, or bf16
    pairs), a view that is not pre-split is split to the same format on the fly. scratch=True: the deterministic split-K
    slab + ordered-sum path that tnb_tracknet_backward uses (slabs and dw start as NaN: every element must be written)."""
    L = lib()
    fmt = BWD_FMT if fmt is None else fmt
    mul = pow2_mul(dz_nhwc) if fmt == 0 else 1.0
    dzs = presplit(dz_nhwc, fmt, mul)
    mul_dev = torch.tensor([mul], device=DEV) if fmt == 0 else None
    mul_ptr = mul_dev.data_ptr() if mul_dev is not None else None
    if scratch:
        dw = torch.full((cout, cin_real, 3, 3), float("nan"), device=DEV)
        ws = torch.full((L.tnb_conv3x3_wgrad_ws_elems(C.byref(view), cout),), float("nan"), device=DEV)
        _lib.check(L.tnb_conv3x3_wgrad_ws(C.byref(view), dzs.data_ptr(), dw.data_ptr(), cout, cin_real, terms, variant,
                                          ws.data_ptr(), fmt, mul_ptr, st()))
        torch.cuda.synchronize()
        return dw
    dw = torch.zeros((cout, cin_real, 3, 3), device=DEV)
    _lib.check(L.tnb_conv3x3_wgrad(C.byref(view), dzs.data_ptr(), dw.data_ptr(), cout, cin_real, terms, variant, fmt,
                                   mul_ptr, st()))
    torch.cuda.synchronize()
    return dw


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def max_abs(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


def predictor_bwd(src, n, h, w, wd, o, dyd, yd, dA, dwp, dbp):
    """tnb_conv1x1_bias_sigmoid_bwd with its workspace (NaN-filled: every partial the final sum reads must be written)."""
    L = lib()
    ws = torch.full((L.tnb_conv1x1_bias_sigmoid_bwd_workspace_bytes(n, h, w, o) // 4,), float("nan"), device=DEV)
    _lib.check(L.tnb_conv1x1_bias_sigmoid_bwd(C.byref(src), n, h, w, wd.data_ptr(), o, dyd.data_ptr(), yd.data_ptr(),
                                              dA.data_ptr(), dwp.data_ptr(), dbp.data_ptr(), ws.data_ptr(), st()))
