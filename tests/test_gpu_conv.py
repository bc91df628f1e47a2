"""-m gpu: the tcgen05 implicit-GEMM kernels (conv fwd / dgrad / wgrad) against the CPU fp32 oracle ops."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tracknetv3_b200 import _lib
from tests import gpu_util as G

pytestmark = pytest.mark.gpu

TOL = {3: 2e-5, 1: 4e-3}  # relative to the output max-abs: 3-term split is fp32-faithful, 1-term is TF32-class


@pytest.fixture(params=[0, 1], ids=["fp16x2^k", "bf16"])
def bwd_fmt(request):
    """format of the backward pass's pre-split operands: fp16 pairs of x * 2^k (the network's backward), bf16 pairs"""
    old, G.BWD_FMT = G.BWD_FMT, request.param
    yield request.param
    G.BWD_FMT = old


def _dz_src(dz_nhwc, cout, h, w, fmt, keep):
    """pre-split gradient tensor as a dgrad view source; fp16: stored multiplied by 2^k, `scale` points at the multiplier"""
    mul = G.pow2_mul(dz_nhwc) if fmt == 0 else 1.0
    ts = G.presplit(dz_nhwc, fmt, mul)
    mul_dev = torch.tensor([mul], device=G.DEV)
    keep += [ts, mul_dev]
    src = _lib.Src(ptr=ts.data_ptr(), scale=mul_dev.data_ptr() if fmt == 0 else None, shift=None, C=cout, Hs=h, Ws=w,
                   mode=_lib.SRC_PRESPLIT)
    return src, ts, mul


def _rand(*shape, seed=0, scale=1.0):
    return (torch.rand(*shape, generator=torch.Generator().manual_seed(seed)) * 2 - 1) * scale


@pytest.mark.parametrize("terms", [3, 1])
@pytest.mark.parametrize("n,h,w,cin,cout", [
    (1, 16, 16, 32, 64),     # one full tile
    (2, 16, 24, 64, 64),     # batch, 3 column tiles of 8
    (1, 20, 40, 64, 128),    # ragged rows (20 = 16 + 4) and columns
    (1, 8, 8, 128, 256),     # image smaller than a tile, BN = 256
    (1, 24, 16, 192, 64),    # 6 K chunks
    (1, 16, 16, 256, 512),   # two N tiles
    (1, 16, 32, 64, 192),    # BN = 192
    (1, 16, 16, 96, 384),    # 2 x BN 192
    (1, 36, 64, 256, 512),   # bottleneck shape: 8-row x 16-column tiles (36 = 4.5 of them), two N tiles
    (1, 40, 16, 64, 64),     # 32-row x 16-column tiles (MT = 4 stacked down the image), ragged rows
    (1, 8, 32, 128, 256),    # one row of 8 x 16 tiles
])
def test_conv3x3_forward_identity_view(n, h, w, cin, cout, terms):
    x = _rand(n, cin, h, w, seed=1)
    wt = _rand(cout, cin, 3, 3, seed=2, scale=0.2)
    ref = F.conv2d(x, wt, padding=1)
    xin = G.nhwc(x)
    view = G.make_view([G.make_src(xin)], n, h, w)
    out, part = G.conv3x3(view, wt.to(G.DEV), cout, terms=terms, stats=True)
    assert G.rel_err(G.nchw(out), ref) < TOL[terms]
    # BatchNorm partials emitted by the epilogue: per-channel sum and sum of squares of the OUTPUT
    s = part.sum(0).cpu()
    assert torch.allclose(s[0], ref.sum((0, 2, 3)), rtol=1e-3, atol=2e-2 * ref.abs().max().item())
    assert torch.allclose(s[1], (ref * ref).sum((0, 2, 3)), rtol=2e-3 if terms == 3 else 2e-2)


def test_conv3x3_first_layer_padding_27_channels():
    """Reference layer 1 has 27 input channels (seq_len 8, bg concat): padded to 32 with zero weights."""
    L = G.lib()
    x = torch.rand(2, 27, 16, 16, generator=torch.Generator().manual_seed(3))
    wt = _rand(64, 27, 3, 3, seed=4, scale=0.2)
    xin = torch.empty(2, 16, 16, 32, device=G.DEV)
    xg = x.to(G.DEV)
    _lib.check(L.tnb_pack_nchw_to_nhwc(xg.data_ptr(), xin.data_ptr(), 2, 27, 16, 16, 32, G.st()))
    assert torch.equal(xin[..., :27].cpu(), x.permute(0, 2, 3, 1)) and xin[..., 27:].abs().max() == 0
    out, _ = G.conv3x3(G.make_view([G.make_src(xin)], 2, 16, 16), wt.to(G.DEV), 64)
    assert G.rel_err(G.nchw(out), F.conv2d(x, wt, padding=1)) < TOL[3]


@pytest.mark.parametrize("mode", ["affine_relu", "pool", "up_concat"])
def test_conv3x3_fused_views(mode, bwd_fmt):
    """BatchNorm-apply + ReLU (+ MaxPool | Upsample + cat) fused into the operand load (model.py:14-15,59-69)."""
    n, h, w = 1, 16, 24
    gen = torch.Generator().manual_seed(5)
    keep = []  # every device tensor whose raw pointer sits in a descriptor must outlive the launch

    def dev(t):
        keep.append(t.contiguous().to(G.DEV))
        return keep[-1]

    if mode == "up_concat":
        z0 = _rand(n, 64, h // 2, w // 2, seed=6); z1 = _rand(n, 32, h, w, seed=7)
        sc0, sh0 = torch.rand(64, generator=gen) + 0.5, _rand(64, seed=8, scale=0.3)
        sc1, sh1 = -(torch.rand(32, generator=gen) + 0.5), _rand(32, seed=9, scale=0.3)  # negative scale too
        a0 = F.relu(z0 * sc0[None, :, None, None] + sh0[None, :, None, None])
        a1 = F.relu(z1 * sc1[None, :, None, None] + sh1[None, :, None, None])
        xin = torch.cat([F.interpolate(a0, scale_factor=2, mode="nearest"), a1], 1)
        srcs = [G.make_src(dev(z0.permute(0, 2, 3, 1)), _lib.SRC_AFFINE_RELU_UP, dev(sc0), dev(sh0)),
                G.make_src(dev(z1.permute(0, 2, 3, 1)), _lib.SRC_AFFINE_RELU, dev(sc1), dev(sh1))]
    else:
        hs, ws = (2 * h, 2 * w) if mode == "pool" else (h, w)
        z0 = _rand(n, 64, hs, ws, seed=6)
        sc0 = torch.where(torch.arange(64) % 3 == 0, -1.0, 1.0) * (torch.rand(64, generator=gen) + 0.5)
        sh0 = _rand(64, seed=8, scale=0.3)
        a0 = F.relu(z0 * sc0[None, :, None, None] + sh0[None, :, None, None])
        xin = F.max_pool2d(a0, 2, 2) if mode == "pool" else a0
        srcs = [G.make_src(dev(z0.permute(0, 2, 3, 1)),
                           _lib.SRC_AFFINE_RELU_POOL if mode == "pool" else _lib.SRC_AFFINE_RELU, dev(sc0), dev(sh0))]
    wt = _rand(64, xin.shape[1], 3, 3, seed=10, scale=0.2)
    out, _ = G.conv3x3(G.make_view(srcs, n, h, w), dev(wt), 64)
    assert G.rel_err(G.nchw(out), F.conv2d(xin, wt, padding=1)) < TOL[3]
    # the same fused view as the wgrad B operand
    dz = _rand(n, 64, h, w, seed=17, scale=1e-5)
    wz = torch.zeros(64, xin.shape[1], 3, 3, requires_grad=True)
    (F.conv2d(xin, wz, padding=1) * dz).sum().backward()
    dw = G.wgrad3x3(G.make_view(srcs, n, h, w), dev(dz.permute(0, 2, 3, 1)), 64, xin.shape[1])
    assert G.rel_err(dw, wz.grad) < 2e-4
    # production backward: the same view materialised once in the pre-split format, wgrad fills are pure copies
    L = G.lib()
    view = G.make_view(srcs, n, h, w)
    vs = torch.empty(n * h * w * xin.shape[1] * 4, dtype=torch.uint8, device=G.DEV)
    _lib.check(L.tnb_view_presplit(C.byref(view), vs.data_ptr(), bwd_fmt, G.st()))
    assert G.max_abs(G.unsplit(vs, (n, h, w, xin.shape[1])), xin.permute(0, 2, 3, 1)) < G.split_tol() * xin.abs().max().item()
    psrc = _lib.Src(ptr=vs.data_ptr(), scale=None, shift=None, C=xin.shape[1], Hs=h, Ws=w, mode=_lib.SRC_PRESPLIT)
    dw2 = G.wgrad3x3(G.make_view([psrc], n, h, w), keep[-1], 64, xin.shape[1])
    assert G.rel_err(dw2, wz.grad) < 2e-4
    del keep


@pytest.mark.parametrize("terms", [3, 1])
@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 16, 16, 64, 64), (2, 20, 24, 192, 64), (1, 16, 16, 128, 256)])
def test_conv3x3_dgrad(n, h, w, cin, cout, terms, bwd_fmt):
    """dIn = conv(dz, rot180(W)^T) through the same kernel: weights packed with mode 1, dz (~1e-6: bf16 pairs keep fp32's
    exponent range, fp16 pairs are stored multiplied by a power of two that the epilogue divides out) supplied in the
    pre-split format so that the operand fill is a pure copy."""
    dz = _rand(n, cout, h, w, seed=11, scale=1e-6)
    wt = _rand(cout, cin, 3, 3, seed=12, scale=0.2)
    ref = F.conv_transpose2d(dz, wt, padding=1)
    t = G.nhwc(dz)
    keep = []
    src, ts, mul = _dz_src(t, cout, h, w, bwd_fmt, keep)
    assert G.max_abs(G.unsplit(ts, t.shape, bwd_fmt, mul), t) < G.split_tol() * 1e-6  # hi + lo: dz to ~2^-17 / 2^-23
    out, _ = G.conv3x3(G.make_view([src], n, h, w), wt.to(G.DEV), cin, terms=terms, fmt=bwd_fmt, mode=1)
    assert G.rel_err(G.nchw(out), ref) < (2e-4 if terms == 3 else 2e-2)
    # identity (fp32) source with on-the-fly bf16 split gives the same numbers as the pre-split bf16 source
    out2, _ = G.conv3x3(G.make_view([G.make_src(t)], n, h, w), wt.to(G.DEV), cin, terms=terms, fmt=1, mode=1)
    assert G.rel_err(out2, out) < (1e-6 if bwd_fmt == 1 else (2e-4 if terms == 3 else 2e-2))


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 16, 16, 64, 64), (2, 20, 24, 128, 256), (1, 24, 40, 512, 512)])
def test_conv3x3_dgrad_fused_bn_reduce(n, h, w, cin, cout):
    """dgrad whose epilogue also produces the BatchNorm-backward reduction of the layer below (sum g, sum g*xhat with
    g = dL/da masked by that layer's ReLU): the gradient itself is unchanged and the column sums match fp64."""
    L = G.lib()
    dz = _rand(n, cout, h, w, seed=31, scale=1e-6)
    wt = _rand(cout, cin, 3, 3, seed=32, scale=0.2)
    din = F.conv_transpose2d(dz, wt, padding=1)                 # dL/da of the producing layer (cin channels)
    z = _rand(n, cin, h, w, seed=33)                            # that layer's raw conv output
    gen = torch.Generator().manual_seed(34)
    scale = torch.where(torch.arange(cin) % 5 == 0, -1.0, 1.0) * (torch.rand(cin, generator=gen) + 0.5)
    shift = _rand(cin, seed=35, scale=0.3)
    mean = _rand(cin, seed=36, scale=0.2)
    invstd = torch.rand(cin, generator=gen) + 0.5
    bc = lambda v: v[None, :, None, None]
    g = torch.where(z * bc(scale) + bc(shift) > 0, din, torch.zeros_like(din)).double()
    ref1 = g.sum((0, 2, 3))
    ref2 = (g * ((z.double() - bc(mean).double()) * bc(invstd).double())).sum((0, 2, 3))
    ts = G.presplit(G.nhwc(dz), 1)  # this (experimental) entry point takes bf16 pairs only
    src = _lib.Src(ptr=ts.data_ptr(), scale=None, shift=None, C=cout, Hs=h, Ws=w, mode=_lib.SRC_PRESPLIT)
    view = G.make_view([src], n, h, w)
    wp = G.pack_weights(wt.to(G.DEV), 1, 1)
    rows = L.tnb_conv3x3_dgrad_bnreduce_rows(n, h, w, cout, cin, 3)
    out = torch.full((n, h, w, cin), float("nan"), device=G.DEV)
    part = torch.full((rows, 2, cin), float("nan"), device=G.DEV)
    dev = [t.to(G.DEV).contiguous() for t in (G.nhwc(z).cpu(), scale, shift, mean, invstd)]
    _lib.check(L.tnb_conv3x3_dgrad_bnreduce(C.byref(view), wp.data_ptr(), out.data_ptr(), part.data_ptr(), cin, 3,
                                            *[t.data_ptr() for t in dev], G.st()))
    torch.cuda.synchronize()
    assert G.rel_err(G.nchw(out), din) < 2e-4
    sums = part.double().sum(0).cpu()
    # a few ReLU masks sit within rounding of zero and dL/da carries the 2e-5 dgrad error: compare against the scale
    # of the absolute column sums
    tol1 = 3e-4 * g.abs().sum((0, 2, 3)) + 1e-12
    assert ((sums[0] - ref1).abs() <= tol1).all()
    tol2 = 3e-4 * (g.abs() * ((z.double() - bc(mean).double()) * bc(invstd).double()).abs()).sum((0, 2, 3)) + 1e-12
    assert ((sums[1] - ref2).abs() <= tol2).all()


@pytest.mark.parametrize("terms", [3, 1])
@pytest.mark.parametrize("n,h,w,cin,cin_real,cout", [
    (1, 8, 16, 32, 27, 64),      # first layer: 27 real channels, one K tile
    (2, 16, 32, 64, 64, 64),     # Cout 64 (half of the 128 accumulator rows unused)
    (1, 20, 24, 64, 64, 128),    # ragged tiles
    (1, 16, 16, 192, 192, 64),   # NT = 96, two ci tiles
    (1, 8, 16, 128, 128, 256),   # two co tiles, NT = 128
    (2, 12, 40, 256, 256, 128),  # ragged 4x16 K tiles, two ci tiles of 128
])
def test_conv3x3_wgrad(n, h, w, cin, cin_real, cout, terms, bwd_fmt):
    x = _rand(n, cin, h, w, seed=13)
    x[:, cin_real:] = 0
    dz = _rand(n, cout, h, w, seed=14, scale=1e-5)
    xr = x[:, :cin_real].clone().requires_grad_(False)
    wt = torch.zeros(cout, cin_real, 3, 3, requires_grad=True)
    (F.conv2d(xr, wt, padding=1) * dz).sum().backward()
    t, d = G.nhwc(x), G.nhwc(dz)
    dw = G.wgrad3x3(G.make_view([G.make_src(t)], n, h, w), d, cout, cin_real, terms=terms)
    assert G.rel_err(dw, wt.grad) < (2e-4 if terms == 3 else 2e-2)
    dw_ws = G.wgrad3x3(G.make_view([G.make_src(t)], n, h, w), d, cout, cin_real, terms=terms, scratch=True)
    assert G.rel_err(dw_ws, wt.grad) < (2e-4 if terms == 3 else 2e-2)


@pytest.mark.parametrize("terms", [3, 1])
@pytest.mark.parametrize("n,h,w,cin,cin_real", [
    (1, 8, 16, 32, 27),       # first layer: 32-channel tile, rows {dy 0 | 1 | 2 | unused}
    (2, 16, 32, 64, 64),      # 64-channel tile, rows {dy | dy + 1}
    (1, 21, 37, 64, 64),      # ragged tiles on both axes
    (2, 12, 40, 192, 192),    # up_block_3.conv_1: three ci tiles
    (1, 4, 16, 128, 128),     # a single K tile per CTA
])
def test_conv3x3_wgrad_tap_stacked(n, h, w, cin, cin_real, terms, bwd_fmt):
    """Cout = 64 with a pre-split view takes the tap-stacked kernel (two filter rows per MMA); the generic kernel on
    the same operands (variant bit 32) must agree with it and both with the oracle op."""
    cout = 64
    x = _rand(n, cin, h, w, seed=23)
    x[:, cin_real:] = 0
    dz = _rand(n, cout, h, w, seed=24, scale=1e-5)
    wt = torch.zeros(cout, cin_real, 3, 3, requires_grad=True)
    (F.conv2d(x[:, :cin_real], wt, padding=1) * dz).sum().backward()
    t, d = G.nhwc(x), G.nhwc(dz)
    ts = G.presplit(t)
    src = _lib.Src(ptr=ts.data_ptr(), scale=None, shift=None, C=cin, Hs=h, Ws=w, mode=_lib.SRC_PRESPLIT)
    dw = G.wgrad3x3(G.make_view([src], n, h, w), d, cout, cin_real, terms=terms)
    tol = 2e-4 if terms == 3 else 2e-2
    assert G.rel_err(dw, wt.grad) < tol
    dw_generic = G.wgrad3x3(G.make_view([src], n, h, w), d, cout, cin_real, terms=terms, variant=32)
    assert G.rel_err(dw_generic, wt.grad) < tol
    assert G.rel_err(dw, dw_generic) < (1e-5 if terms == 3 else 1e-3)
    # tap-major accumulation buffer + scatter (what the network's backward pass uses), both kernels
    for variant in (0, 32):
        dw_ws = G.wgrad3x3(G.make_view([src], n, h, w), d, cout, cin_real, terms=terms, variant=variant, scratch=True)
        assert G.rel_err(dw_ws, wt.grad) < tol


@pytest.mark.parametrize("terms", [3, 1])
@pytest.mark.parametrize("n,h,w,cin,cout", [
    (1, 8, 16, 128, 256),     # one pair per filter row, two K tiles
    (2, 12, 40, 256, 256),    # ragged 4x16 K tiles, two input-channel tiles
    (1, 20, 24, 128, 512),    # two pairs along the output channels
    (3, 16, 48, 384, 256),    # three input-channel tiles, K range split over several CTAs
])
def test_conv3x3_wgrad_cta_pair(n, h, w, cin, cout, terms, bwd_fmt):
    """Cout % 256 == 0 with pre-split operands and a 128-channel input tile runs on CTA pairs (tcgen05 cta_group::2:
    each CTA loads its own 128 dz rows and half of the view tile); the single-CTA kernel on the same operands (variant
    bit 64) must agree with it and both with the oracle op."""
    x = _rand(n, cin, h, w, seed=33)
    dz = _rand(n, cout, h, w, seed=34, scale=1e-5)
    wt = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    (F.conv2d(x, wt, padding=1) * dz).sum().backward()
    t, d = G.nhwc(x), G.nhwc(dz)
    ts = G.presplit(t)
    src = _lib.Src(ptr=ts.data_ptr(), scale=None, shift=None, C=cin, Hs=h, Ws=w, mode=_lib.SRC_PRESPLIT)
    tol = 2e-4 if terms == 3 else 2e-2
    dw = G.wgrad3x3(G.make_view([src], n, h, w), d, cout, cin, terms=terms)
    assert G.rel_err(dw, wt.grad) < tol
    dw_single = G.wgrad3x3(G.make_view([src], n, h, w), d, cout, cin, terms=terms, variant=64)
    assert G.rel_err(dw_single, wt.grad) < tol
    assert G.rel_err(dw, dw_single) < (1e-5 if terms == 3 else 1e-3)
    dw_ws = G.wgrad3x3(G.make_view([src], n, h, w), d, cout, cin, terms=terms, scratch=True)
    assert G.rel_err(dw_ws, wt.grad) < tol


@pytest.mark.parametrize("c_up,c_skip,cout", [(128, 64, 64), (64, 64, 64), (256, 128, 128), (512, 256, 256)])
def test_conv3x3_wgrad_concat_of_half_resolution_source(c_up, c_skip, cout, bwd_fmt):
    """Decoder concat view [upsample(x), skip] given as two pre-split sources, the first one at half resolution
    (TNB_SRC_PRESPLIT_UP: the fill reads pixel (h/2, w/2)): Cout 64 runs the tap-stacked kernel with per-tile source
    selection, larger Cout the generic kernel (one launch per source)."""
    n, h, w = 2, 12, 24
    lo = _rand(n, c_up, h // 2, w // 2, seed=41)
    sk = _rand(n, c_skip, h, w, seed=42)
    xin = torch.cat([F.interpolate(lo, scale_factor=2, mode="nearest"), sk], 1)
    dz = _rand(n, cout, h, w, seed=43, scale=1e-5)
    wt = torch.zeros(cout, c_up + c_skip, 3, 3, requires_grad=True)
    (F.conv2d(xin, wt, padding=1) * dz).sum().backward()
    ps_lo, ps_sk = G.presplit(G.nhwc(lo)), G.presplit(G.nhwc(sk))
    s0 = _lib.Src(ptr=ps_lo.data_ptr(), scale=None, shift=None, C=c_up, Hs=h // 2, Ws=w // 2, mode=_lib.SRC_PRESPLIT_UP)
    s1 = _lib.Src(ptr=ps_sk.data_ptr(), scale=None, shift=None, C=c_skip, Hs=h, Ws=w, mode=_lib.SRC_PRESPLIT)
    dw = G.wgrad3x3(G.make_view([s0, s1], n, h, w), G.nhwc(dz), cout, c_up + c_skip)
    assert G.rel_err(dw, wt.grad) < 2e-4
    if cout == 64:  # the generic kernel on the same operands
        dw2 = G.wgrad3x3(G.make_view([s0, s1], n, h, w), G.nhwc(dz), cout, c_up + c_skip, variant=32)
        assert G.rel_err(dw2, wt.grad) < 2e-4


def test_conv3x3_full_resolution_layer_vs_torch_cuda():
    """down_block_1.conv_2 shape at the reference resolution (64->64 @ 288x512), checked on-device against
    torch's fp32 convolution (TF32 disabled) - the oracle op, just executed on the GPU for speed."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = _rand(1, 64, 288, 512, seed=15).to(G.DEV)
    wt = _rand(64, 64, 3, 3, seed=16, scale=0.1).to(G.DEV)
    ref = F.conv2d(x, wt, padding=1)
    t = x.permute(0, 2, 3, 1).contiguous()
    out, part = G.conv3x3(G.make_view([G.make_src(t)], 1, 288, 512), wt, 64, stats=True)
    assert G.rel_err(G.nchw(out), ref) < 2e-5
    assert torch.allclose(part.sum(0)[0], ref.sum((0, 2, 3)), rtol=1e-3, atol=1.0)


@pytest.mark.parametrize("n,h,w,c,cout", [
    (2, 32, 48, 27, 64),      # the network's first layer: 27 real channels padded to 32, one chunk
    (1, 20, 40, 12, 64),      # ragged rows and columns: the padding ring AND the tile overhang are TMA zero-fill
    (1, 16, 16, 40, 128),     # two 32-channel chunks (cpad 64), BN = 128
])
def test_conv3x3_forward_tensor_tma_input(n, h, w, c, cout):
    """TNB_SRC_PLANAR16: the network input packed as planar fp16 (hi, lo) planes and staged by tensor-TMA
    (cp.async.bulk.tensor, padding ring zero-filled by the hardware). Same split, same products: the result must equal the
    gather path on the fp32 NHWC tensor BIT FOR BIT, and both the oracle op."""
    L = G.lib()
    cpad = (c + 31) // 32 * 32
    x = torch.rand(n, c, h, w, generator=torch.Generator().manual_seed(51))
    wt = _rand(cout, c, 3, 3, seed=52, scale=0.2)
    xd = x.to(G.DEV).contiguous()
    planar = torch.full((n * h * w * cpad * 4,), 255, dtype=torch.uint8, device=G.DEV)
    nhwc = torch.full((n, h, w, cpad), float("nan"), device=G.DEV)
    _lib.check(L.tnb_pack_nchw_to_planar16(xd.data_ptr(), planar.data_ptr(), nhwc.data_ptr(), n, c, h, w, cpad, G.st()))
    torch.cuda.synchronize()
    ref_nhwc = torch.zeros(n, h, w, cpad)
    ref_nhwc[..., :c] = x.permute(0, 2, 3, 1)
    assert torch.equal(nhwc.cpu(), ref_nhwc)
    # the planar tensor: [N][chunk][term][plane][H][W][8] fp16, hi + lo == x to 2^-22
    pl = planar.view(torch.float16).reshape(n, cpad // 32, 2, 4, h, w, 8).float()
    back = (pl[:, :, 0] + pl[:, :, 1]).permute(0, 3, 4, 1, 2, 5).reshape(n, h, w, cpad)  # -> [n, h, w, chunk, plane, 8]
    assert G.max_abs(back, ref_nhwc) < 3e-7
    wpad = torch.zeros(cout, cpad, 3, 3)
    wpad[:, :c] = wt
    wdev = wpad.to(G.DEV)
    src_t = _lib.Src(ptr=planar.data_ptr(), scale=None, shift=None, C=cpad, Hs=h, Ws=w, mode=_lib.SRC_PLANAR16)
    out_t, part_t = G.conv3x3(G.make_view([src_t], n, h, w), wdev, cout, stats=True)
    out_g, part_g = G.conv3x3(G.make_view([G.make_src(nhwc)], n, h, w), wdev, cout, stats=True)
    assert torch.equal(out_t, out_g) and torch.equal(part_t, part_g)
    assert G.rel_err(G.nchw(out_t), F.conv2d(x, wt, padding=1)) < TOL[3]
