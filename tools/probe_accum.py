"""Is the fp32 accumulation of tcgen05.mma (TMEM) round-to-nearest or truncating? Exactly representable operands (products
and the fp64 reference sum are exact), so every bit of error is the accumulator's. Signed relative errors: truncation shows
as a systematic deficit that grows with the number of MMAs chained into one accumulator; round-to-nearest as a zero-mean
random walk. torch's fp32 convolution gradient (cuDNN, TF32 off: FFMA, round-to-nearest) on the same data is the yardstick.
usage (GPU box): python tools/probe_accum.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import gpu_util as G  # noqa: E402
from tracknetv3_b200 import _lib  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def exact_vals(shape, gen, signed):
    k = torch.randint(0, 256, shape, generator=gen).double()
    v = 1.0 + k / 256.0  # 9 significant bits: exact in fp16, exact as a bf16 (hi, lo) pair
    if signed:
        v = v * (torch.randint(0, 2, shape, generator=gen).double() * 2 - 1)
    return v


def report(name, got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    rel = (got - ref) / ref.abs().max()
    inner = rel[..., 1, 1] if rel.dim() == 4 and rel.shape[-1] == 3 else rel
    print(f"  {name:44s} signed mean {inner.mean().item():+.3e}   rms {inner.pow(2).mean().sqrt().item():.3e}   max |err| {rel.abs().max().item():.3e}")


def wgrad_case(n, h, w, cin, cout, signed, fmt):
    gen = torch.Generator().manual_seed(1)
    x = torch.ones(n, cin, h, w, dtype=torch.float64)
    dz = exact_vals((n, cout, h, w), gen, signed)
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    (F.conv2d(x, wt, padding=1) * dz).sum().backward()
    ref = wt.grad
    xs = G.presplit(G.nhwc(x.float()), fmt)
    src = _lib.Src(ptr=xs.data_ptr(), scale=None, shift=None, C=cin, Hs=h, Ws=w, mode=_lib.SRC_PRESPLIT)
    dw = G.wgrad3x3(G.make_view([src], n, h, w), G.nhwc(dz.float()), cout, cin, scratch=True, fmt=fmt)
    report(f"wgrad {cin}->{cout} {n}x{h}x{w} {'signed' if signed else 'positive'} fmt {fmt}: tcgen05", dw, ref)
    w32 = torch.zeros(cout, cin, 3, 3, device="cuda", requires_grad=True)
    (F.conv2d(x.float().cuda(), w32, padding=1) * dz.float().cuda()).sum().backward()
    report("                                   torch fp32 (cuDNN)", w32.grad, ref)


def fwd_case(n, h, w, cin, cout, signed):
    gen = torch.Generator().manual_seed(2)
    x = torch.ones(n, cin, h, w, dtype=torch.float64)
    wt = exact_vals((cout, cin, 3, 3), gen, signed)
    ref = F.conv2d(x, wt, padding=1)
    xin, wdev = G.nhwc(x.float()), wt.float().cuda()  # raw pointers sit in the descriptors: keep the tensors alive
    view = G.make_view([G.make_src(xin)], n, h, w)
    out, _ = G.conv3x3(view, wdev, cout)
    got, r = G.nchw(out)[:, :, 1:-1, 1:-1], ref[:, :, 1:-1, 1:-1]
    report(f"fwd {cin}->{cout} (K = {9 * cin}) {'signed' if signed else 'positive'}: tcgen05", got, r)
    report("                                   torch fp32 (cuDNN)", F.conv2d(x.float().cuda(), wt.float().cuda(), padding=1)[:, :, 1:-1, 1:-1], r)


def fwd_rounding_case(cin, cout, signed, terms=3):
    """x = 1, weights with 22 significant bits (exact as an fp16 (hi, lo) pair after the 2^10 pre-scale): every product is
    exact, but the running sum needs more than 24 bits, so each accumulation rounds. Truncation: the result falls short
    of the exact sum by ~(number of accumulations) * ulp / 2; round-to-nearest: zero-mean."""
    gen = torch.Generator().manual_seed(3)
    n, h, w = 1, 32, 32
    x = torch.ones(n, cin, h, w, dtype=torch.float64)
    k = torch.randint(0, 2 ** 21, (cout, cin, 3, 3), generator=gen).double()
    wt = (2.0 ** 21 + k) / 2.0 ** 22 / 64.0  # [1/128, 1/64): like real weights
    if signed:
        wt = wt * (torch.randint(0, 2, wt.shape, generator=gen).double() * 2 - 1)
    ref = F.conv2d(x, wt, padding=1)
    mag = F.conv2d(x, wt.abs(), padding=1)
    xin, wdev = G.nhwc(x.float()), wt.float().cuda()
    assert (wdev.double().cpu() == wt).all()
    view = G.make_view([G.make_src(xin)], n, h, w)
    out, _ = G.conv3x3(view, wdev, cout, terms=terms)
    t32 = F.conv2d(x.float().cuda(), wdev, padding=1)
    for name, got in ((f"tcgen05 ({terms} terms)", G.nchw(out)), ("torch fp32 (cuDNN)", t32)):
        e = ((got.double().cpu() - ref) / mag)[:, :, 1:-1, 1:-1]
        s = (e * torch.sign(ref[:, :, 1:-1, 1:-1])).mean().item()
        print(f"  fwd {cin}->{cout} K = {9 * cin:5d} {'signed  ' if signed else 'positive'} {name:20s}: error / sum |terms|: "
              f"mean toward larger magnitude {s:+.3e}   rms {e.pow(2).mean().sqrt().item():.3e}")


if __name__ == "__main__":
    print("forward, sums that need rounding (K = 9 * Cin, K / 16 accumulations per term):")
    for cin, cout in [(32, 64), (64, 64), (256, 256), (768, 256)]:
        for signed in (False, True):
            fwd_rounding_case(cin, cout, signed)
    fwd_rounding_case(768, 256, False, terms=1)
    print("exactly representable sums:")
    print("weight gradient (K = pixels of the CTA's split-K range):")
    for shape in [(1, 16, 16, 64, 64), (2, 96, 160, 64, 64), (2, 96, 160, 128, 256), (10, 288, 512, 64, 64), (10, 72, 128, 256, 256)]:
        for signed in (False, True):
            wgrad_case(*shape, signed, 1)
    wgrad_case(2, 96, 160, 64, 64, False, 0)
    print("forward (K = 9 * Cin):")
    for cin, cout in [(64, 64), (256, 256), (768, 256)]:
        for signed in (False, True):
            fwd_case(1, 32, 32, cin, cout, signed)
