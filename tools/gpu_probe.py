"""Bring-up probe for the tcgen05 kernels (run on the GPU box): structured inputs whose outputs reveal
WHICH (pixel, channel) every accumulator element actually read, for each descriptor variant.
usage: python tools/gpu_probe.py conv|wgrad <variant>"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as G  # noqa: E402


def probe_conv(variant, terms=3):
    n, h, w, cin, cout = 1, 16, 16, 32, 64
    x = torch.zeros(n, cin, h, w)
    for c in range(cin):
        x[0, c] = (torch.arange(h * w).reshape(h, w) * cin + c + 1).float()
    wt = torch.zeros(cout, cin, 3, 3)
    for co in range(32):
        wt[co, co, 1, 1] = 1.0          # centre tap identity
        wt[32 + co, co, 0, 0] = 1.0     # tap (dy=0,dx=0): reads pixel (h-1, w-1)
    ref = F.conv2d(x, wt, padding=1)
    xin = G.nhwc(x)
    out, part = G.conv3x3(G.make_view([G.make_src(xin)], n, h, w), wt.to(G.DEV), cout, terms=terms, variant=variant,
                          stats=True)
    out = G.nchw(out).cpu()
    bad = (out != ref)
    print(f"[conv variant {variant} terms {terms}] mismatches {int(bad.sum())} / {bad.numel()} "
          f"max_abs {float((out - ref).abs().nan_to_num(1e30).max()):.4g}")
    if bad.any():
        idx = bad.nonzero()[:24]
        for (_, co, hh, ww) in idx.tolist():
            v = out[0, co, hh, ww].item()
            src = "nan" if v != v else (f"pix {int(v - 1) // cin} ch {int(v - 1) % cin}" if v >= 1 else f"val {v}")
            print(f"   out[co={co} h={hh} w={ww}] = {v}  expected {ref[0, co, hh, ww].item()}  -> reads {src}")
    s = part.sum(0).cpu()
    print("   stats sum ok:", bool(torch.allclose(s[0], ref.sum((0, 2, 3)), rtol=1e-4)))
    # random data check too
    x = torch.rand(2, 64, 20, 24) - 0.5
    wt = (torch.rand(128, 64, 3, 3) - 0.5) * 0.2
    ref = F.conv2d(x, wt, padding=1)
    t = G.nhwc(x)
    out, _ = G.conv3x3(G.make_view([G.make_src(t)], 2, 20, 24), wt.to(G.DEV), 128, terms=terms, variant=variant)
    print(f"   random 64->128 20x24 rel err {G.rel_err(G.nchw(out), ref):.3e}")


def probe_wgrad(variant, terms=3):
    n, h, w, cin, cout = 1, 8, 16, 32, 64
    x = torch.zeros(n, cin, h, w)
    for c in range(cin):
        x[0, c] = (torch.arange(h * w).reshape(h, w) * cin + c + 1).float()
    dz = torch.zeros(n, cout, h, w)
    p0 = (3, 5)
    for co in range(cout):
        dz[0, co, p0[0], p0[1]] = 1.0 if co % 2 == 0 else 2.0
    xr = x.clone()
    wt = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    (F.conv2d(xr, wt, padding=1) * dz).sum().backward()
    ref = wt.grad
    tx, tz = G.nhwc(x), G.nhwc(dz)  # keep alive: descriptors hold raw pointers
    dw = G.wgrad3x3(G.make_view([G.make_src(tx)], n, h, w), tz, cout, cin, terms=terms, variant=variant).cpu()
    bad = dw != ref
    print(f"[wgrad variant {variant} terms {terms}] mismatches {int(bad.sum())} / {bad.numel()} "
          f"max_abs {float((dw - ref).abs().nan_to_num(1e30).max()):.4g}")
    if bad.any():
        for (co, ci, dy, dx) in bad.nonzero()[:24].tolist():
            v = dw[co, ci, dy, dx].item()
            print(f"   dw[co={co} ci={ci} dy={dy} dx={dx}] = {v} expected {ref[co, ci, dy, dx].item()}")
    x = torch.rand(2, 64, 16, 32) - 0.5
    dz = (torch.rand(2, 128, 16, 32) - 0.5) * 1e-4
    wt = torch.zeros(128, 64, 3, 3, requires_grad=True)
    (F.conv2d(x, wt, padding=1) * dz).sum().backward()
    tx, tz = G.nhwc(x), G.nhwc(dz)
    dw = G.wgrad3x3(G.make_view([G.make_src(tx)], 2, 16, 32), tz, 128, 64, terms=terms, variant=variant)
    print(f"   random 64->128 16x32 rel err {G.rel_err(dw, wt.grad):.3e}")


if __name__ == "__main__":
    kind, variant = sys.argv[1], int(sys.argv[2])
    print(torch.cuda.get_device_name(0))
    {"conv": probe_conv, "wgrad": probe_wgrad}[kind](variant)
