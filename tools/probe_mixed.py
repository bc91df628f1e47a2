"""Probe: does tcgen05.mma kind::f16 accept A = bf16 with B = fp16 (independent descriptor fields)?"""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as G
n, h, w, cin, cout = 2, 16, 32, 64, 128
x = torch.rand(n, cin, h, w) - 0.5
dz = (torch.rand(n, cout, h, w) - 0.5) * 1e-5
wt = torch.zeros(cout, cin, 3, 3, requires_grad=True)
(F.conv2d(x, wt, padding=1) * dz).sum().backward()
tx, tz = G.nhwc(x), G.nhwc(dz)
for variant in (0, 4):
    try:
        dw = G.wgrad3x3(G.make_view([G.make_src(tx)], n, h, w), tz, cout, cin, variant=variant)
        print(f"variant {variant}: rel err {G.rel_err(dw, wt.grad):.3e}")
    except Exception as e:
        print(f"variant {variant}: FAILED {e}")
