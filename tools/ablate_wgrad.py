"""Bottleneck ablation of the wgrad kernels (pre-split operands). variant 0 = what the network runs (CTA pairs where
Cout % 256 == 0, the tap-stacked kernel for Cout = 64); bit 64 = single-CTA kernel, on which the ablation bits act:
4 = no MMA, 8 = no fill, 16 = no atomics."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as G
from tracknetv3_b200 import _lib
import ctypes as C
L = G.lib()

def time_wgrad(n, h, w, cin, cout, variant, terms=3, reps=5):
    x = torch.rand(n, h, w, cin, device="cuda")
    dz = (torch.rand(n, h, w, cout, device="cuda") - 0.5) * 1e-5
    xs, dzs = G.presplit(x, 1), G.presplit(dz, 1)  # bf16 pairs
    src = _lib.Src(ptr=xs.data_ptr(), scale=None, shift=None, C=cin, Hs=h, Ws=w, mode=_lib.SRC_PRESPLIT)
    view = G.make_view([src], n, h, w)
    dw = torch.zeros(cout, cin, 3, 3, device="cuda")
    def run():
        _lib.check(L.tnb_conv3x3_wgrad(C.byref(view), dzs.data_ptr(), dw.data_ptr(), cout, cin, terms, variant, 1, None, G.st()))
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

if __name__ == "__main__":
    shapes = [(10, 72, 128, 256, 256), (10, 144, 256, 128, 128), (10, 288, 512, 64, 64), (10, 288, 512, 192, 64), (10, 36, 64, 512, 512)]
    names = {0: "full (as shipped)", 64: "single-CTA full", 68: "single-CTA no-MMA", 72: "single-CTA no-fill", 80: "single-CTA no-atomics"}
    for shp in shapes:
        n, h, w, cin, cout = shp
        gf = 2.0 * n * h * w * cin * cout * 9 / 1e9
        print(f"shape {shp}: {gf:.1f} GFLOP algorithmic")
        for v, nm in names.items():
            for terms in (3,):
                ms = time_wgrad(n, h, w, cin, cout, v, terms)
                print(f"   {nm:20s} terms={terms}: {ms:7.3f} ms  ({gf / ms:8.1f} TFLOP/s-alg)")
