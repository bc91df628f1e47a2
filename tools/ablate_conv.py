"""Bottleneck ablation of the conv kernel on the GPU: time a layer shape with the MMA issue, the operand gather or the
epilogue stores switched off (variant bits 4 / 8 / 16; results are garbage in those modes, only time matters)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as G  # noqa: E402
from tracknetv3_b200 import _lib  # noqa: E402
import ctypes as C  # noqa: E402

L = G.lib()


def time_conv(n, h, w, cin, cout, variant, terms=3, reps=5):
    x = torch.rand(n, h, w, cin, device="cuda")
    sc = torch.rand(cin, device="cuda") + 0.5
    sh = torch.rand(cin, device="cuda") - 0.5
    wt = (torch.rand(cout, cin, 3, 3, device="cuda") - 0.5) * 0.1
    view = G.make_view([G.make_src(x, _lib.SRC_AFFINE_RELU, sc, sh)], n, h, w)
    wp = G.pack_weights(wt, 0, 0)
    out = torch.empty(n, h, w, cout, device="cuda")
    rows = L.tnb_conv3x3_stat_rows(n, h, w, cin, cout, terms)
    part = torch.empty(rows, 2, cout, device="cuda")
    def run():
        _lib.check(L.tnb_conv3x3_fwd(C.byref(view), wp.data_ptr(), out.data_ptr(), part.data_ptr(), cout, terms, 0,
                                     variant, G.st()))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
    shapes = [(10, 288, 512, 64, 64), (10, 144, 256, 128, 128), (10, 72, 128, 256, 256), (10, 288, 512, 192, 64)]
    names = {0: "full", 4: "no-MMA", 8: "no-gather", 16: "no-store", 12: "barriers+epilogue only", 24: "MMA only",
             28: "barriers only"}
    for shp in shapes:
        n, h, w, cin, cout = shp
        gf = 2.0 * n * h * w * cin * cout * 9 / 1e9
        print(f"shape {shp}: {gf:.1f} GFLOP algorithmic")
        for v, nm in names.items():
            for terms in (3, 1):
                ms = time_conv(n, h, w, cin, cout, v, terms)
                print(f"   {nm:24s} terms={terms}: {ms:7.3f} ms  ({gf / ms:8.1f} TFLOP/s-alg)")
