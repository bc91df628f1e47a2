"""Per-role timeline of the tcgen05 conv kernel (debug build: `make -C tracknetv3_b200/csrc TRACE=1`): clock64 stamps of
CTA 0's MMA issuer, weight loader, epilogue warp 2 and producer thread 0 for its first tiles, printed in SM clocks relative
to the start of the launch. Shows which role waits for which. usage: python tools/trace_conv.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tracknetv3_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "tracknetv3_b200", "libtracknet_b200_trace.so")
from tests import gpu_util as G  # noqa: E402

L = G.lib()
L.tnb_debug_set_trace.restype = C.c_int
L.tnb_debug_set_trace.argtypes = [C.c_void_p]
NT, NE = 24, 32


def run(name, n, h, w, cin, cout, dgrad=False):
    torch.manual_seed(0)
    wt = (torch.rand(cout, cin, 3, 3, device="cuda") - 0.5) * 0.1 if not dgrad else (torch.rand(cin, cout, 3, 3, device="cuda") - 0.5) * 0.1
    if dgrad:   # view = pre-split dz with `cin` channels (the layer's cout), output `cout` channels (the layer's cin)
        dz = torch.randn(n, h, w, cin, device="cuda")
        dzs = G.presplit(dz, 1)
        src = _lib.Src(ptr=dzs.data_ptr(), scale=None, shift=None, C=cin, Hs=h, Ws=w, mode=_lib.SRC_PRESPLIT)
        view = G.make_view([src], n, h, w)
        wp = G.pack_weights(wt, 1, 1)
        fmt = 1
    else:
        x = torch.rand(n, h, w, cin, device="cuda")
        sc, sh = torch.rand(cin, device="cuda") + 0.5, torch.rand(cin, device="cuda") - 0.5
        view = G.make_view([G.make_src(x, _lib.SRC_AFFINE_RELU, sc, sh)], n, h, w)
        wp = G.pack_weights(wt, 0, 0)
        fmt = 0
    out = torch.empty(n, h, w, cout, device="cuda")
    rows = L.tnb_conv3x3_stat_rows(n, h, w, cin, cout, 3)
    part = torch.empty(max(rows, 1), 2, cout, device="cuda")
    buf = torch.zeros(4 * NT * NE, dtype=torch.int64, device="cuda")

    def launch():
        _lib.check(L.tnb_conv3x3_fwd(C.byref(view), wp.data_ptr(), out.data_ptr(), part.data_ptr() if not dgrad else None,
                                     cout, 3, fmt, 0, G.st()))
    L.tnb_debug_set_trace(None)
    for _ in range(2):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.tnb_debug_set_trace(buf.data_ptr())
    e0.record(); launch(); e1.record()
    torch.cuda.synchronize()
    L.tnb_debug_set_trace(None)
    t = buf.cpu().reshape(4, NT, NE)
    plan = (C.c_int * 12)()
    L.tnb_conv3x3_plan_query(n, h, w, cin, cout, 3, plan)
    base = int(t[0, 0, 0])
    print(f"\n=== {name}: {n}x{h}x{w} {cin}->{cout}  {e0.elapsed_time(e1):.3f} ms  plan BN {plan[0]} MT {plan[1]} SA {plan[2]} SB {plan[3]} "
          f"G {plan[4]} nbuf {plan[5]} merged {plan[8]} tall {plan[9]}  (clocks relative to CTA 0's first stamp)")
    rel = lambda v: "-" if int(v) == 0 else str(int(v) - base)
    nch = min(cin // 32, 6)
    for k in range(1, 8):
        m = t[0, k]
        print(f" tile {k}: MMA   start {rel(m[0])} tmem_empty {rel(m[1])} | " +
              " | ".join(f"c{c}: fullA {rel(m[2 + 4 * c])} fullB {rel(m[3 + 4 * c])} issued {rel(m[4 + 4 * c])}" for c in range(nch)) +
              f" | tmem_full committed {rel(m[30])}")
        p = t[3, k]
        print(f"         PROD  start {rel(p[0])} bar1 {rel(p[1])} table {rel(p[2])} | " +
              " | ".join(f"c{c}: wait {rel(p[3 + 3 * c])} emptyA {rel(p[4 + 3 * c])} filled {rel(p[5 + 3 * c])}" for c in range(nch)))
        e = t[2, k]
        print(f"         EPI   wait {rel(e[0])} tmem_full {rel(e[1])} drained {rel(e[2])} stats {rel(e[3])}")
        ld = t[1, k]
        print("         LOAD  " + " ".join(f"s{i}: {rel(ld[2 * i])}" for i in range(min(nch * 3, 12))))
    # per-tile period in steady state
    starts = [int(t[0, k, 0]) for k in range(2, 12) if int(t[0, k, 0])]
    if len(starts) > 2:
        print(f" MMA-warp tile period: {(starts[-1] - starts[0]) / (len(starts) - 1):.0f} clocks")


if __name__ == "__main__":
    run("fwd 64->64", 10, 288, 512, 64, 64)
    run("fwd 32->64 (first layer shape)", 10, 288, 512, 32, 64)
    run("fwd 192->64", 10, 288, 512, 192, 64)
    run("fwd 128->128", 10, 144, 256, 128, 128)
    run("dgrad 64->64", 10, 288, 512, 64, 64, dgrad=True)
    run("dgrad 256->256", 10, 72, 128, 256, 256, dgrad=True)
