"""Experiment: thread -> (plane, pixel) mapping of the pre-split cp.async fills in the wgrad kernels.
Generic kernel: variant bits 512 (4 planes per warp instruction) | 1024 (8) | 2048 (16, the shipped geometry through the
new code path); stacked kernel: TNB_WGRAD_MAP (read once per process). Prints time and deviation from variant 0."""
import os, sys
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as G
from tracknetv3_b200 import _lib
L = G.lib()


def run_wgrad(n, h, w, cin, cout, variant, terms=3, reps=5):
    torch.manual_seed(0)
    x = torch.rand(n, h, w, cin, device="cuda")
    dz = (torch.rand(n, h, w, cout, device="cuda") - 0.5) * 1e-5
    xs, dzs = G.presplit(x), G.presplit(dz)
    src = _lib.Src(ptr=xs.data_ptr(), scale=None, shift=None, C=cin, Hs=h, Ws=w, mode=_lib.SRC_PRESPLIT)
    view = G.make_view([src], n, h, w)
    dw = torch.zeros(cout, cin, 3, 3, device="cuda")
    def run():
        _lib.check(L.tnb_conv3x3_wgrad(C.byref(view), dzs.data_ptr(), dw.data_ptr(), cout, cin, terms, variant, G.st()))
    run()
    torch.cuda.synchronize()
    first = dw.clone()
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, first


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "generic"
    if which == "generic":
        shapes = [(10, 72, 128, 256, 256), (10, 144, 256, 128, 128), (10, 36, 64, 512, 512)]
        for shp in shapes:
            gf = 2.0 * shp[0] * shp[1] * shp[2] * shp[3] * shp[4] * 9 / 1e9
            ref = None
            for v, nm in ((0, "shipped"), (512, "4 planes"), (512 | 1024, "8 planes"), (512 | 2048, "16 planes (new path)")):
                ms, dw = run_wgrad(*shp, 32 | v)
                ms_nomma, _ = run_wgrad(*shp, 32 | v | 4)
                if ref is None: ref = dw
                err = ((dw - ref).abs().max() / ref.abs().max()).item()
                print(f"{shp} {nm:22s}: {ms:7.3f} ms ({gf / ms:7.1f} TF/s-alg)  fill-only {ms_nomma:7.3f} ms  rel-dev {err:.2e}", flush=True)
    else:  # stacked kernel: mapping comes from TNB_WGRAD_MAP of this process
        for shp in [(10, 288, 512, 64, 64), (10, 288, 512, 192, 64), (10, 288, 512, 32, 64)]:
            gf = 2.0 * shp[0] * shp[1] * shp[2] * shp[3] * shp[4] * 9 / 1e9
            ms, dw = run_wgrad(*shp, 0)
            print(f"stacked TNB_WGRAD_MAP={os.environ.get('TNB_WGRAD_MAP', '0')} {shp}: {ms:7.3f} ms ({gf / ms:7.1f} TF/s-alg) "
                  f"checksum {dw.double().sum().item():.9e} absmax {dw.abs().max().item():.6e}", flush=True)
