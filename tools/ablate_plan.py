"""Tile-plan experiment for the conv kernel: time forward-configuration layers under the current plan; run once with
TNB_CONV_PLAN=0 and once with TNB_CONV_PLAN=1 (half-width tiles, TMEM accumulator double-buffered for BN >= 128)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ablate_conv import time_conv  # noqa: E402

if __name__ == "__main__":
    shapes = [(10, 144, 256, 128, 128), (10, 72, 128, 256, 256), (10, 36, 64, 512, 512), (10, 144, 256, 384, 128),
              (10, 72, 128, 768, 256), (10, 72, 128, 128, 256)]
    print("TNB_CONV_PLAN =", os.environ.get("TNB_CONV_PLAN", "0"))
    for shp in shapes:
        n, h, w, cin, cout = shp
        gf = 2.0 * n * h * w * cin * cout * 9 / 1e9
        ms = time_conv(n, h, w, cin, cout, 0, 3, reps=10)
        print(f"   shape {shp}: {ms:7.3f} ms  ({gf / ms:8.1f} TFLOP/s-alg)")
