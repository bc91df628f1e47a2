"""Ablation of the CTA-pair wgrad kernel: variant bits 4 = no MMA, 8 = no fill, bits 7-8 = plane-stride padding mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.ablate_wgrad import time_wgrad

if __name__ == "__main__":
    for shp in [(10, 72, 128, 256, 256), (10, 36, 64, 512, 512), (10, 72, 128, 768, 256)]:
        gf = 2.0 * shp[0] * shp[1] * shp[2] * shp[3] * shp[4] * 9 / 1e9
        print(f"shape {shp}: {gf:.1f} GFLOP algorithmic")
        for pad in (0, 1, 2, 3):
            row = []
            for nm, v in (("full", 0), ("no-MMA", 4), ("no-fill", 8)):
                ms = time_wgrad(*shp, v | (pad << 7), 3)
                row.append(f"{nm} {ms:6.3f} ms ({gf / ms:6.1f} TF/s)")
            print(f"   pad mode {pad}: " + "   ".join(row), flush=True)
