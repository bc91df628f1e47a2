"""Measurement only (never on the product path): the reference architecture on stock torch-CUDA (cuDNN/ATen) on the same
box, same workload and timed region as bench.py `value` (fwd + WBCE + bwd, batch resident in HBM) - the "reference's own
torch-CUDA FPS" of BASELINE.json's north_star. bench.py prints the same three numbers in its JSON line
(`torch_cuda_baseline`); this script is the stand-alone form.
usage: python tools/torch_cuda_baseline.py [steps] [warmup]  -> one JSON line per variant."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench as B  # synthetic_batch only
from tools import ref_arch

if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    warmup = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    x, y = B.synthetic_batch(B.BATCH, 13)
    x, y = x.cuda(), y.cuda()
    for name, (tf32, benchmark, channels_last, deterministic) in ref_arch.VARIANTS.items():
        ms = ref_arch.time_variant(name, x, y, B.IN_DIM, B.OUT_DIM, steps, warmup)
        print(json.dumps({"impl": "torch-cuda", "variant": name, "ms_per_step": ms,
                          "frames_per_s": B.BATCH * B.SEQ_LEN / ms * 1e3, "allow_tf32": tf32, "cudnn_benchmark": benchmark,
                          "channels_last": channels_last, "cudnn_deterministic": deterministic,
                          "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}), flush=True)
