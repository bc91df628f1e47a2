"""Minimal driver for ncu: W warm-up + K train steps (fwd + WBCE + bwd) of the bench workload, nothing else.
usage: python tools/profile_step.py [warmup] [steps] [precision] [batch] [variant]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracknetv3_b200 as T  # noqa: E402
import bench  # noqa: E402

warmup = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32x3"
batch = int(sys.argv[4]) if len(sys.argv) > 4 else bench.BATCH
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
torch.manual_seed(13)
model = T.TrackNet(bench.IN_DIM, bench.OUT_DIM, precision=precision).cuda().train()
model._variant = variant
x, y = bench.synthetic_batch(batch, 13)
x, y = x.cuda(), y.cuda()
for i in range(warmup + steps):
    for p in model.parameters():
        p.grad = None
    loss = T.WBCELoss(model(x), y)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
