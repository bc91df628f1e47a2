"""Diagnostic: the smoke() configuration against the fp32 AND fp64 oracle, per parameter (is a deviation ours or the
conditioning of the tiny problem?). usage: python tools/diag_smoke.py [variant]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tracknetv3_b200 as T  # noqa: E402
from oracle import tracknet_oracle as O  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (64, 96)
torch.manual_seed(1)
model = T.TrackNet(12, 4).cuda().train()
model._variant = variant
sd = O.init_tracknet_state(1, 12, 4)
gen = torch.Generator().manual_seed(2)
x = torch.rand(2, 12, h, w, generator=gen)
y = (torch.rand(2, 4, h, w, generator=gen) > 0.98).float()
y_pred = model(x.cuda())
T.WBCELoss(y_pred, y.cuda()).backward()
r_pred, r_loss, r32 = O.tracknet_loss_and_grads(sd, x, y, True)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in O.init_tracknet_state(1, 12, 4).items()}
_, _, r64 = O.tracknet_loss_and_grads(sd64, x.double(), y.double(), True)
rel = lambda a, b: ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()
print(f"variant {variant} {h}x{w} env PLAN={os.environ.get('TNB_CONV_PLAN')} CA={os.environ.get('TNB_CPASYNC_CA')} "
      f"GRAPHS={os.environ.get('TNB_GRAPHS')}: heatmap err {(y_pred.detach().cpu() - r_pred).abs().max().item():.2e}")
worst = 0
for k, p in model.named_parameters():
    a, b, c = rel(p.grad, r32[k]), rel(p.grad, r64[k]), rel(r32[k], r64[k])
    flag = " <-- exceeds 3x yardstick + 2e-2" if b > 3 * c + 2e-2 else ""
    if a > 2e-2 or flag:
        print(f"   {k:38s} vs fp32 {a:.3e}  vs fp64 {b:.3e}  fp32-vs-fp64 {c:.3e}{flag}")
    worst = max(worst, b)
print("   worst vs fp64", worst)
