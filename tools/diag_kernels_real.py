"""The backward kernels alone on REAL tensors: activations and conv-output gradients taken from the fp64 oracle at step 10
of the 20-step Adam trajectory (tools/diag_adam.py), rounded to fp32 and fed to tnb_conv3x3_wgrad / the dgrad convolution;
the reference is the fp64 result on the same fp32-rounded inputs, the yardstick torch's fp32 (cuDNN, TF32 off) on the same
inputs. Random test tensors do not have the structure that matters here: dz sums to zero per channel (BatchNorm backward)
while the activations have a large positive mean, so the weight gradient is the small remainder of a cancelling sum.
usage (GPU box): python tools/diag_kernels_real.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tracknet_oracle as O  # noqa: E402
from tests import gpu_util as G  # noqa: E402
from tests.test_gpu_tracknet import _disc_labels  # noqa: E402
from tracknetv3_b200 import _lib  # noqa: E402

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
gen = torch.Generator().manual_seed(32)
batches = [(torch.rand(2, 12, 96, 160, generator=gen).to(DEV), _disc_labels(2, 4, 96, 160, gen).to(DEV)) for _ in range(4)]

# fp64 oracle, 10 Adam steps, then one recorded forward / backward
sd = {k: (v.double() if v.is_floating_point() else v).to(DEV) for k, v in O.init_tracknet_state(31, 12, 4).items()}
pkeys = [k for k in sd if k.endswith(("conv.weight", "bn.weight", "bn.bias")) or k.startswith("predictor.")]
params = [sd[k].clone().requires_grad_(True) for k in pkeys]
opt = torch.optim.Adam(params, lr=1e-3)
rec = []
real_conv2d = F.conv2d


def recording_conv2d(x, w, *a, **k):
    z = real_conv2d(x, w, *a, **k)
    if w.shape[-1] == 3:
        z.retain_grad()
        rec.append((x, w, z))
    return z


for step in range(11):
    x, y = batches[step % 4]
    work = dict(sd)
    work.update(dict(zip(pkeys, params)))
    opt.zero_grad()
    if step == 10:
        O.F.conv2d = recording_conv2d
    loss = O.wbce_loss(O.tracknet_forward(work, x.double(), True), y.double())
    loss.backward()
    O.F.conv2d = real_conv2d
    if step < 10:
        opt.step()
    for k in sd:
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            sd[k] = work[k]


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


print("layer (3x3 conv call index): cin->cout @HxW | cancellation sum|dz*a| / |sum dz*a| (median over weights) | wgrad distance to fp64: "
      "ours bf16 pairs | ours fp16 pairs | torch fp32 || dgrad: ours bf16 | ours fp16 | torch fp32")
for idx in (16, 15, 14, 12, 9, 6, 3, 1):
    xin, w, z = rec[idx]
    a32, dz32, w32 = xin.detach().float(), z.grad.float(), w.detach().float()
    n, cin, h, wd = a32.shape
    cout = w32.shape[0]
    a64, dz64, w64 = a32.double(), dz32.double(), w32.double()
    # fp64 references on the fp32-rounded inputs
    wz = torch.zeros_like(w64, requires_grad=True)
    (real_conv2d(a64, wz, padding=1) * dz64).sum().backward()
    dw_ref = wz.grad
    wz = torch.zeros_like(w64, requires_grad=True)
    (real_conv2d(a64.abs(), wz, padding=1) * dz64.abs()).sum().backward()
    cancel = (wz.grad / dw_ref.abs().clamp_min(1e-300)).median().item()
    din_ref = F.conv_transpose2d(dz64, w64, padding=1)
    # torch fp32
    wz = torch.zeros_like(w32, requires_grad=True)
    (real_conv2d(a32, wz, padding=1) * dz32).sum().backward()
    dw_t = wz.grad
    din_t = F.conv_transpose2d(dz32, w32, padding=1)
    res = {}
    cpad = (cin + 31) // 32 * 32
    a_nhwc = G.nhwc(F.pad(a32, (0, 0, 0, 0, 0, cpad - cin)))
    dz_nhwc = G.nhwc(dz32)
    for fmt in (1, 0):
        xs = G.presplit(a_nhwc, fmt)
        src = _lib.Src(ptr=xs.data_ptr(), scale=None, shift=None, C=cpad, Hs=h, Ws=wd, mode=_lib.SRC_PRESPLIT)
        res[("w", fmt)] = G.wgrad3x3(G.make_view([src], n, h, wd), dz_nhwc, cout, cin, scratch=True, fmt=fmt)
        mul = G.pow2_mul(dz_nhwc) if fmt == 0 else 1.0
        ts = G.presplit(dz_nhwc, fmt, mul)
        mul_dev = torch.tensor([mul], device=DEV)
        dsrc = _lib.Src(ptr=ts.data_ptr(), scale=mul_dev.data_ptr() if fmt == 0 else None, shift=None, C=cout, Hs=h, Ws=wd,
                        mode=_lib.SRC_PRESPLIT)
        if cin % 32 == 0:
            out, _ = G.conv3x3(G.make_view([dsrc], n, h, wd), w32.contiguous(), cin, fmt=fmt, mode=1)
            res[("d", fmt)] = G.nchw(out)
    d = lambda f: f"{rel(res[('d', f)], din_ref):.2e}" if ("d", f) in res else "   -    "
    print(f"  {idx:2d}: {cin:3d}->{cout:3d} @{h}x{wd} | {cancel:9.1f} | {rel(res[('w', 1)], dw_ref):.2e} | {rel(res[('w', 0)], dw_ref):.2e} | "
          f"{rel(dw_t, dw_ref):.2e} || {d(1)} | {d(0)} | {rel(din_t, din_ref):.2e}")
