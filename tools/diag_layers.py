"""Layer by layer: the CUDA path's intermediate tensors (tnb_tracknet_debug_layer) against the fp64 oracle's at step 10 of
the 20-step Adam trajectory - raw conv output z, BatchNorm scale / shift, dz (gradient w.r.t. z) and din (gradient w.r.t.
the layer input) - next to the fp32 oracle's distance from fp64 for the same tensors. Finds the first tensor where the
CUDA path is further from fp64 than plain fp32 is. usage (GPU box): python tools/diag_layers.py"""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracknetv3_b200 as T  # noqa: E402
from oracle import tracknet_oracle as O  # noqa: E402
from tests.test_gpu_tracknet import _disc_labels  # noqa: E402
from tracknetv3_b200 import _lib  # noqa: E402
from tracknetv3_b200.model import _cfg  # noqa: E402

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
gen = torch.Generator().manual_seed(32)
batches = [(torch.rand(2, 12, 96, 160, generator=gen).to(DEV), _disc_labels(2, 4, 96, 160, gen).to(DEV)) for _ in range(4)]
real_conv2d = F.conv2d


def record(state, dtype, x, y):
    """oracle forward / backward with every 3x3 conv's (input, z) recorded; returns [(input, z, dz, din)] per layer"""
    rec = []

    def conv2d(xx, w, *a, **k):
        z = real_conv2d(xx, w, *a, **k)
        if w.shape[-1] == 3:
            z.retain_grad()
            if xx.requires_grad:
                xx.retain_grad()
            rec.append((xx, z))
        return z

    sd = {k: (v.to(dtype) if v.is_floating_point() else v).clone() for k, v in state.items()}
    for k in sd:
        if k.endswith(("conv.weight", "bn.weight", "bn.bias")) or k.startswith("predictor."):
            sd[k].requires_grad_(True)
    O.F.conv2d = conv2d
    try:
        loss = O.wbce_loss(O.tracknet_forward(sd, x.to(dtype), True), y.to(dtype))
        loss.backward()
    finally:
        O.F.conv2d = real_conv2d
    return [(xx.detach(), z.detach(), z.grad, xx.grad) for xx, z in rec], sd


# the fp64 trajectory's state after 10 steps
sd = {k: (v.double() if v.is_floating_point() else v).to(DEV) for k, v in O.init_tracknet_state(31, 12, 4).items()}
pkeys = [k for k in sd if k.endswith(("conv.weight", "bn.weight", "bn.bias")) or k.startswith("predictor.")]
params = [sd[k].clone().requires_grad_(True) for k in pkeys]
opt = torch.optim.Adam(params, lr=1e-3)
for step in range(10):
    x, y = batches[step % 4]
    work = dict(sd)
    work.update(dict(zip(pkeys, params)))
    opt.zero_grad()
    O.wbce_loss(O.tracknet_forward(work, x.double(), True), y.double()).backward()
    opt.step()
    for k in sd:
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            sd[k] = work[k]
state = dict(sd)
state.update({k: p.detach() for k, p in zip(pkeys, params)})
state = {k: (v.float().double() if v.is_floating_point() else v) for k, v in state.items()}  # the fp32-representable state
x, y = batches[10 % 4]
r64, _ = record(state, torch.float64, x, y)
r32, _ = record(state, torch.float32, x, y)

m = T.TrackNet(12, 4).to(DEV).train()
m.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in state.items()})
T.WBCELoss(m(x), y).backward()
torch.cuda.synchronize()
lib = _lib.load()
n, _, h, w = x.shape
cfg = _cfg(n, h, w, 12, 4, True, m.precision, m._variant)
ws = m._ws_saved


def tensor_at(ptr, shape, dtype=torch.float32):
    numel = 1
    for s in shape:
        numel *= s
    nbytes = numel * torch.empty((), dtype=dtype).element_size()
    off = ptr - ws.data_ptr()
    assert 0 <= off and off + nbytes <= ws.numel(), (off, nbytes, ws.numel())
    return ws[off:off + nbytes].view(dtype).reshape(shape)


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300)).item()


print("distance to the fp64 oracle, max-norm relative to max |fp64 tensor|: ours / oracle fp32")
print(f"{'layer':>5s} {'shape':>16s} | {'z':>19s} | {'dz':>19s} | {'din':>19s}")
for l in range(17):
    ptrs = (C.c_void_p * 8)()
    dims = (C.c_int * 5)()
    _lib.check(lib.tnb_tracknet_debug_layer(C.byref(cfg), ws.data_ptr(), l, ptrs, dims))
    H, W, cin, cout, fmt = list(dims)
    z = tensor_at(ptrs[0], (n, H, W, cout)).permute(0, 3, 1, 2)
    raw = tensor_at(ptrs[5], (n, H, W, 2, cout), torch.float16 if fmt == 0 else torch.bfloat16).double()
    mul = tensor_at(ptrs[7], (1,)).item() if ptrs[7] else 1.0
    dz = ((raw[..., 0, :] + raw[..., 1, :]) / mul).permute(0, 3, 1, 2)
    xin64, z64, dz64, din64 = r64[l]
    xin32, z32, dz32, din32 = r32[l]
    line = f"{l:5d} {cin:4d}->{cout:3d}@{H}x{W:<4d} | {rel(z, z64):.2e} / {rel(z32, z64):.2e} | {rel(dz, dz64):.2e} / {rel(dz32, dz64):.2e} | "
    if ptrs[6] and din64 is not None:
        din = tensor_at(ptrs[6], (n, H, W, cin)).permute(0, 3, 1, 2)
        # the oracle's input gradient is w.r.t. the materialised view (after pool / upsample / concat): same shape
        line += f"{rel(din, din64):.2e} / {rel(din32, din64):.2e}"
    else:
        line += "        -"
    print(line)
