"""Where does the 20-step Adam trajectory of the CUDA path leave the oracle's? (tests/test_gpu_tracknet.py::
test_twenty_adam_steps_track_the_oracle ends ~8 % above the fp64 curve.) Separates the optimizer from the gradients:
  1. our model + FusedAdam | our model + torch.optim.Adam | oracle fp32 + torch Adam | oracle fp32 + FusedAdam | oracle fp64
  2. per-parameter distance to the fp64 gradient at step 0 and along the fp64 trajectory (steps 5, 10): ours vs oracle fp32
  3. oracle fp32 with additive gradient noise eps * max|g| (what level of noise reproduces our curve?)
usage (GPU box): python tools/diag_adam.py [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracknetv3_b200 as T  # noqa: E402
from oracle import tracknet_oracle as O  # noqa: E402
from tests.test_gpu_tracknet import _disc_labels  # noqa: E402

DEV = "cuda"
PREC = sys.argv[1] if len(sys.argv) > 1 else "fp32x3"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
gen = torch.Generator().manual_seed(32)
batches = [(torch.rand(2, 12, 96, 160, generator=gen).to(DEV), _disc_labels(2, 4, 96, 160, gen).to(DEV)) for _ in range(4)]
STEPS = 20


def pkeys_of(sd):
    return [k for k in sd if k.endswith(("conv.weight", "bn.weight", "bn.bias")) or k.startswith("predictor.")]


def oracle_run(dtype, make_opt, noise=0.0, snapshots=None, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    sd = {k: (v.to(dtype) if v.is_floating_point() else v).to(DEV) for k, v in O.init_tracknet_state(31, 12, 4).items()}
    pkeys = pkeys_of(sd)
    params = [sd[k].clone().requires_grad_(True) for k in pkeys]
    opt = make_opt(params)
    losses = []
    for step in range(STEPS):
        x, y = batches[step % 4]
        work = dict(sd)
        work.update(dict(zip(pkeys, params)))
        if snapshots is not None and step in snapshots:
            snapshots[step] = {k: v.detach().clone() for k, v in work.items()}
        opt.zero_grad()
        loss = O.wbce_loss(O.tracknet_forward(work, x.to(dtype), True), y.to(dtype))
        loss.backward()
        if noise > 0:
            for p in params:
                p.grad += noise * p.grad.abs().max() * torch.randn(p.grad.shape, generator=g, device=DEV, dtype=dtype)
        opt.step()
        for k in sd:
            if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                sd[k] = work[k]
        losses.append(loss.item())
    return losses


def our_run(make_opt):
    torch.manual_seed(31)
    m = T.TrackNet(12, 4, precision=PREC).to(DEV).train()
    m.load_state_dict(O.init_tracknet_state(31, 12, 4))
    opt = make_opt(list(m.parameters()))
    losses = []
    for step in range(STEPS):
        x, y = batches[step % 4]
        opt.zero_grad()
        loss = T.WBCELoss(m(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return losses


def grads_at(state, step, dtype, ours):
    x, y = batches[step % 4]
    if ours:
        m = T.TrackNet(12, 4, precision=PREC).to(DEV).train()
        m.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in state.items()})
        T.WBCELoss(m(x), y).backward()
        return {k: p.grad.double() for k, p in m.named_parameters()}
    sd = {k: (v.to(dtype) if v.is_floating_point() else v).clone() for k, v in state.items()}
    pk = pkeys_of(sd)
    for k in pk:
        sd[k].requires_grad_(True)
    O.wbce_loss(O.tracknet_forward(sd, x.to(dtype), True), y.to(dtype)).backward()
    return {k: sd[k].grad.double() for k in pk}


adam_t = lambda ps: torch.optim.Adam(ps, lr=1e-3)
adam_f = lambda ps: T.FusedAdam(ps, lr=1e-3)
snaps = {0: None, 5: None, 10: None}
r64 = oracle_run(torch.float64, adam_t, snapshots=snaps)
runs = {
    "ours + FusedAdam": our_run(adam_f),
    "ours + torch Adam": our_run(adam_t),
    "oracle fp32 + torch Adam": oracle_run(torch.float32, adam_t),
    "oracle fp32 + FusedAdam": oracle_run(torch.float32, adam_f),
}
# the reference's own GPU numerics: torch defaults allow TF32 in cuDNN convolutions (train.py never turns it off)
torch.backends.cudnn.allow_tf32 = True
runs["oracle fp32, cuDNN TF32 allowed (reference default)"] = oracle_run(torch.float32, adam_t)
torch.backends.cudnn.allow_tf32 = False
SHORT = os.environ.get("DIAG_ADAM_SHORT") == "1"  # trajectories only
for eps in (() if SHORT else (1e-5, 1e-4, 3e-4, 1e-3)):
    for seed in (0, 1):
        runs[f"oracle fp32 + noise {eps:g} seed {seed}"] = oracle_run(torch.float32, adam_t, noise=eps, seed=seed)
print(f"precision {PREC}; fp64 curve: " + " ".join(f"{v:.5f}" for v in r64[::3]) + f" ... {r64[-1]:.6f}")
for name, l in runs.items():
    print(f"{name:52s} rel. distance to fp64 at steps 4 / 9 / 14 / 19: " + "  ".join(f"{l[s] / r64[s] - 1:+.2e}" for s in (4, 9, 14, 19)))

if SHORT:
    sys.exit(0)
with torch.no_grad():  # forward distance at the initial state (train-mode BatchNorm)
    st0 = snaps[0]
    x0 = batches[0][0]
    h64 = O.tracknet_forward({k: v.clone() for k, v in st0.items()}, x0.double(), True)
    h32 = O.tracknet_forward({k: (v.float() if v.is_floating_point() else v).clone() for k, v in st0.items()}, x0, True)
    m0 = T.TrackNet(12, 4, precision=PREC).to(DEV).train()
    m0.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in st0.items()})
    ho = m0(x0).double()
    print(f"\nheatmap max-abs distance to the fp64 oracle: ours {(ho - h64).abs().max().item():.3e} | oracle fp32 {(h32.double() - h64).abs().max().item():.3e}")

for step, state in snaps.items():
    g64 = grads_at(state, step, torch.float64, False)
    g32 = grads_at(state, step, torch.float32, False)
    go = grads_at(state, step, None, True)
    torch.backends.cudnn.allow_tf32 = True
    gtf = grads_at(state, step, torch.float32, False)
    torch.backends.cudnn.allow_tf32 = False
    print(f"\nper-parameter distance to the fp64 gradient at the fp64 trajectory's step {step} (max-norm, relative to max |g64|): ours | oracle fp32 | oracle fp32 with cuDNN TF32 | ratio of L2 norms ours/g64")
    # what Adam sees: its first steps move every element by ~lr * sign(g) - the share of elements whose sign disagrees
    # with the fp64 gradient, and the median elementwise relative error, over all 53 tensors
    tot = sum(v.numel() for v in g64.values())
    for name, gg in (("ours", go), ("oracle fp32", g32), ("oracle fp32 + cuDNN TF32", gtf)):
        flips = sum(((gg[k] * g64[k]) < 0).sum().item() for k in g64)
        relerr = torch.cat([((gg[k] - g64[k]).abs() / g64[k].abs().clamp_min(1e-300)).flatten() for k in g64])
        small = sum((gg[k].abs() < 1e-8).sum().item() for k in g64)
        print(f"  {name:26s} sign(g) != sign(g64): {flips / tot:.3e} of {tot} elements; elementwise relative error: median "
              f"{relerr.median().item():.2e}, 90th percentile {relerr.kthvalue(int(0.9 * tot)).values.item():.2e}; |g| < Adam's eps: {small / tot:.3e}")
    for k in g64:
        d = g64[k].abs().max().item()
        eo = (go[k] - g64[k]).abs().max().item() / d
        e32 = (g32[k] - g64[k]).abs().max().item() / d
        flag = "  <<<" if eo > 3 * e32 + 1e-4 else ""
        etf = (gtf[k] - g64[k]).abs().max().item() / d
        print(f"  {k:34s} {eo:.2e} | {e32:.2e} | {etf:.2e} | {go[k].norm().item() / g64[k].norm().item():.5f}{flag}")
