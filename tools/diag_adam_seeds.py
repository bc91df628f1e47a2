"""Is the gap between the CUDA path's 20-step Adam trajectory and the fp64 oracle's a bias or one draw of a chaotic process?
Five (init, data) seeds; for each the final-loss distance to the fp64 run of ours, of the fp32 oracle (twice: cuDNN picks
its algorithms per run), and of the fp32 oracle with cuDNN TF32. usage (GPU box): python tools/diag_adam_seeds.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracknetv3_b200 as T  # noqa: E402
from oracle import tracknet_oracle as O  # noqa: E402
from tests.test_gpu_tracknet import _disc_labels  # noqa: E402

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
STEPS = 20


def make_batches(seed):
    gen = torch.Generator().manual_seed(seed)
    return [(torch.rand(2, 12, 96, 160, generator=gen).to(DEV), _disc_labels(2, 4, 96, 160, gen).to(DEV)) for _ in range(4)]


def oracle_run(dtype, init_seed, batches):
    sd = {k: (v.to(dtype) if v.is_floating_point() else v).to(DEV) for k, v in O.init_tracknet_state(init_seed, 12, 4).items()}
    pkeys = [k for k in sd if k.endswith(("conv.weight", "bn.weight", "bn.bias")) or k.startswith("predictor.")]
    params = [sd[k].clone().requires_grad_(True) for k in pkeys]
    opt = torch.optim.Adam(params, lr=1e-3)
    losses = []
    for step in range(STEPS):
        x, y = batches[step % 4]
        work = dict(sd)
        work.update(dict(zip(pkeys, params)))
        opt.zero_grad()
        loss = O.wbce_loss(O.tracknet_forward(work, x.to(dtype), True), y.to(dtype))
        loss.backward()
        opt.step()
        for k in sd:
            if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                sd[k] = work[k]
        losses.append(loss.item())
    return losses


def our_run(init_seed, batches, precision="fp32x3"):
    m = T.TrackNet(12, 4, precision=precision).to(DEV).train()
    m.load_state_dict(O.init_tracknet_state(init_seed, 12, 4))
    opt = T.FusedAdam(list(m.parameters()), lr=1e-3)
    losses = []
    for step in range(STEPS):
        x, y = batches[step % 4]
        opt.zero_grad()
        loss = T.WBCELoss(m(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return losses


print("relative distance of the loss to the fp64 oracle's, mean over steps 10-19 / at step 19")
for init_seed, data_seed in [(31, 32), (41, 42), (51, 52), (61, 62), (71, 72)]:
    b = make_batches(data_seed)
    r64 = oracle_run(torch.float64, init_seed, b)
    runs = {"ours fp32x3": our_run(init_seed, b), "ours fp32x3_bwd1": our_run(init_seed, b, "fp32x3_bwd1"),
            "oracle fp32 (a)": oracle_run(torch.float32, init_seed, b), "oracle fp32 (b)": oracle_run(torch.float32, init_seed, b)}
    torch.backends.cudnn.allow_tf32 = True
    runs["oracle fp32 + TF32"] = oracle_run(torch.float32, init_seed, b)
    torch.backends.cudnn.allow_tf32 = False
    line = f"seeds ({init_seed}, {data_seed}) fp64 final {r64[-1]:.5f}: "
    for name, l in runs.items():
        late = sum(l[s] / r64[s] - 1 for s in range(10, 20)) / 10
        line += f"{name} {late:+.3f} / {l[-1] / r64[-1] - 1:+.3f}   "
    print(line, flush=True)
