"""Plain PyTorch-CPU fp32 restatement of the reference's floating-point hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py). Pinned against the real reference (imported from
/root/reference in the build container) by oracle/gen_golden.py -> tests/golden/*.npz, and re-checked
against those fixtures by tests/test_oracle.py on every run. Each function cites the reference lines
it restates. Everything is written functionally on a {state_dict key: tensor} mapping so that it
shares no code with tracknetv3_b200/.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

HEIGHT, WIDTH, SIGMA = 288, 512, 2.5          # reference utils/general.py:15-17
DELTA_T = 1 / math.sqrt(HEIGHT ** 2 + WIDTH ** 2)  # utils/general.py:18
COOR_TH = DELTA_T * 50                        # utils/general.py:19

# (block prefix, number of Conv2DBlocks, in, out) - reference model.py:47-53
TRACKNET_BLOCKS = [("down_block_1", 2, None, 64), ("down_block_2", 2, 64, 128), ("down_block_3", 3, 128, 256),
                   ("bottleneck", 3, 256, 512), ("up_block_1", 3, 768, 256), ("up_block_2", 2, 384, 128),
                   ("up_block_3", 2, 192, 64)]


def tracknet_state_keys(in_dim, out_dim):
    """(key, shape) list in state_dict order (reference model.py:4-55)."""
    keys = []
    for prefix, nconv, cin, cout in TRACKNET_BLOCKS:
        c = in_dim if cin is None else cin
        for i in range(1, nconv + 1):
            p = f"{prefix}.conv_{i}"
            keys += [(f"{p}.conv.weight", (cout, c, 3, 3)), (f"{p}.bn.weight", (cout,)), (f"{p}.bn.bias", (cout,)),
                     (f"{p}.bn.running_mean", (cout,)), (f"{p}.bn.running_var", (cout,)),
                     (f"{p}.bn.num_batches_tracked", ())]
            c = cout
    keys += [("predictor.weight", (out_dim, 64, 1, 1)), ("predictor.bias", (out_dim,))]
    return keys


def init_tracknet_state(seed, in_dim, out_dim):
    """torch-default initialisation in module-construction order (what `TrackNet(in_dim, out_dim)` does
    under torch.manual_seed(seed)): Conv2d kaiming_uniform(a=sqrt(5)), BatchNorm ones/zeros."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in tracknet_state_keys(in_dim, out_dim):
        if key.endswith("conv.weight") or key == "predictor.weight":
            fan_in = shape[1] * shape[2] * shape[3]
            bound = math.sqrt(6.0 / ((1 + 5.0) * fan_in))  # kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), ..)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif key == "predictor.bias":
            bound = 1 / math.sqrt(64)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif key.endswith("bn.weight") or key.endswith("running_var"):
            sd[key] = torch.ones(shape)
        elif key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.int64)
        else:
            sd[key] = torch.zeros(shape)
    return sd


def conv_block(sd, prefix, x, training, eps=1e-5, momentum=0.1):
    """Conv2DBlock: conv3x3(pad 1, no bias) -> BatchNorm2d -> ReLU (reference model.py:12-16).
    Train mode normalises with the biased batch variance and tracks the unbiased one (torch semantics)."""
    z = F.conv2d(x, sd[f"{prefix}.conv.weight"], None, 1, 1)
    g, b = sd[f"{prefix}.bn.weight"], sd[f"{prefix}.bn.bias"]
    if training:
        mean = z.mean(dim=(0, 2, 3))
        var = z.var(dim=(0, 2, 3), unbiased=False)
        with torch.no_grad():
            n = z.numel() / z.shape[1]
            rm, rv = sd[f"{prefix}.bn.running_mean"], sd[f"{prefix}.bn.running_var"]
            rm.mul_(1 - momentum).add_(momentum * mean.detach())
            rv.mul_(1 - momentum).add_(momentum * var.detach() * n / max(n - 1, 1))
            sd[f"{prefix}.bn.num_batches_tracked"] += 1
    else:
        mean, var = sd[f"{prefix}.bn.running_mean"], sd[f"{prefix}.bn.running_var"]
    zhat = (z - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
    return F.relu(zhat * g[None, :, None, None] + b[None, :, None, None])


def _stack(sd, prefix, n, x, training):
    for i in range(1, n + 1):
        x = conv_block(sd, f"{prefix}.conv_{i}", x, training)
    return x


def tracknet_forward(sd, x, training):
    """TrackNet.forward (reference model.py:57-73): encoder with 3 max-pools, decoder with nearest x2
    upsampling and [upsampled, skip] channel concat, 1x1 predictor with bias, sigmoid."""
    x1 = _stack(sd, "down_block_1", 2, x, training)
    x2 = _stack(sd, "down_block_2", 2, F.max_pool2d(x1, 2, 2), training)
    x3 = _stack(sd, "down_block_3", 3, F.max_pool2d(x2, 2, 2), training)
    t = _stack(sd, "bottleneck", 3, F.max_pool2d(x3, 2, 2), training)
    t = _stack(sd, "up_block_1", 3, torch.cat([F.interpolate(t, scale_factor=2, mode="nearest"), x3], 1), training)
    t = _stack(sd, "up_block_2", 2, torch.cat([F.interpolate(t, scale_factor=2, mode="nearest"), x2], 1), training)
    t = _stack(sd, "up_block_3", 2, torch.cat([F.interpolate(t, scale_factor=2, mode="nearest"), x1], 1), training)
    return torch.sigmoid(F.conv2d(t, sd["predictor.weight"], sd["predictor.bias"]))


def wbce_loss(y_pred, y, reduce=True):
    """WBCELoss (reference utils/metric.py:15-20)."""
    loss = -((1 - y_pred) ** 2 * y * torch.log(torch.clamp(y_pred, 1e-7, 1))
             + y_pred ** 2 * (1 - y) * torch.log(torch.clamp(1 - y_pred, 1e-7, 1)))
    return loss.mean() if reduce else loss.flatten(1).mean(1)


def tracknet_loss_and_grads(sd, x, y, training=True):
    """One fwd + WBCE + backward of the train step (reference train.py:92-95). Returns y_pred, loss and
    {key: grad} for the 53 parameters."""
    pkeys = [k for k in sd if k.endswith(("conv.weight", "bn.weight", "bn.bias")) or k.startswith("predictor.")]
    params = {k: sd[k].clone().requires_grad_(True) for k in pkeys}
    work = dict(sd)
    work.update(params)
    y_pred = tracknet_forward(work, x, training)
    loss = wbce_loss(y_pred, y)
    grads = torch.autograd.grad(loss, [params[k] for k in pkeys])
    for k in sd:  # running statistics advanced by the forward
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            sd[k] = work[k]
    return y_pred.detach(), loss.detach(), dict(zip(pkeys, grads))


def mixup(x, y, lamb, index):
    """Sample mixup given the already-drawn lambdas / permutation (reference train.py:34-38; the RNG draw
    np.random.beta + torch.randperm stays on the host in both implementations)."""
    lamb = torch.as_tensor(np.maximum(lamb, 1 - lamb)[:, None, None, None]).float()
    return x * lamb + x[index] * (1 - lamb), y * lamb + y[index] * (1 - lamb)


def label_disc(cx, cy, h=HEIGHT, w=WIDTH, sigma=SIGMA):
    """Binary disc label of radius sigma around (cx, cy) on a 1-based grid; all zero when cx == cy == 0
    (reference dataset.py:401-410). Returned in [0, 1] (the reference stores 0/255 then divides)."""
    if cx == 0 and cy == 0:
        return np.zeros((h, w), dtype=np.float32)
    xs, ys = np.meshgrid(np.linspace(1, w, w), np.linspace(1, h, h))
    d = (ys - (cy + 1)) ** 2 + (xs - (cx + 1)) ** 2
    return (d <= sigma ** 2).astype(np.float32)


def get_ensemble_weight(seq_len, eval_mode):
    """Temporal-ensemble weights (reference test.py:39-48)."""
    if eval_mode == "average":
        return torch.ones(seq_len) / seq_len
    if eval_mode == "weight":
        wgt = torch.ones(seq_len)
        for i in range(math.ceil(seq_len / 2)):
            wgt[i] = i + 1
            wgt[seq_len - i - 1] = i + 1
        return wgt / wgt.sum()
    raise ValueError("Invalid mode")


# ---------------------------------------------------------------------------------------------
# InpaintNet (reference model.py:76-129)
# ---------------------------------------------------------------------------------------------
INPAINT_CONVS = [("down_1.conv", 3, 32), ("down_2.conv", 32, 64), ("down_3.conv", 64, 128),
                 ("buttleneck.conv_1.conv", 128, 256), ("buttleneck.conv_2.conv", 256, 256),
                 ("up_1.conv", 384, 128), ("up_2.conv", 192, 64), ("up_3.conv", 96, 32), ("predictor", 32, 2)]


def init_inpaintnet_state(seed):
    """torch-default Conv1d init in module-construction order."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, cin, cout in INPAINT_CONVS:
        bound = 1 / math.sqrt(cin * 3)
        sd[f"{name}.weight"] = (torch.rand((cout, cin, 3), generator=g) * 2 - 1) * bound
        sd[f"{name}.bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bound
    return sd


def inpaintnet_forward(sd, x, m):
    """InpaintNet.forward (reference model.py:113-129): Conv1d(k=3, 'same') + LeakyReLU(0.01) U-Net over
    cat(coords, mask) with three skip concats, sigmoid head. x (N, L, 2), m (N, L, 1) -> (N, L, 2)."""
    def blk(name, t):
        return F.leaky_relu(F.conv1d(t, sd[f"{name}.weight"], sd[f"{name}.bias"], padding=1), 0.01)
    t = torch.cat([x, m], dim=2).permute(0, 2, 1)
    x1 = blk("down_1.conv", t)
    x2 = blk("down_2.conv", x1)
    x3 = blk("down_3.conv", x2)
    t = blk("buttleneck.conv_2.conv", blk("buttleneck.conv_1.conv", x3))
    t = blk("up_1.conv", torch.cat([t, x3], 1))
    t = blk("up_2.conv", torch.cat([t, x2], 1))
    t = blk("up_3.conv", torch.cat([t, x1], 1))
    t = torch.sigmoid(F.conv1d(t, sd["predictor.weight"], sd["predictor.bias"], padding=1))
    return t.permute(0, 2, 1)


def inpaint_blend(coor_pred, coor_inpaint, mask):
    """predict.py:257-261: keep the inpainted value only where the mask is set, then zero rows whose two
    coordinates are both below COOR_TH."""
    out = coor_inpaint * mask + coor_pred * (1 - mask)
    th = (out[:, :, 0] < COOR_TH) & (out[:, :, 1] < COOR_TH)
    out = out.clone()
    out[th] = 0.0
    return out


def get_random_mask(mask_size, mask_ratio):
    """train.py:42-57: Bernoulli(mask_ratio) mask of shape (N, L, 1), 1 = masked; numpy's global RNG, as there."""
    return torch.from_numpy(np.random.binomial(1, mask_ratio, size=mask_size)).float().unsqueeze(-1)


def inpaintnet_loss_and_grads(sd, coor_pred, coor_gt, vis_gt, mask):
    """The arithmetic of one InpaintNet train step up to the gradients (reference train.py:151-163):
    inpaint_mask = vis_gt AND mask; masked input coordinates are zeroed; MSE between the masked prediction and
    the masked ground truth (mean over ALL N*L*2 entries, nn.MSELoss default). Returns refine, loss, {key: grad}."""
    inpaint_mask = torch.logical_and(vis_gt, mask).int()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    refine = inpaintnet_forward(params, coor_pred * (1 - inpaint_mask), inpaint_mask.to(coor_pred.dtype))
    loss = F.mse_loss(refine * inpaint_mask, coor_gt * inpaint_mask)
    grads = torch.autograd.grad(loss, list(params.values()))
    return refine.detach(), loss.detach(), dict(zip(params.keys(), grads))


def clip_grad_norm(grads, max_norm=1.0):
    """nn.utils.clip_grad_norm_ (train.py:161): global L2 norm, scale by max_norm / (norm + 1e-6) clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, {k: g * coef for k, g in grads.items()}


class TemporalEnsemble:
    """Streaming temporal ensemble of overlapping predictions (reference predict.py:163-209 for heatmaps, :245-301
    for InpaintNet coordinates). Sample s predicts frames s..s+L-1; frame t is the combination of the L samples that
    contain it: uniform mean while fewer than L samples have been seen (:181-183), `weight` afterwards (:184-186),
    and plain means over the remaining samples for the last L-1 frames once `num_sample` is reached (:192-201).
    push(pred (B, L, ...)) returns the ensembled frames this batch completes, (B [+ L-1 at the end], ...)."""

    def __init__(self, seq_len, eval_mode, num_sample):
        self.L, self.num_sample, self.count = seq_len, num_sample, 0
        self.weight = get_ensemble_weight(seq_len, eval_mode)
        self.buffer = None

    def push(self, pred):
        L = self.L
        if self.buffer is None:
            self.buffer = torch.zeros((L - 1,) + tuple(pred.shape[1:]), dtype=pred.dtype)
        buf = torch.cat((self.buffer, pred), dim=0)
        batch_i, frame_i = torch.arange(L), torch.arange(L - 1, -1, -1)
        w = self.weight.reshape((L,) + (1,) * (pred.dim() - 2))
        out = []
        for b in range(pred.shape[0]):
            if self.count < L - 1:
                out.append(buf[batch_i + b, frame_i].sum(0) / (self.count + 1))
            else:
                out.append((buf[batch_i + b, frame_i] * w).sum(0))
            self.count += 1
            if self.count == self.num_sample:
                buf = torch.cat((buf, torch.zeros_like(buf[:L - 1])), dim=0)
                for f in range(1, L):
                    out.append(buf[batch_i + b + f, frame_i].sum(0) / (L - f))
        self.buffer = buf[-(L - 1):] if L > 1 else buf[:0]
        return torch.stack(out)
