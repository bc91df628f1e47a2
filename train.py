"""Hot-path subset of the reference's train.py: `mixup` (:19-40), `get_random_mask` (:42-57), the TrackNet step loop
(:59-121) and the InpaintNet step loop (:123-177).

Dataset / TensorBoard / checkpoint plumbing of the reference's __main__ is out of scope;
`train_tracknet` accepts any iterable yielding the reference's (i, x, y, c, _) tuples and runs the step of
train.py:85-96 on the B200 kernels. `python train.py --synthetic` runs a few steps on random frames.
"""
import argparse

import numpy as np
import torch

from tracknetv3_b200 import _lib
from utils.general import get_model
from utils.metric import WBCELoss


def mixup(x, y, alpha=0.5):
    """Returns mixed inputs, pairs of targets (same RNG draws, in the same order, as reference :32-36;
    the blend itself is one fused kernel per tensor)."""
    lib = _lib.load()
    batch_size = x.size()[0]
    lamb = np.random.beta(alpha, alpha, size=batch_size)
    lamb = np.maximum(lamb, 1 - lamb)
    lamb = torch.from_numpy(lamb).float().to(x.device)
    index = torch.randperm(batch_size).to(x.device)
    outs = []
    for t in (x, y):
        _lib.require_cuda(t)
        t = t.contiguous().float()
        o = torch.empty_like(t)
        _lib.check(lib.tnb_mixup(t.data_ptr(), lamb.data_ptr(), index.data_ptr(), o.data_ptr(), batch_size,
                                 t.numel() // batch_size, _lib.stream_ptr()))
        outs.append(o)
    return outs[0], outs[1]


def train_tracknet(model, optimizer, data_loader, param_dict):
    """ Train TrackNet model for one epoch (step semantics of reference :84-96). Returns the mean loss. """
    model.train()
    epoch_loss = []
    for step, (_, x, y, c, _) in enumerate(data_loader):
        optimizer.zero_grad()
        x, y = x.float().cuda(), y.float().cuda()
        if param_dict['alpha'] > 0:
            x, y = mixup(x, y, param_dict['alpha'])
        y_pred = model(x)
        loss = WBCELoss(y_pred, y)
        epoch_loss.append(loss.item())
        loss.backward()
        optimizer.step()
    return float(np.mean(epoch_loss))


def get_random_mask(mask_size, mask_ratio):
    """ Generate random mask by binomial distribution (1 = masked): same numpy draw as reference :54, (N, L, 1). """
    mask = np.random.binomial(1, mask_ratio, size=mask_size)
    mask = torch.from_numpy(mask).float().cuda().unsqueeze(-1)
    return mask


def train_inpaintnet(model, optimizer, data_loader, param_dict):
    """ Train InpaintNet model for one epoch (step semantics of reference :147-163): random mask AND visibility,
        masked coordinates zeroed, MSE on the masked entries, clip_grad_norm_(1), optimizer step. Returns mean loss.
        The model's forward and backward are one kernel each; the few (N, L, 2)-sized loss ops stay torch ops. """
    model.train()
    epoch_loss = []
    for step, (_, coor_pred, coor_gt, _, vis_gt, _) in enumerate(data_loader):
        optimizer.zero_grad()
        coor_pred, coor_gt, vis_gt = coor_pred.float().cuda(), coor_gt.float().cuda(), vis_gt.float().cuda()
        mask = get_random_mask(mask_size=coor_gt.shape[:2], mask_ratio=param_dict['mask_ratio']).cuda()  # (N, L, 1)
        inpaint_mask = torch.logical_and(vis_gt, mask).int()  # visible and masked area
        coor_pred = coor_pred * (1 - inpaint_mask)  # masked area is set to 0
        refine_coor = model(coor_pred, inpaint_mask)
        loss = torch.nn.MSELoss()(refine_coor * inpaint_mask, coor_gt * inpaint_mask)
        epoch_loss.append(loss.item())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        optimizer.step()
    return float(np.mean(epoch_loss))


def _synthetic_loader(steps, batch_size, seq_len, bg_mode, h=288, w=512):
    in_dim = get_model('TrackNet', seq_len, bg_mode).in_dim
    for i in range(steps):
        x = torch.rand(batch_size, in_dim, h, w)
        y = (torch.rand(batch_size, seq_len, h, w) > 0.999).float()
        yield i, x, y, torch.zeros(batch_size, seq_len, 2), None


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--model_name', type=str, default='TrackNet', choices=['TrackNet'])
    parser.add_argument('--seq_len', type=int, default=8)
    parser.add_argument('--batch_size', type=int, default=10)
    parser.add_argument('--learning_rate', type=float, default=0.001)
    parser.add_argument('--bg_mode', type=str, default='concat', choices=['', 'subtract', 'subtract_concat', 'concat'])
    parser.add_argument('--alpha', type=float, default=0.5)
    parser.add_argument('--seed', type=int, default=13)
    parser.add_argument('--steps', type=int, default=5)
    parser.add_argument('--synthetic', action='store_true', default=True)
    args = parser.parse_args()
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    from tracknetv3_b200 import FusedAdam
    model = get_model(args.model_name, args.seq_len, args.bg_mode).cuda()
    optimizer = FusedAdam(model.parameters(), lr=args.learning_rate)
    loss = train_tracknet(model, optimizer, _synthetic_loader(args.steps, args.batch_size, args.seq_len, args.bg_mode),
                          vars(args))
    print(f'mean loss over {args.steps} synthetic steps: {loss:.6f}')
