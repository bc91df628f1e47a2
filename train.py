"""Hot-path subset of the reference's train.py: `mixup` (:19-40), `get_random_mask` (:42-57), the TrackNet step loop
(:59-121) and the InpaintNet step loop (:123-177).

The reference's datasets and TensorBoard writer are out of scope; `train_tracknet` accepts any iterable
yielding the reference's (i, x, y, c, _) tuples and runs the step of train.py:85-96 on the B200 kernels. The optimizer /
scheduler choices, the checkpoint layout and the resume logic of the reference's __main__ (:236-305) are here as
`make_optimizer`, `make_scheduler`, `checkpoint_dict`, `resume_from` and `fit`; `python train.py` takes the reference's
command line and trains on synthetic batches.

Data parallel (BASELINE configs[2]; the reference itself is single-GPU): launched as
`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 train.py ...` every process drives one GPU
with `--batch_size` samples per step, the replicas start from rank 0's weights (`broadcast_module`), every rank draws its
own batches (seed + rank), the gradients are averaged with ONE in-place NCCL allreduce of the flat bucket per step
(`GradBucket`, torch-DDP semantics: BatchNorm statistics stay per replica), validation and the checkpoints are rank 0's.
"""
import argparse
import os

import numpy as np
import torch

from tracknetv3_b200 import _lib
from tracknetv3_b200.data import ScalarReader
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


def train_tracknet(model, optimizer, data_loader, param_dict, bucket=None):
    """ Train TrackNet model for one epoch (step semantics of reference :84-96). Returns the mean loss.
        bucket (tracknetv3_b200.parallel.GradBucket, optional): data parallel - average the gradients over the ranks
        between backward and the optimizer step. """
    model.train()
    epoch_loss = []
    reader = None  # created with the first CUDA loss (the host-logic tests drive this loop with CPU stand-ins)
    for step, (_, x, y, c, _) in enumerate(data_loader):
        optimizer.zero_grad()
        x, y = x.float().cuda(), y.float().cuda()
        if param_dict['alpha'] > 0:
            x, y = mixup(x, y, param_dict['alpha'])
        y_pred = model(x)
        loss = WBCELoss(y_pred, y)
        if loss.is_cuda:                       # reference: epoch_loss.append(loss.item()) - here without draining the stream
            reader = reader or ScalarReader()
            reader.read(loss)
        else:
            epoch_loss.append(loss.item())
        loss.backward()
        if bucket is not None:
            bucket.allreduce()
        optimizer.step()
        if loss.is_cuda:
            epoch_loss.append(reader.value())
    return float(np.mean(epoch_loss))


def get_random_mask(mask_size, mask_ratio):
    """ Generate random mask by binomial distribution (1 = masked): same numpy draw as reference :54, (N, L, 1). """
    mask = np.random.binomial(1, mask_ratio, size=mask_size)
    mask = torch.from_numpy(mask).float().cuda().unsqueeze(-1)
    return mask


def train_inpaintnet(model, optimizer, data_loader, param_dict, bucket=None):
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
        if bucket is not None:
            bucket.allreduce()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        optimizer.step()
    return float(np.mean(epoch_loss))


def make_optimizer(model, optim, learning_rate, fused=True):
    """ Optimizer choices of reference :238-246: 'Adam' (lr), 'SGD' (lr, momentum 0.9), 'Adadelta' (lr); anything else raises
        ValueError('Invalid optimizer.'). With fused=True 'Adam' is the one-launch FusedAdam, whose state_dict interchanges
        with torch.optim.Adam's (same keys), so checkpoints written either way resume either way. """
    if optim == 'Adam':
        if fused:
            from tracknetv3_b200 import FusedAdam
            return FusedAdam(model.parameters(), lr=learning_rate)
        return torch.optim.Adam(model.parameters(), lr=learning_rate)
    if optim == 'SGD':
        return torch.optim.SGD(model.parameters(), lr=learning_rate, momentum=0.9)
    if optim == 'Adadelta':
        return torch.optim.Adadelta(model.parameters(), lr=learning_rate)
    raise ValueError('Invalid optimizer.')


def make_scheduler(optimizer, lr_scheduler, epochs):
    """ 'StepLR' -> StepLR(step_size=int(epochs / 3), gamma=0.1), '' -> None (reference :248-252). """
    if lr_scheduler == 'StepLR':
        return torch.optim.lr_scheduler.StepLR(optimizer, step_size=int(epochs / 3), gamma=0.1)
    return None


def checkpoint_dict(epoch, max_val_acc, model, optimizer, scheduler, param_dict):
    """ The reference's checkpoint layout (:283-301): epoch, max_val_acc, model, optimizer, scheduler (None without one),
        param_dict - what `--resume_training` and predict.py (:99-108: ckpt['model'], ckpt['param_dict']) read. """
    return dict(epoch=epoch, max_val_acc=max_val_acc, model=model.state_dict(), optimizer=optimizer.state_dict(),
                scheduler=scheduler.state_dict() if scheduler is not None else None, param_dict=param_dict)


def resume_from(ckpt, model, optimizer, scheduler):
    """ Restore model / optimizer / scheduler from a checkpoint dict written by this file or by the reference
        (:255-262). Returns (start_epoch, max_val_acc). """
    model.load_state_dict(ckpt['model'])
    optimizer.load_state_dict(ckpt['optimizer'])
    if scheduler is not None and ckpt.get('scheduler') is not None:
        scheduler.load_state_dict(ckpt['scheduler'])
    return ckpt['epoch'] + 1, ckpt['max_val_acc']


def init_distributed():
    """ (rank, world, local_rank) from the torchrun environment; initialises the process group (NCCL on GPUs, gloo
        without) and binds this process to its GPU. A plain `python train.py` is rank 0 of a world of 1. """
    import os
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank, local = int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
    return rank, world, local


def fit(model, optimizer, scheduler, train_loader_fn, val_loader_fn, param_dict, train_fn, eval_fn,
        start_epoch=0, max_val_acc=0., save_dir=None, log=print, rank=0, world=1):
    """ Epoch loop of reference :268-305 without the TensorBoard writer: train, evaluate, scheduler step, `<model>_best.pt`
        when the validation accuracy does not drop, `<model>_cur.pt` every epoch. The loaders are given as callables
        returning a fresh iterable per epoch. Returns (max_val_acc, history).
        world > 1 (data parallel): the replicas are synchronised to rank 0 first, every step averages the gradients over
        the ranks, rank 0 alone evaluates, logs and writes the checkpoints and shares the accuracy with the others. """
    import os
    name = param_dict['model_name']
    history = []
    bucket = None
    if world > 1:
        import torch.distributed as dist
        from tracknetv3_b200.parallel import GradBucket, broadcast_module
        broadcast_module(model)
        bucket = GradBucket(model, overlap=os.environ.get("TNB_ALLREDUCE_OVERLAP") == "1")  # measured: off is faster
    for epoch in range(start_epoch, param_dict['epochs']):
        train_loss = train_fn(model, optimizer, train_loader_fn(), param_dict, bucket) if world > 1 else \
            train_fn(model, optimizer, train_loader_fn(), param_dict)
        val = [0., 0.]
        if rank == 0:
            val_loss, val_res = eval_fn(model, val_loader_fn(), param_dict)
            val = [val_loss, val_res['accuracy'] if name == 'TrackNet' else val_res['inpaint']['accuracy']]
        if world > 1:
            t = torch.tensor(val, dtype=torch.float64, device=next(model.parameters()).device)
            dist.broadcast(t, src=0)
            val = t.tolist()
        val_loss, cur_val_acc = val
        if scheduler is not None:
            scheduler.step()
        history.append((epoch, train_loss, val_loss, cur_val_acc))
        if rank == 0:
            log(f'Epoch [{epoch + 1} / {param_dict["epochs"]}] train loss {train_loss:.6f} val loss {val_loss:.6f} '
                f'val accuracy {cur_val_acc:.4f}')
        if save_dir is not None and rank == 0:
            if cur_val_acc >= max_val_acc:
                torch.save(checkpoint_dict(epoch, cur_val_acc, model, optimizer, scheduler, param_dict),
                           os.path.join(save_dir, f'{name}_best.pt'))
            torch.save(checkpoint_dict(epoch, max(max_val_acc, cur_val_acc), model, optimizer, scheduler, param_dict),
                       os.path.join(save_dir, f'{name}_cur.pt'))
        max_val_acc = max(max_val_acc, cur_val_acc)
    return max_val_acc, history


def _synthetic_tracknet_loader(steps, batch_size, seq_len, bg_mode, h=288, w=512, seed=0):
    """ Batches in the layout of the reference's heatmap-mode dataset (dataset.py:611-647): (index, x, y, coordinates,
        visibility-unused); y = one radius-2.5 disc per frame (dataset.py:401-410), coordinates normalised to [0, 1]. """
    in_dim = get_model('TrackNet', seq_len, bg_mode).in_dim
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(1, h + 1), torch.arange(1, w + 1), indexing='ij')
    for i in range(steps):
        x = torch.rand(batch_size, in_dim, h, w, generator=g)
        cx = torch.randint(0, w, (batch_size, seq_len), generator=g)
        cy = torch.randint(0, h, (batch_size, seq_len), generator=g)
        y = (((xx - (cx[..., None, None] + 1)) ** 2 + (yy - (cy[..., None, None] + 1)) ** 2) <= 2.5 ** 2).float()
        y = y * ((cx != 0) | (cy != 0))[..., None, None]  # (0, 0) is the reference's "no shuttlecock": an empty map
        c = torch.stack([cx / w, cy / h], -1).float()
        idx = torch.stack([torch.full((batch_size, seq_len), i), torch.arange(seq_len).expand(batch_size, seq_len)
                           + i * batch_size * seq_len + torch.arange(batch_size)[:, None] * seq_len], -1)
        yield idx, x, y, c, None


def _synthetic_inpaintnet_loader(steps, batch_size, seq_len, seed=0):
    """ Batches in the layout of the reference's coordinate-mode dataset (dataset.py:649-690): (index, predicted
        coordinates, true coordinates, predicted visibility, true visibility, inpaint mask); smooth parabolic
        trajectories in [0, 1]^2, the prediction = truth + noise with ~20 % of the frames lost (set to 0), and the
        inpaint mask of `generate_inpaint_mask`'s kind: lost frames between visible ones. """
    g = torch.Generator().manual_seed(seed)
    t = torch.linspace(0, 1, seq_len)
    for i in range(steps):
        x0, vx = torch.rand(batch_size, 1, generator=g) * 0.4 + 0.1, torch.rand(batch_size, 1, generator=g) * 0.4
        y0, vy = torch.rand(batch_size, 1, generator=g) * 0.3 + 0.5, -torch.rand(batch_size, 1, generator=g) * 1.2
        coor = torch.stack([x0 + vx * t, (y0 + vy * t + 1.2 * t * t).clamp(0.1, 0.95)], -1)
        lost = torch.rand(batch_size, seq_len, generator=g) < 0.2
        lost[:, 0] = lost[:, -1] = False
        coor_pred = (coor + 0.003 * torch.randn(coor.shape, generator=g)) * (~lost)[..., None]
        vis = torch.ones(batch_size, seq_len, 1)
        idx = torch.stack([torch.full((batch_size, seq_len), i), torch.arange(seq_len).expand(batch_size, seq_len)
                           + i * batch_size * seq_len + torch.arange(batch_size)[:, None] * seq_len], -1)
        yield idx, coor_pred, coor, (~lost)[..., None].float(), vis, lost[..., None].float()


if __name__ == '__main__':
    import os
    # the reference's command line (:180-199); datasets are out of scope here, so the loaders are either the
    # reference's own `dataset.Shuttlecock_Trajectory_Dataset` when it is importable, or --synthetic_steps random batches
    parser = argparse.ArgumentParser()
    parser.add_argument('--model_name', type=str, default='TrackNet', choices=['TrackNet', 'InpaintNet'], help='model type')
    parser.add_argument('--seq_len', type=int, default=8, help='sequence length of input')
    parser.add_argument('--epochs', type=int, default=3, help='number of epochs')
    parser.add_argument('--batch_size', type=int, default=10, help='batch size of training')
    parser.add_argument('--optim', type=str, default='Adam', choices=['Adam', 'SGD', 'Adadelta'], help='optimizer')
    parser.add_argument('--learning_rate', type=float, default=0.001, help='initial learning rate')
    parser.add_argument('--lr_scheduler', type=str, default='', choices=['', 'StepLR'], help='learning rate scheduler')
    parser.add_argument('--bg_mode', type=str, default='', choices=['', 'subtract', 'subtract_concat', 'concat'])
    parser.add_argument('--alpha', type=float, default=-1, help='alpha of sample mixup, -1 means no mixup')
    parser.add_argument('--frame_alpha', type=float, default=-1)
    parser.add_argument('--mask_ratio', type=float, default=0.3)
    parser.add_argument('--tolerance', type=float, default=4)
    parser.add_argument('--resume_training', action='store_true', default=False)
    parser.add_argument('--seed', type=int, default=13)
    parser.add_argument('--save_dir', type=str, default='exp')
    parser.add_argument('--debug', action='store_true', default=False)
    parser.add_argument('--verbose', action='store_true', default=False)
    parser.add_argument('--synthetic_steps', type=int, default=5, help='random batches per epoch (no dataset in this repo)')
    args = parser.parse_args()
    param_dict = vars(args)
    rank, world, _ = init_distributed()
    np.random.seed(args.seed + rank)      # mixup / random-mask draws differ per replica; the weights come from rank 0
    torch.manual_seed(args.seed + rank)
    os.makedirs(args.save_dir, exist_ok=True)
    ckpt = None
    if args.resume_training:
        path = os.path.join(args.save_dir, f'{args.model_name}_cur.pt')
        assert os.path.exists(path), f'No checkpoint found in {args.save_dir}'
        ckpt = torch.load(path, weights_only=False)
        param_dict = dict(ckpt['param_dict'], resume_training=True, epochs=args.epochs, verbose=args.verbose)
        param_dict.setdefault('synthetic_steps', args.synthetic_steps)
    P = argparse.Namespace(**param_dict)
    if rank == 0:
        print(f'Parameters: {param_dict}' + (f' (data parallel over {world} GPUs, global batch {world * P.batch_size})' if world > 1 else ''))
    from test import eval_tracknet, eval_inpaintnet
    tracknet = P.model_name == 'TrackNet'
    model = get_model(P.model_name, P.seq_len, P.bg_mode).cuda() if tracknet else get_model(P.model_name).cuda()
    optimizer = make_optimizer(model, P.optim, P.learning_rate)
    scheduler = make_scheduler(optimizer, P.lr_scheduler, P.epochs)
    start_epoch, max_val_acc = resume_from(ckpt, model, optimizer, scheduler) if ckpt is not None else (0, 0.)
    if tracknet:
        loaders = lambda seed: (lambda: _synthetic_tracknet_loader(P.synthetic_steps, P.batch_size, P.seq_len, P.bg_mode, seed=seed))
    else:
        loaders = lambda seed: (lambda: _synthetic_inpaintnet_loader(P.synthetic_steps, P.batch_size, P.seq_len, seed=seed))
    # every rank trains on its own batches (seed + rank); the validation batches are rank 0's
    best, _ = fit(model, optimizer, scheduler, loaders(P.seed + 1000 * rank), loaders(P.seed + 1), param_dict,
                  train_tracknet if tracknet else train_inpaintnet, eval_tracknet if tracknet else eval_inpaintnet,
                  start_epoch, max_val_acc, P.save_dir, rank=rank, world=world)
    if rank == 0:
        print(f'best validation accuracy {best:.4f}; checkpoints in {P.save_dir}')
    if world > 1:
        torch.distributed.destroy_process_group()
