"""-m gpu, needs 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): the data-parallel train step on hardware - one
process per GPU, NCCL over NVLink. The reference has no multi-GPU code; the contract is torch DDP's: replicas start
identical, every step's gradients are the MEAN of the per-rank gradients, BatchNorm statistics stay per replica."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import tracknetv3_b200 as T
    import train as TR
    from tracknetv3_b200.parallel import GradBucket, broadcast_module
    r, w, local = TR.init_distributed()
    assert dist.get_backend() == "nccl" and torch.cuda.current_device() == rank
    torch.manual_seed(10 + rank)                                  # different initial weights: broadcast must fix it
    model = T.TrackNet(12, 4).cuda().train()
    broadcast_module(model)
    gen = torch.Generator().manual_seed(20 + rank)                # different batches per rank
    x = torch.rand(2, 12, 64, 96, generator=gen).cuda()
    y = (torch.rand(2, 4, 64, 96, generator=gen) > 0.98).float().cuda()
    T.WBCELoss(model(x), y).backward()
    local_grads = torch.cat([p.grad.flatten() for p in model.parameters()]).clone()
    gathered = [torch.empty_like(local_grads) for _ in range(world)]
    dist.all_gather(gathered, local_grads)
    bucket = GradBucket(model)
    assert bucket._shared_flat([p.grad for p in model.parameters()]) is not None   # in place: no flatten / copy-back
    bucket.allreduce()
    avg = torch.cat([p.grad.flatten() for p in model.parameters()])
    want = sum(gathered) / world
    # overlapped mode: the backward runs as two ranges, the bottleneck / decoder / predictor gradients are all-reduced on a
    # side stream under the encoder's backward. Same weights, same batches: the per-rank gradients are bit-identical to
    # the ones gathered above, so every one of several steps must give exactly the same average
    bucket2 = GradBucket(model, overlap=True)
    assert bucket2.split is not None and model._grad_split is bucket2.split
    worst = 0.0
    for _ in range(6):
        for p in model.parameters():
            p.grad = None
        T.WBCELoss(model(x), y).backward()
        assert bucket2.split.first_param == 3 * 7
        bucket2.allreduce()
        avg2 = torch.cat([p.grad.flatten() for p in model.parameters()])
        worst = max(worst, (avg2 - avg).abs().max().item())
    assert worst == 0.0, worst
    weights = torch.cat([p.detach().flatten() for p in model.parameters()])
    rm = model.down_block_1.conv_1.bn.running_mean.clone()
    q.put((rank, (avg - want).abs().max().item(), want.abs().max().item(), (gathered[0] - gathered[1]).abs().max().item(),
           weights.double().sum().item(), rm.double().sum().item()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_averaged_gradients_equal_the_mean_of_the_rank_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    [p.join(60) for p in procs]
    for rank, err, scale, diff, wsum, rmsum in res:
        assert err <= 1e-7 * scale, (rank, err, scale)           # allreduce(AVG) == mean of the gathered rank gradients
        assert diff > 1e-3 * scale                               # the ranks really had different gradients
    assert res[0][4] == res[1][4]                                # identical replicas (rank 0's weights)
    assert res[0][5] != res[1][5]                                # BatchNorm running statistics stay per replica
