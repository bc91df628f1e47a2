"""Measurement only (never on the product path): the reference ARCHITECTURE written with stock torch modules
(nn.Conv2d / nn.BatchNorm2d / nn.ReLU / nn.MaxPool2d / nn.Upsample / torch.cat - what reference model.py:4-73
instantiates) for the torch-CUDA baselines of bench.py and tools/torch_cuda_baseline.py. /root/reference does not exist
on the GPU box, so its model.py cannot be imported there; this module restates the module graph (not its code) and is
pinned to the real reference in the build container by tests/test_host.py (same state_dict shapes, same outputs)."""
import torch
import torch.nn as nn


def _block(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding="same", bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class Net(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        s = lambda *c: nn.Sequential(*[_block(a, b) for a, b in zip(c[:-1], c[1:])])
        self.d1, self.d2, self.d3, self.bt = s(cin, 64, 64), s(64, 128, 128), s(128, 256, 256, 256), s(256, 512, 512, 512)
        self.u1, self.u2, self.u3 = s(768, 256, 256, 256), s(384, 128, 128), s(192, 64, 64)
        self.pred = nn.Conv2d(64, cout, 1)
        self.pool, self.up = nn.MaxPool2d(2, 2), nn.Upsample(scale_factor=2)

    def forward(self, x):
        x1 = self.d1(x); x2 = self.d2(self.pool(x1)); x3 = self.d3(self.pool(x2)); t = self.bt(self.pool(x3))
        t = self.u1(torch.cat([self.up(t), x3], 1)); t = self.u2(torch.cat([self.up(t), x2], 1))
        t = self.u3(torch.cat([self.up(t), x1], 1))
        return torch.sigmoid(self.pred(t))


def wbce(p, y):
    """utils/metric.py:15-20 with reduce=True"""
    return (-((1 - p) ** 2 * y * torch.log(torch.clamp(p, 1e-7, 1))
              + p ** 2 * (1 - y) * torch.log(torch.clamp(1 - p, 1e-7, 1)))).mean()


VARIANTS = {
    # name: (allow_tf32, cudnn.benchmark, channels_last, cudnn.deterministic)
    # the reference as it runs (train.py:205 sets cudnn.deterministic; torch's defaults allow TF32 in cuDNN convolutions)
    "defaults_tf32_deterministic": (True, False, False, True),
    # the numerics class of this repo's fp32x3 path
    "strict_fp32": (False, False, False, True),
    # everything torch offers short of changing the model: TF32 + autotuned algorithms + NHWC
    "best_effort": (True, True, True, False),
}


def time_variant(name, x, y, in_dim, out_dim, steps, warmup):
    """ms per step of forward + WBCE + backward on resident tensors (the timed region of bench.py's `value`)."""
    tf32, benchmark, channels_last, deterministic = VARIANTS[name]
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark,
             torch.backends.cudnn.deterministic)
    try:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = benchmark
        torch.backends.cudnn.deterministic = deterministic
        torch.manual_seed(13)
        net = Net(in_dim, out_dim).cuda().train()
        if channels_last:
            net = net.to(memory_format=torch.channels_last)
            x = x.contiguous(memory_format=torch.channels_last)

        def step():
            for p in net.parameters():
                p.grad = None
            wbce(net(x), y).backward()

        for _ in range(warmup):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    finally:
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark,
         torch.backends.cudnn.deterministic) = saved
        net = None
        torch.cuda.empty_cache()
