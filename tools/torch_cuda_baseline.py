"""Measurement only (never on the product path): the reference architecture written with stock torch modules
(nn.Conv2d / nn.BatchNorm2d / nn.ReLU / nn.MaxPool2d / nn.Upsample / torch.cat, i.e. what reference model.py:4-73
instantiates) run on torch-CUDA (cuDNN/ATen) on the same box, same workload and timed region as bench.py `value`
(fwd + WBCE + bwd, batch resident in HBM). This is the "reference's own torch-CUDA FPS" of BASELINE.json's north_star.
usage: python tools/torch_cuda_baseline.py [steps] [warmup]  -> one JSON line per variant."""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench as B  # synthetic_batch only


def block(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding="same", bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class Net(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        s = lambda *c: nn.Sequential(*[block(a, b) for a, b in zip(c[:-1], c[1:])])
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
    return (-((1 - p) ** 2 * y * torch.log(torch.clamp(p, 1e-7, 1)) + p ** 2 * (1 - y) * torch.log(torch.clamp(1 - p, 1e-7, 1)))).mean()


def run(name, steps, warmup, tf32, benchmark, channels_last):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = benchmark
    torch.manual_seed(13)
    net = Net(B.IN_DIM, B.OUT_DIM).cuda().train()
    x, y = B.synthetic_batch(B.BATCH, 13)
    x, y = x.cuda(), y.cuda()
    if channels_last:
        net = net.to(memory_format=torch.channels_last); x = x.contiguous(memory_format=torch.channels_last)

    def step():
        for p in net.parameters():
            p.grad = None
        wbce(net(x), y).backward()

    for _ in range(warmup):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"impl": "torch-cuda", "variant": name, "ms_per_step": ms, "frames_per_s": B.BATCH * B.SEQ_LEN / ms * 1e3,
                      "allow_tf32": tf32, "cudnn_benchmark": benchmark, "channels_last": channels_last,
                      "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}), flush=True)


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    warmup = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    run("torch defaults (cudnn TF32 allowed), NCHW", steps, warmup, True, False, False)
    run("strict fp32 (allow_tf32=False), NCHW", steps, warmup, False, False, False)
    run("best-effort: TF32 + cudnn.benchmark + channels_last", steps, warmup, True, True, True)
    run("strict fp32 + cudnn.benchmark + channels_last", steps, warmup, False, True, True)
