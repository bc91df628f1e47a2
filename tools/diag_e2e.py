"""Where do the 0.8 ms between bench.py's `value` (batch resident in HBM) and `e2e` (pinned uint8 host data every step) go?
Times, on the bench workload: the staging alone (FramePreprocessor + label_discs), the resident step, the e2e step with
and without the loss read-back. usage (GPU box): python tools/diag_e2e.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import tracknetv3_b200 as T  # noqa: E402

torch.manual_seed(13)
model = T.TrackNet(bench.IN_DIM, bench.OUT_DIM).cuda().train()
frames_h, median_h, centers_h = bench.synthetic_host_batch(bench.BATCH, 13)
fp_, mp_, cp_ = frames_h.pin_memory(), median_h.pin_memory(), centers_h.pin_memory()
fp = T.FramePreprocessor(bench.H, bench.W, bench.H, bench.W)


def stage(f, m, c):
    return fp.process(f, fp.prepare_median(m), bg_mode='concat'), T.label_discs(c, bench.H, bench.W)


fd, md, cd = fp_.cuda(), mp_.cuda(), cp_.cuda()
x_dev, y_dev = stage(fd, md, cd)


def timed(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def step(x, y, read=False):
    for p in model.parameters():
        p.grad = None
    loss = T.WBCELoss(model(x), y)
    loss.backward()
    return loss.item() if read else loss


print("stage (device-resident uint8 -> x, y)      %.3f ms" % timed(lambda: stage(fd, md, cd)))
print("  FramePreprocessor.process alone            %.3f ms" % timed(lambda: fp.process(fd, fp.prepare_median(md), bg_mode='concat')))
print("  label_discs alone                          %.3f ms" % timed(lambda: T.label_discs(cd, bench.H, bench.W)))
print("resident step                               %.3f ms" % timed(lambda: step(x_dev, y_dev)))
print("resident step + loss.item()                 %.3f ms" % timed(lambda: step(x_dev, y_dev, True)))
print("stage + step (device-resident uint8)        %.3f ms" % timed(lambda: step(*stage(fd, md, cd))))
print("stage + step + loss.item()                  %.3f ms" % timed(lambda: step(*stage(fd, md, cd), True)))
state = {}


def e2e(read):
    xd, yd = stage(*next(state["l"]))
    return step(xd, yd, read)


def run_e2e(read, n=30):
    state["l"] = T.DevicePrefetcher(iter([(fp_, mp_, cp_)] * (n + 5)))
    return timed(lambda: e2e(read), n=n, warm=5)


print("prefetcher + stage + step                   %.3f ms" % run_e2e(False))
print("prefetcher + stage + step + loss.item()     %.3f ms" % run_e2e(True))
