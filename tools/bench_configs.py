"""Measurement of BASELINE.json configs[3] (predict.py inference path) and configs[4] (resolution sweep) on one B200.
Not the headline bench (bench.py is); prints one JSON line per configuration. usage: python tools/bench_configs.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tracknetv3_b200 as T  # noqa: E402
import bench as B            # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def inference_path(bs=32, seq_len=8, h=288, w=512, steps=10):
    """configs[3]: TrackNet fwd (eval) -> temporal ensemble -> heatmap decode, then InpaintNet on the decoded
    trajectory; frames/s = bs * seq_len / (sum of the stages) as SURVEY.md 8(d) C4 defines it."""
    torch.manual_seed(0)
    net = T.TrackNet(27, seq_len).cuda().eval()
    inp = T.InpaintNet().cuda().eval()
    x = torch.rand(bs, 27, h, w, device="cuda")
    coor = torch.rand(bs, 16, 2, device="cuda"); mask = (torch.rand(bs, 16, 1, device="cuda") < 0.3).float()
    out = {}
    with torch.no_grad():
        y = net(x)
        out["tracknet_fwd_ms"] = timed(lambda: net(x), steps, 3)
        # decode cost depends on the amount of foreground: time it on heatmaps as a trained model emits them (one small
        # blob per frame, ~15 % empty) and on the random-init output (about half the pixels above 0.5: worst case)
        xb, yb = B.synthetic_batch(bs, 3)
        blobs = (yb.cuda() * 0.9 + 0.02)
        out["decode_ms"] = timed(lambda: T.decode_heatmaps(blobs), steps, 3)
        out["decode_dense_random_ms"] = timed(lambda: T.decode_heatmaps(y), steps, 3)

        def ens_step():
            e = T.TemporalEnsemble(seq_len, "weight", 10 ** 9)
            e.sample_count = 100  # steady state (general case)
            e.state = y[:seq_len - 1]
            return e.push(y)
        out["ensemble_ms"] = timed(ens_step, steps, 3)
        ens = ens_step()
        out["ensemble_decode_ms"] = timed(lambda: T.decode_heatmaps(ens.unsqueeze(1)), steps, 3)
        from utils.general import COOR_TH
        out["inpaintnet_fwd_ms"] = timed(lambda: inp.rectify(coor * (1 - mask), mask, COOR_TH), steps, 3)  # fwd + blend + threshold
    # GPU input pipeline: bs x seq_len 720p RGB frames -> Pillow-exact 288x512 -> (bs, 27, H, W)
    fp = T.FramePreprocessor(720, 1280, h, w)
    frames = torch.randint(0, 256, (bs, seq_len, 720, 1280, 3), dtype=torch.uint8, device="cuda")
    med = fp.prepare_median(frames[0, 0].cpu().numpy())
    out["frame_preprocess_720p_ms"] = timed(lambda: fp.process(frames, med), steps, 3)
    out["frame_preprocess_input_GBps"] = frames.numel() / (out["frame_preprocess_720p_ms"] * 1e-3) / 1e9
    total = out["tracknet_fwd_ms"] + out["decode_ms"] + out["inpaintnet_fwd_ms"]
    out.update(config="configs[3]: predict path bs=32 seq_len=8 288x512 (nonoverlap: fwd + decode + InpaintNet)",
               frames_per_s=bs * seq_len / total * 1e3,
               fwd_tflops=bs * 227.606e9 / (out["tracknet_fwd_ms"] * 1e-3) / 1e12)
    total_e = out["tracknet_fwd_ms"] + out["ensemble_ms"] + out["ensemble_decode_ms"] + out["inpaintnet_fwd_ms"]
    out["frames_per_s_temporal_ensemble"] = bs / total_e * 1e3  # sliding_step 1: each sample completes one new frame
    print(json.dumps(out), flush=True)


def resolution_sweep(bs=8, steps=5):
    """configs[4]: train step (fwd + WBCE + bwd) at 288x512 / 360x640 / 544x960 (540 is not poolable 3x: the
    reference raises there too), seq_len 8, bs 8."""
    for h, w in ((288, 512), (360, 640), (544, 960)):
        torch.manual_seed(0)
        net = T.TrackNet(27, 8).cuda().train()
        x = torch.rand(bs, 27, h, w, device="cuda")
        y = (torch.rand(bs, 8, h, w, device="cuda") > 0.999).float()

        def step():
            for p in net.parameters():
                p.grad = None
            T.WBCELoss(net(x), y).backward()
        ms = timed(step, steps, 2)
        flops = bs * 678.232e9 * (h * w) / (288 * 512)
        print(json.dumps({"config": f"configs[4]: train step bs={bs} seq_len=8 {h}x{w}", "ms_per_step": ms,
                          "frames_per_s": bs * 8 / ms * 1e3, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12}), flush=True)
        del net, x, y
        torch.cuda.empty_cache()


def small_batch_latency(steps=30):
    """configs[0] (seq_len 4, bg none, bs 1) eval forward latency and a bs-1 train step: ~40 / ~150 dependent launches of
    short kernels, where launch latency shows (TNB_PDL / TNB_GRAPHS A/B: both are read once per process)."""
    torch.manual_seed(0)
    net = T.TrackNet(12, 4).cuda().eval()
    x = torch.rand(1, 12, 288, 512, device="cuda")
    with torch.no_grad():
        ms_eval = timed(lambda: net(x), steps, 5)
    net.train()
    y = (torch.rand(1, 4, 288, 512, device="cuda") > 0.999).float()

    def step():
        for p in net.parameters():
            p.grad = None
        T.WBCELoss(net(x), y).backward()
    ms_train = timed(step, steps, 5)
    print(json.dumps({"config": "configs[0]: seq_len=4 bs=1 288x512", "eval_forward_ms": ms_eval, "train_step_ms": ms_train,
                      "TNB_PDL": os.environ.get("TNB_PDL", "default"), "TNB_GRAPHS": os.environ.get("TNB_GRAPHS", "default")}),
          flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "latency":
        small_batch_latency()
    else:
        inference_path()
        resolution_sweep()
        small_batch_latency()
