#!/usr/bin/env python
"""bench.py — TrackNet train-step throughput on B200 (BASELINE.json metric) and its CPU reference arm.

  python bench.py --gpus N --steps K --warmup W                 # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): TrackNet seq_len 8, bg_mode concat (in 27 / out 8 channels), batch 10
per GPU, 288x512, one "step" = forward + WBCELoss + backward (reference train.py:92-95) on synthetic frames
and binary-disc labels (reference dataset.py:401-410). frames = batch x seq_len.

Printed JSON (one line, rank 0):
  value        frames/s with the fp32 input stack and labels resident in HBM (CUDA events, max over ranks);
  e2e          the same step through the public modules starting from HOST data every step: uint8 frames + integer
               label centres in pinned memory -> H2D on a copy stream (DevicePrefetcher) -> FramePreprocessor (resize /
               stack / 255 on the GPU) -> label_discs -> TrackNet -> WBCELoss -> backward, the loss read back by the host
               every step (ScalarReader: a side-stream copy that does not drain the compute stream);
  roofline     the tcgen05 conv kernel's algorithmic TFLOP/s from per-launch CUDA events (the same steps repeated once
               more right after the timed region, which itself runs as CUDA-graph replays) against the measured dense
               bf16 peak; `traffic` and `ncu` (tensor-pipe-active % per kernel, time-weighted over the step: the second half of
               BASELINE.json's metric) from the committed ncu pass, refused (null) when the kernel sources changed since;
  train_step   the secondary region of SURVEY.md 8(d): zero_grad + mixup + forward + loss + .item() + backward + FusedAdam
               (reference train.py:85-96), with the Adam / mixup kernels' own HBM fractions;
  torch_cuda_baseline  the reference architecture on stock torch-CUDA on the same box right after (the >= 6x target's
               denominator): torch defaults as the reference runs them, strict fp32, best-effort torch;
  alt_precision a second, labelled line: the same step with precision="fp32x3_bwd1" (forward unchanged, backward as one
               fp16 pass); never the headline, see profiles/r2_numerics.md;
  cpu_baseline the oracle port of the reference on the host cores (a reported baseline, not the target).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec TrackNet seq_len=8 288x512 bs=10 fwd+bwd"
SEQ_LEN, IN_DIM, OUT_DIM, BATCH, H, W = 8, 27, 8, 10, 288, 512
WORKLOAD = "TrackNet seq_len=8 bg=concat (27->8 ch) 288x512 fwd+WBCE+bwd, BASELINE configs[1]"

# (cin, cout, level) of the 17 3x3 convolutions, reference model.py:47-53 (SURVEY.md §8a layer table)
LAYERS = [(IN_DIM, 64, 0), (64, 64, 0), (64, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 2), (256, 256, 2),
          (256, 512, 3), (512, 512, 3), (512, 512, 3), (768, 256, 2), (256, 256, 2), (256, 256, 2), (384, 128, 1),
          (128, 128, 1), (192, 64, 0), (64, 64, 0)]
# sources whose change invalidates the committed per-kernel DRAM traffic of the conv kernel
TRAFFIC_SOURCES = ["conv_kernel.inc", "conv.cu", "conv_lean.cu", "common.cuh", "igemm.cuh"]


def conv_flops(n, cin, cout, level):
    return 2.0 * n * (H >> level) * (W >> level) * cin * cout * 9


def synthetic_host_batch(n, seed):
    """What a loader holds on the host for one batch: uint8 RGB frames (n, L, H, W, 3), one uint8 median frame
    (H, W, 3) and the integer label centres (n, L, 2), ~15 % of them (0, 0) = "no shuttlecock" (dataset.py:402-403)."""
    import numpy as np
    import torch
    rng = np.random.default_rng(seed)
    frames = torch.from_numpy(rng.integers(0, 256, size=(n, SEQ_LEN, H, W, 3), dtype=np.uint8))
    median = torch.from_numpy(rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8))
    centers = np.stack([rng.integers(0, W, size=(n, SEQ_LEN)), rng.integers(0, H, size=(n, SEQ_LEN))], -1).astype(np.int32)
    centers[rng.random((n, SEQ_LEN)) < 0.15] = 0
    return frames, median, torch.from_numpy(centers)


def synthetic_batch(n, seed):
    """The same batch as fp32 tensors on the host, the way the reference's dataset hands it to the step (x = stacked
    frames / 255 with the median first, y = radius-2.5 discs): input of the CPU arm and of the torch-CUDA baseline."""
    import numpy as np
    import torch
    from oracle.tracknet_oracle import label_disc  # label rule only (dataset.py:401-410); not on the timed path
    frames, median, centers = synthetic_host_batch(n, seed)
    x = torch.cat([median.permute(2, 0, 1).unsqueeze(0).expand(n, 3, H, W),
                   frames.permute(0, 1, 4, 2, 3).reshape(n, SEQ_LEN * 3, H, W)], dim=1).double().div(255.).float()
    y = np.zeros((n, OUT_DIM, H, W), dtype=np.float32)
    for i in range(n):
        for f in range(OUT_DIM):
            cx, cy = int(centers[i, f, 0]), int(centers[i, f, 1])
            if cx != 0 or cy != 0:
                y[i, f] = label_disc(cx, cy)
    return x, torch.from_numpy(y)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    hbm=float(d.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md, sustained)")


def sources_sha():
    h = hashlib.sha256()
    for fn in TRAFFIC_SOURCES:
        h.update(open(os.path.join(ROOT, "tracknetv3_b200", "csrc", fn), "rb").read())
    return h.hexdigest()[:16]


def committed_traffic(precision):
    """Mean DRAM read+write bytes per conv launch from the newest committed ncu pass - only if that pass was taken on
    the kernel sources of THIS tree (profiles/kernel_metrics_r*.json carry the hash); otherwise (None, why)."""
    cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("kernel_metrics_r") and f.endswith(".json"))
    if not cands or precision != "fp32x3":
        return None, "no committed ncu pass for this precision"
    path = os.path.join("profiles", cands[-1])
    km = json.load(open(os.path.join(ROOT, path)))
    meta = km.get("_meta", {})
    if meta.get("sources_sha") != sources_sha():
        return None, f"{path} is stale: taken on other kernel sources (sha {meta.get('sources_sha')} != {sources_sha()})"
    conv = [v for k, v in km.items() if k.startswith("conv3x3")]
    traffic = sum(v["dram_bytes_per_launch"] * v["launches"] for v in conv) / sum(v["launches"] for v in conv)
    return traffic, f"{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per conv launch; HEAD {meta.get('head')})"


def committed_ncu_summary(precision):
    """What the same committed ncu pass says besides the traffic (None when it is stale): tensor-pipe-active % of the
    conv (forward + dgrad) and wgrad kernels, time-weighted over the step's launches, and the DRAM GB/s of the
    BatchNorm-backward kernels - the `conv tensor-pipe %` half of BASELINE.json's metric. ncu numbers (serialised,
    cold-cache replays), not taken in this run: they describe the kernels, the timed region describes the step."""
    traffic, _ = committed_traffic(precision)
    if traffic is None:
        return None
    cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("kernel_metrics_r") and f.endswith(".json"))
    km = json.load(open(os.path.join(ROOT, "profiles", cands[-1])))
    km.pop("_meta", None)

    def weighted(prefix):
        sel = [v for k, v in km.items() if k.startswith(prefix)]
        ms = sum(v["ms"] for v in sel)
        return sum(v["tensor_pipe_pct"] * v["ms"] for v in sel) / ms if ms > 0 else None

    bn = [v for k, v in km.items() if k.startswith("bn_bwd")]
    bn_bytes = sum(v["dram_bytes_per_launch"] * v["launches"] for v in bn)
    bn_ms = sum(v["ms"] for v in bn)
    return {"source": os.path.join("profiles", cands[-1]),
            "conv_fwd_dgrad_tensor_pipe_active_pct": weighted("conv3x3"),
            "wgrad_tensor_pipe_active_pct": weighted("wgrad3x3"),
            "by_kernel_tensor_pipe_active_pct": {k: v["tensor_pipe_pct"] for k, v in km.items() if v["tensor_pipe_pct"] > 0},
            "bn_bwd_dram_gbs": bn_bytes / bn_ms / 1e6 if bn_ms > 0 else None}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nm in enumerate(names):
            if any("Active" in r[5 + i] and "Not" not in r[5 + i] for r in rows):
                out["reasons"].append(nm)
        return out


def host_threads():
    """Threads for the CPU arm: every core this process may run on, whatever the launcher put into OMP_NUM_THREADS
    (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_step_rate(steps, warmup, batch, threads=None):
    """The reference algorithm (oracle port: plain torch CPU fp32) on `batch` samples of the same 288x512 seq_len-8
    workload per step. Returns (frames/s, ms/step, threads)."""
    import torch
    from oracle import tracknet_oracle as O
    if threads is not None:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    sd = O.init_tracknet_state(13, IN_DIM, OUT_DIM)
    x, y = synthetic_batch(batch, 13)
    for _ in range(warmup):
        O.tracknet_loss_and_grads(sd, x, y, True)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.tracknet_loss_and_grads(sd, x, y, True)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch * SEQ_LEN / dt, dt * 1e3, cores


def run_reference_arm(args, rank):
    """`--impl reference`: the reference's algorithm on the host cores (the reference is pure Python / PyTorch with no
    package to install: the arm runs the oracle port, pinned to the real reference by tests/golden). Same metric, unit
    and config as the GPU arm; K timed + W warm-up steps as asked; each step is a bounded sample of the bs-10 workload -
    as many samples per step (10, 5, 2 or 1) as keep the whole run within a few minutes on this host."""
    if rank != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)  # torchrun's per-worker default of 1 is not this arm's thread count
    threads = host_threads()
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    batch = args.cpu_batch
    if batch <= 0:
        _, ms1, _ = cpu_reference_step_rate(1, 1, 1, threads)     # probe: one sample, after one warm-up
        budget_s = 150.0
        batch = 1
        for b in (10, 5, 2):
            if (steps + warmup) * b * ms1 * 1e-3 <= budget_s:
                batch = b
                break
    fps, ms, cores = cpu_reference_step_rate(steps, warmup, batch, threads)
    sample = (f"bs={batch} of the bs={BATCH} workload per step ({steps} timed + {warmup} warm-up steps), torch CPU fp32 "
              f"(oracle port of the reference), {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * args.gpus, "per_gpu_batch": BATCH,
                       "parallelism": f"dp{args.gpus}", "precision": args.precision,
                       "l2": "working set ~10 GB per step >> 126 MB L2 (inputs larger than L2; no explicit flush)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp32x3", choices=["fp32x3", "tf32like", "fp32x3_bwd1"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the torch-CUDA reference-architecture timing")
    ap.add_argument("--no-alt-precision", action="store_true", help="skip the second line (precision fp32x3_bwd1)")
    ap.add_argument("--per-launch", action="store_true", help="print the mean duration of every profiled launch to stderr")
    ap.add_argument("--variant", type=int, default=0, help="kernel experiment bits (tnb_tracknet_cfg_t.variant)")
    ap.add_argument("--cpu-batch", type=int, default=0, help="samples per step of the CPU arm (0 = as many of 10 as fit the time budget)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist
    import tracknetv3_b200 as T
    from tracknetv3_b200 import _lib
    from tracknetv3_b200.parallel import GradBucket, broadcast_module
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback for the product path"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    lib = _lib.load()
    warmup = max(args.warmup, 3)

    torch.manual_seed(13)  # reference train.py:195 default seed
    np.random.seed(13)
    model = T.TrackNet(IN_DIM, OUT_DIM, precision=args.precision).cuda().train()
    model._variant = args.variant
    if world > 1:
        broadcast_module(model)
    bucket = GradBucket(model, overlap=os.environ.get("TNB_ALLREDUCE_OVERLAP") == "1")  # measured: off is faster
    frames_h, median_h, centers_h = synthetic_host_batch(BATCH, 13 + rank)
    frames_pin, median_pin, centers_pin = frames_h.pin_memory(), median_h.pin_memory(), centers_h.pin_memory()
    fp = T.FramePreprocessor(H, W, H, W)

    def stage(frames_d, median_d, centers_d, out=(None, None)):
        """host layout -> the tensors the step consumes, on the device: resize (identity at 288x512) / stack / 255, and
        the label discs from their centres"""
        x = fp.process(frames_d, fp.prepare_median(median_d), bg_mode='concat', out=out[0])
        return x, T.label_discs(centers_d, H, W, out=out[1])

    x_dev, y_dev = stage(frames_pin.cuda(), median_pin.cuda(), centers_pin.cuda())

    def step_resident():
        for p in model.parameters():
            p.grad = None
        loss = T.WBCELoss(model(x_dev), y_dev)
        loss.backward()
        bucket.allreduce()
        return loss

    # e2e: every step's inputs start in pinned host memory (uint8 frames, the median frame, int32 label centres) and
    # are copied H2D inside the timed region (side stream, overlapped with the previous step's compute by the package's
    # DevicePrefetcher); preprocessing and labels run on the GPU; the loss is read back every step
    e2e_state = {"loader": None}
    reader = T.ScalarReader()
    # two preallocated (x, y) sets cycled by the staging: no allocator traffic (and no moving buffers under the CUDA-graph
    # cache) inside the timed region
    stage_out = [(torch.empty_like(x_dev), torch.empty_like(y_dev)) for _ in range(2)]
    e2e_state["n"] = 0

    def step_e2e():
        for p in model.parameters():
            p.grad = None
        e2e_state["n"] += 1
        xd, yd = stage(*next(e2e_state["loader"]), out=stage_out[e2e_state["n"] & 1])
        loss = T.WBCELoss(model(xd), yd)
        reader.read(loss)      # D2H copy of the step result (train.py:94) on a side stream, behind the loss kernel only
        loss.backward()
        bucket.allreduce()
        return reader.value()  # the host has this step's loss before the next step starts; the backward is not waited for

    def host_batches(count):
        for _ in range(count):
            yield (frames_pin, median_pin, centers_pin)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, before=None):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        if before is not None:
            before()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(warmup):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    # Per-launch durations of the tensor-core kernels: the SAME steps once more, right after the timed region, with a
    # CUDA event pair around every launch (tnb_profile_enable). They cannot be taken inside the timed region itself:
    # there the library replays each pass as one CUDA graph, and per-launch events would split the graph.
    prof_steps = min(args.steps, 5)
    lib.tnb_profile_enable(1)
    timed(step_resident, prof_steps)
    lib.tnb_profile_enable(0)
    maxrec = 4096
    desc = (C.c_int * (6 * maxrec))()
    kms = (C.c_float * maxrec)()
    nrec = lib.tnb_profile_collect(maxrec, desc, kms)

    e2e_warm = 10  # untimed warm-up of the e2e path: every launch-argument set of the loop (two input sets x the few
                   # addresses torch's allocator cycles for the heatmaps and their gradient) has to be seen twice before
                   # the library replays it as a CUDA graph; `cuda_graph_activity` reports what was left for the timed region
    # ONE prefetcher for warm-up and timed steps, as in a real loop in steady state: every timed step's batch was copied
    # while the step before it computed, and every timed step issues the copy of the batch after it (the last one copies
    # a batch nobody consumes) - K host->device copies inside the timed region. A fresh prefetcher per region would move
    # its buffers, and with them the addresses torch's allocator hands the step: launch-argument sets the CUDA-graph cache
    # has not seen, i.e. eager launches and captures inside the timed region.
    e2e_state["loader"] = T.DevicePrefetcher(host_batches(e2e_warm + args.steps + 1))
    for _ in range(e2e_warm):
        step_e2e()

    def start_loader():
        pass

    def graph_stats():
        g4 = (C.c_longlong * 4)()
        lib.tnb_graph_stats(g4)
        return list(g4)

    gs_before_e2e = graph_stats()
    ms_e2e = timed(step_e2e, args.steps, before=start_loader)
    gs_after_e2e = graph_stats()
    # CUDA-graph cache activity inside the e2e timed region: [captures, replays, eager calls, failures] (a capture or an
    # eager call there means a launch-argument set the cache had not seen twice: buffers that moved)
    e2e_graph_delta = [b - a for a, b in zip(gs_before_e2e, gs_after_e2e)]

    # ---- secondary region: the whole reference train step, train.py:85-96 ----
    from train import mixup
    opt = T.FusedAdam(model.parameters(), lr=1e-3)

    def step_train():
        opt.zero_grad()
        x, y = mixup(x_dev, y_dev, 0.5)
        loss = T.WBCELoss(model(x), y)
        v = loss.item()
        loss.backward()
        bucket.allreduce()
        opt.step()
        return v

    for _ in range(3):
        step_train()
    ms_train = timed(step_train, args.steps)

    def kernel_alone(fn, iters=10):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ms_adam = kernel_alone(opt.step)
    ms_mixup = kernel_alone(lambda: mixup(x_dev, y_dev, 0.5))
    nparams = sum(p.numel() for p in model.parameters())
    adam_bytes = 28.0 * nparams                                   # p, g, m, v read; p, m, v written (fp32)
    mixup_bytes = 12.0 * (x_dev.numel() + y_dev.numel())          # x[i], x[perm[i]] read, out written, for x and y

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    frames = BATCH * SEQ_LEN * world * args.steps
    value = frames / (ms_total * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)

    peaks = load_peaks()
    kinds = {0: "conv3x3 fwd", 1: "conv3x3 dgrad", 2: "conv3x3 wgrad", 3: "bn_bwd", 4: "predictor"}
    per_kind = {}
    for i in range(max(nrec, 0)):
        k, n, hh, ww, cin, cout = (desc[6 * i + j] for j in range(6))
        fl = 0.0
        if k in (0, 1, 2):
            cin_alg = IN_DIM if (k != 1 and cin == 32) else cin  # first layer: 27 real channels padded to 32
            fl = 2.0 * n * hh * ww * cin_alg * cout * 9
        e = per_kind.setdefault(k, [0.0, 0.0, 0])
        e[0] += kms[i]; e[1] += fl; e[2] += 1
    if args.per_launch:  # diagnosis: median duration of every profiled launch of a step, in launch order (stderr)
        per_step = max(nrec, 0) // prof_steps
        for j in range(per_step):
            k, n, hh, ww, cin, cout = (desc[6 * j + t] for t in range(6))
            runs = sorted(kms[j + per_step * r] for r in range(prof_steps))
            ms = runs[len(runs) // 2]  # median over the profiled steps
            fl = 2.0 * n * hh * ww * (IN_DIM if (k != 1 and cin == 32) else cin) * cout * 9 if k in (0, 1, 2) else 0.0
            print(f"launch {j:3d} {kinds.get(k, k):14s} n={n} {hh}x{ww} {cin}->{cout}: {ms:.4f} ms (min {runs[0]:.4f} max {runs[-1]:.4f})"
                  + (f"  {fl / ms / 1e9:7.1f} TFLOP/s-alg" if fl else ""), file=sys.stderr)
    breakdown = {kinds.get(k, str(k)): {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[2] / prof_steps,
                                        "tflops": (v[1] / (v[0] * 1e-3) / 1e12) if v[0] > 0 and v[1] > 0 else None}
                 for k, v in sorted(per_kind.items())}
    conv_ms = sum(per_kind.get(k, [0, 0, 0])[0] for k in (0, 1))
    conv_fl = sum(per_kind.get(k, [0, 0, 0])[1] for k in (0, 1))
    conv_launches = sum(per_kind.get(k, [0, 0, 0])[2] for k in (0, 1))
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    fwd_terms, bwd_terms = {"fp32x3": (3, 3), "tf32like": (1, 1), "fp32x3_bwd1": (3, 1)}[args.precision]
    # executed MMA FLOPs per algorithmic FLOP of the profiled conv launches (forward + dgrad)
    fl_f, fl_d = per_kind.get(0, [0, 0, 0])[1], per_kind.get(1, [0, 0, 0])[1]
    terms = (fwd_terms * fl_f + bwd_terms * fl_d) / max(fl_f + fl_d, 1.0)
    traffic, traffic_src = committed_traffic(args.precision)
    roofline = {"bound": "tensor", "kernel": "conv3x3_kernel / conv3x3_lean_kernel (forward + dgrad launches)", "achieved": achieved,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peaks["source"], "avg_launch_ms": conv_ms / max(conv_launches, 1),
                "algorithmic_flops_per_launch": conv_fl / max(conv_launches, 1),
                "mma_flops_per_algorithmic_flop": terms,
                "frac_executed": achieved * terms / peaks["tflops"],
                "timing": f"CUDA event pair around every conv launch, {prof_steps} steps run right after the timed region "
                          "(the timed region itself replays CUDA graphs)",
                "whole_step_frac": (84.78e9 * value / world) / (peaks["tflops"] * 1e12),
                "ncu": committed_ncu_summary(args.precision)}

    cfg = _lib.TrackNetCfg(n=BATCH, h=H, w=W, in_dim=IN_DIM, out_dim=OUT_DIM, training=1, fwd_terms=fwd_terms,
                           bwd_terms=bwd_terms, variant=args.variant, bn_eps=1e-5, bn_momentum=0.1)
    launches = (lib.tnb_tracknet_num_launches(C.byref(cfg), 0) + lib.tnb_tracknet_num_launches(C.byref(cfg), 1)
                + 3) * args.steps  # + WBCE forward (2 kernels) and backward (1)

    train_step = {"region": "zero_grad + mixup + forward + WBCE + loss.item() + backward + FusedAdam.step (reference train.py:85-96)",
                  "value": frames / (ms_train * 1e-3), "unit": "frames/s", "ms_per_step": ms_train / args.steps,
                  "adam_kernel": {"ms": ms_adam, "algorithmic_bytes": adam_bytes, "gbs": adam_bytes / ms_adam / 1e6,
                                  "frac_of_hbm_peak": adam_bytes / ms_adam / 1e6 / peaks["hbm"]},
                  "mixup_kernel": {"ms": ms_mixup, "algorithmic_bytes": mixup_bytes, "gbs": mixup_bytes / ms_mixup / 1e6,
                                   "frac_of_hbm_peak": mixup_bytes / ms_mixup / 1e6 / peaks["hbm"],
                                   "note": "two launches (x, y) + the host RNG draws of train.py:32-35"}}

    # second, clearly labelled line: the same step with the backward pass as ONE fp16 pass on
    # power-of-two scaled gradients (forward - heatmaps, loss - bit-identical to the headline). Never the headline:
    # profiles/r2_numerics.md has the gradient table and the 5-seed Adam trajectories that say what it costs (nothing
    # measurable) - the BASELINE metric stays on the fp32-faithful 3-term products in both directions.
    alt = None
    if args.precision == "fp32x3" and world == 1 and not args.no_alt_precision:
        m1 = T.TrackNet(IN_DIM, OUT_DIM, precision="fp32x3_bwd1").cuda().train()
        m1.load_state_dict(model.state_dict())

        def step_alt():
            for p in m1.parameters():
                p.grad = None
            T.WBCELoss(m1(x_dev), y_dev).backward()

        for _ in range(3):
            step_alt()
        n_alt = min(args.steps, 10)
        ms_alt = timed(step_alt, n_alt) / n_alt
        alt = {"precision": "fp32x3_bwd1", "dtype": "forward fp16x3 (fp32-faithful, identical heatmaps), backward single fp16 "
               "pass on 2^k-scaled gradients (11-bit operands; TF32, the reference's GPU default, has 10)",
               "ms_per_step": ms_alt, "value": BATCH * SEQ_LEN / ms_alt * 1e3, "unit": "frames/s", "steps": n_alt,
               "evidence": "profiles/r2_numerics.md"}
        del m1
        torch.cuda.empty_cache()

    torch_base = None
    if not args.no_torch_baseline and world == 1:
        # the reference architecture on stock torch-CUDA, same box, same batch, same timed region; our tensors first make
        # room (the workspaces are cached in the model)
        from tools import ref_arch
        xt, yt = x_dev.clone(), y_dev.clone()
        model._ws_saved = model._ws_scratch = None
        torch.cuda.empty_cache()
        torch_base = {"region": "forward + WBCE + backward, batch resident in HBM (= `value`)", "torch": torch.__version__,
                      "cudnn": torch.backends.cudnn.version()}
        for name in ref_arch.VARIANTS:
            try:
                ms = ref_arch.time_variant(name, xt, yt, IN_DIM, OUT_DIM, 10, 3)
                torch_base[name] = {"ms_per_step": ms, "frames_per_s": BATCH * SEQ_LEN / ms * 1e3,
                                    "this_repo_over_it": ms / (ms_total / args.steps)}
            except Exception as e:  # the baseline must never take the bench line down
                torch_base[name] = {"error": f"{type(e).__name__}: {e}"[:200]}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = host_threads()
        _, ms1, _ = cpu_reference_step_rate(1, 1, 1, threads)
        batch = 2 if 4 * 2 * ms1 * 1e-3 <= 25.0 else 1       # ~10-30 s of CPU work
        fps, ms_cpu, cores = cpu_reference_step_rate(3, 1, batch, threads)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"bs={batch} of the bs={BATCH} workload, 3 timed + 1 warm-up train steps, torch CPU fp32 (oracle port)"}

    gs = (C.c_longlong * 4)()
    lib.tnb_graph_stats(gs)
    nbytes_in = frames_pin.numel() + median_pin.numel() + centers_pin.numel() * 4
    dtypes = {"fp32x3": "fp16x3 / bf16x3 (3-term hi/lo split operands, fp32 accumulate: fp32-faithful)",
              "tf32like": "fp16 / bf16 single pass, fp32 accumulate (TF32-class)",
              "fp32x3_bwd1": "forward fp16x3 (fp32-faithful), backward single fp16 pass on 2^k-scaled gradients (TF32-class)"}
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtypes[args.precision], "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "per_gpu_batch": BATCH,
                       "parallelism": f"dp{world}", "precision": args.precision,
                       "l2": "working set ~10 GB per step >> 126 MB L2 (inputs larger than L2; no explicit flush)"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": 4,
                    "cuda_graph_activity": dict(zip(("captured", "replayed_calls", "stream_launched_calls", "capture_failures"),
                                                    e2e_graph_delta)),
                    "path": "pinned uint8 frames + median + int32 label centres -> DevicePrefetcher (copy stream) -> "
                            "FramePreprocessor + label_discs (GPU) -> TrackNet -> WBCELoss -> backward; the loss is read back every "
                            "step by ScalarReader (side-stream D2H copy behind the loss kernel, host waits for it each step)"},
            "gpu_launches": launches,
            "cuda_graphs": {"captured": gs[0], "replayed_calls": gs[1], "stream_launched_calls": gs[2], "capture_failures": gs[3]},
            "roofline": roofline, "kernel_breakdown": breakdown, "train_step": train_step,
            "alt_precision": alt, "torch_cuda_baseline": torch_base, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
