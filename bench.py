#!/usr/bin/env python
"""bench.py — TrackNet train-step throughput on B200 (BASELINE.json metric) and its CPU reference arm.

  python bench.py --gpus N --steps K --warmup W                 # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): TrackNet seq_len 8, bg_mode concat (in 27 / out 8 channels), batch 10
per GPU, 288x512, one "step" = forward + WBCELoss + backward (reference train.py:92-95) on synthetic frames
and binary-disc labels (reference dataset.py:401-410). frames = batch x seq_len.

Printed JSON (one line, rank 0): value = frames/s with inputs resident in HBM; e2e = the same step through
the reference-facing modules with pinned-host inputs copied H2D every step and loss.item() read back;
roofline = the tcgen05 conv kernel's algorithmic TFLOP/s from per-launch CUDA events (the same steps repeated once
more right after the timed region, which itself runs as CUDA-graph replays) against the measured dense bf16 peak; cpu_baseline = the oracle port of the reference on the host cores.
"""
import argparse
import ctypes as C
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

# (cin, cout, level) of the 17 3x3 convolutions, reference model.py:47-53 (SURVEY.md §8a layer table)
LAYERS = [(IN_DIM, 64, 0), (64, 64, 0), (64, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 2), (256, 256, 2),
          (256, 512, 3), (512, 512, 3), (512, 512, 3), (768, 256, 2), (256, 256, 2), (256, 256, 2), (384, 128, 1),
          (128, 128, 1), (192, 64, 0), (64, 64, 0)]


def conv_flops(n, cin, cout, level):
    return 2.0 * n * (H >> level) * (W >> level) * cin * cout * 9


def synthetic_batch(n, seed):
    """x ~ U[0,1) like /255 frames; y = radius-2.5 binary discs at random centres, ~15% empty maps."""
    import numpy as np
    import torch
    from oracle.tracknet_oracle import label_disc  # label rule only (dataset.py:401-410); not on the timed path
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, IN_DIM, H, W, generator=g)
    rng = np.random.default_rng(seed)
    y = np.zeros((n, OUT_DIM, H, W), dtype=np.float32)
    for i in range(n):
        for f in range(OUT_DIM):
            if rng.random() > 0.15:
                y[i, f] = label_disc(int(rng.integers(0, W)), int(rng.integers(0, H)))
    return x, torch.from_numpy(y)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    hbm=float(d.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md, sustained)")


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


def cpu_reference_step_rate(steps, warmup, batch=1):
    """The reference algorithm (oracle port: plain torch CPU fp32, all host threads) on a bounded sample:
    `batch` samples of the same 288x512 seq_len-8 workload per step. Returns (frames/s, ms/step, cores)."""
    import torch
    from oracle import tracknet_oracle as O
    # torch's default intra-op pool = the physical cores it detects; oversubscribing all SMT siblings of a big
    # host makes these small convolutions slower, so the default is "all the threads torch will use"
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
    if rank != 0:
        return
    steps, warmup = min(args.steps, 8), min(args.warmup, 2)
    fps, ms, cores = cpu_reference_step_rate(steps, warmup, batch=1)
    sample = f"bs=1 of the bs={BATCH} workload per step ({steps} timed + {warmup} warm-up steps), torch CPU fp32, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "TrackNet seq_len=8 bg=concat 288x512 fwd+WBCE+bwd (configs[1])",
                       "global_batch": 1, "note": "reference is pure Python/PyTorch: CPU arm = oracle port"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp32x3", choices=["fp32x3", "tf32like"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-launch", action="store_true", help="print the mean duration of every profiled launch to stderr")
    ap.add_argument("--variant", type=int, default=0, help="kernel experiment bits (tnb_tracknet_cfg_t.variant)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank)

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
    model = T.TrackNet(IN_DIM, OUT_DIM, precision=args.precision).cuda().train()
    model._variant = args.variant
    if world > 1:
        broadcast_module(model)
    bucket = GradBucket(model)
    x_host, y_host = synthetic_batch(BATCH, 13 + rank)
    x_pin, y_pin = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_pin.cuda(), y_pin.cuda()

    def step_resident():
        for p in model.parameters():
            p.grad = None
        loss = T.WBCELoss(model(x_dev), y_dev)
        loss.backward()
        bucket.allreduce()
        return loss

    # e2e: every step's inputs start in pinned host memory and are copied H2D inside the timed region (side stream,
    # overlapped with the previous step's compute by the package's DevicePrefetcher); the loss is read back every step
    e2e_state = {"loader": None}

    def step_e2e():
        for p in model.parameters():
            p.grad = None
        xd, yd = next(e2e_state["loader"])
        loss = T.WBCELoss(model(xd), yd)
        loss.backward()
        bucket.allreduce()
        return loss.item()  # D2H read of the step result, like train.py:94

    def host_batches(count):
        for _ in range(count):
            yield (x_pin, y_pin)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
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
    e2e_state["loader"] = T.DevicePrefetcher(host_batches(1))
    step_e2e()  # untimed warm-up of the e2e path

    def timed_e2e(steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        e2e_state["loader"] = T.DevicePrefetcher(host_batches(steps))  # the first copy is issued inside the region
        for _ in range(steps):
            step_e2e()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    ms_e2e = timed_e2e(args.steps)

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
    terms = 3 if args.precision == "fp32x3" else 1
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "kernel_metrics_r1.json")
    if os.path.exists(prof) and args.precision == "fp32x3":  # dram read+write per conv launch from the committed ncu pass
        km = json.load(open(prof))
        conv = [v for k, v in km.items() if k.startswith("conv3x3_kernel")]
        traffic = sum(v["dram_bytes_per_launch"] * v["launches"] for v in conv) / sum(v["launches"] for v in conv)
        traffic_src = "profiles/kernel_metrics_r1.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per conv launch)"
    roofline = {"bound": "tensor", "kernel": "conv3x3_kernel (forward + dgrad launches)", "achieved": achieved,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peaks["source"], "avg_launch_ms": conv_ms / max(conv_launches, 1),
                "algorithmic_flops_per_launch": conv_fl / max(conv_launches, 1),
                "mma_flops_per_algorithmic_flop": terms,
                "frac_executed": achieved * terms / peaks["tflops"],
                "timing": f"CUDA event pair around every conv launch, {prof_steps} steps run right after the timed region "
                          "(the timed region itself replays CUDA graphs)",
                "whole_step_frac": (84.78e9 * value / world) / (peaks["tflops"] * 1e12)}

    cfg = _lib.TrackNetCfg(n=BATCH, h=H, w=W, in_dim=IN_DIM, out_dim=OUT_DIM, training=1, fwd_terms=terms,
                           bwd_terms=terms, variant=args.variant, bn_eps=1e-5, bn_momentum=0.1)
    launches = (lib.tnb_tracknet_num_launches(C.byref(cfg), 0) + lib.tnb_tracknet_num_launches(C.byref(cfg), 1)
                + 3) * args.steps  # + WBCE forward (2 kernels) and backward (1)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        fps, ms_cpu, cores = cpu_reference_step_rate(3, 1, batch=1)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "bs=1 of the bs=10 workload, 3 timed + 1 warm-up train steps, torch CPU fp32 (oracle port)"}

    gs = (C.c_longlong * 4)()
    lib.tnb_graph_stats(gs)
    nbytes_in = x_pin.numel() * 4 + y_pin.numel() * 4
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "fp16x3 (3-term hi/lo split operands, fp32 accumulate: fp32-faithful)" if terms == 3
            else "fp16 single pass, fp32 accumulate (TF32-class)",
            "data": "synthetic",
            "config": {"workload": "TrackNet seq_len=8 bg=concat (27->8 ch) 288x512 fwd+WBCE+bwd, BASELINE configs[1]",
                       "global_batch": BATCH * world, "per_gpu_batch": BATCH, "parallelism": f"dp{world}",
                       "precision": args.precision,
                       "l2": "working set ~10 GB per step >> 126 MB L2 (inputs larger than L2; no explicit flush)"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "cuda_graphs": {"captured": gs[0], "replayed_calls": gs[1], "stream_launched_calls": gs[2], "capture_failures": gs[3]},
            "roofline": roofline, "kernel_breakdown": breakdown, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
