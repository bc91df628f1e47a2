"""CPU tests of the host-side mirror of the reference interface: module tree / state_dict layout,
get_model table and errors, data-parallel sharding + gradient bucket (gloo, world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tracknet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_layout_matches_reference():
    import tracknetv3_b200 as T
    m = T.TrackNet(27, 8)
    sd = m.state_dict()
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == O.tracknet_state_keys(27, 8)
    assert len(sd) == 104 and sum(p.numel() for p in m.parameters()) == 11341000
    assert sd["bottleneck.conv_2.bn.num_batches_tracked"].dtype == torch.int64
    assert [t.data_ptr() for t in m._state_tensors()] == [v.data_ptr() for v in sd.values()]
    i = T.InpaintNet()
    assert list(i.state_dict().keys()) == list(O.init_inpaintnet_state(0).keys())
    assert sum(p.numel() for p in i.parameters()) == 520610
    # torch-default init under the same seed reproduces the oracle's (and hence the reference's) init
    torch.manual_seed(5)
    m2 = T.TrackNet(12, 4)
    ref = O.init_tracknet_state(5, 12, 4)
    assert all(torch.equal(v.float(), ref[k].float()) for k, v in m2.state_dict().items())


def test_get_model_table_and_errors():
    from utils.general import get_model, COOR_TH, HEIGHT, WIDTH
    assert get_model("TrackNet", 8, "concat").in_dim == 27
    assert get_model("TrackNet", 8, "subtract").in_dim == 8
    assert get_model("TrackNet", 8, "subtract_concat").in_dim == 32
    assert get_model("TrackNet", 4, "").in_dim == 12 and get_model("TrackNet", 4, "").out_dim == 4
    assert type(get_model("InpaintNet")).__name__ == "InpaintNet"
    with pytest.raises(ValueError, match="Invalid model name"):
        get_model("YOLO")
    assert (HEIGHT, WIDTH) == (288, 512) and abs(COOR_TH - 50 / np.sqrt(288 ** 2 + 512 ** 2)) < 1e-12


def test_predict_coordinate_branch_and_errors():
    import predict as P
    idx = torch.tensor([[[0, 0], [0, 1], [0, 2]], [[0, 2], [0, 3], [0, 3]]])
    c = torch.tensor([[[0.5, 0.5], [0.0, 0.0], [0.25, 0.75]], [[0.9, 0.9], [0.1, 0.2], [0.3, 0.3]]])
    out = P.predict(idx, c_pred=c, img_scaler=(2.5, 2.5))
    # frame 2 is repeated at the start of sample 1 -> that sample stops immediately (reference predict.py:46-67)
    assert out["Frame"] == [0, 1, 2] and out["X"] == [640, 0, 320] and out["Y"] == [360, 0, 540]
    assert out["Visibility"] == [1, 0, 1]
    with pytest.raises(ValueError, match="Invalid input"):
        P.predict(idx)


def test_shard_batch_partitions_exactly():
    from tracknetv3_b200.parallel import shard_batch
    for n, w in ((80, 8), (10, 4), (7, 2), (3, 8)):
        ranges = [shard_batch(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tracknetv3_b200.parallel import GradBucket, broadcast_module
    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4))
    broadcast_module(net)
    w0 = torch.cat([p.detach().flatten() for p in net.parameters()]).clone()
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    GradBucket(net).allreduce()
    g = torch.cat([p.grad.flatten() for p in net.parameters()])
    # gradients handed out as slices of one allocation (what TrackNet's backward does): reduced in place, no copies
    flat = torch.cat([torch.full((p.numel(),), float(rank + 1) * (i + 1)) for i, p in enumerate(net.parameters())])
    off = 0
    for p in net.parameters():
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    bucket = GradBucket(net, overlap=True)  # asked for, not available (gloo, a module without a two-range backward): one call
    assert bucket.split is None and not hasattr(net, "_grad_split")
    assert bucket._shared_flat([p.grad for p in net.parameters()]) is not None
    bucket.allreduce()
    assert torch.equal(flat, g) and torch.equal(torch.cat([p.grad.flatten() for p in net.parameters()]), g)
    q.put((rank, w0.numpy(), g.numpy()))
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    [p.join(60) for p in procs]
    assert np.array_equal(res[0][1], res[1][1])              # identical replicas after broadcast
    assert np.array_equal(res[0][2], res[1][2])              # identical averaged gradients
    # rank grads were (r+1)*(i+1): mean over ranks = 1.5*(i+1)
    assert np.allclose(np.unique(res[0][2]), [1.5, 3.0, 4.5, 6.0])


def test_resample_tables_match_pillow_restatement():
    """Host-side coefficient tables of the GPU resize (tracknetv3_b200.frames) == the oracle's restatement of
    Pillow's precompute_coeffs / normalize_coeffs_8bpc, which tests/test_oracle.py pins against real Pillow."""
    import tracknetv3_b200 as T
    from oracle import pil_resize_oracle as P
    for a, b in ((720, 288), (1280, 512), (1080, 288), (50, 100), (31, 5), (97, 40)):
        b1, k1 = T.resample_table(a, b)
        b2, k2 = P.precompute_coeffs(a, b)
        assert np.array_equal(b1, b2) and np.array_equal(k1, k2), (a, b)
        assert (k1.sum(1) - (1 << 22)).__abs__().max() <= k1.shape[1]      # weights sum to 1 in 22-bit fixed point
    b, k = T.resample_table(288, 288)                                      # identity pass
    assert np.array_equal(b[:, 0], np.arange(288)) and (b[:, 1] == 1).all() and (k == 1 << 22).all()


def test_evaluate_coordinate_mode_matches_reference_fixture():
    """test.evaluate in coordinate mode is host arithmetic on (N, L, 2) arrays (no GPU involved): check it against the
    reference's evaluate (fixture) here; the heatmap mode is a -m gpu test."""
    import test as TT
    g = np.load(os.path.join(ROOT, "tests", "golden", "evaluate.npz"))
    d = TT.evaluate(torch.from_numpy(g["indices"]), c_true=torch.from_numpy(g["c_true"]),
                    c_pred=torch.from_numpy(g["c_pred"]), output_gt=True)
    for k in d:
        assert np.array_equal(np.asarray(d[k], dtype=np.float64), g["co/" + k]), k
    with pytest.raises(ValueError):
        TT.evaluate(torch.from_numpy(g["indices"]))


def test_generate_inpaint_mask_matches_reference_fixture():
    """test.generate_inpaint_mask (host logic between the TrackNet and InpaintNet passes, predict.py:216) vs the
    reference's function executed on 300 random trajectories by oracle/gen_golden.py."""
    import test as TT
    g = np.load(os.path.join(ROOT, "tests", "golden", "inpaint_mask.npz"))
    for y, vis, mask, th in zip(g["y"], g["vis"], g["mask"], g["th"]):
        n = int((vis >= 0).sum())
        got = TT.generate_inpaint_mask({"Y": y[:n].tolist(), "Visibility": vis[:n].tolist()}, th_h=float(th))
        assert got == mask[:n].tolist(), (y[:n], vis[:n])
    assert g["mask"].max() == 1 and (g["mask"] == 1).sum() > 200      # the fixture does exercise marked runs


def test_train_main_plumbing_optimizer_scheduler_checkpoint_resume(tmp_path):
    """Host side of the reference's train.py __main__ (:236-305): optimizer / scheduler choices, checkpoint layout,
    best / current checkpoint policy and resume. Runs on CPU with stand-in train / eval functions."""
    import train as TR
    import tracknetv3_b200 as T
    net = torch.nn.Linear(4, 2)
    assert isinstance(TR.make_optimizer(net, "Adam", 1e-3), T.FusedAdam)
    assert type(TR.make_optimizer(net, "Adam", 1e-3, fused=False)) is torch.optim.Adam
    sgd = TR.make_optimizer(net, "SGD", 0.1)
    assert type(sgd) is torch.optim.SGD and sgd.defaults["momentum"] == 0.9 and sgd.defaults["lr"] == 0.1
    assert type(TR.make_optimizer(net, "Adadelta", 1.0)) is torch.optim.Adadelta
    with pytest.raises(ValueError, match="Invalid optimizer"):
        TR.make_optimizer(net, "LAMB", 1e-3)
    assert TR.make_scheduler(sgd, "", 30) is None
    sch = TR.make_scheduler(sgd, "StepLR", 7)
    assert sch.step_size == 2 and sch.gamma == 0.1  # int(epochs / 3), reference :250

    ck = TR.checkpoint_dict(3, 0.5, net, sgd, None, {"model_name": "TrackNet"})
    assert set(ck) == {"epoch", "max_val_acc", "model", "optimizer", "scheduler", "param_dict"} and ck["scheduler"] is None

    # epoch loop: accuracy 0.2, 0.6, 0.6, 0.4 -> best is rewritten at epochs 0, 1, 2 (>=), not at 3; cur every epoch
    accs = [0.2, 0.6, 0.6, 0.4, 0.9, 0.1]
    calls = {"train": 0, "eval": 0}

    def train_fn(model, optimizer, loader, param_dict):
        calls["train"] += 1
        optimizer.zero_grad()
        model(torch.ones(1, 4)).sum().backward()
        optimizer.step()
        return 1.0 / calls["train"]

    def eval_fn(model, loader, param_dict):
        acc = accs[calls["eval"]]
        calls["eval"] += 1
        return 0.5, {"accuracy": acc}

    pd = {"model_name": "TrackNet", "epochs": 4, "tolerance": 4}
    opt = TR.make_optimizer(net, "SGD", 0.1)
    sch = TR.make_scheduler(opt, "StepLR", 4)  # step_size 1: lr x0.1 per epoch
    best, hist = TR.fit(net, opt, sch, lambda: [], lambda: [], pd, train_fn, eval_fn, save_dir=str(tmp_path), log=lambda s: None)
    assert best == 0.6 and [h[0] for h in hist] == [0, 1, 2, 3]
    cur = torch.load(tmp_path / "TrackNet_cur.pt", weights_only=False)
    bst = torch.load(tmp_path / "TrackNet_best.pt", weights_only=False)
    assert cur["epoch"] == 3 and bst["epoch"] == 2 and cur["max_val_acc"] == 0.6 and bst["max_val_acc"] == 0.6
    assert abs(opt.param_groups[0]["lr"] - 0.1 * 0.1 ** 4) < 1e-12

    # resume: a fresh model / optimizer / scheduler continue at epoch 4 with the restored weights, momentum and lr
    net2 = torch.nn.Linear(4, 2)
    opt2 = TR.make_optimizer(net2, "SGD", 0.1)
    sch2 = TR.make_scheduler(opt2, "StepLR", 4)
    start, mx = TR.resume_from(cur, net2, opt2, sch2)
    assert (start, mx) == (4, 0.6)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    assert abs(opt2.param_groups[0]["lr"] - opt.param_groups[0]["lr"]) < 1e-15 and sch2.last_epoch == sch.last_epoch
    mb = [opt.state[p]["momentum_buffer"] for p in net.parameters()]
    mb2 = [opt2.state[p]["momentum_buffer"] for p in net2.parameters()]
    assert all(torch.equal(a, b) for a, b in zip(mb, mb2))
    pd["epochs"] = 6
    best2, hist2 = TR.fit(net2, opt2, sch2, lambda: [], lambda: [], pd, train_fn, eval_fn, start, mx, str(tmp_path), log=lambda s: None)
    assert [h[0] for h in hist2] == [4, 5] and best2 == 0.9
    assert torch.load(tmp_path / "TrackNet_best.pt", weights_only=False)["epoch"] == 4


def test_synthetic_tracknet_loader_matches_the_reference_batch_layout():
    import train as TR
    from oracle.tracknet_oracle import label_disc
    batches = list(TR._synthetic_tracknet_loader(2, 3, 4, "concat", h=32, w=64, seed=1))
    assert len(batches) == 2
    idx, x, y, c, _ = batches[1]
    assert idx.shape == (3, 4, 2) and x.shape == (3, 15, 32, 64) and y.shape == (3, 4, 32, 64) and c.shape == (3, 4, 2)
    assert len({int(v) for v in idx[..., 1].flatten()}) == 12  # distinct frame ids inside a batch
    cx, cy = int(round(float(c[1, 2, 0]) * 64)), int(round(float(c[1, 2, 1]) * 32))
    assert np.array_equal(y[1, 2].numpy(), label_disc(cx, cy, h=32, w=64))  # the label rule of dataset.py:401-410
    assert 0 < y[1, 2].sum() <= 21  # a radius-2.5 disc has 21 pixels, fewer when clipped by the border


def test_predict_csv_writer_and_frame_reader(tmp_path):
    """Host IO around predict.run_video: the csv equals what the reference's pandas writer produces; an .mp4 written with
    OpenCV comes back frame by frame (skipped when this OpenCV build has no mp4 encoder)."""
    import pandas as pd
    import predict as P
    d = {"Frame": [0, 1, 2], "X": [640, 0, 17], "Y": [360, 0, 5], "Visibility": [1, 0, 1]}
    P.write_pred_csv(d, str(tmp_path / "a.csv"))
    ref = pd.DataFrame({"Frame": d["Frame"], "Visibility": d["Visibility"], "X": d["X"], "Y": d["Y"]}).to_csv(index=False)
    assert open(tmp_path / "a.csv").read() == ref
    with pytest.raises(AssertionError, match="Invalid video file format"):
        P.generate_frames("clip.avi")
    import cv2
    path = str(tmp_path / "clip.mp4")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 25, (64, 48))
    if not wr.isOpened():
        pytest.skip("no mp4 encoder in this OpenCV build")
    for i in range(5):
        wr.write(np.full((48, 64, 3), 40 * i, np.uint8))
    wr.release()
    frames = P.generate_frames(path)
    assert len(frames) == 5 and frames[0].shape == (48, 64, 3)
    assert abs(int(frames[3].mean()) - 120) <= 3


def test_predict_load_models_reads_the_reference_checkpoint_layout(tmp_path):
    import predict as P
    import train as TR
    from utils.general import get_model
    torch.manual_seed(3)
    tn, ip = get_model("TrackNet", 4, "subtract"), get_model("InpaintNet")
    pd_t = {"model_name": "TrackNet", "seq_len": 4, "bg_mode": "subtract"}
    pd_i = {"model_name": "InpaintNet", "seq_len": 16, "bg_mode": ""}
    torch.save(TR.checkpoint_dict(0, 0.1, tn, torch.optim.SGD(tn.parameters(), lr=0.1), None, pd_t), tmp_path / "t.pt")
    torch.save(TR.checkpoint_dict(0, 0.1, ip, torch.optim.SGD(ip.parameters(), lr=0.1), None, pd_i), tmp_path / "i.pt")
    pd_i["seq_len"] = 12
    torch.save(TR.checkpoint_dict(0, 0.1, ip, torch.optim.SGD(ip.parameters(), lr=0.1), None, pd_i), tmp_path / "i.pt")
    t2, i2, seq_len, bg_mode, inpaint_seq_len = P.load_models(str(tmp_path / "t.pt"), str(tmp_path / "i.pt"))
    assert (seq_len, bg_mode, t2.in_dim, t2.out_dim, inpaint_seq_len) == (4, "subtract", 4, 4, 12)
    assert all(torch.equal(a, b) for a, b in zip(tn.state_dict().values(), t2.state_dict().values()))
    assert all(torch.equal(a, b) for a, b in zip(ip.state_dict().values(), i2.state_dict().values()))
    assert P.load_models(str(tmp_path / "t.pt"))[1] is None


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys, the
    same metric / unit / config family as the GPU arm, zero transfer bytes, and a cpu_baseline that describes itself."""
    import json
    import subprocess
    # --cpu-batch 1: one sample per step (the default takes as many of the 10 as fit its time budget on the host)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-batch", "1"], capture_output=True, text=True, cwd=ROOT, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))   # what torchrun exports to its workers
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["value"] - 8 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]  # bs 1 x seq_len 8 per step
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "bs=1" in cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0))              # all host threads, whatever OMP_NUM_THREADS said
    assert d["steps"] == 1 and d["warmup"] == 0                     # K and W as asked
    assert d["config"]["workload"] == bench.WORKLOAD and d["config"]["global_batch"] == bench.BATCH  # the GPU arm's config
    # non-zero ranks of a torchrun launch stay silent
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--cpu-batch", "1"],
                        capture_output=True, text=True, cwd=ROOT, timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_predict_windows_match_the_reference_dataset():
    """predict._windows (the sampling of input sequences over a clip / a trajectory) against the index tables the
    reference's Shuttlecock_Trajectory_Dataset built for the same arguments (tests/golden/predict_windows.npz,
    oracle/gen_predict_flow.py): sliding step 1 and seq_len, with and without padding, clips shorter than a window."""
    import predict as P
    g = np.load(os.path.join(ROOT, "tests", "golden", "predict_windows.npz"))
    assert len(g.files) == 48
    for key in g.files:
        n, L, step, pad = (int(v) for v in key.split("_"))
        win = P._windows(n, L, step, bool(pad))
        want = g[key]
        assert tuple(win.shape) == want.shape[:2], key
        assert np.array_equal(P._as_indices(win).numpy(), want), key


def test_ref_arch_is_the_reference_model():
    """tools/ref_arch.Net (the torch-CUDA baseline of bench.py; /root/reference is absent on the GPU box) is the
    reference's module graph: with the fixture's weights it reproduces the stored output of the REAL reference on the C1
    input (tests/golden/tracknet_c1.npz), train-mode BatchNorm."""
    from tools import ref_arch
    from oracle import tracknet_oracle as O
    g = np.load(os.path.join(ROOT, "tests", "golden", "tracknet_c1.npz"))
    seed = int(g["seed"])
    sd = O.init_tracknet_state(seed, 12, 4)
    gen = torch.Generator().manual_seed(seed)
    for key, shape in O.tracknet_state_keys(12, 4):      # replay the generator stream of gen_golden (weights, then x)
        if key.endswith("conv.weight") or key.startswith("predictor."):
            torch.rand(shape, generator=gen)
    x = torch.rand(tuple(g["shape"]), generator=gen)
    net = ref_arch.Net(12, 4).train()
    own = net.state_dict()
    assert [tuple(v.shape) for v in own.values()] == [tuple(v.shape) for v in sd.values()]   # same 104 tensors, same order
    net.load_state_dict(dict(zip(own.keys(), sd.values())))
    with torch.no_grad():
        y = net(x)
    assert np.abs(y.numpy() - g["y_train"]).max() < 5e-5


def _dp_fit_worker(rank, world, port, save_dir, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import train as TR
    r, w, _ = TR.init_distributed()                 # gloo here (no GPU): the same code path torchrun drives with NCCL
    assert (r, w) == (rank, world)
    torch.manual_seed(50 + rank)                     # replicas start different: fit() must synchronise them to rank 0
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    evals = []

    def train_fn(model, optimizer, loader, param_dict, bucket):
        losses = []
        for x, y in loader:
            optimizer.zero_grad()
            loss = ((model(x) - y) ** 2).mean()
            loss.backward()
            bucket.allreduce()
            optimizer.step()
            losses.append(loss.item())
        return float(np.mean(losses))

    def eval_fn(model, loader, param_dict):
        evals.append(rank)
        return 0.25, {"accuracy": 0.5 + 0.1 * len(evals)}

    def loader():                                    # every rank draws its own batches
        g = torch.Generator().manual_seed(7 + 1000 * rank)
        return [(torch.randn(4, 6, generator=g), torch.randn(4, 1, generator=g)) for _ in range(3)]

    pd = {"model_name": "TrackNet", "epochs": 2}
    best, hist = TR.fit(net, opt, None, loader, loader, pd, train_fn, eval_fn, save_dir=save_dir, log=lambda s: None,
                        rank=rank, world=world)
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    q.put((rank, flat.numpy(), best, [h[3] for h in hist], len(evals)))
    dist.destroy_process_group()


def test_data_parallel_fit_gloo_world2(tmp_path):
    """train.fit with world 2 (gloo on CPU; NCCL under torchrun on GPUs): replicas synchronised to rank 0, gradients
    averaged every step (so the replicas stay identical although their batches differ), rank 0 alone evaluates and writes
    the checkpoints, and every rank learns the validation accuracy."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_fit_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    [p.join(60) for p in procs]
    assert np.array_equal(res[0][1], res[1][1])                      # identical replicas after 6 averaged steps
    assert res[0][2] == res[1][2] == 0.7 and res[0][3] == res[1][3] == [0.6, 0.7]
    assert (res[0][4], res[1][4]) == (2, 0)                          # only rank 0 evaluated
    ck = torch.load(tmp_path / "TrackNet_cur.pt", weights_only=False)
    assert ck["epoch"] == 1 and ck["max_val_acc"] == 0.7 and (tmp_path / "TrackNet_best.pt").exists()
    # the checkpoint holds the synchronised weights
    assert np.array_equal(torch.cat([v.flatten() for v in ck["model"].values()]).numpy(), res[0][1])
