"""Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE (read from /root/reference, never copied).

Run in the build container only (the GPU box has no /root/reference):  python oracle/gen_golden.py
  * model.py and utils/metric.py import cleanly (they need only torch) and are executed as-is;
  * test.py cannot be imported (pycocotools / parse are not installed), so `predict_location` and
    `get_ensemble_weight` are extracted from its source with `ast` at generation time and exec'd against
    the real OpenCV; train.py's `mixup` likewise.
The fixtures pin oracle/ (tests/test_oracle.py) and, through it, the CUDA path (tests/test_gpu_*.py).
"""
import ast
import importlib.util
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def extract_functions(path, names, env):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), env)
    return env


def grad_stats(g):
    g = g.detach().double().flatten()
    idx = torch.linspace(0, g.numel() - 1, 16).long()
    return np.concatenate([[g.sum().item(), g.abs().sum().item(), g.abs().max().item()], g[idx].numpy()])


def gen_inpaint_train(refmodel):
    """One InpaintNet train step (reference train.py:147-166) on the REAL reference module, CPU fp32: random
    Bernoulli mask AND visibility -> zero the masked coordinates -> forward -> MSE on the masked entries ->
    backward -> clip_grad_norm_(1) -> Adam(lr 1e-3). The RNG draw of the mask (get_random_mask, train.py:42-57,
    numpy binomial) is made here and stored, the step arithmetic runs on the stored tensors."""
    torch.manual_seed(5)
    np.random.seed(5)
    net = refmodel.InpaintNet().train()
    n, L = 6, 16
    coor_gt = torch.rand(n, L, 2)
    vis_gt = (torch.rand(n, L, 1) < 0.85).float()
    coor_gt = coor_gt * vis_gt
    coor_pred = (coor_gt + 0.02 * torch.randn(n, L, 2)).clamp(0, 1) * vis_gt
    mask = torch.from_numpy(np.random.binomial(1, 0.3, size=(n, L))).float().unsqueeze(-1)
    inpaint_mask = torch.logical_and(vis_gt, mask).int()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    opt.zero_grad()
    x_in = coor_pred * (1 - inpaint_mask)
    refine = net(x_in, inpaint_mask)
    loss = torch.nn.MSELoss()(refine * inpaint_mask, coor_gt * inpaint_mask)
    loss.backward()
    grads = {k: p.grad.detach().clone().numpy() for k, p in net.named_parameters()}
    total_norm = torch.nn.utils.clip_grad_norm_(net.parameters(), 1)
    clipped = {k: p.grad.detach().clone().numpy() for k, p in net.named_parameters()}
    opt.step()
    after = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    out = {"seed": 5, "coor_gt": coor_gt.numpy(), "coor_pred": coor_pred.numpy(), "vis_gt": vis_gt.numpy(),
           "mask": mask.numpy(), "refine": refine.detach().numpy(), "loss": loss.item(),
           "total_norm": float(total_norm)}
    for k in grads:  # full gradients; clipped gradients and updated parameters as (sum, |sum|, max, 16 samples)
        out["grad/" + k] = grads[k]
        out["clipped/" + k] = grad_stats(torch.from_numpy(clipped[k]))
        out["after/" + k] = grad_stats(torch.from_numpy(after[k]))
    np.savez_compressed(f"{OUT}/inpaintnet_train.npz", **out)


def gen_tracknet_eval_step(refmodel, refmetric):
    """fwd + WBCE + backward through the REAL reference TrackNet in eval() mode (BatchNorm layers frozen at their running
    statistics; plain autograd, reference model.py:4-16), after five train-mode forwards have moved those statistics off
    their initial (0, 1): pins the oracle's eval-mode step, which in turn pins tnb_tracknet_cfg_t.training = 2."""
    torch.manual_seed(23)
    m = refmodel.TrackNet(27, 8)
    x = torch.rand(2, 27, 32, 64)
    y = (torch.rand(2, 8, 32, 64) > 0.97).float()
    m.train()
    with torch.no_grad():
        for _ in range(5):
            m(x)
    m.eval()
    before = {k: v.clone() for k, v in m.state_dict().items() if "running" in k or "tracked" in k}
    y_pred = m(x)
    loss = refmetric.WBCELoss(y_pred, y)
    loss.backward()
    assert all(torch.equal(v, before[k]) for k, v in m.state_dict().items() if k in before)
    names = [n for n, _ in m.named_parameters()]
    sd = m.state_dict()
    np.savez_compressed(f"{OUT}/tracknet_eval_step.npz", seed=23, x=x.numpy(), y=y.numpy(), y_pred=y_pred.detach().numpy(),
                        loss=loss.item(), grad_stats=np.stack([grad_stats(p.grad) for _, p in m.named_parameters()]),
                        grad_names=np.array(names),
                        grad_first=m.down_block_1.conv_1.conv.weight.grad.numpy(),
                        grad_last=m.up_block_3.conv_2.conv.weight.grad.numpy(),
                        grad_bn_w=m.bottleneck.conv_2.bn.weight.grad.numpy(),
                        grad_bn_b=m.down_block_2.conv_1.bn.bias.grad.numpy(),
                        grad_pred_w=m.predictor.weight.grad.numpy(),
                        running_mean_mid=sd["bottleneck.conv_1.bn.running_mean"].numpy(),
                        running_var_last=sd["up_block_3.conv_2.bn.running_var"].numpy(),
                        nbt=int(sd["bottleneck.conv_2.bn.num_batches_tracked"]))


def _find_stream_loop(tree, names):
    """The `for step, (<names>) in enumerate(tqdm(data_loader))` loop of predict.py's __main__ whose body holds the
    per-sample `for b in range(b_size)` ensemble loop."""
    for node in ast.walk(tree):
        if isinstance(node, ast.For) and isinstance(node.target, ast.Tuple) and len(node.target.elts) == 2 and \
                isinstance(node.target.elts[1], ast.Tuple) and \
                [getattr(e, "id", None) for e in node.target.elts[1].elts] == names and \
                any(isinstance(n, ast.For) and getattr(n.target, "id", None) == "b" for n in ast.walk(node)):
            return node
    raise RuntimeError("ensemble loop not found in the reference")


def gen_temporal_ensemble():
    """EXECUTE the reference's streaming temporal-ensemble loops (predict.py:168-209 heatmaps, :252-301 coordinates)
    as they are, lifted out of its __main__ with ast, against fake models that return prepared predictions and a fake
    `predict` that records what the loop hands to the decoder. Small maps (6x10) keep the fixture tiny."""
    src = open(f"{REF}/predict.py").read()
    tree = ast.parse(src)
    env0 = {"torch": torch, "np": np, "math": math, "tqdm": lambda x: x}
    extract_functions(f"{REF}/test.py", {"get_ensemble_weight"}, env0)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # the loops call .cuda() on their inputs; there is no GPU here
    out = {}
    try:
        for mode in ("weight", "average"):
            for L, video_len, bs in ((8, 21, 4), (4, 4, 3), (3, 11, 16)):
                H, W = 6, 10
                num_sample = video_len - L + 1
                g = torch.Generator().manual_seed(100 + L + video_len)
                preds = torch.rand(num_sample, L, H, W, generator=g)
                idx = torch.stack([torch.stack([torch.zeros(L), torch.arange(s, s + L).float()], 1) for s in range(num_sample)])
                batches = [(idx[a:a + bs], torch.arange(a, min(a + bs, num_sample))) for a in range(0, num_sample, bs)]
                rec = []
                env = dict(env0)
                env.update(HEIGHT=H, WIDTH=W, seq_len=L, num_sample=num_sample, sample_count=0, buffer_size=L - 1,
                           batch_i=torch.arange(L), frame_i=torch.arange(L - 1, -1, -1),
                           y_pred_buffer=torch.zeros((L - 1, L, H, W)), weight=env0["get_ensemble_weight"](L, mode),
                           data_loader=batches, tracknet=lambda x: preds[x.long()], img_scaler=(1, 1),
                           tracknet_pred_dict={},
                           predict=lambda i, y_pred=None, c_pred=None, img_scaler=None: rec.append((i.clone(), y_pred.clone())) or {})
                loop = _find_stream_loop(tree, ["i", "x"])
                exec(compile(ast.Module(body=[loop], type_ignores=[]), f"{REF}/predict.py", "exec"), env)
                key = f"hm_{mode}_{L}_{video_len}_{bs}"
                out[key + "/preds"] = preds.numpy()
                out[key + "/ens"] = torch.cat([r[1] for r in rec]).numpy()
                out[key + "/idx"] = torch.cat([r[0] for r in rec]).numpy()
                out[key + "/counts"] = np.array([len(r[1]) for r in rec])
        # coordinate ensemble around InpaintNet (predict.py:252-301): fake inpaintnet returns prepared coordinates
        for mode in ("weight",):
            for L, n, bs in ((16, 40, 16), (5, 5, 2)):
                g = torch.Generator().manual_seed(200 + L + n)
                coor = torch.rand(n, L, 2, generator=g)
                coor[torch.rand(n, L, generator=g) < 0.2] = 0.01  # some below COOR_TH
                pred_in = torch.rand(n, L, 2, generator=g)
                msk = (torch.rand(n, L, 1, generator=g) < 0.4).float()
                idx = torch.stack([torch.stack([torch.zeros(L), torch.arange(s, s + L).float()], 1) for s in range(n)])
                order = iter(range(0, n, bs))
                batches = [(idx[a:a + bs], pred_in[a:a + bs], msk[a:a + bs]) for a in range(0, n, bs)]
                calls = {"a": 0}

                def fake_inpaint(c, m, _c=calls, _coor=coor, _bs=bs):
                    a = _c["a"]; _c["a"] += _bs
                    return _coor[a:a + c.shape[0]]
                rec = []

                class _DS:
                    def __len__(self):
                        return n
                env = dict(env0)
                env.update(seq_len=L, COOR_TH=50 / math.sqrt(288 ** 2 + 512 ** 2), data_loader=batches, dataset=_DS(),
                           num_sample=n, sample_count=0, buffer_size=L - 1, batch_i=torch.arange(L),
                           frame_i=torch.arange(L - 1, -1, -1), coor_inpaint_buffer=torch.zeros((L - 1, L, 2)),
                           weight=env0["get_ensemble_weight"](L, mode), inpaintnet=fake_inpaint, img_scaler=(1, 1),
                           inpaint_pred_dict={},
                           predict=lambda i, y_pred=None, c_pred=None, img_scaler=None: rec.append((i.clone(), c_pred.clone())) or {})
                loop = _find_stream_loop(tree, ["i", "coor_pred", "inpaint_mask"])
                exec(compile(ast.Module(body=[loop], type_ignores=[]), f"{REF}/predict.py", "exec"), env)
                key = f"co_{mode}_{L}_{n}_{bs}"
                out[key + "/coor"] = coor.numpy(); out[key + "/pred_in"] = pred_in.numpy(); out[key + "/mask"] = msk.numpy()
                out[key + "/ens"] = torch.cat([r[1] for r in rec]).numpy()
                out[key + "/counts"] = np.array([len(r[1]) for r in rec])
    finally:
        torch.Tensor.cuda = real_cuda
    np.savez_compressed(f"{OUT}/temporal_ensemble.npz", **out)


def gen_evaluate():
    """EXECUTE the reference's `evaluate` (test.py:81-221) - lifted out with ast, together with the helpers it calls
    (`predict_location`, utils.general `to_img` / `to_img_format`) - on small synthetic heatmap and coordinate batches
    that hit every outcome: TP, TN, FP1 (too far), FP2 (ball predicted, none there), FN, padded duplicate frames."""
    import cv2
    env = {"np": np, "cv2": cv2, "torch": torch, "math": math, "HEIGHT": 288, "WIDTH": 512}
    extract_functions(f"{REF}/utils/general.py", {"to_img", "to_img_format"}, env)
    extract_functions(f"{REF}/test.py", {"predict_location", "evaluate"}, env)
    env["pred_types"] = ['TP', 'TN', 'FP1', 'FP2', 'FN']
    env["pred_types_map"] = {t: i for i, t in enumerate(env["pred_types"])}
    H, W, N, L = 36, 64, 3, 4
    rng = np.random.default_rng(21)
    y_true = np.zeros((N, L, H, W), np.float32)
    y_pred = rng.random((N, L, H, W)).astype(np.float32) * 0.4          # background below the 0.5 threshold

    def disc(a, cx, cy, r, v):
        yy, xx = np.ogrid[:H, :W]
        a[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = v
    centres = [(10, 8), (30, 20), (50, 12), (20, 30)]
    for n in range(N):
        for f in range(L):
            k = (n * L + f) % 6
            cx, cy = centres[(n + f) % 4]
            if k in (0, 1, 2, 4):
                disc(y_true[n, f], cx, cy, 2, 1.0)
            if k in (0, 1):
                disc(y_pred[n, f], cx + (1 if k else 0), cy, 2, 0.9)     # TP (0 or 1 px off)
            elif k == 2:
                disc(y_pred[n, f], (cx + 20) % W, cy, 2, 0.8)            # FP1: far from the truth
            elif k == 3:
                disc(y_pred[n, f], cx, cy, 3, 0.7)                       # FP2: nothing there
            # k == 4: FN, k == 5: TN
    y_pred[0, 1][2:4, 40:43] = 0.95                                      # a second, smaller blob: largest bbox wins
    idx = np.zeros((N, L, 2), np.int64)
    for n in range(N):
        for f in range(L):
            idx[n, f] = (0, min(n * L + f, N * L - 3))                   # last sample repeats frames: padded -> break
    out = {"y_true": y_true, "y_pred": y_pred, "indices": idx}
    for name, kw in (("plain", {}), ("bbox_gt", {"output_bbox": True, "output_gt": True}),
                     ("scaled", {"img_scaler": (2.5, 2.5), "tolerance": 1.0, "output_gt": True})):
        d = env["evaluate"](torch.from_numpy(idx), y_true=torch.from_numpy(y_true.copy()),
                            y_pred=torch.from_numpy(y_pred.copy()), **kw)
        for k, v in d.items():
            out[f"hm_{name}/{k}"] = np.asarray(v, dtype=np.float64)
    c_true = rng.random((N, L, 2)).astype(np.float32)
    c_pred = c_true + rng.normal(0, 0.004, (N, L, 2)).astype(np.float32)
    c_true[0, 0] = 0; c_pred[0, 0] = 0          # TN
    c_true[0, 1] = 0                            # FP2
    c_pred[0, 2] = 0                            # FN
    c_pred[1, 0] += 0.2                         # FP1
    out["c_true"], out["c_pred"] = c_true.copy(), c_pred.copy()
    d = env["evaluate"](torch.from_numpy(idx), c_true=torch.from_numpy(c_true.copy()),
                        c_pred=torch.from_numpy(c_pred.copy()), output_gt=True)
    for k, v in d.items():
        out[f"co/{k}"] = np.asarray(v, dtype=np.float64)
    np.savez_compressed(f"{OUT}/evaluate.npz", **out)


def gen_input_pipeline():
    """EXECUTE the reference's frame preprocessing - `Video_IterableDataset.__process__` (dataset.py:783-812), lifted
    out of its class with ast and run against the real Pillow of this image - for bg_mode '' and 'concat', plus the
    median preparation of dataset.py:102-107. Small frames keep the fixture small; the resize arithmetic does not
    depend on the size (tests/test_oracle.py also pins the restatement against Pillow at 720p -> 288x512)."""
    from PIL import Image
    tree = ast.parse(open(f"{REF}/dataset.py").read())
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "__process__":
            fn = node
    env = {"np": np, "Image": Image}
    fn.name = "process"
    exec(compile(ast.Module(body=[fn], type_ignores=[]), f"{REF}/dataset.py", "exec"), env)
    rng = np.random.default_rng(31)
    out = {}
    for name, (h0, w0, H, W, L) in {"down": (90, 160, 36, 64, 3), "odd": (77, 123, 40, 56, 2), "up": (20, 30, 36, 64, 2)}.items():
        imgs = rng.integers(0, 256, (L, h0, w0, 3), dtype=np.uint8)
        imgs[0, 10:30, 20:50] = 255; imgs[1, :, ::2] = 0          # hard edges: over/undershoot gets clipped
        med_src = np.median(imgs, 0)
        med = np.moveaxis(np.array(Image.fromarray(med_src.astype('uint8')).resize(size=(W, H))), -1, 0)  # dataset.py:104-107

        class S:  # stand-in for the dataset object
            pass
        for bg in ("", "concat", "subtract", "subtract_concat"):
            s_ = S(); s_.bg_mode, s_.HEIGHT, s_.WIDTH, s_.seq_len = bg, H, W, L
            s_.median = med if bg == "concat" else med_src      # dataset.py:103-109: resized CHW only for 'concat'
            with np.errstate(invalid="ignore"):
                out[f"{name}/{bg or 'none'}"] = env["process"](s_, imgs)
        out[f"{name}/imgs"], out[f"{name}/median_src"], out[f"{name}/median"] = imgs, med_src, med
    np.savez_compressed(f"{OUT}/input_pipeline.npz", **out)


def gen_inpaint_mask():
    """EXECUTE the reference's `generate_inpaint_mask` (test.py:223-258) on random trajectories: visibility runs of
    random length (gaps at the start, in the middle, at the end, gaps of length 1, all visible, none visible) with
    y coordinates on both sides of the height threshold."""
    env = {"np": np}
    extract_functions(f"{REF}/test.py", {"generate_inpaint_mask"}, env)
    rng = np.random.default_rng(41)
    ys, vs, ms, ths = [], [], [], []
    for case in range(300):
        n = int(rng.integers(1, 40))
        vis = np.ones(n, dtype=np.int64)
        pos = 0
        while pos < n:
            run = int(rng.integers(1, 8))
            vis[pos:pos + run] = rng.integers(0, 2)
            pos += run
        if case % 17 == 0:
            vis[:] = case % 2
        y = rng.integers(0, 288, n) * vis
        th = float(rng.choice([30.0, 14.4, 100.0]))
        m = env["generate_inpaint_mask"]({"Y": y.tolist(), "Visibility": vis.tolist()}, th_h=th)
        pad = lambda a: np.pad(np.asarray(a), (0, 40 - n), constant_values=-1)
        ys.append(pad(y)); vs.append(pad(vis)); ms.append(pad(m)); ths.append(th)
    np.savez_compressed(f"{OUT}/inpaint_mask.npz", y=np.stack(ys), vis=np.stack(vs), mask=np.stack(ms), th=np.array(ths))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    refmodel = load_module("ref_model", f"{REF}/model.py")
    if len(sys.argv) > 1 and sys.argv[1] == "inpaint_train":  # regenerate only this fixture
        gen_inpaint_train(refmodel)
        print("inpaintnet_train.npz written")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "inpaint_mask":
        gen_inpaint_mask()
        print("inpaint_mask.npz written")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "input_pipeline":
        gen_input_pipeline()
        print("input_pipeline.npz written")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "evaluate":
        gen_evaluate()
        print("evaluate.npz written")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "temporal_ensemble":
        gen_temporal_ensemble()
        print("temporal_ensemble.npz written")
        return
    refmetric = load_module("ref_metric", f"{REF}/utils/metric.py")
    if len(sys.argv) > 1 and sys.argv[1] == "eval_step":
        gen_tracknet_eval_step(refmodel, refmetric)
        print("tracknet_eval_step.npz written")
        return
    import cv2
    env = {"np": np, "cv2": cv2, "torch": torch, "math": math}
    extract_functions(f"{REF}/test.py", {"predict_location", "get_ensemble_weight"}, env)
    extract_functions(f"{REF}/train.py", {"mixup"}, env)

    # ---- C1: TrackNet forward, seq_len 4, bg none, bs 1, 288x512 (BASELINE.json configs[0]) ----
    torch.manual_seed(13)
    m = refmodel.TrackNet(12, 4)
    x = torch.rand(1, 12, 288, 512)
    with torch.no_grad():
        m.train()
        y_train = m(x)
        m.eval()  # running stats now hold one momentum-0.1 update
        y_eval = m(x)
    np.savez_compressed(f"{OUT}/tracknet_c1.npz", seed=13, in_dim=12, out_dim=4, shape=np.array(x.shape),
                        y_train=y_train.numpy(), y_eval=y_eval.numpy(),
                        x_checksum=np.array([x.double().sum().item(), x[0, 3, 17, 99].item()]))

    # ---- small train step: seq_len 8, bg concat (in 27, out 8), bs 2, 32x64 ----
    torch.manual_seed(21)
    m = refmodel.TrackNet(27, 8)
    x = torch.rand(2, 27, 32, 64)
    y = (torch.rand(2, 8, 32, 64) > 0.97).float() * torch.rand(2, 8, 32, 64).clamp_min(0.3)
    m.train()
    y_pred = m(x)
    loss = refmetric.WBCELoss(y_pred, y)
    loss_ns = refmetric.WBCELoss(y_pred, y, reduce=False)
    loss.backward()
    names = [n for n, _ in m.named_parameters()]
    gstats = np.stack([grad_stats(p.grad) for _, p in m.named_parameters()])
    sd = m.state_dict()
    np.savez_compressed(f"{OUT}/tracknet_step.npz", seed=21, x=x.numpy(), y=y.numpy(), y_pred=y_pred.detach().numpy(),
                        loss=loss.item(), loss_per_sample=loss_ns.detach().numpy(), grad_stats=gstats,
                        grad_names=np.array(names),
                        grad_first=m.down_block_1.conv_1.conv.weight.grad.numpy(),
                        grad_last=m.up_block_3.conv_2.conv.weight.grad.numpy(),
                        grad_pred_w=m.predictor.weight.grad.numpy(), grad_pred_b=m.predictor.bias.grad.numpy(),
                        grad_bn_w=m.up_block_1.conv_1.bn.weight.grad.numpy(),
                        running_mean_first=sd["down_block_1.conv_1.bn.running_mean"].numpy(),
                        running_var_last=sd["up_block_3.conv_2.bn.running_var"].numpy(),
                        nbt=int(sd["bottleneck.conv_2.bn.num_batches_tracked"]))

    # ---- WBCE incl. clamp edges ----
    torch.manual_seed(3)
    p = torch.rand(3, 2, 16, 24)
    p.view(-1)[:6] = torch.tensor([0.0, 1.0, 1e-8, 1 - 1e-8, 1e-7, 0.5])
    t = (torch.rand(3, 2, 16, 24) > 0.9).float()
    t.view(-1)[:6] = torch.tensor([1.0, 0.0, 1.0, 0.0, 0.3, 0.7])
    p.requires_grad_(True)
    l1 = refmetric.WBCELoss(p, t)
    (g1,) = torch.autograd.grad(l1, p)
    l2 = refmetric.WBCELoss(p, t, reduce=False)
    wts = torch.tensor([1.0, -2.0, 0.5])
    (g2,) = torch.autograd.grad((l2 * wts).sum(), p)
    np.savez_compressed(f"{OUT}/wbce.npz", p=p.detach().numpy(), y=t.numpy(), loss=l1.item(), grad=g1.numpy(),
                        loss_ns=l2.detach().numpy(), gout_ns=wts.numpy(), grad_ns=g2.numpy())

    # ---- decode: crafted + random masks, real OpenCV through the reference's predict_location ----
    rng = np.random.default_rng(0)
    H, W = 48, 64
    masks = []
    def blank(): return np.zeros((H, W), np.uint8)
    masks.append(blank())                                            # empty
    a = blank(); a[10:13, 5:9] = 255; masks.append(a)               # one blob
    a = blank(); a[5:8, 5:8] = 255; a[20:23, 30:33] = 255; a[40:43, 10:13] = 255; masks.append(a)  # 3-way area tie
    a = blank(); a[5:8, 5:8] = 255; a[8, 8] = 255; masks.append(a)  # diagonal touch (8-connectivity)
    a = blank(); a[5:15, 5:15] = 255; a[7:13, 7:13] = 0; a[9:11, 9:11] = 255; masks.append(a)  # ring + nested dot
    a = blank(); a[0:2, 0:3] = 255; a[H - 2:H, W - 3:W] = 255; masks.append(a)  # image corners, tie
    a = blank(); a[3, 2:40] = 255; a[10:20, 50] = 255; masks.append(a)  # thin lines: 38x1 vs 1x10
    a = blank(); a[:, :] = 255; masks.append(a)                      # full
    a = blank(); a[::2, ::2] = 255; masks.append(a)                  # isolated pixels, massive tie
    a = blank(); a[::2, :] = 255; masks.append(a)                    # stripes (tie)
    for dens in (0.02, 0.1, 0.3, 0.5, 0.7):
        for _ in range(6):
            masks.append(((rng.random((H, W)) < dens) * 255).astype(np.uint8))
    for _ in range(10):                                              # few gaussian-ish blobs, realistic
        a = blank()
        for _ in range(rng.integers(1, 4)):
            cy, cx, r = rng.integers(0, H), rng.integers(0, W), rng.integers(1, 5)
            yy, xx = np.ogrid[:H, :W]
            a[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 255
        masks.append(a)
    masks = np.stack(masks)
    boxes = np.array([env["predict_location"](mm) for mm in masks], dtype=np.int32)
    np.savez_compressed(f"{OUT}/decode_golden.npz", masks=masks, boxes=boxes, cv2_version=cv2.__version__)

    # ---- InpaintNet forward ----
    torch.manual_seed(7)
    net = refmodel.InpaintNet().eval()
    coor = torch.rand(4, 16, 2)
    mask = (torch.rand(4, 16, 1) < 0.3).float()
    coor = coor * (1 - mask)
    with torch.no_grad():
        out = net(coor, mask)
    np.savez_compressed(f"{OUT}/inpaintnet.npz", seed=7, coor=coor.numpy(), mask=mask.numpy(), out=out.numpy())

    gen_tracknet_eval_step(refmodel, refmetric)
    gen_inpaint_train(refmodel)
    gen_temporal_ensemble()
    gen_evaluate()
    gen_input_pipeline()
    gen_inpaint_mask()

    # ---- small host-side pieces: ensemble weights, mixup ----
    ew = {f"weight_{L}": env["get_ensemble_weight"](L, "weight").numpy() for L in (1, 4, 7, 8)}
    ew["average_8"] = env["get_ensemble_weight"](8, "average").numpy()
    np.random.seed(11); torch.manual_seed(11)
    xm, ym = torch.rand(4, 3, 8, 8), torch.rand(4, 2, 8, 8)
    np.random.seed(11); torch.manual_seed(11)
    lamb = np.random.beta(0.5, 0.5, size=4); index = torch.randperm(4)  # the draws mixup() makes, in its order
    np.random.seed(11); torch.manual_seed(11)
    x_mix, y_mix = env["mixup"](xm, ym, 0.5)
    np.savez_compressed(f"{OUT}/host_pieces.npz", x=xm.numpy(), y=ym.numpy(), lamb=lamb, index=index.numpy(),
                        x_mix=x_mix.numpy(), y_mix=y_mix.numpy(), **ew)
    print("golden fixtures written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
