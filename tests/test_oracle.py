"""CPU tests: the oracle restatement reproduces the fixtures generated from the REAL reference
(oracle/gen_golden.py), i.e. the oracle is pinned. No GPU, no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle as D
from oracle import tracknet_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_tracknet_train_step_matches_reference(golden_dir):
    g = _load(golden_dir, "tracknet_step.npz")
    sd = O.init_tracknet_state(int(g["seed"]), 27, 8)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    y_pred, loss, grads = O.tracknet_loss_and_grads(sd, x, y, training=True)
    assert np.abs(y_pred.numpy() - g["y_pred"]).max() < 2e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-6 * max(1, abs(float(g["loss"])))
    names = [str(n) for n in g["grad_names"]]
    assert names == list(grads.keys())  # parameters() order == oracle key order
    for k, ref in (("down_block_1.conv_1.conv.weight", "grad_first"), ("up_block_3.conv_2.conv.weight", "grad_last"),
                   ("predictor.weight", "grad_pred_w"), ("predictor.bias", "grad_pred_b"),
                   ("up_block_1.conv_1.bn.weight", "grad_bn_w")):
        scale = np.abs(g[ref]).max()
        assert np.abs(grads[k].numpy() - g[ref]).max() <= 2e-4 * scale + 1e-10, k
    for i, k in enumerate(names):  # every one of the 53 gradients, via its statistics
        gs = g["grad_stats"][i]
        mine = grads[k].double().flatten()
        assert abs(mine.abs().sum().item() - gs[1]) <= 2e-3 * gs[1] + 1e-12, k
    assert np.abs(sd["down_block_1.conv_1.bn.running_mean"].numpy() - g["running_mean_first"]).max() < 1e-6
    assert np.abs(sd["up_block_3.conv_2.bn.running_var"].numpy() - g["running_var_last"]).max() < 1e-5
    assert int(sd["bottleneck.conv_2.bn.num_batches_tracked"]) == int(g["nbt"]) == 1


def test_tracknet_eval_mode_step_matches_reference(golden_dir):
    """The oracle's eval-mode step (BatchNorm frozen at its running statistics) against the REAL reference module run the
    same way (oracle/gen_golden.py eval_step): what tests/test_gpu_tracknet.py checks training = 2 against."""
    g = _load(golden_dir, "tracknet_eval_step.npz")
    sd = O.init_tracknet_state(int(g["seed"]), 27, 8)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    with torch.no_grad():
        for _ in range(5):
            O.tracknet_forward(sd, x, True)
    assert np.abs(sd["bottleneck.conv_1.bn.running_mean"].numpy() - g["running_mean_mid"]).max() < 1e-6
    assert np.abs(sd["up_block_3.conv_2.bn.running_var"].numpy() - g["running_var_last"]).max() < 1e-5
    frozen = {k: v.clone() for k, v in sd.items() if "running" in k or "tracked" in k}
    y_pred, loss, grads = O.tracknet_loss_and_grads(sd, x, y, training=False)
    assert all(torch.equal(sd[k], v) for k, v in frozen.items())       # an eval-mode step advances nothing
    assert int(sd["bottleneck.conv_2.bn.num_batches_tracked"]) == int(g["nbt"]) == 5
    assert np.abs(y_pred.numpy() - g["y_pred"]).max() < 2e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-6 * max(1, abs(float(g["loss"])))
    names = [str(n) for n in g["grad_names"]]
    assert names == list(grads.keys())
    for k, ref in (("down_block_1.conv_1.conv.weight", "grad_first"), ("up_block_3.conv_2.conv.weight", "grad_last"),
                   ("bottleneck.conv_2.bn.weight", "grad_bn_w"), ("down_block_2.conv_1.bn.bias", "grad_bn_b"),
                   ("predictor.weight", "grad_pred_w")):
        scale = np.abs(g[ref]).max()
        assert np.abs(grads[k].numpy() - g[ref]).max() <= 2e-4 * scale + 1e-10, k
    for i, k in enumerate(names):
        gs = g["grad_stats"][i]
        mine = grads[k].double().flatten()
        assert abs(mine.abs().sum().item() - gs[1]) <= 2e-3 * gs[1] + 1e-12, k


@pytest.mark.parametrize("h,w", [(32, 48)])
def test_tracknet_small_forward_is_deterministic(h, w):
    sd = O.init_tracknet_state(1, 12, 4)
    x = torch.rand(1, 12, h, w, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        a = O.tracknet_forward(dict(sd), x, False)
        b = O.tracknet_forward(dict(sd), x, False)
    assert a.shape == (1, 4, h, w) and torch.equal(a, b)


def test_tracknet_c1_full_size_matches_reference(golden_dir):
    """BASELINE.json configs[0]: seq_len 4, bg none, bs 1, 288x512 - train-mode and eval-mode heatmaps."""
    g = _load(golden_dir, "tracknet_c1.npz")
    seed = int(g["seed"])
    sd = O.init_tracknet_state(seed, 12, 4)
    gen = torch.Generator().manual_seed(seed)
    # gen_golden consumed the default generator for the weights first, then drew x: replay the same stream
    for key, shape in O.tracknet_state_keys(12, 4):
        if key.endswith("conv.weight") or key.startswith("predictor."):
            torch.rand(shape, generator=gen)
    x = torch.rand(tuple(g["shape"]), generator=gen)
    assert abs(x.double().sum().item() - g["x_checksum"][0]) < 1e-6 and x[0, 3, 17, 99].item() == g["x_checksum"][1]
    with torch.no_grad():
        y_train = O.tracknet_forward(sd, x, True)
        y_eval = O.tracknet_forward(sd, x, False)
    assert np.abs(y_train.numpy() - g["y_train"]).max() < 5e-5
    assert np.abs(y_eval.numpy() - g["y_eval"]).max() < 5e-5


def test_wbce_matches_reference(golden_dir):
    g = _load(golden_dir, "wbce.npz")
    p = torch.from_numpy(g["p"]).requires_grad_(True)
    y = torch.from_numpy(g["y"])
    loss = O.wbce_loss(p, y)
    (grad,) = torch.autograd.grad(loss, p)
    assert abs(loss.item() - float(g["loss"])) < 1e-7
    assert np.abs(grad.numpy() - g["grad"]).max() < 1e-9
    assert np.abs(O.wbce_loss(p, y, reduce=False).detach().numpy() - g["loss_ns"]).max() < 1e-7


def test_inpaintnet_matches_reference(golden_dir):
    g = _load(golden_dir, "inpaintnet.npz")
    sd = O.init_inpaintnet_state(int(g["seed"]))
    out = O.inpaintnet_forward(sd, torch.from_numpy(g["coor"]), torch.from_numpy(g["mask"]))
    assert np.abs(out.numpy() - g["out"]).max() < 1e-6


def test_host_pieces_match_reference(golden_dir):
    g = _load(golden_dir, "host_pieces.npz")
    for L in (1, 4, 7, 8):
        assert np.array_equal(O.get_ensemble_weight(L, "weight").numpy(), g[f"weight_{L}"])
    assert np.array_equal(O.get_ensemble_weight(8, "average").numpy(), g["average_8"])
    xm, ym = O.mixup(torch.from_numpy(g["x"]), torch.from_numpy(g["y"]), g["lamb"], torch.from_numpy(g["index"]))
    assert np.abs(xm.numpy() - g["x_mix"]).max() < 1e-7 and np.abs(ym.numpy() - g["y_mix"]).max() < 1e-7


def test_label_disc_shape_and_area():
    d = O.label_disc(100, 50)
    assert d.shape == (288, 512) and d.sum() == 21 and d[50, 100] == 1  # radius-2.5 disc has 21 pixels
    assert O.label_disc(0, 0).sum() == 0


def test_decode_oracle_matches_cv2_fixtures(golden_dir):
    g = _load(golden_dir, "decode_golden.npz")
    for i, (m, box) in enumerate(zip(g["masks"], g["boxes"])):
        assert tuple(D.predict_location(m)) == tuple(int(v) for v in box), f"mask {i}"
    assert D.center((3, 4, 5, 6)) == (5, 7)


def test_decode_oracle_matches_live_cv2():
    """When OpenCV is importable (it is in this image) pin the restatement on fresh random masks as well."""
    cv2 = pytest.importorskip("cv2")
    from hypothesis import given, settings, strategies as st
    import hypothesis.extra.numpy as hnp

    def cv2_ref(hm):  # the reference's rule (test.py:59-77), stated on cv2 primitives
        if hm.max() == 0:
            return (0, 0, 0, 0)
        cnts, _ = cv2.findContours(hm.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        rects = [cv2.boundingRect(c) for c in cnts]
        best = rects[0]
        for r in rects[1:]:
            if r[2] * r[3] > best[2] * best[3]:
                best = r
        return tuple(best)

    @settings(max_examples=150, deadline=None)
    @given(hnp.arrays(np.uint8, st.tuples(st.integers(1, 24), st.integers(1, 24)), elements=st.sampled_from([0, 0, 255])))
    def check(mask):
        assert tuple(D.predict_location(mask)) == cv2_ref(mask)

    check()


def test_inpaintnet_train_step_matches_reference(golden_dir):
    """oracle restatement of train.py:147-166 vs the real reference module's step (oracle/gen_golden.py)."""
    g = _load(golden_dir, "inpaintnet_train.npz")
    sd = O.init_inpaintnet_state(int(g["seed"]))
    t = lambda k: torch.from_numpy(g[k])
    refine, loss, grads = O.inpaintnet_loss_and_grads(sd, t("coor_pred"), t("coor_gt"), t("vis_gt"), t("mask"))
    assert (refine - t("refine")).abs().max() < 1e-6
    assert abs(loss.item() - float(g["loss"])) < 1e-7
    for k, gr in grads.items():
        ref = t("grad/" + k)
        assert (gr - ref).abs().max() <= 1e-5 * ref.abs().max() + 1e-9, k
    total, clipped = O.clip_grad_norm(grads, 1.0)
    assert abs(total.item() - float(g["total_norm"])) < 1e-5 * float(g["total_norm"])
    for k, gr in clipped.items():
        st = g["clipped/" + k]
        assert abs(gr.double().sum().item() - st[0]) <= 1e-5 * st[1] + 1e-9, k


def test_temporal_ensemble_matches_reference_loops(golden_dir):
    """oracle TemporalEnsemble vs the reference's own streaming loops executed by oracle/gen_golden.py."""
    g = _load(golden_dir, "temporal_ensemble.npz")
    keys = sorted({k.split("/")[0] for k in g.files})
    assert len(keys) == 8
    for key in keys:
        kind, mode, L, n, bs = key.split("_")
        L, n, bs = int(L), int(n), int(bs)
        if kind == "hm":
            preds = torch.from_numpy(g[key + "/preds"])
            num_sample = preds.shape[0]
        else:
            coor, pred_in, msk = (torch.from_numpy(g[key + "/" + k]) for k in ("coor", "pred_in", "mask"))
            preds = O.inpaint_blend(pred_in, coor, msk)   # predict.py:257-261 happens before buffering
            num_sample = n
        ens = O.TemporalEnsemble(L, mode, num_sample)
        outs = [ens.push(preds[a:a + bs]) for a in range(0, num_sample, bs)]
        assert [len(o) for o in outs] == list(g[key + "/counts"])
        out = torch.cat(outs)
        ref = torch.from_numpy(g[key + "/ens"]).reshape(out.shape)
        if kind == "co":  # predict.py:289-291: the ensembled coordinates are thresholded again
            th = (out[:, 0] < O.COOR_TH) & (out[:, 1] < O.COOR_TH)
            out[th] = 0.0
        assert torch.equal(out, ref), key


def test_evaluate_matches_reference(golden_dir):
    """oracle evaluate() vs the reference's own evaluate executed by oracle/gen_golden.py (all five outcome types,
    padded duplicate frames, bbox / confidence / ground-truth outputs, image scaling)."""
    from oracle import decode_oracle as D
    g = _load(golden_dir, "evaluate.npz")
    cases = {"hm_plain": {}, "hm_bbox_gt": {"output_bbox": True, "output_gt": True},
             "hm_scaled": {"img_scaler": (2.5, 2.5), "tolerance": 1.0, "output_gt": True}}
    for name, kw in cases.items():
        d = D.evaluate(g["indices"], y_true=g["y_true"], y_pred=g["y_pred"], **kw)
        keys = sorted(k.split("/")[1] for k in g.files if k.startswith(name + "/"))
        assert sorted(d.keys()) == keys
        for k in keys:
            assert np.array_equal(np.asarray(d[k], dtype=np.float64), g[f"{name}/{k}"]), (name, k)
    d = D.evaluate(g["indices"], c_true=g["c_true"], c_pred=g["c_pred"], output_gt=True)
    for k in d:
        assert np.array_equal(np.asarray(d[k], dtype=np.float64), g[f"co/{k}"]), k
    assert set(g["hm_plain/Type"].tolist()) == {0.0, 1.0, 2.0, 3.0, 4.0}


def test_input_pipeline_matches_reference_and_pillow(golden_dir):
    """oracle restatement of Pillow's bicubic resize + the reference's frame stacking vs (i) the reference's own
    `__process__` executed with real Pillow by oracle/gen_golden.py, (ii) real Pillow itself when importable."""
    from oracle import pil_resize_oracle as P
    g = _load(golden_dir, "input_pipeline.npz")
    for name in ("down", "odd", "up"):
        imgs, med = g[f"{name}/imgs"], g[f"{name}/median"]
        H, W = med.shape[1:]
        assert np.array_equal(P.process_frames(imgs, None, W, H), g[f"{name}/none"])
        assert np.array_equal(P.process_frames(imgs, med, W, H), g[f"{name}/concat"])
        for bg in ("subtract", "subtract_concat"):
            got = P.process_frames(imgs, None, W, H, bg_mode=bg, median_src=g[f"{name}/median_src"])
            assert np.array_equal(got, g[f"{name}/{bg}"]), (name, bg)
        med2 = np.moveaxis(P.resize_bicubic_u8(g[f"{name}/median_src"].astype("uint8"), W, H), -1, 0)
        assert np.array_equal(med2, med)
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    for h, w, oh, ow in ((720, 1280, 288, 512), (97, 131, 40, 64), (50, 70, 100, 90)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(P.resize_bicubic_u8(img, ow, oh), np.array(Image.fromarray(img).resize(size=(ow, oh))))
