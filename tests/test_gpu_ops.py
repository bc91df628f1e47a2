"""-m gpu: the HBM-bound kernels (BN, predictor, loss, mixup, Adam, decode, InpaintNet) through the C ABI,
against the CPU oracle and the fixtures generated from the real reference."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import tracknetv3_b200 as T
from tracknetv3_b200 import _lib
from oracle import decode_oracle as D
from oracle import tracknet_oracle as O
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_wbce_forward_backward_vs_reference_fixture(golden_dir):
    g = _load(golden_dir, "wbce.npz")
    p = torch.from_numpy(g["p"]).to(G.DEV).requires_grad_(True)
    y = torch.from_numpy(g["y"]).to(G.DEV)
    loss = T.WBCELoss(p, y)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    assert np.abs(p.grad.cpu().numpy() - g["grad"]).max() < 1e-8 + 1e-5 * np.abs(g["grad"]).max()
    p2 = torch.from_numpy(g["p"]).to(G.DEV).requires_grad_(True)
    l2 = T.WBCELoss(p2, y, reduce=False)
    assert l2.shape == (3,) and np.abs(l2.detach().cpu().numpy() - g["loss_ns"]).max() < 1e-6
    (l2 * torch.from_numpy(g["gout_ns"]).to(G.DEV)).sum().backward()
    assert np.abs(p2.grad.cpu().numpy() - g["grad_ns"]).max() < 1e-8 + 1e-5 * np.abs(g["grad_ns"]).max()


def test_wbce_full_size_vs_oracle():
    gen = torch.Generator().manual_seed(0)
    p = torch.rand(2, 8, 288, 512, generator=gen)
    y = (torch.rand(2, 8, 288, 512, generator=gen) > 0.999).float()
    ref = O.wbce_loss(p, y).item()
    got = T.WBCELoss(p.to(G.DEV), y.to(G.DEV)).item()
    assert abs(got - ref) < 2e-6 * abs(ref)


def test_mixup_vs_reference_fixture(golden_dir):
    g = _load(golden_dir, "host_pieces.npz")
    L = G.lib()
    lam = torch.from_numpy(np.maximum(g["lamb"], 1 - g["lamb"])).float().to(G.DEV)
    idx = torch.from_numpy(g["index"]).to(G.DEV)
    for src, ref in (("x", "x_mix"), ("y", "y_mix")):
        x = torch.from_numpy(g[src]).to(G.DEV)
        out = torch.empty_like(x)
        _lib.check(L.tnb_mixup(x.data_ptr(), lam.data_ptr(), idx.data_ptr(), out.data_ptr(), x.shape[0],
                               x[0].numel(), G.st()))
        assert np.abs(out.cpu().numpy() - g[ref]).max() < 1e-6


def test_fused_adam_matches_torch_adam():
    torch.manual_seed(0)
    shapes = [(64, 27, 3, 3), (64,), (512, 512, 3, 3), (8, 64, 1, 1), (8,)]
    ref_p = [torch.randn(s).requires_grad_(True) for s in shapes]
    my_p = [p.detach().clone().to(G.DEV).requires_grad_(True) for p in ref_p]
    ref_opt = torch.optim.Adam(ref_p, lr=1e-3)
    my_opt = T.FusedAdam(my_p, lr=1e-3)
    for step in range(4):
        for rp, mp in zip(ref_p, my_p):
            g = torch.randn(rp.shape, generator=torch.Generator().manual_seed(step * 10 + rp.numel() % 7)) * 10 ** (-step)
            rp.grad = g.clone()
            mp.grad = g.to(G.DEV)
        ref_opt.step(); my_opt.step()
    for rp, mp in zip(ref_p, my_p):
        assert G.max_abs(mp, rp) < 2e-6



def test_fused_adam_checkpoint_interchanges_with_torch_adam():
    """optimizer.state_dict() round trip in both directions (reference train.py:258-259 saves it, :286-301 resumes):
    two steps with one optimizer, load its state into the other kind, two more steps - same parameters as four steps of
    torch.optim.Adam."""
    shapes = [(64, 27, 3, 3), (64,), (8, 64, 1, 1)]

    def grads(step):
        return [torch.randn(s, generator=torch.Generator().manual_seed(100 * step + i)) * 0.1 for i, s in enumerate(shapes)]

    torch.manual_seed(1)
    init = [torch.randn(s) for s in shapes]
    ref_p = [p.clone().requires_grad_(True) for p in init]
    ref = torch.optim.Adam(ref_p, lr=1e-3)
    for step in range(4):
        for p, g in zip(ref_p, grads(step)):
            p.grad = g
        ref.step()
    for first_fused in (True, False):
        pa = [p.clone().to(G.DEV).requires_grad_(True) for p in init]
        make = lambda fused, ps: T.FusedAdam(ps, lr=1e-3) if fused else torch.optim.Adam(ps, lr=1e-3)
        opt = make(first_fused, pa)
        for step in range(2):
            for p, g in zip(pa, grads(step)):
                p.grad = g.to(G.DEV)
            opt.step()
        ckpt = opt.state_dict()
        pb = [p.detach().clone().requires_grad_(True) for p in pa]
        opt2 = make(not first_fused, pb)
        opt2.load_state_dict(ckpt)
        for step in range(2, 4):
            for p, g in zip(pb, grads(step)):
                p.grad = g.to(G.DEV)
            opt2.step()
        for rp, mp in zip(ref_p, pb):
            assert G.max_abs(mp, rp) < 2e-6, first_fused

def test_decode_vs_cv2_fixture_and_oracle(golden_dir):
    g = _load(golden_dir, "decode_golden.npz")
    masks = torch.from_numpy(g["masks"]).to(G.DEV)
    boxes = T.decode_heatmaps(masks).cpu().numpy()
    assert np.array_equal(boxes, g["boxes"])  # bit-exact vs real OpenCV through the reference's rule
    assert T.predict_location(g["masks"][2]) == tuple(int(v) for v in g["boxes"][2])
    # float heatmaps, threshold strictly > 0.5 (predict.py:35)
    hm = torch.rand(3, 4, 40, 56, generator=torch.Generator().manual_seed(1))
    hm[0, 0] = 0.5
    got = T.decode_heatmaps(hm.to(G.DEV)).cpu().numpy()
    assert np.array_equal(got, D.decode_batch(hm.numpy()))
    assert tuple(got[0, 0]) == (0, 0, 0, 0)


def test_decode_full_size_planted_blobs():
    """288x512 maps (C4 size): planted boxes incl. an area tie resolved towards the bottom-most start pixel."""
    hm = torch.zeros(4, 288, 512)
    hm[0, 100:105, 200:207] = 0.9                    # single 7x5
    hm[1, 10:14, 10:14] = 0.8; hm[1, 250:254, 400:404] = 0.7   # tie 4x4: bottom one wins
    hm[2, 287, 511] = 1.0                            # last pixel
    got = T.decode_heatmaps(hm.to(G.DEV)).cpu().numpy().tolist()
    assert got == [[200, 100, 7, 5], [400, 250, 4, 4], [511, 287, 1, 1], [0, 0, 0, 0]]
    dense = torch.rand(2, 288, 512, generator=torch.Generator().manual_seed(2))
    assert np.array_equal(T.decode_heatmaps(dense.to(G.DEV)).cpu().numpy(), D.decode_batch(dense.numpy()))


def test_inpaintnet_vs_reference_fixture_and_oracle(golden_dir):
    g = _load(golden_dir, "inpaintnet.npz")
    torch.manual_seed(int(g["seed"]))
    net = T.InpaintNet().to(G.DEV).eval()
    with torch.no_grad():
        out = net(torch.from_numpy(g["coor"]).to(G.DEV), torch.from_numpy(g["mask"]).to(G.DEV))
    assert np.abs(out.cpu().numpy() - g["out"]).max() < 1e-5
    sd = O.init_inpaintnet_state(3)
    net.load_state_dict(sd)
    x = torch.rand(32, 16, 2, generator=torch.Generator().manual_seed(4))
    m = (torch.rand(32, 16, 1, generator=torch.Generator().manual_seed(5)) < 0.3).float()
    with torch.no_grad():
        out = net((x * (1 - m)).to(G.DEV), m.to(G.DEV))
    assert G.max_abs(out, O.inpaintnet_forward(sd, x * (1 - m), m)) < 1e-5
    # the kernels read raw fp32 pointers: a converted model must be refused, not misread
    with pytest.raises(RuntimeError, match="contiguous fp32"):
        net.half()(x.to(G.DEV), m.to(G.DEV))
    with pytest.raises(RuntimeError, match=r"\(N, L, 2\)"):
        net.float()(x[:, :, :1].to(G.DEV), m.to(G.DEV))


def test_inpaintnet_train_step_vs_reference_fixture(golden_dir):
    """train.py:147-166 on the CUDA path vs the step run on the real reference module (oracle/gen_golden.py):
    output, loss, all 18 gradients, clip_grad_norm_ total norm and the Adam-updated parameters."""
    import train as TR
    g = _load(golden_dir, "inpaintnet_train.npz")
    t = lambda k: torch.from_numpy(g[k]).to(G.DEV)
    net = T.InpaintNet().to(G.DEV).train()
    net.load_state_dict(O.init_inpaintnet_state(int(g["seed"])))
    opt = T.FusedAdam(net.parameters(), lr=1e-3)
    opt.zero_grad()
    inpaint_mask = torch.logical_and(t("vis_gt"), t("mask")).int()
    refine = net(t("coor_pred") * (1 - inpaint_mask), inpaint_mask)
    loss = torch.nn.MSELoss()(refine * inpaint_mask, t("coor_gt") * inpaint_mask)
    loss.backward()
    assert G.max_abs(refine.detach(), torch.from_numpy(g["refine"])) < 1e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    for k, p in net.named_parameters():
        ref = torch.from_numpy(g["grad/" + k])
        err = (p.grad.cpu() - ref).abs().max().item()
        assert err <= 2e-5 * ref.abs().max().item() + 1e-9, (k, err)
    total = torch.nn.utils.clip_grad_norm_(net.parameters(), 1)
    assert abs(total.item() - float(g["total_norm"])) < 1e-5 * float(g["total_norm"])
    opt.step()
    for k, v in net.state_dict().items():
        st = g["after/" + k]
        assert abs(v.double().sum().item() - st[0]) <= 1e-5 * st[1] + 1e-7, k
    # the drop-in loop itself (random mask drawn inside, like the reference): loss goes down on a fixed batch
    np.random.seed(0)
    batch = (None, t("coor_pred").cpu(), t("coor_gt").cpu(), None, t("vis_gt").cpu(), None)
    losses = [TR.train_inpaintnet(net, opt, [batch] * 20, {"mask_ratio": 0.3, "verbose": False}) for _ in range(3)]
    assert losses[-1] < losses[0]


def test_inpaintnet_backward_vs_oracle_shapes():
    """Backward kernel on other (N, L) including L = 1, odd L and the input-coordinate gradient."""
    for n, l, seed in ((1, 1, 0), (3, 7, 1), (32, 16, 2), (5, 28, 3)):
        sd = O.init_inpaintnet_state(seed)
        gen = torch.Generator().manual_seed(seed)
        x = torch.rand(n, l, 2, generator=gen)
        m = (torch.rand(n, l, 1, generator=gen) < 0.4).float()
        w = torch.rand(n, l, 2, generator=gen)
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xr = x.clone().requires_grad_(True)
        ref = torch.autograd.grad((O.inpaintnet_forward(params, xr, m) * w).sum(), list(params.values()) + [xr])
        net = T.InpaintNet().to(G.DEV).train()
        net.load_state_dict(sd)
        xd = x.to(G.DEV).requires_grad_(True)
        (net(xd, m.to(G.DEV)) * w.to(G.DEV)).sum().backward()
        for (k, p), r in zip(net.named_parameters(), ref[:-1]):
            assert (p.grad.cpu() - r).abs().max() <= 2e-5 * r.abs().max() + 1e-8, (n, l, k)
        assert (xd.grad.cpu() - ref[-1]).abs().max() <= 2e-5 * ref[-1].abs().max() + 1e-8
    with pytest.raises(RuntimeError):
        net = T.InpaintNet().to(G.DEV).train()
        net(torch.rand(1, 40, 2, device=G.DEV), torch.ones(1, 40, 1, device=G.DEV)).sum().backward()


def test_bn_finalize_and_predictor():
    L = G.lib()
    gen = torch.Generator().manual_seed(6)
    n, h, w, c, o = 2, 8, 12, 64, 8
    z = torch.randn(n, c, h, w, generator=gen) * 2 + 0.5
    gamma, beta = torch.rand(c, generator=gen) + 0.5, torch.randn(c, generator=gen) * 0.1
    rm, rv = torch.randn(c, generator=gen), torch.rand(c, generator=gen) + 0.5
    part = torch.stack([z.sum((0, 2, 3)), (z * z).sum((0, 2, 3))])[None].contiguous().to(G.DEV)  # one "tile"
    bn = torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta); bn.running_mean.copy_(rm); bn.running_var.copy_(rv)
    a_ref = F.relu(bn(z))
    dev = [t.clone().to(G.DEV) for t in (gamma, beta, rm, rv)]
    outs = [torch.empty(c, device=G.DEV) for _ in range(4)]
    _lib.check(L.tnb_bn_finalize(part.data_ptr(), 1, float(n * h * w), dev[0].data_ptr(), dev[1].data_ptr(),
                                 dev[2].data_ptr(), dev[3].data_ptr(), 0.1, 1e-5, 1, outs[0].data_ptr(),
                                 outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), c, G.st()))
    assert G.max_abs(dev[2], bn.running_mean) < 1e-5 and G.max_abs(dev[3], bn.running_var) < 1e-4
    a = F.relu(z * outs[0].cpu()[None, :, None, None] + outs[1].cpu()[None, :, None, None])
    assert G.max_abs(a, a_ref) < 2e-5
    # predictor forward / backward on relu(bn(z))
    wp = (torch.rand(o, c, 1, 1, generator=gen) - 0.5).requires_grad_(True)
    bp = (torch.rand(o, generator=gen) - 0.5).requires_grad_(True)
    a_leaf = a_ref.detach().clone().requires_grad_(True)
    y_ref = torch.sigmoid(F.conv2d(a_leaf, wp, bp))
    dy = torch.randn(n, o, h, w, generator=gen)
    y_ref.backward(dy)
    zt = G.nhwc(z)
    src = G.make_src(zt, _lib.SRC_AFFINE_RELU, outs[0], outs[1])
    y = torch.empty(n, o, h, w, device=G.DEV)
    wd, bd = wp.detach().to(G.DEV), bp.detach().to(G.DEV)
    _lib.check(L.tnb_conv1x1_bias_sigmoid_fwd(C.byref(src), n, h, w, wd.data_ptr(), bd.data_ptr(), o, y.data_ptr(), G.st()))
    assert G.max_abs(y, y_ref) < 1e-5
    dA = torch.empty(n, h, w, c, device=G.DEV); dwp = torch.empty(o, c, device=G.DEV); dbp = torch.empty(o, device=G.DEV)
    dyd = dy.to(G.DEV)
    G.predictor_bwd(src, n, h, w, wd, o, dyd, y, dA, dwp, dbp)
    assert G.rel_err(G.nchw(dA), a_leaf.grad) < 1e-4
    assert G.rel_err(dwp, wp.grad.reshape(o, c)) < 1e-4 and G.rel_err(dbp, bp.grad) < 1e-4


@pytest.mark.parametrize("n,h,w,o,affine", [(2, 8, 12, 8, True), (1, 7, 9, 4, True), (3, 5, 11, 12, False), (1, 40, 72, 8, True),
                                            (2, 9, 13, 20, True), (1, 16, 24, 37, False)])
def test_predictor_backward_shapes(n, h, w, o, affine):
    """sigmoid + 1x1 conv backward (two streaming kernels): pixel counts that are not multiples of the 32-pixel warp
    chunk, out_dim 4 / 8 / 12 (both template instances) and 20 / 37 (several groups of 16 output channels: the
    reference takes any seq_len), BN+ReLU and identity activation sources."""
    L = G.lib()
    gen = torch.Generator().manual_seed(60 + o)
    z = torch.randn(n, 64, h, w, generator=gen)
    sc, sh = torch.rand(64, generator=gen) + 0.5, torch.randn(64, generator=gen) * 0.3
    a = F.relu(z * sc[None, :, None, None] + sh[None, :, None, None]) if affine else z
    wp = (torch.rand(o, 64, 1, 1, generator=gen) - 0.5).requires_grad_(True)
    bp = (torch.rand(o, generator=gen) - 0.5).requires_grad_(True)
    a_leaf = a.detach().clone().requires_grad_(True)
    y_ref = torch.sigmoid(F.conv2d(a_leaf, wp, bp))
    dy = torch.randn(n, o, h, w, generator=gen)
    y_ref.backward(dy)
    zt, scd, shd = G.nhwc(z), sc.to(G.DEV), sh.to(G.DEV)
    src = G.make_src(zt, _lib.SRC_AFFINE_RELU, scd, shd) if affine else G.make_src(zt)
    yd, dyd, wd = y_ref.detach().to(G.DEV).contiguous(), dy.to(G.DEV), wp.detach().to(G.DEV)
    yf = torch.full((n, o, h, w), float("nan"), device=G.DEV)
    bd = bp.detach().to(G.DEV)
    _lib.check(L.tnb_conv1x1_bias_sigmoid_fwd(C.byref(src), n, h, w, wd.data_ptr(), bd.data_ptr(), o, yf.data_ptr(), G.st()))
    assert G.max_abs(yf, y_ref) < 1e-5
    dA = torch.full((n, h, w, 64), float("nan"), device=G.DEV)
    dwp = torch.full((o, 64), float("nan"), device=G.DEV); dbp = torch.full((o,), float("nan"), device=G.DEV)
    G.predictor_bwd(src, n, h, w, wd, o, dyd, yd, dA, dwp, dbp)
    dw1, db1 = dwp.clone(), dbp.clone()
    G.predictor_bwd(src, n, h, w, wd, o, dyd, yd, dA, dwp, dbp)
    assert torch.equal(dw1, dwp) and torch.equal(db1, dbp)  # block partials summed in block order: no atomics
    assert G.rel_err(G.nchw(dA), a_leaf.grad) < 1e-4
    assert G.rel_err(dwp, wp.grad.reshape(o, 64)) < 1e-4 and G.rel_err(dbp, bp.grad) < 1e-4


@pytest.mark.parametrize("consumers", ["same", "pool+skip", "up"])
def test_bn_relu_backward_with_gradient_routing(consumers):
    """BN+ReLU backward fused with MaxPool / Upsample / cat gradient routing vs torch autograd (CPU fp32)."""
    L = G.lib()
    gen = torch.Generator().manual_seed(7)
    n, h, w, c = 2, 8, 12, 64
    z = (torch.randn(n, c, h, w, generator=gen)).requires_grad_(True)
    gamma = (torch.rand(c, generator=gen) + 0.5) * torch.where(torch.arange(c) % 5 == 0, -1.0, 1.0)
    beta = torch.randn(c, generator=gen) * 0.2
    a = F.relu(F.batch_norm(z, None, None, gamma, beta, True, 0.1, 1e-5))
    srcs, keep = [], []
    if consumers == "same":
        d0 = torch.randn(n, c, h, w, generator=gen)
        a.backward(d0)
        t = G.nhwc(d0); keep.append(t)
        srcs.append(_lib.GradSrc(t.data_ptr(), c, 0, _lib.GRAD_SAME, h, w))
    elif consumers == "pool+skip":
        dpool = torch.randn(n, c, h // 2, w // 2, generator=gen)
        dskip = torch.randn(n, 32 + c, h, w, generator=gen)  # skip occupies channels [32, 32+c) of a concat
        (F.max_pool2d(a, 2, 2) * dpool).sum().backward(retain_graph=True)
        (a * dskip[:, 32:]).sum().backward()
        t0, t1 = G.nhwc(dpool), G.nhwc(dskip); keep += [t0, t1]
        srcs.append(_lib.GradSrc(t0.data_ptr(), c, 0, _lib.GRAD_POOL, h // 2, w // 2))
        srcs.append(_lib.GradSrc(t1.data_ptr(), 32 + c, 32, _lib.GRAD_SAME, h, w))
    else:
        dup = torch.randn(n, c + 16, 2 * h, 2 * w, generator=gen)  # upsampled part = channels [0, c)
        (F.interpolate(a, scale_factor=2, mode="nearest") * dup[:, :c]).sum().backward()
        t0 = G.nhwc(dup); keep.append(t0)
        srcs.append(_lib.GradSrc(t0.data_ptr(), c + 16, 0, _lib.GRAD_UP, 2 * h, 2 * w))
    with torch.no_grad():
        mean = z.mean((0, 2, 3)); var = z.var((0, 2, 3), unbiased=False)
        invstd = 1 / torch.sqrt(var + 1e-5)
        scale = gamma * invstd; shift = beta - mean * scale
    dev = {k: v.contiguous().to(G.DEV) for k, v in dict(scale=scale, shift=shift, mean=mean, invstd=invstd).items()}
    zt = G.nhwc(z.detach())
    rows = L.tnb_bn_bwd_blocks(n, h, w, c)
    part = torch.zeros(rows, 2, c, device=G.DEV); sums = torch.zeros(2, c, device=G.DEV)
    dz = torch.full((n, h, w, c), float("nan"), device=G.DEV)
    dgamma = torch.empty(c, device=G.DEV); dbeta = torch.empty(c, device=G.DEV)
    args = _lib.BnBwd()
    for i, s in enumerate(srcs):
        args.g[i] = s
    args.ng = len(srcs)
    args.z = zt.data_ptr()
    args.scale, args.shift, args.mean, args.invstd = (dev[k].data_ptr() for k in ("scale", "shift", "mean", "invstd"))
    args.N, args.H, args.W, args.C = n, h, w, c
    args.part, args.sums, args.dz = part.data_ptr(), sums.data_ptr(), dz.data_ptr()
    args.inv_count = 1.0 / (n * h * w)
    amax = torch.zeros(1, device=G.DEV)
    args.amax = amax.data_ptr()
    args.dz_format = 0
    _lib.check(L.tnb_bn_relu_bwd_reduce(C.byref(args), G.st()))
    _lib.check(L.tnb_bn_relu_bwd_finalize(part.data_ptr(), rows, c, sums.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), G.st()))
    _lib.check(L.tnb_bn_relu_bwd_apply(C.byref(args), G.st()))
    torch.cuda.synchronize()
    assert G.rel_err(G.nchw(dz), z.grad) < 2e-4
    assert amax.item() == dz.abs().max().item()
    # production formats: the same dz emitted pre-split for the dgrad / wgrad operand fills - bf16 (hi, lo) pairs
    # (dz_format 1), or fp16 pairs of dz * 2^k with k derived from max |g|, which the reduction pass measures (dz_format 2)
    # ... and, in the same pass, the layer's own activation relu(bn(z)) pre-split: the next layer's wgrad operand
    act_ref = F.relu(z.detach() * scale[None, :, None, None] + shift[None, :, None, None])
    for dz_format in (1, 2):
        fmt = 1 if dz_format == 1 else 0
        dzs = torch.zeros(n * h * w * c * 4, dtype=torch.uint8, device=G.DEV)
        acts = torch.zeros(n * h * w * c * 4, dtype=torch.uint8, device=G.DEV)
        gm = torch.zeros(2, device=G.DEV)  # [max |g|, multiplier]
        args.dz, args.dz_format, args.amax = dzs.data_ptr(), dz_format, None
        args.act_presplit = acts.data_ptr()
        mul = 1.0
        if dz_format == 2:
            args.gmax, args.dz_mul = gm.data_ptr(), gm.data_ptr() + 4
            _lib.check(L.tnb_bn_relu_bwd_reduce(C.byref(args), G.st()))  # measures max |g| (partials as before)
        _lib.check(L.tnb_bn_relu_bwd_apply(C.byref(args), G.st()))
        torch.cuda.synchronize()
        if dz_format == 2:
            gmax, mul = gm.tolist()
            assert gmax > 0 and mul == 2.0 ** round(np.log2(mul))  # a power of two ...
            assert 2.0 ** 7 <= mul * gmax * scale.abs().max().item() < 2.0 ** 8  # ... chosen by the documented rule
            assert dz.abs().max().item() * mul < 65504  # nothing was clamped
        assert G.max_abs(G.unsplit(dzs, (n, h, w, c), fmt, mul), dz) <= G.split_tol(fmt) * dz.abs().max().item()
        assert G.max_abs(G.unsplit(acts, (n, h, w, c), fmt), G.nhwc(act_ref)) <= G.split_tol(fmt) * act_ref.abs().max().item()
    # ... or its 2x2 max-pool at half resolution (act_pool): the operand of a next layer that reads this one through MaxPool2d
    pooled = torch.zeros(n * (h // 2) * (w // 2) * c * 4, dtype=torch.uint8, device=G.DEV)
    args.dz_format, args.act_presplit, args.act_pool = 1, pooled.data_ptr(), 1
    _lib.check(L.tnb_bn_relu_bwd_apply(C.byref(args), G.st()))
    torch.cuda.synchronize()
    pool_ref = F.max_pool2d(act_ref, 2, 2)
    assert G.max_abs(G.unsplit(pooled, (n, h // 2, w // 2, c), 1), G.nhwc(pool_ref)) <= G.split_tol(1) * pool_ref.abs().max().item()
    args.act_pool, args.act_presplit = 0, None
    # d gamma / d beta from the same reductions (checked through a second autograd pass)
    z2 = z.detach().clone(); g2 = gamma.clone().requires_grad_(True); b2 = beta.clone().requires_grad_(True)
    a2 = F.relu(F.batch_norm(z2, None, None, g2, b2, True, 0.1, 1e-5))
    if consumers == "same":
        a2.backward(d0)
    elif consumers == "pool+skip":
        ((F.max_pool2d(a2, 2, 2) * dpool).sum() + (a2 * dskip[:, 32:]).sum()).backward()
    else:
        (F.interpolate(a2, scale_factor=2, mode="nearest") * dup[:, :c]).sum().backward()
    assert G.rel_err(dgamma, g2.grad) < 2e-4 and G.rel_err(dbeta, b2.grad) < 2e-4
    del keep


def test_device_prefetcher_yields_batches_in_order():
    """Batches arrive in order and intact while the consumer keeps the GPU busy (slots are reused every 2nd batch)."""
    host = [(torch.full((256, 1024), float(i)).pin_memory(), torch.arange(4).float().pin_memory() + i) for i in range(7)]
    host.append((torch.full((3, 5), 7.0).pin_memory(), torch.arange(2).float().pin_memory()))  # shape change
    busy = torch.rand(2048, 2048, device="cuda")
    seen = 0
    for i, (a, b) in enumerate(T.DevicePrefetcher(iter(host))):
        acc = a.sum() + b.sum()
        for _ in range(4):
            busy = busy @ busy * 1e-3       # work that is still running when the next copy is issued
        assert a.is_cuda and torch.equal(a.cpu(), host[i][0]) and torch.equal(b.cpu(), host[i][1])
        assert acc.item() == host[i][0].sum().item() + host[i][1].sum().item()
        seen += 1
    assert seen == len(host)


def test_temporal_ensemble_vs_reference_loops(golden_dir):
    """GPU TemporalEnsemble vs the reference's streaming loops (fixture produced by executing predict.py's own
    loops, oracle/gen_golden.py): same frames per batch, values within 1 ulp-ish, decoded boxes identical."""
    g = _load(golden_dir, "temporal_ensemble.npz")
    keys = sorted({k.split("/")[0] for k in g.files})
    for key in keys:
        kind, mode, L, n, bs = key.split("_")
        L, n, bs = int(L), int(n), int(bs)
        if kind == "hm":
            preds = torch.from_numpy(g[key + "/preds"])
            num_sample = preds.shape[0]
        else:
            coor, pred_in, msk = (torch.from_numpy(g[key + "/" + k]) for k in ("coor", "pred_in", "mask"))
            preds = O.inpaint_blend(pred_in, coor, msk)
            num_sample = n
        ens = T.TemporalEnsemble(L, mode, num_sample)
        outs = [ens.push(preds[a:a + bs].to(G.DEV)) for a in range(0, num_sample, bs)]
        assert [len(o) for o in outs] == list(g[key + "/counts"])
        out = torch.cat(outs).cpu()
        ref = torch.from_numpy(g[key + "/ens"]).reshape(out.shape)
        if kind == "co":
            th = (out[:, 0] < O.COOR_TH) & (out[:, 1] < O.COOR_TH)
            out[th] = 0.0
        assert (out - ref).abs().max().item() <= 1.2e-7, key
        print(f"temporal ensemble {key}: bit-exact={torch.equal(out, ref)}")
    with pytest.raises(RuntimeError):
        ens.push(preds[:1].to(G.DEV))  # more samples than num_sample


def test_temporal_ensemble_full_size_and_predict_ensemble():
    """288x512 heatmaps, seq_len 8, bs 32 batches through predict.predict_ensemble: frames, coordinates and
    visibility equal the oracle's ensemble + OpenCV-rule decode."""
    import predict as P
    L, H, W, num_sample, bs = 8, 288, 512, 40, 32
    gen = torch.Generator().manual_seed(9)
    # blobs that move one pixel per frame so that the ensemble of the 8 overlapping predictions stays above 0.5
    preds = torch.zeros(num_sample, L, H, W)
    for s in range(num_sample):
        for f in range(L):
            t = s + f
            cx, cy = 20 + 5 * t, 30 + 3 * t
            preds[s, f, cy - 3:cy + 4, cx - 3:cx + 4] = 0.6 + 0.4 * torch.rand(7, 7, generator=gen)
    idx = torch.stack([torch.stack([torch.zeros(L), torch.arange(s, s + L).float()], 1) for s in range(num_sample)])
    oracle = O.TemporalEnsemble(L, "weight", num_sample)
    ens = T.TemporalEnsemble(L, "weight", num_sample)
    got = {"Frame": [], "X": [], "Y": [], "Visibility": []}
    want = {"Frame": [], "X": [], "Y": [], "Visibility": []}
    for a in range(0, num_sample, bs):
        d = P.predict_ensemble(ens, idx[a:a + bs], y_pred=preds[a:a + bs].to(G.DEV))
        for k in got:
            got[k] += d[k]
        o = oracle.push(preds[a:a + bs])
        boxes = D.decode_batch(o.unsqueeze(1).numpy())
        last = a + bs >= num_sample
        frames = list(range(a, min(a + bs, num_sample))) + (list(range(num_sample, num_sample + L - 1)) if last else [])
        for j, fr in enumerate(frames):
            x, y, w, h = (int(v) for v in boxes[j][0])
            cx, cy = int(x + w / 2), int(y + h / 2)
            want["Frame"].append(fr); want["X"].append(cx); want["Y"].append(cy)
            want["Visibility"].append(0 if cx == 0 and cy == 0 else 1)
    assert got == want
    assert len(got["Frame"]) == num_sample + L - 1 and sum(got["Visibility"]) > 30


def test_evaluate_on_gpu_vs_reference_fixture(golden_dir):
    """test.evaluate (GPU decode + eval_stats kernel) vs the reference's evaluate run by oracle/gen_golden.py: all
    outcome types, bbox / confidence / ground-truth outputs, scaling; plus a full-size batch against the oracle."""
    import test as TT
    g = _load(golden_dir, "evaluate.npz")
    cases = {"hm_plain": {}, "hm_bbox_gt": {"output_bbox": True, "output_gt": True},
             "hm_scaled": {"img_scaler": (2.5, 2.5), "tolerance": 1.0, "output_gt": True}}
    for name, kw in cases.items():
        d = TT.evaluate(torch.from_numpy(g["indices"]), y_true=torch.from_numpy(g["y_true"]).to(G.DEV),
                        y_pred=torch.from_numpy(g["y_pred"]).to(G.DEV), **kw)
        keys = sorted(k.split("/")[1] for k in g.files if k.startswith(name + "/"))
        assert sorted(d.keys()) == keys
        for k in keys:
            assert np.array_equal(np.asarray(d[k], dtype=np.float64), g[f"{name}/{k}"]), (name, k)
    # 288x512, 2 x 8 frames: blobs, empties, a prediction without ground truth
    xb, yb = __import__("bench").synthetic_batch(2, 5)
    gen = torch.Generator().manual_seed(6)
    yp = yb * (0.6 + 0.4 * torch.rand(yb.shape, generator=gen)) + 0.3 * torch.rand(yb.shape, generator=gen)
    yp[0, 0] = 0.1; yp[1, 7, 100:104, 200:205] = 0.97
    idx = torch.stack([torch.stack([torch.zeros(8), torch.arange(8 * s, 8 * s + 8).float()], 1) for s in range(2)])
    got = TT.evaluate(idx, y_true=yb.to(G.DEV), y_pred=yp.to(G.DEV), output_bbox=True, output_gt=True)
    want = D.evaluate(idx.numpy(), y_true=yb.numpy(), y_pred=yp.numpy(), output_bbox=True, output_gt=True)
    assert got == want


def test_frame_preprocessing_vs_reference_fixture_and_pillow_oracle(golden_dir):
    """GPU resize + normalise + stack vs (i) the reference's `__process__` run with real Pillow (fixture),
    (ii) the oracle restatement of Pillow at the real sizes (720p -> 288x512, identity, 360x640): bit-exact uint8
    resampling, outputs equal to float32(frames / 255.)."""
    from oracle import pil_resize_oracle as P
    g = _load(golden_dir, "input_pipeline.npz")
    for name in ("down", "odd", "up"):
        imgs, med = g[f"{name}/imgs"], g[f"{name}/median"]
        H, W = med.shape[1:]
        fp = T.FramePreprocessor(imgs.shape[1], imgs.shape[2], H, W)
        x = torch.from_numpy(imgs).to(G.DEV).unsqueeze(0)                  # (1, L, Hs, Ws, 3)
        assert torch.equal(fp.process(x).cpu()[0], torch.from_numpy(g[f"{name}/none"]).float())
        m = fp.prepare_median(g[f"{name}/median_src"])
        assert torch.equal(m.cpu(), torch.from_numpy(med))
        assert torch.equal(fp.process(x, m).cpu()[0], torch.from_numpy(g[f"{name}/concat"]).float())
        for bg in ("subtract", "subtract_concat"):   # difference image incl. numpy's wrap-around uint8 cast
            got = fp.process(x, g[f"{name}/median_src"], bg_mode=bg).cpu()[0]
            assert torch.equal(got, torch.from_numpy(g[f"{name}/{bg}"]).float()), (name, bg)
    rng = np.random.default_rng(8)
    for hs, ws, h, w, n, l in ((720, 1280, 288, 512, 2, 3), (288, 512, 288, 512, 1, 2), (720, 1280, 360, 640, 1, 2),
                               (288, 400, 288, 512, 1, 1)):
        imgs = rng.integers(0, 256, (n, l, hs, ws, 3), dtype=np.uint8)
        fp = T.FramePreprocessor(hs, ws, h, w)
        got = fp.process(torch.from_numpy(imgs).to(G.DEV)).cpu()
        for i in range(n):
            want = torch.from_numpy(P.process_frames(imgs[i], None, w, h)).float()
            assert torch.equal(got[i], want), (hs, ws, h, w, i)
    with pytest.raises(RuntimeError):
        fp.process(torch.zeros(1, 1, 10, 10, 3, dtype=torch.uint8, device=G.DEV))
    with pytest.raises(ValueError):
        fp.process(torch.zeros(1, 1, 288, 400, 3, dtype=torch.uint8, device=G.DEV), bg_mode="subtract")


def test_run_video_all_bg_modes_cover_every_frame():
    """predict.run_video with the two subtract modes (float64 median with x.5 values from the GPU median kernel): every
    frame of the video gets exactly one prediction. (The composition itself is pinned bit for bit to the reference's
    __main__ in tests/test_gpu_predict_flow.py.)"""
    import predict as P
    from utils.general import get_model
    torch.manual_seed(0)
    t = 12
    video = torch.randint(0, 40, (t, 90, 160, 3), dtype=torch.uint8)
    for bg_mode in ('subtract', 'subtract_concat'):
        p3, none = P.run_video(video, get_model('TrackNet', 8, bg_mode).to(G.DEV).eval(), None, bg_mode=bg_mode, batch_size=3)
        assert none is None and p3['Frame'] == list(range(t))


def test_eval_loops_vs_oracle():
    """test.eval_tracknet / eval_inpaintnet (validation loops of the reference, test.py:308-443) on prepared
    predictions: loss = mean of the per-batch losses, confusion counts = the oracle's evaluate() over all frames."""
    import test as TT
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "evaluate.npz"))
    y_true, y_pred, idx = (torch.from_numpy(g[k]) for k in ("y_true", "y_pred", "indices"))

    class Stub(torch.nn.Module):  # a "model" that returns the prepared heatmaps of the batch it is given
        def forward(self, x):
            return y_pred[x[:, 0, 0, 0].long().cpu()].to(G.DEV)
    batches = [(idx[a:a + 2], torch.arange(a, min(a + 2, 3)).float().reshape(-1, 1, 1, 1), y_true[a:a + 2], None, None)
               for a in (0, 2)]
    loss, res = TT.eval_tracknet(Stub(), batches, {"tolerance": 4, "verbose": False})
    want = np.zeros(5); losses = []
    for i, x, y, _, _ in batches:
        sel = x[:, 0, 0, 0].long()
        want += TT.get_eval_res(D.evaluate(i.numpy(), y_true=y.numpy(), y_pred=y_pred[sel].numpy(), tolerance=4))
        losses.append(O.wbce_loss(y_pred[sel], y).item())
    assert [res[k] for k in ("TP", "TN", "FP1", "FP2", "FN")] == want.tolist()
    assert abs(loss - np.mean(losses)) < 1e-6
    assert res["accuracy"] == (want[0] + want[1]) / want.sum()
    # InpaintNet loop with the real module (random weights) against the oracle forward + oracle evaluate
    torch.manual_seed(2)
    net = T.InpaintNet().to(G.DEV)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(3)
    coor = torch.rand(4, 16, 2, generator=gen)
    mask = (torch.rand(4, 16, 1, generator=gen) < 0.3).float()
    coor_pred = (coor + 0.003 * torch.randn(4, 16, 2, generator=gen)) * (1 - mask)
    idx2 = torch.stack([torch.stack([torch.zeros(16), torch.arange(16 * s, 16 * s + 16).float()], 1) for s in range(4)])
    loss2, res2 = TT.eval_inpaintnet(net, [(idx2, coor_pred, coor, None, None, mask)], {"tolerance": 4, "verbose": False})
    ci = O.inpaintnet_forward(sd, coor_pred, mask)
    ci = ci * mask + coor_pred * (1 - mask)
    assert abs(loss2 - F.mse_loss(ci * mask, coor * mask).item()) < 1e-6
    ci = O.inpaint_blend(coor_pred, O.inpaintnet_forward(sd, coor_pred, mask), mask)
    for t, (ct, cp) in {"inpaint": (coor, ci), "reconstruct": (coor_pred, ci), "baseline": (coor, coor_pred)}.items():
        w = TT.get_eval_res(D.evaluate(idx2.numpy(), c_true=ct.numpy(), c_pred=cp.numpy(), tolerance=4))
        assert [res2[t][k] for k in ("TP", "TN", "FP1", "FP2", "FN")] == w.tolist(), t


def test_scalar_reader_returns_the_value_without_waiting_for_later_work():
    """ScalarReader: the host gets the scalar that was ready at read(), whatever is enqueued on the compute stream
    afterwards (the value is not the tensor's LATER content, and value() does not wait for the later kernels)."""
    import time
    reader = T.ScalarReader()
    t = torch.full((1,), 3.5, device=G.DEV)
    big = torch.rand(8192, 8192, device=G.DEV)
    torch.cuda.synchronize()
    reader.read(t)
    t.add_(1.0)                                   # later work on the compute stream must not leak into the value
    for _ in range(20):
        big = big @ big.clamp(-1e-3, 1e-3)        # ~20 x 1.1 TFLOP of later work
    t0 = time.perf_counter()
    v = reader.value()
    waited = time.perf_counter() - t0
    busy = not torch.cuda.current_stream().query()
    torch.cuda.synchronize()
    assert v == 3.5 and t.item() == 4.5
    assert busy and waited < 0.05, (busy, waited)  # the matmuls were still running when the value arrived
    with pytest.raises(RuntimeError):
        reader.value()
    with pytest.raises(RuntimeError):
        reader.read(torch.zeros(2, device=G.DEV))
