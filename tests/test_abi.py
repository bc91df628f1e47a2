"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports exactly what the header
declares, and the size/plan queries (pure host arithmetic) behave. No compute calls."""
import ctypes as C
import os
import re

import pytest

from tracknetv3_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tracknet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tnb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tracknet_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms  # the ctypes table binds the whole header, nothing else


def test_abi_version_and_error_channel():
    lib = _lib.load()
    assert lib.tnb_abi_version() == 2
    cfg = _lib.TrackNetCfg(n=1, h=540, w=960, in_dim=27, out_dim=8, training=1, fwd_terms=3, bwd_terms=3,
                           variant=0, bn_eps=1e-5, bn_momentum=0.1)
    assert lib.tnb_tracknet_workspace_bytes(C.byref(cfg)) == 0  # 540 is not divisible by 8: the reference raises too
    assert b"divisible by 8" in lib.tnb_last_error()


def test_workspace_and_plan_queries():
    lib = _lib.load()
    cfg = _lib.TrackNetCfg(n=10, h=288, w=512, in_dim=27, out_dim=8, training=1, fwd_terms=3, bwd_terms=3,
                           variant=0, bn_eps=1e-5, bn_momentum=0.1)
    train_bytes = lib.tnb_tracknet_workspace_bytes(C.byref(cfg))
    cfg.training = 0
    eval_bytes = lib.tnb_tracknet_workspace_bytes(C.byref(cfg))
    assert 0 < eval_bytes < train_bytes < 16 * 2 ** 30
    assert lib.tnb_conv3x3_wpack_elems(27, 64) == 64 * 32 * 9 * 2
    # 64-wide tiles keep two accumulator halves per M tile (x_hi * [w_hi | w_lo] in one MMA): 16x16-pixel CTA tiles;
    # 128-wide tiles: 16x16 as well (2 buffers x 2 M tiles x 128 columns)
    assert lib.tnb_conv3x3_stat_rows(10, 288, 512, 64, 64, 3) == 10 * 18 * 32
    assert lib.tnb_conv3x3_stat_rows(10, 144, 256, 128, 128, 3) == 10 * 9 * 16
    # eval forward: pack_input, ONE weight-pack launch, (conv, bn_finalize) x 17, predictor
    assert lib.tnb_tracknet_num_launches(C.byref(cfg), 0) == 37
    cfg.training = 1
    assert lib.tnb_tracknet_num_launches(C.byref(cfg), 0) == 38       # + the num_batches_tracked counters
    # backward: predictor (dA, dW, final sum) 3, (reduce, finalize, apply, wgrad) x 17, no view pass at all (the network
    # input is pre-split by the forward's pack launch, every other wgrad operand comes out of a BatchNorm-backward apply
    # pass), dgrad 16, ordered split-K sums 17
    assert lib.tnb_tracknet_num_launches(C.byref(cfg), 1) == 104
    cfg.training = 2                                                   # eval() with gradients enabled: running statistics,
    assert lib.tnb_tracknet_workspace_bytes(C.byref(cfg)) == train_bytes   # the backward state of a training forward,
    assert lib.tnb_tracknet_num_launches(C.byref(cfg), 0) == 37        # no counter launch,
    assert lib.tnb_tracknet_num_launches(C.byref(cfg), 1) == 104       # the same backward launches
    cfg.training = 3
    assert lib.tnb_tracknet_workspace_bytes(C.byref(cfg)) == 0 and b"training must be" in lib.tnb_last_error()
    cfg.training = 1
    cfg.variant = 16384                                                # skip halves of the 3 decoder concats by view passes
    assert lib.tnb_tracknet_num_launches(C.byref(cfg), 1) == 107
    cfg.variant = 0
    cfg.out_dim = 20                                                   # predictor kernels take 16 output channels per launch
    assert lib.tnb_tracknet_num_launches(C.byref(cfg), 0) == 39
    assert lib.tnb_heatmap_decode_workspace_bytes(256, 288, 512) == 256 * 5 * 288 * 512 * 4


def test_backward_range_split_and_debug_layer_queries():
    """Host-only queries of the data-parallel / debugging entry points (no device work)."""
    lib = _lib.load()
    assert lib.tnb_tracknet_grad_split_layer() == 7   # bottleneck.conv_1: 9.59 M of 11.34 M parameters lie behind it
    import tracknetv3_b200 as T
    m = T.TrackNet(27, 8)
    params = list(m.parameters())
    tail = sum(p.numel() for p in params[3 * 7:])
    assert len(params) == 53 and sum(p.numel() for p in params) == 11341000 and tail == 9590536
    cfg = _lib.TrackNetCfg(n=2, h=64, w=96, in_dim=27, out_dim=8, training=1, fwd_terms=3, bwd_terms=3, variant=0,
                           bn_eps=1e-5, bn_momentum=0.1)
    nbytes = lib.tnb_tracknet_workspace_bytes(C.byref(cfg))
    ptrs, dims = (C.c_void_p * 8)(), (C.c_int * 5)()
    seen = []
    for layer, (hh, ww, cin, cout) in ((0, (64, 96, 32, 64)), (7, (8, 12, 256, 512)), (10, (16, 24, 768, 256)), (16, (64, 96, 64, 64))):
        # workspace = NULL: the "pointers" are the byte offsets of the layer's tensors inside a workspace
        _lib.check(lib.tnb_tracknet_debug_layer(C.byref(cfg), None, layer, ptrs, dims))
        assert tuple(dims)[:4] == (hh, ww, cin, cout) and dims[4] == 1          # bf16 pairs in the 3-term backward
        offs = [ptrs[i] or 0 for i in range(8)]
        assert 0 < offs[0] < offs[5] < nbytes and (layer == 0) == (offs[6] == 0) and offs[7] == 0
        seen.append(offs[0])
    assert seen == sorted(seen)                                                   # layers are laid out in order
    cfg.bwd_terms = 1                                                             # single-pass backward: fp16 pairs + multiplier
    _lib.check(lib.tnb_tracknet_debug_layer(C.byref(cfg), None, 3, ptrs, dims))
    assert dims[4] == 0 and (ptrs[7] or 0) > 0
    assert lib.tnb_tracknet_debug_layer(C.byref(cfg), None, 17, ptrs, dims) != 0  # no such layer


def test_no_oracle_or_torch_fallback_in_product():
    """The product path must not import the oracle, and must fail loudly without CUDA."""
    import torch
    import tracknetv3_b200 as T
    pkg = os.path.join(ROOT, "tracknetv3_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("no oracle", ""), fn
    m = T.TrackNet(12, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 12, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.WBCELoss(torch.rand(1, 1, 4, 4), torch.rand(1, 1, 4, 4))


def test_missing_library_fails_loudly():
    """Without the built extension every op refuses to run (no eager / CPU substitute): the loader raises with the build
    instruction. Checked in a fresh interpreter whose TNB_LIBRARY points at a file that does not exist."""
    import subprocess
    import sys
    code = ("import torch, tracknetv3_b200 as T\n"
            "for call in (lambda: T._lib.load(), lambda: T.decode_heatmaps(torch.zeros(1, 4, 4)),\n"
            "             lambda: T.FusedAdam([torch.nn.Parameter(torch.zeros(2))]).step()):\n"
            "    try:\n"
            "        call()\n"
            "    except RuntimeError as e:\n"
            "        assert 'not built' in str(e) and 'no CPU / PyTorch fallback' in str(e), e\n"
            "    else:\n"
            "        raise SystemExit('an op ran without the library')\n"
            "print('refused')\n")
    env = dict(os.environ, TNB_LIBRARY=os.path.join(ROOT, "tracknetv3_b200", "no_such_library.so"))
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("refused"), r.stdout + r.stderr


_LAYER_SHAPES = [(27, 64, 0), (64, 64, 0), (64, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 2), (256, 512, 3),
                 (512, 512, 3), (768, 256, 2), (384, 128, 1), (192, 64, 0)]  # (cin, cout, level) of TrackNet's 3x3 layers


def _plans(lib, hw_list=((288, 512), (360, 640), (544, 960), (64, 96)), terms_list=(3, 1)):
    out = (C.c_int * 12)()
    for h, w in hw_list:
        for cin, cout, level in _LAYER_SHAPES:
            for side in ((cin, cout), (cout, cin)):  # forward, and dgrad (K side = cout, N side = cin)
                if side[1] == 27:
                    continue  # the first layer has no dgrad
                for terms in terms_list:
                    assert lib.tnb_conv3x3_plan_query(10, h >> level, w >> level, side[0], side[1], terms, out) == 0
                    yield (h, w, side, terms), list(out)


def test_conv_plans_fit_the_sm_for_every_layer_and_resolution():
    """Launch plans of every TrackNet layer (forward and dgrad orientation) at the reference resolution, the sweep
    resolutions and the smoke size: shared memory within the 227 KB opt-in limit, TMEM within 512 columns (a power of
    two), at least two slots in every ring, two accumulator buffers whenever they fit, 64-wide tiles merged."""
    lib = _lib.load()
    n = 0
    for key, (bn, mt, sa, sb, g, nbuf, tmem, smem, merged, tall, pair, layout) in _plans(lib):
        n += 1
        assert key[2][1] % bn == 0 and bn in (32, 64, 128, 192, 256), key
        assert 0 < smem <= 232448, (key, smem)
        accw = 2 * bn if (merged and key[3] > 1) else bn
        assert nbuf * mt * accw <= tmem <= 512 and tmem & (tmem - 1) == 0, key
        assert nbuf == (2 if 2 * mt * accw <= 512 else 1), key
        assert sa >= 2 and sb >= 2 and g in (1, 3) and sb <= 8, key
        assert merged == (1 if bn == 64 else 0) and pair == 0 and layout == merged, key  # experiments are off by default
        assert tall in (0, 1)
    assert n == (len(_LAYER_SHAPES) * 2 - 1) * 2 * 4
