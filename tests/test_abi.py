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
    assert lib.tnb_abi_version() == 1
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
    assert lib.tnb_tracknet_num_launches(C.byref(cfg), 0) == 53
    assert lib.tnb_heatmap_decode_workspace_bytes(256, 288, 512) == 256 * 5 * 288 * 512 * 4


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
