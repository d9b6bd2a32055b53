"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol the header declares,
rejects bad shapes host-side with the documented codes, and the Python layer refuses CPU tensors
(there is no CPU fallback).  No compute call reaches a GPU here."""
import ctypes as C
import importlib
import os
import re

import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def cabi():
    import __graft_entry__ as ge
    ge.build()
    return importlib.import_module("x-as-supervision_b200._cabi")


def test_library_exports_every_declared_symbol(cabi):
    header = open(os.path.join(ROOT, "include", "xsup_b200.h")).read()
    declared = set(re.findall(r"\b(xsup_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(cabi.SYMBOLS)
    for name in declared:
        assert hasattr(cabi.lib, name), name
    assert cabi.lib.xsup_abi_version() == 1


def test_struct_layouts_match_header(cabi):
    assert C.sizeof(cabi.Shape) == 9 * 4
    assert C.sizeof(cabi.Cam) == 5 * C.sizeof(C.c_void_p)
    assert C.sizeof(cabi.LossCfg) == 13 * 4
    assert C.sizeof(cabi.Xchg) == 24 and cabi.lib.xsup_xchg_floats(8) == 2 * 8 * 64


def test_strides(cabi):
    s = cabi.make_shape(2, 17, 64, 64, 64, 3, 15, torch.float32)
    assert cabi.lib.xsup_stats_stride(s) == 80            # 4 + 64 + 3*3 -> 77 -> 80
    assert cabi.lib.xsup_coef_stride(s) == 72
    assert cabi.lib.xsup_stats_stride(s) % 4 == 0
    assert cabi.lib.xsup_stats_floats(s) == 2 * 17 * 80 + 16 and cabi.lib.xsup_coef_floats(s) == 2 * 17 * 72 + 16


@pytest.mark.parametrize("kw,code", [
    (dict(D=32, W=64), -1),            # D != W
    (dict(NS=14), -1),                 # even window
    (dict(NH=63), -1),                 # NH > D-2
    (dict(NH=0), -1),
    (dict(dtype=7), -4),
])
def test_bad_shapes_are_rejected_before_launch(cabi, kw, code):
    f = dict(B=1, K=17, D=64, H=64, W=64, NH=3, NS=15, dtype=0, head=0)
    f.update(kw)
    s = cabi.Shape(*[f[n] for n in ("B", "K", "D", "H", "W", "NH", "NS", "dtype", "head")])
    buf = (C.c_float * 64)()
    p = C.addressof(buf)
    rc = cabi.lib.xsup_integral_fwd(p, p, p, None, p, s, None)
    assert rc == code
    assert cabi.lib.xsup_last_error()
    with pytest.raises(RuntimeError):
        cabi.check(rc, "xsup_integral_fwd")


def test_null_and_alignment(cabi):
    s = cabi.make_shape(1, 17, 64, 64, 64, 3, 15, torch.float32)
    buf = (C.c_float * 64)()
    p = C.addressof(buf)
    assert cabi.lib.xsup_integral_fwd(None, p, p, None, p, s, None) == -3
    aligned = (p + 15) // 16 * 16
    assert cabi.lib.xsup_integral_fwd(aligned + 4, p, p, None, aligned, s, None) == -2


def test_loss_cfg_validation(cabi):
    buf = (C.c_float * 64)()
    p = C.addressof(buf)
    cam = cabi.Cam(p, p, p, p, p)
    cfg = cabi.LossCfg(4, 40, 3, 256, 256, 2000.0, 1.0, 0, 0, 0, 0, 0, 4)       # K > 32
    assert cabi.lib.xsup_reproj_loss_fwd(p, p, cam, p, p, p, cfg, None) == -1
    cfg = cabi.LossCfg(4, 16, 3, 256, 256, 2000.0, 1.0, 0.1, 0.1, 0, 1, 0, 4)   # symmetry needs K >= 17
    assert cabi.lib.xsup_reproj_loss_fwd(p, p, cam, p, p, p, cfg, None) == -1
    cfg = cabi.LossCfg(4, 18, 3, 256, 256, 2000.0, 1.0, 0.1, 0.1, 0, 1, 2, 4)   # symmetry per joint undefined
    assert cabi.lib.xsup_reproj_select(p, p, p, p, p, p, cfg, None) == -1


def test_no_cpu_fallback(cabi):
    ops = importlib.import_module("x-as-supervision_b200.ops")
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.integral_multi_head(torch.zeros(1, 18 * 16, 16, 16), 18, 3, 5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.find_peak(torch.rand(2, 3, 16), 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "x-as-supervision_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "xsup_oracle" not in src and "oracle/" not in src, f
