"""The C-ABI library loads and exports every symbol include/triplane_b200.h declares; argument validation that
needs no GPU (no compute calls).  CPU."""
import ctypes as C
import os
import re

import pytest

from triplaneturbo_b200 import _cabi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions():
    src = open(os.path.join(ROOT, "include", "triplane_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _cabi.load()


def test_header_and_binding_agree():
    assert header_functions() == sorted(_cabi.SIGNATURES)
    # tt_config travels by pointer through the ABI: 12 scalars + the optional image shape, 4 bytes each (the .cu asserts 56 too)
    assert C.sizeof(_cabi.TTConfig) == 56
    src = open(os.path.join(ROOT, "include", "triplane_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", src[src.index("/* Scalars of the path"):src.index("} tt_config;")], flags=re.S)
    fields = re.findall(r"\b(?:int32_t|float)\s+([a-z_A-Z, ]+);", body)
    names = [n.strip() for f in fields for n in f.split(",")]
    assert names == [n for n, _ in _cabi.TTConfig._fields_], names


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.tt_version() == 100


def test_size_queries_and_offsets(lib):
    for C_ in (8, 16, 32, 40, 64):
        assert lib.tt_wgrad_floats(C_) == 64 * C_ + 4096 + 64 + 192 * C_ + 4096 + 192
        off = (C.c_int64 * 6)()
        assert lib.tt_wgrad_offsets(C_, off) == 0
        assert list(off) == [0, 64 * C_, 64 * C_ + 4096, 64 * C_ + 4160, 256 * C_ + 4160, 256 * C_ + 8256]
        assert lib.tt_wpack_floats(C_) > lib.tt_wgrad_floats(C_)
    assert lib.tt_wpack_floats(12) == 0
    cfg = _cabi.TTConfig(8, 16, 2, 5, 1.0, 0.5, 100.0, 1.0, 0.1, 4.0, 0.05, 0)
    # 7 seed floats per sample + compaction lists + hidden-gradient planes [P][3][R*R][64]
    assert lib.tt_render_bwd_scratch_floats(C.byref(cfg), 10, 7) >= 630 + 2 * 3 * 256 * 64
    assert lib.tt_render_bwd_scratch_floats(None, 10, 7) >= 630
    assert lib.tt_geometry_bwd_scratch_floats(C.byref(cfg), 10) >= 180 + 2 * 3 * 256 * 64
    assert lib.tt_sample_scratch_floats(10, 16) >= 340


def test_argument_errors_are_reported_not_crashed(lib):
    cfg = _cabi.TTConfig(12, 16, 1, 1, 1.0, 0.5, 100.0, 1.0, 0.1, 4.0, 0.05, 0)      # unsupported C
    rc = lib.tt_geometry_fwd(None, None, C.byref(cfg), None, 0, 0, None, None, None, None, None, None, None)
    assert rc == -1 and b"channel" in lib.tt_last_error()
    rc = lib.tt_repack_planes(None, 1, 8, 0, 0, 8, 16, None, None)
    assert rc == -1
    rc = lib.tt_pack_weights(*([None] * 9), 32, None, None)
    assert rc == -1


def test_package_refuses_cpu_tensors():
    import torch
    from triplaneturbo_b200 import ops
    with pytest.raises(_cabi.TTError):
        ops.repack_planes(torch.zeros(1, 6, 8, 4, 4))
