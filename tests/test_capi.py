"""The C-ABI library loads, exports every symbol include/hoisdf_b200.h declares, and validates arguments
(no kernel is launched here: argument checks return before any CUDA call)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hoisdf_b200.h")).read()
    return sorted(set(re.findall(r"\b(hoisdf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_built):
    lib = C.CDLL(lib_built)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libhoisdf_b200.so does not export %s" % n


def test_binding_covers_header(lib_built):
    from hoisdf_b200 import _capi
    assert set(_capi.SIGNATURES) == set(declared_symbols())
    assert _capi.lib.hoisdf_abi_version() == _capi.ABI_VERSION


def test_argument_validation_without_gpu(lib_built):
    from hoisdf_b200 import _capi
    lib = _capi.lib
    a = _capi.LinearArgs()            # all NULL
    assert lib.hoisdf_linear_fwd(C.byref(a), None) == -1          # HOISDF_E_NULL
    assert lib.hoisdf_gather_fwd(None, None, 0, None, 0, 0, 0, None, 0, None, 0, None) == -1
    assert lib.hoisdf_attention_fwd(None, 0, None, None, 0, None, 0, 1, 4, 1, 1, 1, None, None, 0, None) == -1
    assert lib.hoisdf_lattice_count(None, None, None, 3.1, 1, 64, None, None, None) == -1
    a.x, a.w, a.y = 16, 16, 16        # non-NULL, aligned dummies; bad K alignment must be rejected before launch
    a.m, a.n, a.k, a.ldx, a.ldw, a.ldy = 4, 4, 6, 8, 8, 4
    assert lib.hoisdf_linear_fwd(C.byref(a), None) == -3          # HOISDF_E_ALIGN
    a.k = 4
    a.ldx = 2
    assert lib.hoisdf_linear_fwd(C.byref(a), None) == -3
    assert b"aligned" in lib.hoisdf_status_string(-3)
    with pytest.raises(_capi.HoisdfError):
        _capi.check(-2, "demo")
    assert lib.hoisdf_lattice_chunks(64) == 256


def test_metric_argument_validation(lib_built):
    from hoisdf_b200 import _capi
    lib = _capi.lib
    assert lib.hoisdf_obj_metrics_fwd(None, None, 1, 1, None, None, 1, None, None, 1, None, None, None, None, None, 0, None) == -1
    assert lib.hoisdf_mesh_metrics_fwd(16, 16, 1, 0, None, None, None, 16, 1 << 20, None) == -2
    assert lib.hoisdf_mesh_metrics_fwd(16, 16, 1, 1000, None, None, None, 16, 8, None) == -2      # workspace too small
    assert lib.hoisdf_hand_joint_metrics_fwd(None, None, 1, 21, None, None, None, None) == -1
    assert lib.hoisdf_obj_metrics_workspace_bytes(32, 1000) == 32 * 4 * 14 * 4


def test_no_cpu_fallback(lib_built):
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    import torch
    from hoisdf_b200 import ops
    with pytest.raises(RuntimeError):
        ops.sdf_pad_input(torch.zeros(4, 289))
