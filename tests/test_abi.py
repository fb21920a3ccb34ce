"""The C-ABI library loads without a GPU and exports every symbol include/kindel_b200.h declares."""
import ctypes
import os
import re

from helpers import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kindel_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kdl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    from kindel_b200 import _ffi

    names = declared_symbols()
    assert len(names) >= 14
    assert sorted(_ffi.EXPORTED_SYMBOLS) == names
    lib = ctypes.CDLL(_ffi.lib_path())
    for n in names:
        assert getattr(lib, n) is not None


def test_abi_version_and_status_strings():
    from kindel_b200 import _ffi

    lib = _ffi.load()
    assert lib.kdl_abi_version() == 2
    assert _ffi.status_string(0) == "ok"
    assert "IndexError" in _ffi.status_string(_ffi.KDL_ERR_INDEX)
    assert "KeyError" in _ffi.status_string(_ffi.KDL_ERR_KEY)
    assert lib.kdl_launch_count() == 0 or lib.kdl_launch_count() > 0


def test_struct_layouts_match_the_header():
    from kindel_b200 import _ffi

    # kdl_batch: 2 x i64, 4 ptr, 6 x i32, 3 ptr, 2 x i64, 3 ptr
    assert ctypes.sizeof(_ffi.KdlBatch) == 8 * (2 + 4 + 3 + 3 + 2 + 3)
    assert ctypes.sizeof(_ffi.KdlDiag) == 24


def test_engine_refuses_to_run_without_cuda():
    import pytest
    import torch

    from kindel_b200 import engine

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.require_cuda()
