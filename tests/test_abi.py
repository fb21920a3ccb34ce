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
    assert lib.kdl_abi_version() == 1
    assert _ffi.status_string(0) == "ok"
    assert "IndexError" in _ffi.status_string(_ffi.KDL_ERR_INDEX)
    assert "KeyError" in _ffi.status_string(_ffi.KDL_ERR_KEY)
    assert lib.kdl_launch_count() == 0 or lib.kdl_launch_count() > 0


def test_struct_layouts_match_the_header():
    from kindel_b200 import _ffi

    # kdl_batch: 3 x i64, 6 ptr, 2 x i32, 3 ptr, i64, 2 ptr = 17 eight-byte words - 1 (two i32 share one)
    assert ctypes.sizeof(_ffi.KdlBatch) == 8 * 18
    assert ctypes.sizeof(_ffi.KdlDiag) == 24


def test_engine_refuses_to_run_without_cuda():
    import pytest
    import torch

    from kindel_b200 import engine

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.require_cuda()


def test_gpu_validated_kernels_unchanged():
    """profiles/r01_gpu_validated_kernels.txt lists the kernels (SASS hashes) that ran on the B200; later work added
    experimental instantiations and host-emulation guards around them, none of which may change their machine
    code while there is no GPU to re-validate it on.  Needs cuobjdump (CUDA toolkit); skipped without it."""
    import shutil
    import sys

    import pytest

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_hashes

    recorded = sass_hashes.read_recorded(os.path.join(ROOT, "profiles", "r01_gpu_validated_kernels.txt"))
    current = sass_hashes.kernel_hashes()
    assert len(recorded) >= 17
    changed = [name for name, h in recorded.items() if current.get(name) != h]
    assert not changed, ("GPU-validated kernels changed (re-validate on the device, then regenerate the list): %s"
                         % ", ".join(changed))
