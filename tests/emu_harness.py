"""Builds and drives the kernel emulator (tests/emu/): test infrastructure.

`tests/emu/emu_pileup.cpp` compiles the SOURCE of K0 + the tile-owner kernels (K1f, K1x) for the host on top
of `tests/emu/cuda_emu.h`, a functional model of the CUDA execution model (fibres per thread, warp
collectives, shared memory, mbarrier / bulk copy / cp.async with late completion).  `run_pileup` runs one
kernel over a batch held in numpy arrays and returns the count table."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
OUT_DIR = os.path.join(EMU_DIR, "_build")
LIB = os.path.join(OUT_DIR, "libkdl_emu.so")
CUDA_INCLUDE = os.environ.get("CUDA_INCLUDE", "/usr/local/cuda/include")
K1F, K1X, K1F_LEAN, K1W, K1W2 = 0, 1, 2, 3, 4

_lib = None


def available() -> bool:
    return shutil.which("g++") is not None and os.path.exists(os.path.join(CUDA_INCLUDE, "cuda_runtime.h"))


def _sources():
    csrc = os.path.join(ROOT, "kindel_b200", "csrc")
    return [os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(EMU_DIR, "emu_pileup.cpp"),
            os.path.join(csrc, "kdl_common.cuh"), os.path.join(csrc, "pileup_tiled.cu"),
            os.path.join(csrc, "pileup_wide.cu"), os.path.join(csrc, "pileup_ws.cu"),
            os.path.join(ROOT, "include", "kindel_b200.h")]


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = _sources()
    if not (os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in src)):
        os.makedirs(OUT_DIR, exist_ok=True)
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-I", CUDA_INCLUDE, "-I", os.path.join(ROOT, "include"),
               os.path.join(EMU_DIR, "emu_pileup.cpp"), "-o", LIB]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building the kernel emulator failed:\n" + res.stdout + res.stderr)
    from kindel_b200 import _ffi

    lib = C.CDLL(LIB)
    lib.emu_last_error.restype = C.c_char_p
    lib.emu_pileup.restype = C.c_int
    lib.emu_pileup.argtypes = [C.POINTER(_ffi.KdlBatch), C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                               C.c_longlong, C.c_int, C.c_int, C.c_int]
    _lib = lib
    return lib


def run_pileup(batch, variant: int, fresh: bool, grid: int = 5, tile_lo: int = 0, n_tiles: int = None,
               counts: np.ndarray = None) -> np.ndarray:
    """K0 + one tile-owner kernel over tiles [tile_lo, tile_lo + n_tiles) of `batch` (a bamio.ReadBatch).

    fresh=True: the weight columns of `counts` hold garbage on entry (the kernel must overwrite them);
    fresh=False: the kernel adds to what is there.  Returns int32 [19, n_slots]."""
    from kindel_b200 import engine

    lib = load()
    st, keep = engine.host_struct(batch)
    n_slots = int(batch.n_slots)
    if n_tiles is None:
        n_tiles = n_slots // 512 - tile_lo
    if counts is None:
        counts = np.zeros((19, n_slots), dtype=np.int32)
        if fresh:
            counts[0:5] = 0x5A5A5A5A
    index = np.zeros(8 * (n_slots // 512), dtype=np.uint32)
    rc = lib.emu_pileup(C.byref(st), counts.ctypes.data, n_slots, index.ctypes.data, tile_lo, n_tiles, variant,
                        1 if fresh else 0, grid)
    del keep
    if rc:
        raise RuntimeError(lib.emu_last_error().decode())
    return counts
