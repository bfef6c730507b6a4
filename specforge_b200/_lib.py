"""ctypes loader for libspecforge_b200.so (the C-ABI CUDA library).

There is deliberately no fallback: if the shared library is missing or a symbol cannot be
resolved the import fails loudly (the product path never routes through oracle/ or PyTorch ops).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_int, c_int64, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspecforge_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")


class SfError(RuntimeError):
    """Raised when a libspecforge_b200 entry point returns a non-zero status."""


def build_library(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into specforge_b200/libspecforge_b200.so (nvcc cross-compiles on CPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libspecforge_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the EAGLE3 hot path)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(l: ctypes.CDLL) -> None:
    l.sf_last_error.restype = c_char_p
    l.sf_version.restype = c_char_p
    l.sf_launch_count.restype = c_longlong
    l.sf_launch_count_reset.restype = None
    l.sf_gemm_bf16.restype = c_int
    l.sf_gemm_bf16.argtypes = [
        c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int64,
        c_int, c_int, c_int, c_int, c_int, c_void_p,
    ]


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().sf_last_error().decode("utf-8", "replace")
        raise SfError(f"{what} failed with status {status}: {msg}")
