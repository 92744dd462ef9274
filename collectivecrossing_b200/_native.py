"""Loader of the in-tree C-ABI library ``csrc/libccb200.so`` (built by ``make -C csrc`` or
``__graft_entry__.build()``).  There is NO fallback: if the library is missing or a call fails,
the product raises."""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

from . import _abi

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = Path(os.environ.get("CCB200_LIB", CSRC / "libccb200.so"))  # override: A/B builds of the same ABI
_lib = None


class NativeError(RuntimeError):
    """A C-ABI call returned a negative status that has no closer Python equivalent."""


def build(force: bool = False) -> Path:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    # `all`: the library and its checked twin (device-side bounds asserts, tests/test_gpu_checked_build.py)
    cmd = ["make", "-s", "-j", str(min(8, os.cpu_count() or 1)), "-C", str(CSRC), "all"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True)
    return LIB_PATH


def library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or python -c 'import __graft_entry__ as g; "
                "g.build()').  collectivecrossing_b200 has no CPU or PyTorch fallback."
            )
        lib = _abi.bind(C.CDLL(str(LIB_PATH)))
        if lib.cc_abi_version() != _abi.ABI_VERSION:
            raise ImportError(f"{LIB_PATH} has ABI version {lib.cc_abi_version()}, expected {_abi.ABI_VERSION}: rebuild it")
        _lib = lib
    return _lib


def last_error() -> str:
    return library().cc_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a cc_status to the exception the reference would raise for the same condition."""
    if rc == _abi.OK:
        return
    msg = last_error()
    if rc in (_abi.ERR_INVALID_ACTION, _abi.ERR_INVALID_ARG):
        raise ValueError(msg)  # reference: ValueError (collectivecrossing.py:701-711)
    if rc == _abi.ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == _abi.ERR_NOMEM:
        raise MemoryError(msg)
    raise NativeError(f"cc_status {rc}: {msg}")
