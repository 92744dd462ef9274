"""The C-ABI library loads and exports every symbol include/ccb200.h declares (no GPU needed)."""

import ctypes as C
import re
from pathlib import Path

from collectivecrossing_b200 import _abi, _native

HEADER = Path(__file__).resolve().parents[1] / "include" / "ccb200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.library()
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in ccb200.h but not exported"
    assert set(names) == set(_abi.EXPORTS), set(names) ^ set(_abi.EXPORTS)
    assert lib.cc_abi_version() == _abi.ABI_VERSION


def test_struct_layouts_match_header():
    assert C.sizeof(_abi.CCConfig) == 14 * 4 + 4 * 8
    assert C.sizeof(_abi.CCStepIO) == 8 * 8 + 4 * 4
    assert C.sizeof(_abi.CCStats) == 8 * 8
    text = HEADER.read_text()
    for name, value in (("CC_O_ALIVE_PREV", _abi.O_ALIVE_PREV), ("CC_O_OBS_PRESENT", _abi.O_OBS_PRESENT),
                        ("CC_E_WAS_RESET", _abi.E_WAS_RESET), ("CC_I_AT_DESTINATION", _abi.I_AT_DESTINATION),
                        ("CC_MAX_AGENTS", _abi.MAX_AGENTS)):
        assert re.search(rf"{name}\s*=?\s*{value}\b", text), name


def test_error_paths_without_gpu():
    lib = _native.library()
    assert lib.cc_create(None, 1, 0, 0, 0, None) == _abi.ERR_INVALID_ARG
    assert b"null" in lib.cc_last_error()
    assert lib.cc_num_envs(None) == 0 and lib.cc_obs_len(None) == 0
    lib.cc_destroy(None)


def test_oracle_and_product_share_the_abi_version():
    import oracle

    assert oracle.lib().cc_oracle_abi_version() == _native.library().cc_abi_version()
