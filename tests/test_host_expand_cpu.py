"""cc_expand_obs_host (host-side row expansion of the compact observation table) against the oracle's rows.
No GPU: the function is pure data movement on host threads; the table comes from the oracle's CC_OBS_TABLE output."""

import ctypes as C

import numpy as np
import pytest
from cases import crew_config, large_config, readme_config, readme_crew

from collectivecrossing_b200 import _abi, _native
from collectivecrossing_b200.lowering import lower_config


def _expand(low, table, dtype, threads):
    lib = _native.library()
    n, a = table.shape[0], low.num_agents
    out = np.full((n, a, low.obs_len), 77, dtype)
    code = _abi.OBS_FP32 if dtype == np.float32 else _abi.OBS_INT8
    rc = lib.cc_expand_obs_host(C.byref(low), n, table.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), code, threads)
    assert rc == 0, lib.cc_last_error()
    return out


@pytest.mark.parametrize("make_cfg,n", [(readme_config, 5000), (lambda: readme_crew(3, 2), 3001), (lambda: crew_config(7, 5), 700),
                                        (lambda: large_config(30), 67), (lambda: crew_config(60, 40), 9), (lambda: crew_config(1, 0), 33)])
@pytest.mark.parametrize("threads", [1, 3, 0])
def test_expanded_rows_equal_oracle_rows(make_cfg, n, threads):
    import oracle

    cfg = make_cfg()
    low = lower_config(cfg)
    orc = oracle.OracleEnvs(low, n, seed=4)
    orc.reset()
    for _ in range(7):
        orc.step(policy="greedy", auto_reset=True, obs_dtype=_abi.OBS_NONE)
    table = orc.observe(_abi.OBS_TABLE)
    assert table.shape == (n, low.num_agents, 4) and table.dtype == np.int8
    for dtype, code in ((np.float32, _abi.OBS_FP32), (np.int8, _abi.OBS_INT8)):
        want = orc.observe(code)
        got = _expand(low, table, dtype, threads)
        assert np.array_equal(got, want), f"{dtype.__name__} rows differ"


def test_expand_writes_only_its_range_and_handles_unaligned_destinations():
    """Streaming stores work on 16-byte units: the bytes around an unaligned destination must stay untouched."""
    import oracle

    low = lower_config(readme_crew(3, 2))   # 5 agents: an env block (5 x 26 bytes) is not a multiple of 16
    n = 4097
    orc = oracle.OracleEnvs(low, n, seed=1)
    orc.reset()
    table, want = orc.observe(_abi.OBS_TABLE), orc.observe(_abi.OBS_INT8)
    lib = _native.library()
    nbytes = want.size
    for shift in (0, 1, 7, 15):
        raw = np.full(nbytes + 64, 0x5A, np.uint8)
        dst = raw[16 + shift:16 + shift + nbytes]
        rc = lib.cc_expand_obs_host(C.byref(low), n, table.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), _abi.OBS_INT8, 4)
        assert rc == 0
        assert np.array_equal(dst.view(np.int8).reshape(want.shape), want), shift
        assert (raw[:16 + shift] == 0x5A).all() and (raw[16 + shift + nbytes:] == 0x5A).all(), shift


def test_expand_rejects_bad_arguments():
    lib = _native.library()
    low = lower_config(readme_config())
    buf = np.zeros(64, np.int8)
    p = buf.ctypes.data_as(C.c_void_p)
    assert lib.cc_expand_obs_host(C.byref(low), 1, None, p, _abi.OBS_FP32, 1) == _abi.ERR_INVALID_ARG
    assert lib.cc_expand_obs_host(C.byref(low), 1, p, p, _abi.OBS_TABLE, 1) == _abi.ERR_INVALID_ARG
    assert lib.cc_expand_obs_host(C.byref(low), 0, p, p, _abi.OBS_FP32, 1) == _abi.OK
