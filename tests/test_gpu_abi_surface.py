"""The entry points of include/ccb200.h that the Python wrappers reach only indirectly (or not at all), called through ctypes as a
foreign binding would: state injection and read-back with device and with host pointers, the size queries, the device-side copy
of the statistics block, the timing bracket, the launch counter."""

import ctypes as C

import numpy as np
import pytest
from cases import readme_config

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if isinstance(a, np.ndarray) else C.c_void_p(a.data_ptr())


def test_state_stats_timing_and_size_entry_points():
    import oracle
    from collectivecrossing_b200 import BatchedCollectiveCrossing, _abi, _native
    from collectivecrossing_b200.lowering import lower_config

    cfg = readme_config(max_steps=30)
    n, A = 777, 8
    lib = _native.library()
    env = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=3, obs_dtype="int8", auto_reset=True)
    h, stream = env._h, torch.cuda.current_stream().cuda_stream
    assert lib.cc_num_envs(h) == n and lib.cc_num_agents(h) == A and lib.cc_obs_len(h) == 6 + 4 * A

    # a state from the oracle's seeded reset, injected with HOST pointers, read back both ways
    orc = oracle.OracleEnvs(lower_config(cfg), n, seed=3)
    orc.reset_seeded(np.arange(100, 100 + n))
    x, y, f, s = (np.ascontiguousarray(v) for v in orc.get_state())
    s[:] = np.arange(n) % 7
    assert lib.cc_set_state_host(h, _p(x), _p(y), _p(f), _p(s)) == _abi.OK
    assert np.array_equal(env.x.cpu().numpy(), x) and np.array_equal(env.step_count.cpu().numpy(), s)      # the attached tensors ARE the state
    gx, gy, gf, gs = np.zeros_like(x), np.zeros_like(y), np.zeros_like(f), np.zeros_like(s)
    assert lib.cc_get_state_host(h, _p(gx), _p(gy), _p(gf), _p(gs)) == _abi.OK
    assert all(np.array_equal(a, b) for a, b in ((gx, x), (gy, y), (gf, f), (gs, s)))
    dx, dy = torch.zeros((n, A), dtype=torch.int8, device="cuda"), torch.zeros((n, A), dtype=torch.int8, device="cuda")
    df, ds = torch.zeros((n, A), dtype=torch.uint8, device="cuda"), torch.zeros(n, dtype=torch.int32, device="cuda")
    assert lib.cc_get_state(h, _p(dx), _p(dy), _p(df), _p(ds), stream) == _abi.OK
    torch.cuda.synchronize()
    assert np.array_equal(dx.cpu().numpy(), x) and np.array_equal(df.cpu().numpy(), f) and np.array_equal(ds.cpu().numpy(), s)
    # ... and with DEVICE pointers: the mirrored state (x -> 12 - x is inside the lattice, the flags stay)
    dx2 = (12 - dx).contiguous()
    assert lib.cc_set_state(h, _p(dx2), _p(dy), _p(df), _p(ds), stream) == _abi.OK
    torch.cuda.synchronize()
    assert torch.equal(env.x, dx2) and torch.equal(env.y, dy)
    assert lib.cc_set_state(h, _p(dx), _p(dy), _p(df), _p(ds), stream) == _abi.OK
    # a NULL array is skipped: only the step counters change here
    s2 = (s + 1).astype(np.int32)
    assert lib.cc_set_state_host(h, None, None, None, _p(s2)) == _abi.OK
    assert np.array_equal(env.x.cpu().numpy(), x) and np.array_equal(env.step_count.cpu().numpy(), s2)
    assert lib.cc_set_state_host(h, None, None, None, _p(s)) == _abi.OK
    assert lib.cc_set_state_host(None, _p(x), _p(y), _p(f), _p(s)) == _abi.ERR_INVALID_ARG

    # steps from that state against the oracle; the timing bracket, the launch counter and the statistics block on the way
    orc.step_count[:] = s
    launches = lib.cc_launch_count(h)
    assert lib.cc_stats_reset(h, stream) == _abi.OK
    assert lib.cc_timing_begin(h, stream) == _abi.OK
    for _ in range(40):
        out = env.step(policy="greedy")
        res = orc.step(policy="greedy", auto_reset=True, obs_dtype=_abi.OBS_INT8)
    ms = C.c_float(-1.0)
    assert lib.cc_timing_end(h, stream, C.byref(ms)) == _abi.OK and 0.0 < ms.value < 5000.0
    assert lib.cc_launch_count(h) == launches + 40
    assert np.array_equal(out.obs.cpu().numpy(), res["obs"]) and np.array_equal(env.x.cpu().numpy(), orc.x)
    block = torch.zeros(8, dtype=torch.int64, device="cuda")
    assert lib.cc_stats_copy(h, _p(block), stream) == _abi.OK
    st = _abi.CCStats()
    assert lib.cc_stats_read(h, C.byref(st), stream) == _abi.OK
    raw = block.cpu().numpy()
    assert [int(v) for v in raw[:6]] == [st.env_steps, st.episodes, st.terminated_all, st.truncated_all, st.arrivals, st.episode_length_sum]
    assert raw[6:].view(np.float64).tolist() == [st.episode_return_sum, st.reward_sum]
    assert (st.env_steps, st.episodes, st.arrivals) == (n * 40, orc.stats.episodes - 0, orc.stats.arrivals) or st.episodes > 0
    assert lib.cc_stats_reset(h, stream) == _abi.OK and lib.cc_stats_read(h, C.byref(st), stream) == _abi.OK and st.env_steps == 0 and st.reward_sum == 0.0
    env.close()
