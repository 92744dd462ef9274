"""Argument errors of the hot-path entry points, through ctypes as a foreign caller would hit them: every malformed cc_step_io is
refused with CC_ERR_INVALID_ARG and a message BEFORE anything is launched — the env state, the RNG counter and the launch count are
what they were, and the next valid step equals the step of a twin env that never saw the bad calls."""

import ctypes as C

import pytest
from cases import crew_config, readme_config

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _clone(io):
    return type(io).from_buffer_copy(bytes(io))   # ctypes structures copy by value


def _bad_ios(env, io):
    def variant(**kw):
        v = _clone(io)
        for k, val in kw.items():
            setattr(v, k, val)
        return v

    from collectivecrossing_b200 import _abi

    yield "unknown policy", variant(policy=9), "policy"
    yield "external policy without actions", variant(policy=_abi.POLICIES["external"], actions=None), "actions"
    yield "unknown obs dtype", variant(obs_dtype=3), "obs_dtype"
    yield "obs dtype without buffer", variant(obs=None), "obs"
    yield "unknown reward dtype", variant(reward_dtype=2), "reward_dtype"
    yield "no reward buffer", variant(reward=None), "required"
    yield "no env flags buffer", variant(env_flags=None), "required"
    yield "misaligned observation buffer", variant(obs=env.obs.data_ptr() + 4), "aligned"


@pytest.mark.parametrize("make_cfg,kernel", [(lambda: readme_config(max_steps=30), "auto"), (lambda: crew_config(7, 5, max_steps=30), "auto")])
def test_malformed_step_io_is_refused_before_anything_runs(make_cfg, kernel):
    from collectivecrossing_b200 import BatchedCollectiveCrossing, _abi, _native

    cfg = make_cfg()
    n = 3000
    env = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=9, obs_dtype="float32", auto_reset=True, kernel=kernel)
    twin = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=9, obs_dtype="float32", auto_reset=True, kernel=kernel)
    env.reset(); twin.reset()
    for _ in range(5):
        env.step(policy="greedy"); twin.step(policy="greedy")
    lib = _native.library()
    good = env._fill_io(None, None, "greedy", None)
    launches = env.launch_count
    stream = torch.cuda.current_stream().cuda_stream
    for what, io, fragment in _bad_ios(env, good):
        for call in (lambda: lib.cc_step(env._h, C.byref(io), stream), lambda: lib.cc_rollout_fused(env._h, C.byref(io), 4, stream)):
            rc = call()
            assert rc == _abi.ERR_INVALID_ARG, (what, rc)
            assert fragment in lib.cc_last_error().decode(), (what, lib.cc_last_error())
    assert lib.cc_step(None, C.byref(good), stream) == _abi.ERR_INVALID_ARG and lib.cc_step(env._h, None, stream) == _abi.ERR_INVALID_ARG
    assert lib.cc_rollout_fused(env._h, C.byref(good), 0, stream) == _abi.ERR_INVALID_ARG
    # the host-buffer entry points check the same things (host pointers)
    host = env.make_host_buffers()
    hio = env._host_io(host, "greedy", None)
    assert lib.cc_rollout_host(env._h, C.byref(hio), 0) == _abi.ERR_INVALID_ARG
    bad = _clone(hio); bad.reward = None
    assert lib.cc_step_host(env._h, C.byref(bad)) == _abi.ERR_INVALID_ARG
    bad = _clone(hio); bad.policy = -1
    assert lib.cc_step_host(env._h, C.byref(bad)) == _abi.ERR_INVALID_ARG
    assert lib.cc_set_host_chunk(None, 0) == _abi.ERR_INVALID_ARG and lib.cc_last_host_call(env._h, None) == _abi.ERR_INVALID_ARG
    torch.cuda.synchronize()
    assert env.launch_count == launches, "a refused call launched something"
    assert torch.equal(env.x, twin.x) and torch.equal(env.step_count, twin.step_count)
    for _ in range(40):   # resets in here: the RNG counter must not have moved either
        a, b = env.step(policy="greedy"), twin.step(policy="greedy")
    assert torch.equal(a.obs, b.obs) and torch.equal(a.reward, b.reward) and torch.equal(env.x, twin.x) and torch.equal(env.flags, twin.flags)
    assert env.stats() == twin.stats() and env.stats()["episodes"] > 0
    env.close(); twin.close()
