"""GPU: the compact observation table (CC_OBS_TABLE), the pipelined host path (cc_step_host / cc_rollout_host, chunks over
three streams, optional host-side row expansion) and the handle's ordering / checkpoint contracts — against the oracle and
against the device path."""

import numpy as np
import pytest
from cases import crew_config, large_config, readme_config, readme_crew
from helpers import random_states

from collectivecrossing_b200 import _abi
from collectivecrossing_b200.lowering import lower_config

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def make_env(cfg, n, **kw):
    from collectivecrossing_b200 import BatchedCollectiveCrossing

    return BatchedCollectiveCrossing(cfg, n, "cuda:0", **kw)


TABLE_CASES = {
    "A8_readme": (lambda: readme_config(max_steps=30), 1031),
    "A5_small_lattice": (lambda: readme_crew(3, 2, term="all"), 258),
    "A4": (lambda: crew_config(3, 1, max_steps=40, reward="simple_distance", term="all"), 333),
    "A1": (lambda: crew_config(1, 0, max_steps=25), 37),
    "A12": (lambda: crew_config(7, 5, max_steps=40), 130),
    "A64_large": (lambda: large_config(40), 19),
    "A100": (lambda: crew_config(60, 40, max_steps=30, reward="constant_negative"), 9),
}


# ---- CC_OBS_TABLE -----------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["auto", "lanes"])
@pytest.mark.parametrize("case", list(TABLE_CASES))
def test_table_mode_matches_oracle(case, kernel):
    """The compact table through both mappings (reset, step with auto-reset, observe, fused rollout), and its host-side
    expansion against the oracle's float32 / int8 rows."""
    import oracle

    make_cfg, n = TABLE_CASES[case]
    cfg = make_cfg()
    low = lower_config(cfg)
    env = make_env(cfg, n, seed=6, global_env_offset=11, obs_dtype="table", auto_reset=True, with_info=True, kernel=kernel)
    orc = oracle.OracleEnvs(low, n, seed=6, global_env_offset=11)
    assert np.array_equal(env.reset().cpu().numpy(), orc.reset(obs_dtype=_abi.OBS_TABLE))
    for t in range(45):
        out = env.step(policy="waiting" if t % 2 else "random")
        res = orc.step(policy="waiting" if t % 2 else "random", auto_reset=True, obs_dtype=_abi.OBS_TABLE)
        assert out.obs.shape == (n, low.num_agents, 4) and out.obs.dtype == torch.int8
        assert np.array_equal(out.obs.cpu().numpy(), res["obs"]), f"{case}/{kernel}/t={t}: table"
        assert np.array_equal(out.reward.cpu().numpy(), res["reward"]) and np.array_equal(out.agent_flags.cpu().numpy(), res["agent_flags"])
        assert np.array_equal(out.env_flags.cpu().numpy(), res["env_flags"])
    if low.num_agents <= 8:
        assert env.last_kernel == ("threads" if kernel == "auto" else "lanes")
    assert np.array_equal(env.observe().cpu().numpy(), orc.observe(_abi.OBS_TABLE))
    # rows rebuilt on the host from the table == the rows the oracle builds
    table = out.obs.cpu()
    assert np.array_equal(env.expand_table_host(table).numpy(), orc.observe(_abi.OBS_FP32))
    assert np.array_equal(env.expand_table_host(table, dtype=torch.int8, n_threads=2).numpy(), orc.observe(_abi.OBS_INT8))
    # fused rollout, time-major tables
    traj = env.rollout_trajectory(6, policy="greedy")
    for t in range(6):
        res = orc.step(policy="greedy", auto_reset=True, obs_dtype=_abi.OBS_TABLE)
        assert np.array_equal(traj["obs"][t].cpu().numpy(), res["obs"]), f"{case}/{kernel}: fused slice {t}"
    env.check_error()
    env.close()


def test_table_rows_equal_kernel_rows_at_scale():
    """262,144 README envs: expanding the table on the host gives exactly the float32 tensor the kernel writes."""
    cfg = readme_config(max_steps=25)
    n = 1 << 18
    a = make_env(cfg, n, seed=2, obs_dtype="float32")
    b = make_env(cfg, n, seed=2, obs_dtype="table")
    a.reset(); b.reset()
    for _ in range(33):
        oa = a.step(policy="waiting")
        ob = b.step(policy="waiting")
    rows = b.expand_table_host(ob.obs.cpu())
    assert torch.equal(rows, oa.obs.cpu())
    a.close(); b.close()


# ---- pipelined host path ------------------------------------------------------------------------------
@pytest.mark.parametrize("obs", ["float32", "int8", "table", None])
@pytest.mark.parametrize("chunk", [0, 32, 4096])
def test_step_host_chunks_equal_device_step(obs, chunk):
    """cc_step_host with external actions: any chunking (one chunk, many ragged chunks) returns what cc_step returns,
    with no synchronisation by the caller between reset() / step() on the torch stream and the host call."""
    cfg = readme_config(max_steps=20)
    n = 10_000 + 17
    dev_obs = obs or "none"
    a = make_env(cfg, n, seed=1, obs_dtype=dev_obs, with_info=True)
    b = make_env(cfg, n, seed=1, obs_dtype=dev_obs, with_info=True)
    b.set_host_chunk(chunk)
    a.reset(); b.reset()
    host = b.make_host_buffers()
    rng = np.random.default_rng(0)
    for t in range(24):
        acts = torch.from_numpy(rng.integers(0, 5, size=(n, 8)).astype(np.int8))
        if t % 5 == 4:   # device steps in between: the host call must order itself behind them
            a.step(policy="greedy"); b.step(policy="greedy")
        out = a.step(acts.cuda())
        host["actions"].copy_(acts)
        b.step_host(host)
        if obs is not None:
            assert torch.equal(out.obs.cpu(), host["obs"]), (t, "obs")
        assert torch.equal(out.reward.cpu(), host["reward"]) and torch.equal(out.agent_flags.cpu(), host["agent_flags"]), t
        assert torch.equal(out.env_flags.cpu(), host["env_flags"]) and torch.equal(out.agent_info.cpu(), host["agent_info"]), t
        assert torch.equal(out.actions.cpu(), host["actions_out"]), t
    assert torch.equal(a.x, b.x) and torch.equal(a.flags, b.flags) and a.step_counter == b.step_counter
    assert a.stats() == b.stats()
    b.check_error()
    a.close(); b.close()


@pytest.mark.parametrize("case,obs", [("A12", "int8"), ("A64_large", "float32"), ("A5_small_lattice", "float32")])
def test_step_host_other_crews_with_dict_order(case, obs):
    """Crews the lane-group kernel serves, with a per-env move order (the caller's dict order)."""
    make_cfg, n = TABLE_CASES[case]
    cfg = make_cfg()
    n = n * 7
    a = make_env(cfg, n, seed=3, obs_dtype=obs)
    b = make_env(cfg, n, seed=3, obs_dtype=obs)
    b.set_host_chunk(96)
    a.reset(); b.reset()
    A = a.num_agents
    host = b.make_host_buffers(pinned=False)
    rng = np.random.default_rng(1)
    import ctypes as C
    for t in range(8):
        acts = torch.from_numpy(rng.integers(0, 5, size=(n, A)).astype(np.int8))
        order = torch.from_numpy(np.stack([rng.permutation(A) for _ in range(n)]).astype(np.int8))
        out = a.step(acts.cuda(), order=order.cuda())
        host["actions"].copy_(acts)
        io = b._host_io(host, "external", None)
        io.order = order.data_ptr()
        from collectivecrossing_b200 import _native
        _native.check(b._lib.cc_step_host(b._h, C.byref(io)))
        assert torch.equal(out.obs.cpu(), host["obs"]) and torch.equal(out.reward.cpu(), host["reward"]), t
        assert torch.equal(out.agent_flags.cpu(), host["agent_flags"]) and torch.equal(out.env_flags.cpu(), host["env_flags"]), t
    assert torch.equal(a.x, b.x)
    a.close(); b.close()


@pytest.mark.parametrize("expand_threads", [0, 3, -1, -2, None])
@pytest.mark.parametrize("obs", ["float32", "int8"])
def test_host_expansion_delivers_the_same_rows(obs, expand_threads):
    """cc_set_host_expand: the table crosses PCIe and the rows are rebuilt in the caller's buffer — same bytes as the
    rows the kernel writes.  (-1: all host threads; -2 / None: the automatic choice of a new handle.)"""
    cfg = readme_config(max_steps=20)
    n = 40_000
    a = make_env(cfg, n, seed=8, obs_dtype=obs)
    b = make_env(cfg, n, seed=8, obs_dtype=obs)
    if expand_threads is not None:
        b.set_host_expand(expand_threads)
    with pytest.raises(ValueError, match="cc_set_host_expand"):
        b.set_host_expand(-3)
    b.set_host_chunk(8192)
    a.reset(); b.reset()
    host = b.make_host_buffers()
    for t in range(12):
        out = a.step(policy="waiting")
        b.step_host(host, policy="waiting")
        assert torch.equal(out.obs.cpu(), host["obs"]) and torch.equal(out.reward.cpu(), host["reward"]), t
    call = b.last_host_call()
    rows = host["obs"].numel() * host["obs"].element_size()
    assert call["chunks"] == -(-n // 8192) and call["chunk_envs"] == 8192 and call["h2d_bytes"] == 0     # the policy runs in the kernel
    if expand_threads in (None, -2):   # the automatic choice: float32 rows of >= 16 MiB on a host with >= 8 threads
        import os
        assert (call["expand_threads"] > 0) == (len(os.sched_getaffinity(0)) >= 8 and rows >= 16 << 20)   # (crew of 8: int8 rows qualify too)
    else:
        assert (call["expand_threads"] > 0) == (expand_threads != 0)
    per_env = (4 * 8 if call["expand_threads"] else rows // n) + 4 * 8 + 8 + 8 + 1     # table or rows, rewards, flags, actions, env flags
    assert call["d2h_bytes"] == n * per_env
    T = 5
    hostT = b.make_host_buffers(n_steps=T)
    traj = a.rollout_trajectory(T, policy="greedy")
    b.rollout_host(hostT, T, policy="greedy")
    assert torch.equal(traj["obs"].cpu(), hostT["obs"]) and torch.equal(traj["env_flags"].cpu(), hostT["env_flags"])
    a.close(); b.close()


@pytest.mark.parametrize("case,policy,obs,chunk", [("A8_readme", "greedy", "float32", 0), ("A8_readme", "external", "table", 2048),
                                                   ("A8_readme", "waiting", "int8", 1000), ("A12", "greedy", "int8", 64), ("A5_small_lattice", "random", "float32", 0)])
def test_rollout_host_equals_device_rollout(case, policy, obs, chunk):
    """cc_rollout_host: T steps per env, time-major host outputs == cc_rollout_fused on the device (one fused launch per chunk
    for crews of at most 8, one launch per step and chunk otherwise)."""
    make_cfg, n = TABLE_CASES[case]
    cfg = make_cfg()
    n = n * 9 + 5
    T = 11
    a = make_env(cfg, n, seed=5, global_env_offset=100, obs_dtype=obs, with_info=True)
    b = make_env(cfg, n, seed=5, global_env_offset=100, obs_dtype=obs, with_info=True)
    b.set_host_chunk(chunk)
    a.reset(); b.reset()
    A = a.num_agents
    host = b.make_host_buffers(n_steps=T)
    acts = None
    if policy == "external":
        acts = torch.from_numpy(np.random.default_rng(2).integers(0, 5, size=(T, n, A)).astype(np.int8))
        host["actions"].copy_(acts)
    for rep in range(2):
        traj = a.rollout_trajectory(T, policy=policy, actions=None if acts is None else acts.cuda())
        b.rollout_host(host, T, policy=policy)
        for k in ("obs", "reward", "agent_flags", "agent_info", "env_flags"):
            assert torch.equal(traj[k].cpu(), host[k]), (rep, k)
        assert torch.equal(traj["actions"].cpu(), host["actions_out"]), rep
    assert torch.equal(a.x, b.x) and torch.equal(a.step_count, b.step_count) and a.step_counter == b.step_counter
    sa, sb = a.stats(), b.stats()
    for k in ("env_steps", "episodes", "arrivals", "episode_length_sum"):
        assert sa[k] == sb[k], k
    b.check_error()
    a.close(); b.close()


def test_host_step_after_torch_side_state_writes_needs_no_sync():
    """set_state is a torch copy on the current stream the library never sees: step_host orders itself behind it
    (cc_order_after) — compared with a device step from the same injected state, many times over."""
    cfg = readme_config()
    n = 50_000
    rng = np.random.default_rng(3)
    a = make_env(cfg, n, obs_dtype="table", auto_reset=False)
    b = make_env(cfg, n, obs_dtype="table", auto_reset=False)
    host = b.make_host_buffers()
    for rep in range(10):
        x, y, f, s = random_states(cfg, n, rng)
        st = [torch.from_numpy(v).cuda() for v in (x, y, f, s)]
        a.set_state(*st); b.set_state(*st)
        out = a.step(policy="greedy")
        b.step_host(host, policy="greedy")
        assert torch.equal(out.obs.cpu(), host["obs"]) and torch.equal(out.reward.cpu(), host["reward"]), rep
    a.close(); b.close()


# ---- checkpoint with the numpy-compatible generators ---------------------------------------------------
def test_checkpoint_carries_the_seeded_generators():
    """reset(seed) ... reset() ... [checkpoint] ... reset(): the resumed handle's unseeded reset continues the stream."""
    cfg = readme_config(max_steps=15)
    n = 257
    seeds = torch.arange(1000, 1000 + n, dtype=torch.int64, device="cuda")
    a = make_env(cfg, n, obs_dtype="int8")
    assert a.get_state()["rng"] is None
    a.reset_seeded(seeds)
    a.rollout(9, policy="greedy")
    a.reset_seeded(None)
    snap = a.get_state()
    assert snap["rng"] is not None and tuple(snap["rng"].shape) == (n, 6)
    a.rollout(4, policy="waiting")
    want_obs = a.reset_seeded(None).clone()
    want = (a.x.clone(), a.y.clone())
    b = make_env(cfg, n, obs_dtype="int8")
    with pytest.raises(ValueError, match="seed them first"):
        b.reset_seeded(None)
    b.load_state(snap)
    b.rollout(4, policy="waiting")
    got_obs = b.reset_seeded(None)
    assert torch.equal(want_obs, got_obs) and torch.equal(want[0], b.x) and torch.equal(want[1], b.y)
    a.close(); b.close()


def test_auto_reset_attempt_cap_is_the_same_in_both_mappings():
    """A tram without an interior cell: auto-reset cannot place the exiting agent.  Both mappings and the oracle stop at
    the same attempt, keep that candidate and raise the sticky error."""
    import oracle
    from cases import unchecked
    from collectivecrossing_b200.truncated_configs import MaxStepsTruncatedConfig

    cfg = unchecked(width=6, height=4, division_y=2, tram_door_left=0, tram_door_right=0, tram_length=1, num_boarding_agents=1,
                    num_exiting_agents=1, exiting_destination_area_y=0, boarding_destination_area_y=4,
                    truncated_config=MaxStepsTruncatedConfig(max_steps=1))
    low = lower_config(cfg)
    n = 6
    x = np.tile(np.array([[1, 3]], np.int8), (n, 1)); y = np.tile(np.array([[0, 3]], np.int8), (n, 1))
    f = np.ones((n, 2), np.uint8); s = np.zeros(n, np.int32)
    orc = oracle.OracleEnvs(low, n, seed=2)
    orc.set_state(x, y, f, s)
    res = orc.step(policy="random", auto_reset=True, obs_dtype=_abi.OBS_NONE, check=False)
    assert res["status"] == _abi.ERR_RESET_STUCK
    for kernel in ("threads", "lanes"):
        env = make_env(cfg, n, seed=2, obs_dtype="none", auto_reset=True, kernel=kernel)
        env.set_state(*(torch.from_numpy(v).cuda() for v in (x, y, f, s)))
        out = env.step(policy="random")
        assert env.last_kernel == kernel
        assert np.array_equal(env.x.cpu().numpy(), orc.x) and np.array_equal(env.y.cpu().numpy(), orc.y), kernel
        assert np.array_equal(out.env_flags.cpu().numpy(), res["env_flags"]), kernel
        with pytest.raises(Exception, match="no free valid cell"):
            env.check_error()
        env.close()


@pytest.mark.parametrize("obs,expand", [("float32", -2), ("table", -2), ("int8", 0), ("float32", 0)])
def test_pageable_buffers_go_through_pinned_mirrors_with_the_same_bytes(obs, expand):
    """Host buffers in ordinary (numpy / malloc) memory: the outputs land in pinned mirrors of the handle and its host threads copy
    them out, the actions are gathered into a mirror first — the bytes the caller sees are those of the device step."""
    cfg = readme_config(max_steps=25)
    n = 150_000
    a = make_env(cfg, n, seed=4, obs_dtype=obs, with_info=True)
    b = make_env(cfg, n, seed=4, obs_dtype=obs, with_info=True)
    b.set_host_expand(expand)
    a.reset(); b.reset()
    host = b.make_host_buffers(pinned=False)
    assert not host["reward"].is_pinned()
    for t in range(8):
        acts = a.policy_actions("greedy" if t % 2 else "waiting").clone()
        out = a.step(acts)
        host["actions"].copy_(acts)
        b.step_host(host)
        assert torch.equal(out.obs.cpu(), host["obs"]) and torch.equal(out.reward.cpu(), host["reward"]), t
        assert torch.equal(out.agent_flags.cpu(), host["agent_flags"]) and torch.equal(out.env_flags.cpu(), host["env_flags"]), t
        assert torch.equal(out.agent_info.cpu(), host["agent_info"]) and torch.equal(acts.cpu(), host["actions_out"]), t
    T = 3
    hostT = b.make_host_buffers(pinned=False, n_steps=T)
    traj = a.rollout_trajectory(T, policy="waiting")
    b.rollout_host(hostT, T, policy="waiting")
    for k in ("obs", "reward", "agent_flags", "agent_info", "env_flags"):
        assert torch.equal(traj[k].cpu(), hostT[k]), k
    assert torch.equal(traj["actions"].cpu(), hostT["actions_out"])
    assert torch.equal(a.x, b.x) and torch.equal(a.flags, b.flags)
    a.close(); b.close()
