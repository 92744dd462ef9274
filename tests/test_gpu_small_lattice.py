"""GPU: the small-lattice thread-per-env kernel (cc_kernel_tpe2.cuh: one-byte cells, per-cell table entries with the reward,
SWAR flag logic, TMA int8 images) against the oracle, step by step and field by field, for every crew size, output mode and
action source it serves — and against cc_step_tpe_kernel through the CCB200_TPE2 switch."""

import numpy as np
import pytest
from cases import cassette_config, readme_config, readme_crew
from helpers import random_states

from collectivecrossing_b200 import _abi
from collectivecrossing_b200.lowering import lower_config

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

OBS_CODE = {"none": _abi.OBS_NONE, "table": _abi.OBS_TABLE, "int8": _abi.OBS_INT8}


def make_env(cfg, n, **kw):
    from collectivecrossing_b200 import BatchedCollectiveCrossing

    return BatchedCollectiveCrossing(cfg, n, "cuda:0", **kw)


CREWS = [(1, 0), (0, 1), (1, 1), (2, 1), (2, 2), (3, 2), (4, 2), (4, 3), (5, 3), (8, 0), (0, 8)]


@pytest.mark.parametrize("policy", ["greedy", "waiting", "random", "external"])
@pytest.mark.parametrize("obs", ["none", "table"])
@pytest.mark.parametrize("crew", CREWS)
def test_small_lattice_kernel_matches_oracle(crew, obs, policy):
    import oracle

    b, e = crew
    reward = ["default", "simple_distance", "binary", "constant_negative"][(b + 2 * e) % 4]
    cfg = readme_crew(b, e, reward=reward, term="all" if (b + e) % 2 else "individual", max_steps=17)
    low = lower_config(cfg)
    n = 32 * 9 + 5
    rng = np.random.default_rng(b * 16 + e)
    env = make_env(cfg, n, seed=4, global_env_offset=9, obs_dtype=obs, auto_reset=True, with_info=True, kernel="threads")
    orc = oracle.OracleEnvs(low, n, seed=4, global_env_offset=9)
    x, y, f, s = random_states(cfg, n, rng, step_hi=10)
    env.set_state(*(torch.from_numpy(v).cuda() for v in (x, y, f, s)))
    orc.set_state(x, y, f, s)
    A = b + e
    for t in range(40):
        acts = rng.integers(0, 5, size=(n, A)).astype(np.int8) if policy == "external" else None
        out = env.step(None if acts is None else torch.from_numpy(acts).cuda(), policy=policy)
        res = orc.step(acts, policy=policy, auto_reset=True, obs_dtype=OBS_CODE[obs])
        tag = f"crew {crew} {obs} {policy} t={t}"
        for k, u, v in (("x", env.x, orc.x), ("y", env.y, orc.y), ("flags", env.flags, orc.flags), ("step", env.step_count, orc.step_count),
                        ("episode_return", env.episode_return, orc.episode_return)):
            assert np.array_equal(u.cpu().numpy(), v), f"{tag}: state {k}"
        for k, u in (("reward", out.reward), ("agent_flags", out.agent_flags), ("agent_info", out.agent_info), ("env_flags", out.env_flags)):
            assert np.array_equal(u.cpu().numpy(), res[k]), f"{tag}: {k}"
        assert np.array_equal(out.actions.cpu().numpy(), res["actions_out"]), f"{tag}: actions"
        if obs != "none":
            assert np.array_equal(out.obs.cpu().numpy(), res["obs"]), f"{tag}: obs"
    assert env.last_kernel_name == f"ccb::cc_step_tpe2_kernel<{A},{OBS_CODE[obs]}>"
    st = env.stats()
    assert st["episodes"] == orc.stats.episodes > 0 and st["arrivals"] == orc.stats.arrivals
    assert st["reward_sum"] == pytest.approx(orc.stats.reward_sum, rel=1e-9)
    env.check_error()
    env.close()


@pytest.mark.parametrize("policy", ["greedy", "waiting", "random", "external"])
@pytest.mark.parametrize("make_cfg", [lambda: readme_config(max_steps=19), lambda: readme_config("simple_distance", "all", 23, distance_penalty_factor=0.3)])
def test_small_lattice_int8_rows_match_oracle(make_cfg, policy):
    """int8 rows of 8 agents: per-thread 304-byte images, one bulk copy per 32 envs (ragged last group included)."""
    import oracle

    cfg = make_cfg()
    low = lower_config(cfg)
    n = 32 * 40 + 13
    rng = np.random.default_rng(7)
    env = make_env(cfg, n, seed=2, obs_dtype="int8", auto_reset=True, with_info=True)
    orc = oracle.OracleEnvs(low, n, seed=2)
    assert np.array_equal(env.reset().cpu().numpy(), orc.reset())
    for t in range(45):
        acts = rng.integers(0, 5, size=(n, 8)).astype(np.int8) if policy == "external" else None
        out = env.step(None if acts is None else torch.from_numpy(acts).cuda(), policy=policy)
        res = orc.step(acts, policy=policy, auto_reset=True, obs_dtype=_abi.OBS_INT8)
        assert np.array_equal(out.obs.cpu().numpy(), res["obs"]), (policy, t)
        assert np.array_equal(out.reward.cpu().numpy(), res["reward"]) and np.array_equal(out.agent_flags.cpu().numpy(), res["agent_flags"]), (policy, t)
        assert np.array_equal(out.env_flags.cpu().numpy(), res["env_flags"]) and np.array_equal(out.agent_info.cpu().numpy(), res["agent_info"]), (policy, t)
    assert env.last_kernel_name == "ccb::cc_step_tpe2_kernel<8,1>"
    assert np.array_equal(env.x.cpu().numpy(), orc.x) and np.array_equal(env.flags.cpu().numpy(), orc.flags)
    # fused: 9 steps in one launch, time-major rows
    traj = env.rollout_trajectory(9, policy="waiting")
    for t in range(9):
        res = orc.step(policy="waiting", auto_reset=True, obs_dtype=_abi.OBS_INT8)
        assert np.array_equal(traj["obs"][t].cpu().numpy(), res["obs"]), t
        assert np.array_equal(traj["reward"][t].cpu().numpy(), res["reward"]), t
    env.check_error()
    env.close()


def test_small_lattice_kernel_cassette_geometry_and_invalid_actions():
    """Another small lattice (the reference cassettes' 10x6 grid, 3 agents); out-of-range actions raise and move nothing."""
    import oracle

    cfg = cassette_config()
    low = lower_config(cfg)
    n = 100
    env = make_env(cfg, n, seed=1, obs_dtype="table", auto_reset=False, kernel="threads")
    orc = oracle.OracleEnvs(low, n, seed=1)
    env.reset(); orc.reset()
    rng = np.random.default_rng(0)
    for t in range(30):
        acts = rng.integers(0, 5, size=(n, 3)).astype(np.int8)
        out = env.step(torch.from_numpy(acts).cuda())
        res = orc.step(acts, obs_dtype=_abi.OBS_TABLE)
        assert np.array_equal(out.obs.cpu().numpy(), res["obs"]) and np.array_equal(out.reward.cpu().numpy(), res["reward"]), t
    assert env.last_kernel_name == "ccb::cc_step_tpe2_kernel<3,16>"
    env.check_error()
    before = env.x.clone()
    bad = np.full((n, 3), 4, np.int8)
    bad[7, 1], bad[50, 2] = 5, -3
    env.step(torch.from_numpy(bad).cuda())
    assert torch.equal(before, env.x)
    with pytest.raises(ValueError, match="Invalid action"):
        env.check_error()
    env.close()
