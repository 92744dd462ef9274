"""GPU parity AT THE SIZES BASELINE.json names: the device runs the full batch, the C oracle follows blocks of envs spread over
the batch (first, middle, last, odd offsets).  Possible because every random draw is keyed on the GLOBAL env index: an oracle
created with global_env_offset = k reproduces env k of the big batch — reset placement, random actions and auto-reset included.
Every step, the blocks' positions, flags, rewards, observations and env flags are compared bit for bit."""

import numpy as np
import pytest
from cases import large_config, readme_config

from collectivecrossing_b200 import _abi
from collectivecrossing_b200.lowering import lower_config

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

OBS_CODE = {"float32": _abi.OBS_FP32, "int8": _abi.OBS_INT8, "table": _abi.OBS_TABLE}


def _strided_oracle_check(cfg, n, obs, policy, steps, block, offsets, seed=17):
    import oracle
    from collectivecrossing_b200 import BatchedCollectiveCrossing

    low = lower_config(cfg)
    env = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=seed, obs_dtype=obs, auto_reset=True, with_info=True)
    env.reset()
    orcs = []
    for off in offsets:
        o = oracle.OracleEnvs(low, block, seed=seed, global_env_offset=off)
        o.reset()
        orcs.append(o)
        assert np.array_equal(env.x[off:off + block].cpu().numpy(), o.x), f"reset placement of envs {off}.."
    resets = 0
    for t in range(steps):
        out = env.step(policy=policy)
        for off, o in zip(offsets, orcs):
            res = o.step(policy=policy, auto_reset=True, obs_dtype=OBS_CODE[obs])
            sl = slice(off, off + block)
            tag = f"t={t} envs {off}..{off + block}"
            assert np.array_equal(env.x[sl].cpu().numpy(), o.x) and np.array_equal(env.y[sl].cpu().numpy(), o.y), f"{tag}: positions"
            assert np.array_equal(env.flags[sl].cpu().numpy(), o.flags) and np.array_equal(env.step_count[sl].cpu().numpy(), o.step_count), f"{tag}: flags / step"
            assert np.array_equal(out.reward[sl].cpu().numpy(), res["reward"]), f"{tag}: rewards"
            assert np.array_equal(out.agent_flags[sl].cpu().numpy(), res["agent_flags"]) and np.array_equal(out.env_flags[sl].cpu().numpy(), res["env_flags"]), f"{tag}: flags out"
            assert np.array_equal(out.agent_info[sl].cpu().numpy(), res["agent_info"]), f"{tag}: infos"
            assert np.array_equal(out.obs[sl].cpu().numpy(), res["obs"]), f"{tag}: observations"
            assert np.array_equal(env.episode_return[sl].cpu().numpy(), o.episode_return), f"{tag}: episode return"
            resets += int((res["env_flags"] & _abi.E_WAS_RESET != 0).sum())
    env.check_error()
    st = env.stats()
    assert st["env_steps"] == steps * n
    name = env.last_kernel_name
    env.close()
    del env, out
    torch.cuda.empty_cache()
    return resets, st, name


@pytest.mark.parametrize("obs", ["int8", "float32"])
def test_config3_one_million_envs_of_64_agents(obs):
    """BASELINE config 3: 64x32 grid, 48 boarding + 16 exiting agents, SimpleDistance reward, AllAtDestination termination,
    1,048,576 envs, random actions; MaxSteps 40 so that three rounds of truncation-driven auto-resets fall inside the 125 steps."""
    n = 1 << 20
    offsets = [0, 32 * 1001, n // 2 - 5, n - 24]
    resets, st, name = _strided_oracle_check(large_config(40), n, obs, "random", 125, 24, offsets)
    assert name.startswith("ccb::cc_kernel<32,2,")
    assert resets == 3 * 24 * len(offsets)            # nobody gets 64 agents home in 40 random steps: every env is truncated thrice
    assert st["episodes"] == st["truncated_all"] == 3 * n


@pytest.mark.parametrize("reward,kw", [("binary", dict(goal_reward=1.0, no_goal_reward=0.0)), ("constant_negative", dict(step_penalty=-1.0))])
def test_config4_four_million_envs_reward_sweep(reward, kw):
    """BASELINE config 4: README grid, Binary / ConstantNegative reward x IndividualAtDestination termination, auto-reset,
    4,194,304 envs, random actions, float32 rows; 125 steps > MaxSteps 100, so every env is truncated and re-placed once."""
    n = 1 << 22
    offsets = [0, 777_777, n // 2 + 31, n - 64]
    resets, st, name = _strided_oracle_check(readme_config(reward, "individual", 100, **kw), n, "float32", "random", 125, 64, offsets)
    assert name == "ccb::cc_step_tpe2_kernel<8,4>"
    assert resets >= 64 * len(offsets)
    assert st["episodes"] >= n


@pytest.mark.parametrize("obs", ["int8", "table"])
def test_config5_shard_two_million_envs_waiting_policy_compact_rows(obs):
    """BASELINE config 5 shard (2,097,152 README envs, waiting policy, auto-reset) in the compact output modes the small-lattice
    kernel serves, with a global offset as rank 5 of 8 would have it."""
    import oracle
    from collectivecrossing_b200 import BatchedCollectiveCrossing

    cfg = readme_config()
    low = lower_config(cfg)
    n, base, block = 1 << 21, 5 * (1 << 21), 48
    env = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=3, global_env_offset=base, obs_dtype=obs, auto_reset=True, with_info=True)
    env.reset()
    offsets = [0, 1_234_567, n - block]
    orcs = [oracle.OracleEnvs(low, block, seed=3, global_env_offset=base + off) for off in offsets]
    for o in orcs:
        o.reset()
    resets = 0
    for t in range(130):
        out = env.step(policy="waiting")
        for off, o in zip(offsets, orcs):
            res = o.step(policy="waiting", auto_reset=True, obs_dtype=OBS_CODE[obs])
            sl = slice(off, off + block)
            assert np.array_equal(out.obs[sl].cpu().numpy(), res["obs"]) and np.array_equal(out.reward[sl].cpu().numpy(), res["reward"]), (t, off)
            assert np.array_equal(out.agent_flags[sl].cpu().numpy(), res["agent_flags"]) and np.array_equal(out.env_flags[sl].cpu().numpy(), res["env_flags"]), (t, off)
            assert np.array_equal(out.actions[sl].cpu().numpy(), res["actions_out"]), (t, off)
            resets += int((res["env_flags"] & _abi.E_WAS_RESET != 0).sum())
    assert env.last_kernel_name == f"ccb::cc_step_tpe2_kernel<8,{OBS_CODE[obs]}>"
    assert resets > 3 * block * len(offsets)          # episodes of the waiting policy last ~31 steps
    env.check_error()
    env.close()
