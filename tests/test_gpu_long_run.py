"""Long runs: thousands of steps with auto-reset (dozens of episodes per env, the Philox counter far from its start, every path of the
placement loop taken many times) must stay on the oracle's trajectory — state, statistics and the last step's outputs."""

import numpy as np
import pytest
from cases import crew_config, readme_config

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("make_cfg,n,policy,steps,fused", [
    (lambda: readme_config(max_steps=60), 700, "greedy", 3000, 50),         # small-lattice thread-per-env kernel, 50 steps per launch
    (lambda: readme_config(max_steps=100, term="all"), 515, "waiting", 2400, 1),   # the same kernel, one launch per step
    (lambda: readme_config(max_steps=45), 300, "random", 2000, 40),
    (lambda: crew_config(7, 5, max_steps=50), 130, "greedy", 1500, 1),      # lane-group kernel
])
def test_long_run_stays_on_the_oracle_trajectory(make_cfg, n, policy, steps, fused):
    import oracle
    from collectivecrossing_b200 import BatchedCollectiveCrossing, _abi
    from collectivecrossing_b200.lowering import lower_config

    cfg = make_cfg()
    env = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=77, global_env_offset=12345, obs_dtype="int8", auto_reset=True, with_info=True)
    orc = oracle.OracleEnvs(lower_config(cfg), n, seed=77, global_env_offset=12345)
    env.reset(); orc.reset()
    done = 0
    while done < steps:
        if fused > 1:
            traj = env.rollout_trajectory(fused, policy=policy)
            last = {k: (None if v is None else v[-1]) for k, v in traj.items()}
            k = fused
        else:
            out = env.step(policy=policy)
            last = dict(obs=out.obs, reward=out.reward, agent_flags=out.agent_flags, agent_info=out.agent_info, env_flags=out.env_flags)
            k = 1
        for _ in range(k):
            res = orc.step(policy=policy, auto_reset=True, obs_dtype=_abi.OBS_INT8)
        done += k
        if done % 500 < k:   # a full comparison every ~500 steps, so that a divergence is reported near where it starts
            assert np.array_equal(env.x.cpu().numpy(), orc.x) and np.array_equal(env.y.cpu().numpy(), orc.y), done
            assert np.array_equal(env.flags.cpu().numpy(), orc.flags) and np.array_equal(env.step_count.cpu().numpy(), orc.step_count), done
            for key in ("obs", "reward", "agent_flags", "agent_info", "env_flags"):
                assert np.array_equal(last[key].cpu().numpy(), res[key]), (done, key)
    env.check_error()
    st = env.stats()
    o = orc.stats
    assert (st["env_steps"], st["episodes"], st["terminated_all"], st["truncated_all"], st["arrivals"], st["episode_length_sum"]) == \
           (o.env_steps, o.episodes, o.terminated_all, o.truncated_all, o.arrivals, o.episode_length_sum)
    assert st["episodes"] > 10 * n
    env.close()
