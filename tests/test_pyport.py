"""The pure-Python port against the unmodified reference (dict-for-dict) and the golden vectors."""

import numpy as np
import pytest
from cases import GOLDEN_CASES, readme_config
from helpers import load_golden

from oracle import refload
from oracle.pyport import PyEnv


def same_step(a, b):
    for da, db in zip(a, b):
        assert set(da) == set(db)
        for k in da:
            if isinstance(da[k], np.ndarray):
                assert da[k].dtype == db[k].dtype and np.array_equal(da[k], db[k])
            else:
                assert da[k] == db[k] and type(da[k] == db[k]) in (bool, np.bool_)


@pytest.mark.skipif(not refload.available(), reason="reference sources not mounted")
@pytest.mark.parametrize("policy", ["random", "greedy", "waiting"])
@pytest.mark.parametrize("reward,term", [("default", "individual"), ("simple_distance", "all"), ("binary", "individual"), ("constant_negative", "all")])
def test_pyport_equals_reference(policy, reward, term):
    ref = refload.load()
    cfg = readme_config(reward, term, max_steps=35)
    renv = ref.CollectiveCrossingEnv(refload.to_reference_config(cfg))
    penv = PyEnv(cfg)
    rng = np.random.default_rng(4)
    rpol = {"greedy": ref.GreedyPolicy, "waiting": ref.WaitingPolicy}.get(policy)
    rpol = rpol(randomness_factor=0.0, seed=42) if rpol else None
    for seed in range(4):
        ro, ri = renv.reset(seed=seed)
        po, pi = penv.reset(seed=seed)
        same_step((ro, ri), (po, pi))
        for t in range(45):
            if rpol is None:
                ids = list(rng.permutation(penv.ids))[: int(rng.integers(0, 9))]
                acts = {str(i): int(rng.integers(0, 5)) for i in ids}
            else:
                acts = {i: int(rpol.get_action(i, None, renv)) for i in renv.agents if renv._agents[i].active}
                assert acts == penv.policy_actions(policy)
            same_step(renv.step(acts), penv.step(acts))
            assert renv.agents == penv.agents


@pytest.mark.parametrize("name", ["readme_random", "readme_waiting", "crew_7_5"])
def test_pyport_replays_golden(name):
    cfg, rec = GOLDEN_CASES[name][0](), load_golden(name)
    T, N, A = rec["actions"].shape
    for n in range(min(N, 4)):
        env = PyEnv(cfg)
        env.reset(seed=int(rec["seeds"][n]))
        assert [env.pos[i][0] for i in env.ids] == rec["init_x"][n].tolist()
        for t in range(T):
            order = [k for k in rec["order"][t, n] if k >= 0]
            obs, rew, term, trunc, _ = env.step({env.ids[k]: int(rec["actions"][t, n, k]) for k in order})
            assert [env.pos[i][0] for i in env.ids] == rec["x"][t, n].tolist()
            assert [env.pos[i][1] for i in env.ids] == rec["y"][t, n].tolist()
            for k, i in enumerate(env.ids):
                assert rew.get(i, 0.0) == rec["reward"][t, n, k]
                assert term[i] == bool(rec["agent_flags"][t, n, k] & 16)
                if i in obs:
                    assert np.array_equal(obs[i], rec["obs"][t, n, k].astype(np.float32))


def test_runner_shaped_sampling_loop_port_equals_reference():
    """The env-runner-shaped sampling loop of tests/rllib_stub.py (used on the GPU against the façade) gives identical
    episodes on the Python port and on the unmodified reference, driven through the reference's own
    examples/training_script.py:26-64 (register_env + policy_mapping_fn + env_config, executed unchanged behind stub ray)."""
    import sys
    import tempfile
    from pathlib import Path

    import rllib_stub
    from oracle import refload
    from oracle.pyport import PyEnv

    if not refload.available():
        pytest.skip("reference neither mounted nor staged")
    rllib_stub.install()
    refload.load()
    with tempfile.TemporaryDirectory() as d:
        src = (refload.extract(Path(d), prefixes=("examples/",)) / "examples" / "training_script.py").read_text()
    rllib_stub.ENV_REGISTRY.clear()
    ns = rllib_stub.run_training_script_head(src)
    theirs = rllib_stub.ENV_REGISTRY["collective_crossing"](ns["env_config"])
    assert type(theirs).__module__ == "collectivecrossing.collectivecrossing"
    from collectivecrossing_b200.configs import CollectiveCrossingConfig
    ours_cfg = CollectiveCrossingConfig(**{k: (type(v).__name__, v.model_dump()) if hasattr(v, "model_dump") else v for k, v in ns["env_config"].items()
                                           if not hasattr(v, "model_dump")},
                                        **_own_strategy_configs(ns["env_config"]))
    policies = {"boarding": rllib_stub.SeededPolicy(1), "exiting": rllib_stub.SeededPolicy(2)}
    want = rllib_stub.sample_episodes(theirs, ns["policy_mapping_fn"], policies, n_steps=250, seed=11)
    got = rllib_stub.sample_episodes(PyEnv(ours_cfg), ns["policy_mapping_fn"], policies, n_steps=250, seed=11)
    assert len(want) >= 2
    rllib_stub.assert_same_episodes(got, want, "python port vs reference")


def _own_strategy_configs(env_config: dict) -> dict:
    """The reference's strategy config objects of an env_config dict rebuilt as OUR classes (same names, same fields)."""
    import collectivecrossing_b200.reward_configs as rc
    import collectivecrossing_b200.terminated_configs as tc
    import collectivecrossing_b200.truncated_configs as uc

    out = {}
    for key, mod in (("reward_config", rc), ("terminated_config", tc), ("truncated_config", uc)):
        if key in env_config:
            v = env_config[key]
            out[key] = getattr(mod, type(v).__name__)(**v.model_dump())
    return out
