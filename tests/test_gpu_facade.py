"""The single-env dict facade (``CollectiveCrossingEnv``) on the GPU: the reference's cassettes,
a dict-for-dict comparison with the pure-Python port on random action dicts, and the behaviours the
reference's own unit tests pin (tests/collectivecrossing/envs/*.py, cited per test)."""

import numpy as np
import pytest
from cases import cassette_config, readme_config
from helpers import load_golden

from collectivecrossing_b200 import _abi
from collectivecrossing_b200.truncated_configs import MaxStepsTruncatedConfig

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def make(cfg):
    from collectivecrossing_b200 import CollectiveCrossingEnv

    return CollectiveCrossingEnv(cfg)


def same_step(a, b):
    for da, db in zip(a, b):
        assert set(da) == set(db), (set(da), set(db))
        for k in da:
            if isinstance(da[k], np.ndarray):
                assert da[k].dtype == db[k].dtype == np.float32 and np.array_equal(da[k], db[k]), k
            else:
                assert da[k] == db[k], (k, da[k], db[k])


@pytest.mark.parametrize("name", ["cassette_basic", "cassette_regression"])
def test_reference_cassettes_through_the_facade(name):
    """test_trajectory_vcr.py:494-593: exact dict equality incl. float64 rewards and infos."""
    rec = load_golden(name)
    env = make(cassette_config())
    ids = env.possible_agents
    obs, infos = env.reset(seed=42)
    assert infos == {a: {"agent_type": a.split("_")[0]} for a in ids}
    for k, a in enumerate(ids):
        assert obs[a].dtype == np.float32 and np.array_equal(obs[a], rec["init_obs"][0, k].astype(np.float32))
    for t in range(rec["actions"].shape[0]):
        order = [k for k in rec["order"][t, 0] if k >= 0]
        o, r, term, trunc, info = env.step({ids[k]: int(rec["actions"][t, 0, k]) for k in order})
        for k, a in enumerate(ids):
            bits = int(rec["agent_flags"][t, 0, k])
            assert (a in r) == bool(bits & _abi.O_ALIVE_PREV) == (a in trunc)
            if a in r:
                assert isinstance(r[a], float) and r[a] == rec["reward"][t, 0, k]  # e.g. -0.30000000000000004
                assert trunc[a] == bool(bits & _abi.O_TRUNC_VALUE)
            assert term[a] == bool(bits & _abi.O_TERM_VALUE)
            assert (a in o) == bool(bits & _abi.O_OBS_PRESENT) == (a in info)
            if a in o:
                assert np.array_equal(o[a], rec["obs"][t, 0, k].astype(np.float32))
                ai = int(rec["agent_info"][t, 0, k])
                assert info[a] == {"agent_type": a.split("_")[0], "in_tram_area": bool(ai & 1), "at_door": bool(ai & 2),
                                   "active": bool(ai & 4), "at_destination": bool(ai & 8)}
        assert term["__all__"] == bool(rec["env_flags"][t, 0] & 1) and trunc["__all__"] == bool(rec["env_flags"][t, 0] & 2)
    env.close()


@pytest.mark.parametrize("reward,term", [("default", "individual"), ("simple_distance", "all"), ("binary", "individual"), ("constant_negative", "all")])
def test_facade_equals_python_port_on_random_action_dicts(reward, term):
    from oracle.pyport import PyEnv

    cfg = readme_config(reward, term, max_steps=25)
    env, port = make(cfg), PyEnv(cfg)
    rng = np.random.default_rng(7)
    for episode, seed in enumerate([3, None, None, 11]):  # None: reset() keeps the generator, like gymnasium
        same_step(env.reset(seed=seed), port.reset(seed=seed))
        for t in range(32):  # runs past the end of the episode
            ids = list(rng.permutation(port.ids))[: int(rng.integers(0, 9))]
            acts = {str(i): int(rng.integers(0, 5)) for i in ids}
            same_step(env.step(acts), port.step(acts))
            assert env.agents == port.agents
    env.close()


@pytest.mark.parametrize("kind", ["greedy", "waiting"])
def test_policy_wrappers_match_python_port(kind):
    from collectivecrossing_b200.baseline_policies import GreedyPolicy, WaitingPolicy
    from oracle.pyport import PyEnv

    cfg = readme_config(max_steps=60)
    env, port = make(cfg), PyEnv(cfg)
    policy = {"greedy": GreedyPolicy, "waiting": WaitingPolicy}[kind](randomness_factor=0.0, seed=42)
    obs, _ = env.reset(seed=5)
    port.reset(seed=5)
    for t in range(60):
        acts = {a: policy.get_action(a, obs.get(a), env) for a in env.agents if env._agents[a].active}
        assert acts == port.policy_actions(kind)
        res = env.step(acts)
        same_step(res, port.step(acts))
        obs = res[0]
        if res[2]["__all__"] or res[3]["__all__"]:
            break
    assert kind == "greedy" or res[2]["__all__"]  # the waiting policy finishes the README config
    env.close()


def test_invalid_actions_and_agents_raise_like_the_reference():
    """test_action_agent_validity.py:59-71 and test_collective_crossing.py:207-235."""
    env = make(readme_config())
    env.reset(seed=1)
    with pytest.raises(ValueError, match="Invalid action"):
        env.step({"boarding_0": 5})
    with pytest.raises(ValueError, match="Invalid action"):
        env.step({"boarding_0": -1})
    with pytest.raises(ValueError, match="Unknown agent ID") as exc:
        env.step({"boarding 0": 1})
    assert "'boarding_0', 'boarding_1'" in str(exc.value) and "'exiting_0'" in str(exc.value)
    for bad in (None, "", "agent_99"):
        with pytest.raises(ValueError, match="Unknown agent ID"):
            env.step({bad: 0})
    env.step({})  # an empty dict is a legal step
    env.close()


def test_observation_structure_and_wait():
    """test_collective_crossing.py:75-113, 280-352, 396-429."""
    env = make(readme_config())
    obs, _ = env.reset(seed=9)
    A = 8
    assert len(obs) == A and all(o.dtype == np.float32 and o.shape == (6 + 4 * A,) for o in obs.values())
    for k, a in enumerate(env.possible_agents):
        o = obs[a]
        assert o[2:6].tolist() == [8.0, 4.0, 7.0, 9.0]  # door centre, division, door left/right (absolute)
        assert o[6 + 4 * k: 10 + 4 * k].tolist() == [-1.0] * 4
        assert np.array_equal(o, env._get_agent_observation(a))
        assert o[0] == env._agents[a].x and o[1] == env._agents[a].y
    before = {a: env._agents[a].position.copy() for a in env.possible_agents}
    env.step({a: 4 for a in env.possible_agents})
    assert all(np.array_equal(before[a], env._agents[a].position) for a in before)
    assert env.observation_space.shape == (38,) and env.action_space.n == 5
    assert env.get_observation_space("boarding_0") is env.observation_space and set(env.action_spaces) == set(env.possible_agents)
    env.close()


def test_state_injection_termination_and_reward_keys():
    """test_collective_crossing.py:116-152, test_rewards.py:222-327: tests overwrite env._agents."""
    for term in ("individual", "all"):
        env = make(readme_config(term=term))
        env.reset(seed=2)
        env._agents["boarding_0"].position = np.array([5, 7])  # one step below the boarding destination row
        _, r, terminated, _, _ = env.step({"boarding_0": 1})
        assert env._agents["boarding_0"].y == 8 and not env._agents["boarding_0"].active
        assert r["boarding_0"] == 15.0 and isinstance(r["boarding_0"], float)
        assert terminated["boarding_0"] == (term == "individual") and not terminated["__all__"]
        _, r2, terminated2, truncated2, obs_info = env.step({})
        if term == "individual":  # done agents get no reward / truncated / obs entry any more
            assert "boarding_0" not in r2 and "boarding_0" not in truncated2 and terminated2["boarding_0"] is True
            assert "boarding_0" not in env.agents
        else:                      # all-at-destination: still alive, keeps collecting the arrival reward
            assert r2["boarding_0"] == 15.0 and terminated2["boarding_0"] is False and "boarding_0" in env.agents
        env.close()


def test_max_steps_one_truncates_everybody():
    """test_rewards.py:330-369."""
    cfg = readme_config().model_copy(update={"truncated_config": MaxStepsTruncatedConfig(max_steps=1)})
    env = make(cfg)
    env.reset(seed=4)
    _, rewards, terminateds, truncateds, _ = env.step({a: 4 for a in env.possible_agents})
    assert truncateds["__all__"] is True and all(truncateds[a] for a in env.possible_agents) and len(rewards) == 8
    assert env.agents == []
    _, rewards, terminateds, truncateds, _ = env.step({})
    assert rewards == {} and truncateds == {"__all__": False} and len(terminateds) == 9
    env.close()


def test_binary_and_constant_rewards_exact():
    """test_rewards.py:78-160: Binary never pays goal_reward; ConstantNegative is the constant."""
    env = make(readme_config("binary", goal_reward=1.0, no_goal_reward=-1.0))
    env.reset(seed=0)
    env._agents["exiting_0"].position = np.array([5, 1])
    _, r, term, _, _ = env.step({"exiting_0": 3})
    assert term["exiting_0"] and r["exiting_0"] == -1.0 and set(r.values()) == {-1.0}
    env.close()
    env = make(readme_config("constant_negative", step_penalty=-2.5))
    env.reset(seed=0)
    _, r, _, _, _ = env.step({})
    assert set(r.values()) == {-2.5} and len(r) == 8
    env.close()


def test_unknown_strategy_names_raise_at_construction():
    from collectivecrossing_b200.reward_configs import CustomRewardConfig

    cfg = readme_config().model_copy(update={"reward_config": CustomRewardConfig(reward_function="mine")})
    with pytest.raises(ValueError, match="Unknown reward function 'mine'"):
        make(cfg)


def test_env_config_dict_like_rllib():
    from collectivecrossing_b200 import CollectiveCrossingEnv

    d = readme_config().model_dump(exclude={"reward_config", "terminated_config", "truncated_config", "observation_config"})
    env = CollectiveCrossingEnv(d)
    obs, _ = env.reset(seed=42)
    assert [int(obs[a][0]) for a in env.possible_agents] == [1, 7, 5, 1, 2, 6, 8, 9]  # SURVEY.md §8c known answer
    env.close()


# ---- the strategy objects (the reference's plugin API), evaluated on the device --------------------------
def _inject(env, ref, rng):
    """Same random valid state into the facade's host records and the pure-Python port."""
    cfg = env.config
    cells = [(x, y) for x in range(cfg.width + 1) for y in range(cfg.height + 1) if ref.valid(x, y)]
    picks = rng.choice(len(cells), size=len(ref.ids), replace=False)
    for a, k in zip(ref.ids, picks):
        x, y = cells[k]
        ref.pos[a], ref.active[a], ref.term[a], ref.trunc[a] = (x, y), True, False, False
        ag = env._agents[a]
        ag.position, ag.active, ag.terminated, ag.truncated = np.array([x, y]), True, False, False


@pytest.mark.parametrize("reward", ["default", "simple_distance", "binary", "constant_negative"])
def test_reward_function_objects_match_python_port(reward):
    """rewards.py:41-182 through ``env._reward_function.calculate_reward`` / ``env._calculate_reward`` on random
    valid states, and through function objects of the OTHER kinds built from their own configs
    (test_rewards.py:34-160: exact constants for binary / constant_negative)."""
    from oracle.pyport import PyEnv

    from collectivecrossing_b200.rewards import REWARD_FUNCTIONS, get_reward_function

    cfg = readme_config(reward)
    env, ref = make(cfg), PyEnv(cfg)
    env.reset(seed=1)
    ref.reset(seed=1)
    assert type(env._reward_function) is REWARD_FUNCTIONS[reward]
    rng = np.random.default_rng(3)
    for _ in range(12):
        _inject(env, ref, rng)
        for a in ref.ids:
            want = float(ref.reward(a))
            got = env._calculate_reward(a)
            assert isinstance(got, float) and got == want, (a, got, want)
            assert env._reward_function.calculate_reward(a, env) == want
    # a function object evaluates with ITS config, whatever the env was built with
    other = readme_config("constant_negative", step_penalty=-2.5)
    fn = get_reward_function(other.reward_config)
    assert all(fn.calculate_reward(a, env) == -2.5 for a in ref.ids)
    # rewards.py:65-66: None for agents that are terminated or truncated (test_rewards.py:222-327, 476-526)
    env._agents["boarding_0"].terminated = True
    env._agents["exiting_0"].truncated = True
    assert env._calculate_reward("boarding_0") is None and env._calculate_reward("exiting_0") is None
    assert env._calculate_reward("boarding_1") is not None
    env.close()


@pytest.mark.parametrize("current_step,max_steps,expected", [(0, 10, False), (5, 10, False), (9, 10, False), (10, 10, True),
                                                             (11, 10, True), (0, 1, False), (1, 1, True), (999, 1000, False), (1000, 1000, True)])
def test_truncated_function_truth_table(current_step, max_steps, expected):
    """test_truncateds.py:14-53: ``calculate_truncated`` is ``step_count >= max_steps``; None for a done agent."""
    from collectivecrossing_b200.truncateds import MaxStepsTruncatedFunction

    env = make(readme_config(max_steps=max_steps))
    env.reset(seed=42)
    assert isinstance(env._truncated_function, MaxStepsTruncatedFunction)
    env._step_count = current_step
    for a in env.possible_agents:
        assert env._truncated_function.calculate_truncated(a, env) is expected
        assert env._calculate_truncated(a) is expected
    env._agents["boarding_0"].terminated = True
    assert env._calculate_truncated("boarding_0") is None
    env.close()


@pytest.mark.parametrize("term", ["individual", "all"])
def test_terminated_function_objects(term):
    """test_terminateds.py:29-133: nobody is terminated after reset; an agent moved to its destination row is
    terminated in the individual mode at once, in the all-at-destination mode only when everybody is there."""
    from collectivecrossing_b200.terminateds import TERMINATED_FUNCTIONS

    cfg = readme_config(term=term)
    env = make(cfg)
    env.reset(seed=42)
    name = {"individual": "individual_at_destination", "all": "all_at_destination"}[term]
    assert type(env._terminated_function) is TERMINATED_FUNCTIONS[name]
    assert not any(env._calculate_terminated(a) for a in env.possible_agents)
    env._agents["boarding_0"].position = np.array([5, cfg.boarding_destination_area_y])
    assert env._calculate_terminated("boarding_0") is (term == "individual")
    assert env._calculate_terminated("exiting_0") is False
    for k, a in enumerate(env.possible_agents):      # everybody onto the destination row (distinct x)
        y = cfg.boarding_destination_area_y if a.startswith("boarding") else cfg.exiting_destination_area_y
        env._agents[a].position = np.array([3 + k if a.startswith("boarding") else k, y])
    assert all(env._calculate_terminated(a) is True for a in env.possible_agents)
    env.close()


def test_observation_function_integration_and_registries():
    """test_collective_crossing.py:260-277, 396-429: the observation function object returns what reset / step
    returned; unknown names raise with the reference's message."""
    from collectivecrossing_b200.observations import DefaultObservationFunction, get_observation_function
    from collectivecrossing_b200.rewards import get_reward_function

    env = make(readme_config())
    obs, _ = env.reset(seed=42)
    assert isinstance(env._observation_function, DefaultObservationFunction)
    assert env._observation_function.observation_config.get_observation_function_name() == "default"
    for a in env.possible_agents:
        assert np.array_equal(obs[a], env._observation_function.get_agent_observation(a, env))
        assert env._observation_function.return_agent_observation_space(a, env).shape == (38,)
    obs, *_ = env.step({a: 4 for a in env.possible_agents})
    for a in obs:
        assert np.array_equal(obs[a], env._observation_function.get_agent_observation(a, env))

    class Invalid:
        def get_observation_function_name(self):
            return "invalid"

        def get_reward_function_name(self):
            return "invalid"

    with pytest.raises(ValueError, match="Unknown observation function 'invalid'"):
        get_observation_function(Invalid())
    with pytest.raises(ValueError, match="Unknown reward function 'invalid'"):
        get_reward_function(Invalid())
    env.close()


def test_np_random_is_the_placement_generator_like_in_the_reference():
    """gymnasium's env.np_random is the generator reset() places agents with: draws a caller makes between resets shift the
    next placement.  The façade's generator lives on the device; env.np_random is a numpy view of its state that is pushed
    back before the next unseeded reset — same numbers, same placements as the unmodified reference."""
    from cases import readme_config
    from oracle import refload

    from collectivecrossing_b200 import CollectiveCrossingEnv

    if not refload.available():
        pytest.skip("reference neither mounted nor staged")
    ref = refload.load()
    cfg = readme_config()
    ours, theirs = CollectiveCrossingEnv(cfg), ref.CollectiveCrossingEnv(refload.to_reference_config(cfg))
    for env in (ours, theirs):
        env.reset(seed=77)
    a, b = ours.np_random.integers(0, 1000, size=5), theirs.np_random.integers(0, 1000, size=5)
    assert np.array_equal(a, b)
    assert ours.np_random.random() == theirs.np_random.random()
    o1, _ = ours.reset()
    o2, _ = theirs.reset()
    assert o1.keys() == o2.keys() and all(np.array_equal(o1[k], o2[k]) for k in o1)
    # and again without touching the generator in between
    o1, _ = ours.reset()
    o2, _ = theirs.reset()
    assert all(np.array_equal(o1[k], o2[k]) for k in o1)
    assert int(ours.np_random.integers(0, 10**6)) == int(theirs.np_random.integers(0, 10**6))
    ours.close()
