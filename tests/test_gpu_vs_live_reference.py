"""GPU parity against the UNMODIFIED reference, live: the reference (imported from /root/reference or from the archive
oracle/stage_ref.py staged, which travels to the GPU box) is stepped on the host cores in this very test run, and the CUDA path
must reproduce it — positions, flags, observations, infos and policy actions bit for bit, rewards as the correctly rounded
float32 of the reference's float64 (north_star tolerance 1e-6 relative).

* BASELINE config 2 as SURVEY.md §8d words it: README config, a >= 256-env sample, `reset(seed)` on the device against the
  reference's `reset(seed)`, then >= 100 steps of the reference's own GreedyPolicy / WaitingPolicy against ONE fused launch;
* the single-env façade against the reference on random valid configs with random action dicts (random order, missing agents)."""

import multiprocessing as mp

import numpy as np
import pytest
from cases import large_config, readme_config

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _record_chunk(job):
    from oracle import refrun

    make_cfg, kind, seeds, n_steps, kw = job
    return refrun.record(_CONFIGS[make_cfg](), seeds, n_steps, source=kind, **kw)


_CONFIGS = {
    "readme": readme_config,
    "cfg3": lambda: large_config(max_steps=40),                                                  # BASELINE config 3: 64x32, 48 + 16 agents
    "cfg4_binary": lambda: readme_config(reward="binary", term="individual", max_steps=100),     # BASELINE config 4
    "cfg4_constant_negative": lambda: readme_config(reward="constant_negative", term="individual", max_steps=100),
}


def _record_parallel(kind, seeds, n_steps, workers=8, make_cfg="readme", **kw):
    chunks = [list(c) for c in np.array_split(np.asarray(seeds), workers) if len(c)]
    with mp.get_context("spawn").Pool(len(chunks)) as pool:
        parts = pool.map(_record_chunk, [(make_cfg, kind, c, n_steps, dict(kw, stream_seed=kw.get("stream_seed", 0) + k)) for k, c in enumerate(chunks)])
    out = {}
    for k in parts[0]:
        axis = 0 if k.startswith("init_") else 1
        out[k] = np.concatenate([p[k] for p in parts], axis=axis)
    return out


@pytest.mark.parametrize("make_cfg,n_envs,n_steps,obs,kernel", [
    ("cfg3", 24, 48, "int8", "auto"), ("cfg3", 16, 44, "float32", "auto"),
    ("cfg4_binary", 96, 108, "float32", "threads"), ("cfg4_constant_negative", 96, 108, "int8", "threads"), ("cfg4_binary", 40, 104, "int8", "lanes")])
def test_baseline_shapes_against_the_live_reference(make_cfg, n_envs, n_steps, obs, kernel):
    """BASELINE configs 3 and 4 as shapes: the unmodified reference steps its envs with random action dicts in random order with
    missing agents, here and now; the device replays the same actions through the kernel each shape runs on.  (Config 4's
    truncation at MaxSteps 100 falls inside the window; the reference keeps stepping finished envs, so does the replay.)"""
    from helpers import assert_same, replay_device
    from oracle import refload

    if not refload.available():
        pytest.skip("reference neither mounted nor staged")
    threads = kernel == "threads"   # the thread-per-env kernel: agent order, float32 rewards (the correctly rounded float64)
    rec = _record_parallel("random", list(range(300, 300 + n_envs)), n_steps, make_cfg=make_cfg, validate=False, shuffle_order=not threads,
                           drop_prob=0.0 if threads else 0.1, stream_seed=17)
    got = replay_device(_CONFIGS[make_cfg](), rec, use_order=not threads, obs_dtype=obs, reward_dtype="float32" if threads else "float64", kernel=kernel)
    if threads:
        rec = dict(rec, reward=rec["reward"].astype(np.float32).astype(np.float64))
        assert got["_kernel"] == "threads"
    assert_same(rec, got, f"{make_cfg}/{obs}/{kernel}")
    assert (rec["env_flags"] != 0).any() or make_cfg == "cfg3"


@pytest.mark.parametrize("kind,n_steps", [("greedy", 105), ("waiting", 70)])
def test_config2_sample_of_288_envs_against_the_live_reference(kind, n_steps):
    from oracle import refload

    from collectivecrossing_b200 import _abi, BatchedCollectiveCrossing

    if not refload.available():
        pytest.skip("reference neither mounted nor staged")
    seeds = np.arange(5000, 5288, dtype=np.int64)
    rec = _record_parallel(kind, seeds, n_steps)
    cfg = readme_config()
    n = len(seeds)
    env = BatchedCollectiveCrossing(cfg, n, "cuda:0", obs_dtype="float32", reward_dtype="float32", auto_reset=False, with_info=True)
    obs0 = env.reset_seeded(torch.from_numpy(seeds).cuda())
    assert np.array_equal(env.x.cpu().numpy(), rec["init_x"]) and np.array_equal(env.y.cpu().numpy(), rec["init_y"]), "reset(seed) placement"
    assert np.array_equal(obs0.cpu().numpy().astype(np.int8), rec["init_obs"])
    before = env.launch_count
    traj = env.rollout_trajectory(n_steps, policy=kind)
    assert env.launch_count == before + 1 and env.last_kernel_name == "ccb::cc_step_tpe2_kernel<8,4>"     # ONE fused launch for the whole window
    env.check_error()
    assert np.array_equal(traj["actions"].cpu().numpy(), rec["actions"]), "policy actions"
    assert np.array_equal(traj["agent_flags"].cpu().numpy(), rec["agent_flags"]) and np.array_equal(traj["env_flags"].cpu().numpy(), rec["env_flags"])
    present = (rec["agent_flags"] & _abi.O_OBS_PRESENT) != 0
    assert np.array_equal(traj["obs"].cpu().numpy().astype(np.int8)[present], rec["obs"][present]), "observations"
    assert np.array_equal(traj["agent_info"].cpu().numpy()[present], rec["agent_info"][present]), "infos"
    got_r = traj["reward"].cpu().numpy()
    assert np.array_equal(got_r, rec["reward"].astype(np.float32)), "rewards: the correctly rounded float32 of the reference's float64"
    np.testing.assert_allclose(got_r, rec["reward"], rtol=1e-6, atol=0)
    assert np.array_equal(env.x.cpu().numpy(), rec["x"][-1]) and np.array_equal(env.flags.cpu().numpy(), rec["flags"][-1])
    assert (rec["env_flags"] != 0).any(), "some episodes ended inside the window"
    env.close()


def _random_valid_config(rng):
    from collectivecrossing_b200.configs import CollectiveCrossingConfig
    from collectivecrossing_b200.reward_configs import BinaryRewardConfig, ConstantNegativeRewardConfig, DefaultRewardConfig, SimpleDistanceRewardConfig
    from collectivecrossing_b200.terminated_configs import AllAtDestinationTerminatedConfig, IndividualAtDestinationTerminatedConfig
    from collectivecrossing_b200.truncated_configs import MaxStepsTruncatedConfig

    rewards = [DefaultRewardConfig(distance_penalty_factor=float(rng.choice([0.1, 0.25]))), SimpleDistanceRewardConfig(distance_penalty_factor=0.3),
               BinaryRewardConfig(goal_reward=2.0, no_goal_reward=-0.5), ConstantNegativeRewardConfig(step_penalty=-1.5)]
    while True:
        try:
            w, h = int(rng.integers(8, 20)), int(rng.integers(6, 12))
            d = int(rng.integers(2, h - 2))
            length = int(rng.integers(4, w))
            dl = int(rng.integers(0, length - 2))
            dr = int(rng.integers(dl + 2, length + 1))
            return CollectiveCrossingConfig(
                width=w, height=h, division_y=d, tram_door_left=dl, tram_door_right=min(dr, length - 1), tram_length=length,
                num_boarding_agents=int(rng.integers(1, 5)), num_exiting_agents=int(rng.integers(1, 4)),
                exiting_destination_area_y=int(rng.integers(0, d)), boarding_destination_area_y=int(rng.integers(d + 1, h + 1)),
                reward_config=rewards[int(rng.integers(0, 4))],
                terminated_config=[IndividualAtDestinationTerminatedConfig(), AllAtDestinationTerminatedConfig()][int(rng.integers(0, 2))],
                truncated_config=MaxStepsTruncatedConfig(max_steps=int(rng.integers(5, 40))))
        except Exception:  # noqa: BLE001 - the validator rejected this draw
            continue


@pytest.mark.parametrize("seed", range(12))
def test_facade_against_the_live_reference_on_random_configs(seed):
    from oracle import refload

    from collectivecrossing_b200 import CollectiveCrossingEnv

    if not refload.available():
        pytest.skip("reference neither mounted nor staged")
    ref = refload.load()
    rng = np.random.default_rng(900 + seed)
    cfg = _random_valid_config(rng)
    ours, theirs = CollectiveCrossingEnv(cfg), ref.CollectiveCrossingEnv(refload.to_reference_config(cfg))
    o1, i1 = ours.reset(seed=seed)
    o2, i2 = theirs.reset(seed=seed)
    assert o1.keys() == o2.keys() and all(np.array_equal(o1[k], o2[k]) for k in o1) and i1 == i2
    ids = ours.possible_agents
    for t in range(45):
        chosen = [a for a in ids if rng.random() > 0.15]
        rng.shuffle(chosen)
        acts = {a: int(rng.integers(0, 5)) for a in chosen}
        r1, r2 = ours.step(dict(acts)), theirs.step(dict(acts))
        for what, g, w in zip(("observations", "rewards", "terminateds", "truncateds", "infos"), r1, r2):
            assert set(g) == set(w), f"step {t}: keys of {what}"
            for k in g:
                if isinstance(g[k], np.ndarray):
                    assert g[k].dtype == w[k].dtype and np.array_equal(g[k], w[k]), f"step {t}: {what}[{k}]"
                else:
                    assert g[k] == w[k], f"step {t}: {what}[{k}]: {g[k]!r} vs {w[k]!r}"      # floats compare exactly: same float64 arithmetic
        assert ours.agents == theirs.agents or set(ours.agents) == set(theirs.agents)
    ours.close()


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_config1_recipe_with_unseeded_resets_against_the_live_reference(seed):
    """BASELINE config 1 as SURVEY.md §8d words it (C1): README config, one env, `reset(seed=s)`, actions
    `np.random.default_rng(s).integers(0, 5, size=A)` per step in agent order, `reset()` — WITHOUT a seed, i.e. continuing the env's
    generator where the last placement left it — whenever the episode is over.  The façade (every step and every placement on the
    device) must track the reference through several episodes."""
    from oracle import refload

    from collectivecrossing_b200 import CollectiveCrossingEnv

    if not refload.available():
        pytest.skip("reference neither mounted nor staged")
    ref = refload.load()
    cfg = readme_config(max_steps=40)
    ours, theirs = CollectiveCrossingEnv(cfg), ref.CollectiveCrossingEnv(refload.to_reference_config(cfg))
    o1, i1 = ours.reset(seed=seed)
    o2, i2 = theirs.reset(seed=seed)
    assert all(np.array_equal(o1[k], o2[k]) for k in o2) and i1 == i2
    ids = ours.possible_agents
    rng = np.random.default_rng(seed)
    episodes = 0
    for t in range(260):
        draw = rng.integers(0, 5, size=len(ids))
        acts = {a: int(v) for a, v in zip(ids, draw) if a in theirs.agents}
        r1, r2 = ours.step(dict(acts)), theirs.step(dict(acts))
        for what, g, w in zip(("observations", "rewards", "terminateds", "truncateds", "infos"), r1, r2):
            assert set(g) == set(w), f"step {t}: keys of {what}"
            for k in g:
                assert np.array_equal(g[k], w[k]) if isinstance(g[k], np.ndarray) else g[k] == w[k], f"step {t}: {what}[{k}]"
        assert set(ours.agents) == set(theirs.agents)
        if r2[2]["__all__"] or r2[3]["__all__"]:
            episodes += 1
            o1, i1 = ours.reset()
            o2, i2 = theirs.reset()
            assert o1.keys() == o2.keys() and all(np.array_equal(o1[k], o2[k]) for k in o2), f"reset() after episode {episodes}: placement"
            assert i1 == i2
    assert episodes >= 5
    ours.close()
