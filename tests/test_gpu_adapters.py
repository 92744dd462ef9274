"""GPU: cassette recording through the facade and from a batched rollout, the vectorised
multi-agent view, rendering — each pinned against the pure-Python port or the facade."""

import numpy as np
import pytest
from cases import cassette_config, readme_config

from collectivecrossing_b200 import cassette

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def action_dicts(ids, n, seed):
    rng = np.random.default_rng(seed)
    return [{a: int(rng.integers(0, 5)) for a in ids} for _ in range(n)]


@pytest.mark.parametrize("make_cfg", [cassette_config, lambda: readme_config("simple_distance", "all", 30)])
def test_facade_cassette_equals_python_port_cassette(make_cfg):
    from oracle.pyport import PyEnv

    from collectivecrossing_b200 import CollectiveCrossingEnv

    cfg = make_cfg()
    env, ref = CollectiveCrossingEnv(cfg), PyEnv(cfg)
    acts = action_dicts(ref.ids, 60, 5)
    got, want = cassette.record_trajectory(env, acts), cassette.record_trajectory(ref, acts)
    assert got == want
    cassette.replay_trajectory(env, want, strict=True)
    env.close()


def test_batched_recorder_exports_the_facade_trajectory():
    """Env k of a batched device rollout, exported as a cassette, is the trajectory the single-env
    facade records from reset(seed=42+k) with the same actions."""
    from collectivecrossing_b200 import BatchedCollectiveCrossing, CollectiveCrossingEnv

    cfg = readme_config(max_steps=25)
    n, k = 64, 37
    env = BatchedCollectiveCrossing(cfg, n, "cuda:0", obs_dtype="float32", reward_dtype="float64", auto_reset=False, with_info=True)
    obs = env.reset_seeded(torch.arange(n, dtype=torch.int64, device="cuda") + cassette.RESET_SEED)
    rec = cassette.BatchedTrajectoryRecorder(env, env_index=k)
    rec.begin(obs)
    rng = np.random.default_rng(8)
    steps = []
    for _ in range(40):
        a = rng.integers(0, 5, size=(n, env.num_agents)).astype(np.int8)
        steps.append(a[k].copy())
        if rec.after_step(env.step(torch.from_numpy(a).cuda())):
            break
    env.check_error()
    got = rec.cassette()
    facade = CollectiveCrossingEnv(cfg)
    ids = facade.possible_agents
    want = cassette.record_trajectory(facade, [{a: int(row[i]) for i, a in enumerate(ids)} for row in steps], seed=cassette.RESET_SEED + k)
    assert len(got["steps"]) == len(want["steps"]) > 5
    assert got == want
    env.close()
    facade.close()


def test_vector_view_matches_facade_dicts_and_policy_slices():
    from collectivecrossing_b200 import CollectiveCrossingEnv
    from collectivecrossing_b200.vector_env import VectorCollectiveCrossing, policy_mapping_fn

    cfg = readme_config(max_steps=20)
    n, k = 33, 32
    vec = VectorCollectiveCrossing(cfg, n, "cuda:0", auto_reset=False)
    obs0 = vec.reset(seed=100)
    assert obs0.shape == (n, 8, 38) and obs0.dtype == torch.float32
    facade = CollectiveCrossingEnv(cfg)
    fobs, _ = facade.reset(seed=100 + k)
    ids = vec.possible_agents
    assert ids == facade.possible_agents and [policy_mapping_fn(a) for a in ids] == ["boarding"] * 5 + ["exiting"] * 3
    assert all(np.array_equal(obs0[k, i].cpu().numpy(), fobs[a]) for i, a in enumerate(ids))
    rng = np.random.default_rng(0)
    for t in range(25):
        a = torch.from_numpy(rng.integers(0, 5, size=(n, 8)).astype(np.int8)).cuda()
        # half of the steps hand the actions over per policy, as two RLlib policies would
        step = vec.step({"boarding": a[:, :5], "exiting": a[:, 5:]} if t % 2 else a)
        want = facade.step({aid: int(a[k, i]) for i, aid in enumerate(ids)})
        got = vec.to_multi_agent_dicts(k)
        for g, w in zip(got, want):
            assert set(g) == set(w)
            for key in g:
                if isinstance(g[key], np.ndarray):
                    assert np.array_equal(g[key], w[key])
                elif isinstance(g[key], float):
                    assert g[key] == pytest.approx(w[key], rel=1e-6)     # float32 batch rewards vs float64 facade
                else:
                    assert g[key] == w[key]
        # masks replace missing keys
        assert int(step.valid[k].sum()) == len(want[1]) and int(step.obs_valid[k].sum()) == len(want[0])
        b = vec.policy_batch(step, "boarding")
        assert b["obs"].shape == (n * 5, 38) and b["rewards"].shape == (n * 5,)
        assert torch.equal(vec.policy_view(step.obs, "exiting"), step.obs[:, 5:])
        assert vec.policy_view(step.obs, "boarding").data_ptr() == step.obs.data_ptr()   # zero-copy
    vec.env.check_error()
    # on-device policy through the same view
    vec.reset(seed=7)
    s = vec.step(policy="waiting")
    assert s.obs.shape == (n, 8, 38) and bool(s.valid.all())
    vec.close()
    facade.close()


def test_render_rgb_array_from_facade_and_batch():
    from collectivecrossing_b200 import BatchedCollectiveCrossing, CollectiveCrossingEnv
    from collectivecrossing_b200.rendering import COLORS, render_batched

    cfg = readme_config()
    env = CollectiveCrossingEnv(cfg)
    env.reset(seed=42)
    img = env.render()
    cell = 32
    assert img.shape == ((cfg.height + 2) * cell, (cfg.width + 2) * cell, 3)
    x, y = (int(v) for v in env._agents["boarding_0"].position)
    assert tuple(int(v) for v in img[(cfg.height + 1 - y) * cell + cell // 2, x * cell + cell // 2]) == COLORS["boarding_agent"]
    with pytest.raises(NotImplementedError):
        env.render("human")
    b = BatchedCollectiveCrossing(cfg, 4, "cuda:0", obs_dtype="none")
    b.reset_seeded(torch.tensor([42, 43, 44, 45], dtype=torch.int64, device="cuda"))
    assert np.array_equal(render_batched(b, 0), img)            # same seed, same picture
    b.close()
    env.close()
