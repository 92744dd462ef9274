"""Record trajectories of the UNMODIFIED reference in the batched device-output format.

Build-container only (needs ``refload.available()``).  Used by ``tests/golden/make_golden.py``
to generate the committed fixtures and by ``tests/test_oracle_vs_reference.py`` to pin the C
oracle directly against the reference on fresh random cases.
"""

from __future__ import annotations

import numpy as np

from collectivecrossing_b200 import _abi

from . import refload

WAIT = 4


def agent_ids(cfg) -> list[str]:
    return [f"boarding_{i}" for i in range(cfg.num_boarding_agents)] + [
        f"exiting_{i}" for i in range(cfg.num_exiting_agents)
    ]


def snapshot(env, ids):
    """(x, y, flags, step) of a reference env in device encoding."""
    x = np.array([env._agents[i].x for i in ids], np.int8)
    y = np.array([env._agents[i].y for i in ids], np.int8)
    f = np.array(
        [
            (_abi.F_ACTIVE if env._agents[i].active else 0)
            | (_abi.F_TERMINATED if env._agents[i].terminated else 0)
            | (_abi.F_TRUNCATED if env._agents[i].truncated else 0)
            for i in ids
        ],
        np.uint8,
    )
    return x, y, f, np.int32(env._step_count)


def encode_step(env, ids, result, obs_len):
    """Reference ``step()`` return value -> device-style arrays for one env."""
    obs, rewards, terminateds, truncateds, infos = result
    A = len(ids)
    reward = np.zeros(A, np.float64)
    aflags = np.zeros(A, np.uint8)
    ainfo = np.zeros(A, np.uint8)
    obs_arr = np.zeros((A, obs_len), np.int8)
    for k, i in enumerate(ids):
        ag = env._agents[i]
        bits = 0
        bits |= _abi.O_ACTIVE if ag.active else 0
        bits |= _abi.O_TERMINATED if ag.terminated else 0
        bits |= _abi.O_TRUNCATED if ag.truncated else 0
        assert (i in rewards) == (i in truncateds), "reward/truncated keys always travel together"
        if i in rewards:
            bits |= _abi.O_ALIVE_PREV
            reward[k] = rewards[i]
        assert i in terminateds, "terminateds has every agent every step"
        bits |= _abi.O_TERM_VALUE if terminateds[i] else 0
        bits |= _abi.O_TRUNC_VALUE if truncateds.get(i, False) else 0
        assert (i in obs) == (i in infos)
        if i in obs:
            bits |= _abi.O_OBS_PRESENT
            o = np.asarray(obs[i])
            assert o.dtype == np.float32 and o.shape == (obs_len,)
            assert np.all(o == np.round(o)) and o.min() >= -1 and o.max() <= 126
            obs_arr[k] = o.astype(np.int8)
            inf = infos[i]
            assert inf["agent_type"] == ag.agent_type.value
            ainfo[k] = (
                (_abi.I_IN_TRAM_AREA if inf["in_tram_area"] else 0)
                | (_abi.I_AT_DOOR if inf["at_door"] else 0)
                | (_abi.I_ACTIVE if inf["active"] else 0)
                | (_abi.I_AT_DESTINATION if inf["at_destination"] else 0)
            )
        aflags[k] = bits
    eflags = np.uint8(
        (_abi.E_TERMINATED_ALL if terminateds["__all__"] else 0) | (_abi.E_TRUNCATED_ALL if truncateds["__all__"] else 0)
    )
    return reward, aflags, ainfo, eflags, obs_arr


def record(cfg, seeds, n_steps, source="random", stream_seed=0, validate=True, shuffle_order=False,
           drop_prob=0.0):
    """Step one reference env per seed for ``n_steps`` (no resets: stepping continues past the
    end of the episode, which the reference allows) and return a dict of arrays:

    init_{x,y,flags,step}        state after ``reset(seed)``
    actions [T,N,A] int8         action per agent (WAIT where the dict had no entry)
    order   [T,N,A] int8         dict order (agent indices, -1 padded)
    x,y,flags [T,N,A], step [T,N]  post-step state
    reward [T,N,A] f64, agent_flags, agent_info [T,N,A] u8, env_flags [T,N] u8, obs [T,N,A,L] i8
    """
    ref = refload.load()
    rcfg = refload.to_reference_config(cfg, validate=validate)
    ids = agent_ids(cfg)
    A, N, T = len(ids), len(seeds), n_steps
    L = 6 + 4 * A
    rng = np.random.default_rng(stream_seed)
    out = dict(
        init_x=np.zeros((N, A), np.int8), init_y=np.zeros((N, A), np.int8), init_flags=np.zeros((N, A), np.uint8),
        init_step=np.zeros(N, np.int32), init_obs=np.zeros((N, A, L), np.int8),
        actions=np.full((T, N, A), WAIT, np.int8), order=np.full((T, N, A), -1, np.int8),
        x=np.zeros((T, N, A), np.int8), y=np.zeros((T, N, A), np.int8), flags=np.zeros((T, N, A), np.uint8),
        step=np.zeros((T, N), np.int32), reward=np.zeros((T, N, A), np.float64),
        agent_flags=np.zeros((T, N, A), np.uint8), agent_info=np.zeros((T, N, A), np.uint8),
        env_flags=np.zeros((T, N), np.uint8), obs=np.zeros((T, N, A, L), np.int8),
    )
    for n, seed in enumerate(seeds):
        env = ref.CollectiveCrossingEnv(rcfg)
        obs, _ = env.reset(seed=int(seed))
        out["init_x"][n], out["init_y"][n], out["init_flags"][n], out["init_step"][n] = snapshot(env, ids)
        for k, i in enumerate(ids):
            out["init_obs"][n, k] = np.asarray(obs[i]).astype(np.int8)
        policy = None
        if source == "greedy":
            policy = ref.GreedyPolicy(randomness_factor=0.0, seed=42)
        elif source == "waiting":
            policy = ref.WaitingPolicy(randomness_factor=0.0, seed=42)
        for t in range(T):
            if policy is None:
                chosen = list(range(A))
                if drop_prob > 0:
                    chosen = [k for k in chosen if rng.random() >= drop_prob]
                if shuffle_order:
                    rng.shuffle(chosen)
                acts = {ids[k]: int(rng.integers(0, 5)) for k in chosen}
            else:
                # scripts/run_greedy_policy_demo.py:71-77: live, active agents in agent order
                acts = {}
                for i in env.agents:
                    if env._agents[i].active:
                        acts[i] = int(policy.get_action(i, None, env))
            for pos, (i, a) in enumerate(acts.items()):
                k = ids.index(i)
                out["actions"][t, n, k] = a
                out["order"][t, n, pos] = k
            res = env.step(acts)
            out["x"][t, n], out["y"][t, n], out["flags"][t, n], out["step"][t, n] = snapshot(env, ids)
            r, af, ai, ef, ob = encode_step(env, ids, res, L)
            out["reward"][t, n], out["agent_flags"][t, n], out["agent_info"][t, n] = r, af, ai
            out["env_flags"][t, n], out["obs"][t, n] = ef, ob
    return out


def replay_with_oracle(cfg, rec, use_order=True, policy="external"):
    """Step the C oracle over a recording's initial states / actions; returns the same arrays."""
    from collectivecrossing_b200.lowering import lower_config

    from . import OracleEnvs

    low = lower_config(cfg)
    T, N, A = rec["actions"].shape
    o = OracleEnvs(low, N)
    o.set_state(rec["init_x"], rec["init_y"], rec["init_flags"], rec["init_step"])
    got = {k: np.zeros_like(v) for k, v in rec.items() if not k.startswith("init_")}
    got["init_obs"] = o.observe()
    for t in range(T):
        res = o.step(
            rec["actions"][t] if policy == "external" else None,
            order=rec["order"][t] if (use_order and policy == "external") else None,
            policy=policy, obs_dtype=_abi.OBS_INT8, reward_dtype=_abi.REWARD_F64,
        )
        got["x"][t], got["y"][t], got["flags"][t], got["step"][t] = o.get_state()
        got["reward"][t], got["agent_flags"][t] = res["reward"], res["agent_flags"]
        got["agent_info"][t], got["env_flags"][t], got["obs"][t] = res["agent_info"], res["env_flags"], res["obs"]
        got["actions"][t] = res["actions_out"]
        got["order"][t] = rec["order"][t]
    return got


def compare(rec, got, what="oracle", policy_actions=False):
    """Field-by-field exact comparison (obs / info only where the reference returned them)."""
    for k in ("x", "y", "flags", "step", "agent_flags", "env_flags"):
        bad = np.argwhere(rec[k] != got[k])
        assert bad.size == 0, f"{what}: {k} differs first at {bad[0]}: ref={rec[k][tuple(bad[0])]} got={got[k][tuple(bad[0])]}"
    assert np.array_equal(rec["reward"], got["reward"]), f"{what}: float64 rewards differ"
    present = (rec["agent_flags"] & _abi.O_OBS_PRESENT) != 0
    assert np.array_equal(rec["agent_info"][present], got["agent_info"][present]), f"{what}: infos differ"
    assert np.array_equal(rec["obs"][present], got["obs"][present]), f"{what}: observations differ"
    assert np.array_equal(rec["init_obs"], got["init_obs"]), f"{what}: reset observations differ"
    if policy_actions:
        assert np.array_equal(rec["actions"], got["actions"]), f"{what}: policy actions differ"
