"""Shared helpers: load golden recordings, replay them through the oracle or the device,
compare field by field."""

from __future__ import annotations

from pathlib import Path

import numpy as np

from collectivecrossing_b200 import _abi
from collectivecrossing_b200.lowering import lower_config

GOLDEN = Path(__file__).resolve().parent / "golden"
CASSETTE_BITS = _abi.O_ALIVE_PREV | _abi.O_TERM_VALUE | _abi.O_TRUNC_VALUE | _abi.O_OBS_PRESENT


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def replay_oracle(cfg, rec, policy="external", use_order=True):
    import oracle

    low = lower_config(cfg)
    T, N, A = rec["actions"].shape
    o = oracle.OracleEnvs(low, N)
    o.set_state(rec["init_x"], rec["init_y"], rec["init_flags"], rec["init_step"])
    got = _blank(rec)
    got["init_obs"] = o.observe()
    for t in range(T):
        res = o.step(rec["actions"][t] if policy == "external" else None,
                     order=rec["order"][t] if (use_order and policy == "external") else None,
                     policy=policy, obs_dtype=_abi.OBS_INT8, reward_dtype=_abi.REWARD_F64)
        got["x"][t], got["y"][t], got["flags"][t], got["step"][t] = o.get_state()
        for k in ("reward", "agent_flags", "agent_info", "env_flags", "obs"):
            got[k][t] = res[k]
        got["actions"][t] = res["actions_out"]
    return got


def replay_device(cfg, rec, policy="external", use_order=True, obs_dtype="int8", reward_dtype="float64", device="cuda:0", kernel="auto"):
    import torch

    from collectivecrossing_b200 import BatchedCollectiveCrossing

    T, N, A = rec["actions"].shape
    env = BatchedCollectiveCrossing(cfg, N, device, obs_dtype=obs_dtype, reward_dtype=reward_dtype, auto_reset=False, with_info=True, kernel=kernel)
    dev = env.device
    env.set_state(torch.from_numpy(rec["init_x"]).to(dev), torch.from_numpy(rec["init_y"]).to(dev),
                  torch.from_numpy(rec["init_flags"]).to(dev), torch.from_numpy(rec["init_step"]).to(dev))
    got = _blank(rec)
    got["init_obs"] = env.observe().cpu().numpy().astype(np.int8)
    acts = torch.from_numpy(rec["actions"]).to(dev)
    order = torch.from_numpy(rec["order"]).to(dev)
    for t in range(T):
        out = env.step(acts[t] if policy == "external" else None,
                       order=order[t] if (use_order and policy == "external") else None, policy=policy)
        got["x"][t], got["y"][t] = env.x.cpu().numpy(), env.y.cpu().numpy()
        got["flags"][t], got["step"][t] = env.flags.cpu().numpy(), env.step_count.cpu().numpy()
        got["reward"][t] = out.reward.cpu().numpy()
        got["agent_flags"][t], got["agent_info"][t] = out.agent_flags.cpu().numpy(), out.agent_info.cpu().numpy()
        got["env_flags"][t] = out.env_flags.cpu().numpy()
        got["obs"][t] = out.obs.cpu().numpy().astype(np.int8)
        got["actions"][t] = out.actions.cpu().numpy()
    env.check_error()
    got["_kernel"] = env.last_kernel
    env.close()
    return got


def _blank(rec):
    T, N, A = rec["actions"].shape
    L = 6 + 4 * A
    return dict(
        x=np.zeros((T, N, A), np.int8), y=np.zeros((T, N, A), np.int8), flags=np.zeros((T, N, A), np.uint8),
        step=np.zeros((T, N), np.int32), reward=np.zeros((T, N, A), np.float64), agent_flags=np.zeros((T, N, A), np.uint8),
        agent_info=np.zeros((T, N, A), np.uint8), env_flags=np.zeros((T, N), np.uint8), obs=np.zeros((T, N, A, L), np.int8),
        actions=np.zeros((T, N, A), np.int8),
    )


def assert_same(rec, got, what, flag_mask=0xFF, policy_actions=False, reward_rtol=0.0):
    """Exact comparison (rewards exact unless ``reward_rtol``); obs / info only where the
    reference returned them; state arrays only if the recording has them (cassettes do not)."""
    for k in ("x", "y", "flags", "step", "env_flags"):
        if k in rec:
            bad = np.argwhere(rec[k] != got[k])
            assert bad.size == 0, f"{what}: {k} first differs at (t,n[,a])={bad[0]}: want {rec[k][tuple(bad[0])]} got {got[k][tuple(bad[0])]}"
    bad = np.argwhere((rec["agent_flags"] & flag_mask) != (got["agent_flags"] & flag_mask))
    assert bad.size == 0, f"{what}: agent_flags first differ at {bad[0]}: want {rec['agent_flags'][tuple(bad[0])]:#x} got {got['agent_flags'][tuple(bad[0])]:#x}"
    if reward_rtol == 0.0:
        assert np.array_equal(rec["reward"], got["reward"]), f"{what}: rewards differ (exact float64 comparison)"
    else:
        np.testing.assert_allclose(got["reward"], rec["reward"], rtol=reward_rtol, atol=0, err_msg=f"{what}: rewards")
    present = (rec["agent_flags"] & _abi.O_OBS_PRESENT) != 0
    assert np.array_equal(rec["agent_info"][present], got["agent_info"][present]), f"{what}: infos differ"
    assert np.array_equal(rec["obs"][present], got["obs"][present]), f"{what}: observations differ"
    assert np.array_equal(rec["init_obs"], got["init_obs"]), f"{what}: reset observations differ"
    if policy_actions:
        assert np.array_equal(rec["actions"], got["actions"]), f"{what}: policy actions differ"


def random_states(cfg, n, rng, step_hi=None):
    """Seeded reset states from the oracle's numpy-exact ``reset(seed)`` (int64 seeds)."""
    import oracle

    low = lower_config(cfg)
    o = oracle.OracleEnvs(low, n)
    o.reset_seeded(rng.integers(0, 2**40, size=n))
    if step_hi:
        o.step_count[:] = rng.integers(0, step_hi, size=n)
    return o.get_state()
