"""Trajectory cassettes in the reference's VCR JSON schema.

The reference records trajectories as JSON "cassettes" and replays them as its regression test
(``tests/collectivecrossing/envs/test_trajectory_vcr.py:40-196``).  This module writes and replays
the same schema for any env with the reference's dict API — the single-env facade
(:class:`collectivecrossing_b200.CollectiveCrossingEnv`, every step on the GPU) — and exports one
env of a batched device rollout in that schema, so that a cassette recorded here can be fed to the
reference's own ``TrajectoryVCR.replay_trajectory`` and vice versa.

Schema (one JSON object)::

    config                 the env config fields, strategy configs as ``model_dump()``
    initial_observations   {agent id: [floats]}      after ``reset(seed=42)``
    initial_infos          {agent id: {...}}
    steps[]                step, actions, active_actions, observations,
                           next_observations, next_rewards, next_terminated, next_truncated, next_infos
"""

from __future__ import annotations

import json
from pathlib import Path
from typing import Any

import numpy as np

from . import _abi

RESET_SEED = 42  # the reference records and replays with reset(seed=42) (test_trajectory_vcr.py:56,147)


def _plain(v: Any) -> Any:
    return bool(v) if isinstance(v, (bool, np.bool_)) else v


def _config_block(config: Any) -> dict:
    """test_trajectory_vcr.py:58-74"""
    return {
        "width": config.width, "height": config.height, "division_y": config.division_y,
        "tram_door_left": config.tram_door_left, "tram_door_right": config.tram_door_right,
        "tram_length": config.tram_length, "num_boarding_agents": config.num_boarding_agents,
        "num_exiting_agents": config.num_exiting_agents,
        "exiting_destination_area_y": config.exiting_destination_area_y,
        "boarding_destination_area_y": config.boarding_destination_area_y,
        "render_mode": config.render_mode,
        "reward_config": config.reward_config.model_dump(),
        "terminated_config": config.terminated_config.model_dump(),
        "truncated_config": config.truncated_config.model_dump(),
    }


def _is_active(env: Any, agent_id: str) -> bool:
    agents = getattr(env, "_agents", None)
    if agents is not None:
        return agent_id in agents and bool(agents[agent_id].active)
    return bool(env.active.get(agent_id, False))  # oracle.pyport.PyEnv keeps plain dicts


def _config_of(env: Any) -> Any:
    return env.config if hasattr(env, "config") else env.c


def record_trajectory(env: Any, actions_sequence: list[dict[str, int]], path: str | Path | None = None, seed: int = RESET_SEED) -> dict:
    """Run ``actions_sequence`` on ``env`` from ``reset(seed)`` and capture everything the
    reference's recorder captures (test_trajectory_vcr.py:40-121).  Actions of inactive agents are
    filtered out before stepping, as there; the recording stops when the episode is over."""
    observations, infos = env.reset(seed=seed)
    traj: dict = {
        "config": _config_block(_config_of(env)),
        "initial_observations": {k: np.asarray(v).tolist() for k, v in observations.items()},
        "initial_infos": {k: {ik: _plain(iv) for ik, iv in v.items()} for k, v in infos.items()},
        "steps": [],
    }
    for step_num, actions in enumerate(actions_sequence):
        active_actions = {a: int(act) for a, act in actions.items() if _is_active(env, a)}
        step = {"step": step_num, "actions": {a: int(v) for a, v in actions.items()}, "active_actions": active_actions,
                "observations": {k: np.asarray(v).tolist() for k, v in observations.items()}}
        observations, rewards, terminated, truncated, infos = env.step(active_actions)
        step["next_observations"] = {k: np.asarray(v).tolist() for k, v in observations.items()}
        step["next_rewards"] = {k: float(v) for k, v in rewards.items()}
        step["next_terminated"] = {k: bool(v) for k, v in terminated.items()}
        step["next_truncated"] = {k: bool(v) for k, v in truncated.items()}
        step["next_infos"] = {k: {ik: _plain(iv) for ik, iv in v.items()} for k, v in infos.items()}
        traj["steps"].append(step)
        if terminated.get("__all__", False) or truncated.get("__all__", False):
            break
    if path is not None:
        Path(path).write_text(json.dumps(traj, indent=2))
    return traj


def load_cassette(path: str | Path) -> dict:
    return json.loads(Path(path).read_text())


def replay_trajectory(env: Any, cassette: dict | str | Path, seed: int = RESET_SEED, strict: bool = False) -> dict:
    """Replay a cassette on ``env`` and assert what the reference's replay asserts
    (test_trajectory_vcr.py:123-196): initial and per-step observations, rewards within 1e-6,
    terminated flags.  ``strict=True`` additionally requires identical key sets, exactly equal
    rewards, truncated flags and infos (the reference's stricter golden comparison, :494-593)."""
    traj = load_cassette(cassette) if not isinstance(cassette, dict) else cassette
    observations, infos = env.reset(seed=seed)
    for agent_id, expected in traj["initial_observations"].items():
        assert agent_id in observations, f"Agent {agent_id} missing in replay"
        np.testing.assert_array_equal(observations[agent_id], np.array(expected), err_msg=f"Initial observation mismatch for {agent_id}")
    if strict:
        assert {k: {ik: _plain(iv) for ik, iv in v.items()} for k, v in infos.items()} == traj["initial_infos"], "initial infos differ"
    for step in traj["steps"]:
        n = step["step"]
        for agent_id, expected in step["observations"].items():
            if agent_id in observations:
                np.testing.assert_array_equal(observations[agent_id], np.array(expected), err_msg=f"Step {n} observation mismatch for {agent_id}")
        # the reference replays the ORIGINAL actions (:162,174); its recorder stepped the filtered ones
        # (:96) — equivalent, because an inactive agent does not move (collectivecrossing.py:398)
        observations, rewards, terminated, truncated, infos = env.step(step["actions"] if not strict else step["active_actions"])
        for agent_id, expected in step["next_observations"].items():
            if agent_id in observations:
                np.testing.assert_array_equal(observations[agent_id], np.array(expected), err_msg=f"Step {n} next observation mismatch for {agent_id}")
        for agent_id, expected in step["next_rewards"].items():
            if agent_id in rewards:
                assert abs(rewards[agent_id] - expected) < 1e-6, f"Step {n} reward mismatch for {agent_id}"
        for agent_id, expected in step["next_terminated"].items():
            if agent_id in terminated:
                assert terminated[agent_id] == expected, f"Step {n} termination mismatch for {agent_id}"
        if strict:
            assert set(observations) == set(step["next_observations"]), f"Step {n}: observation keys differ"
            assert {k: float(v) for k, v in rewards.items()} == step["next_rewards"], f"Step {n}: rewards differ"
            assert {k: bool(v) for k, v in terminated.items()} == step["next_terminated"], f"Step {n}: terminateds differ"
            assert {k: bool(v) for k, v in truncated.items()} == step["next_truncated"], f"Step {n}: truncateds differ"
            assert {k: {ik: _plain(iv) for ik, iv in v.items()} for k, v in infos.items()} == step["next_infos"], f"Step {n}: infos differ"
    return traj


# ---- one env of a batched device rollout, in the same schema ----------------------------------------
class BatchedTrajectoryRecorder:
    """Records env ``env_index`` of a :class:`BatchedCollectiveCrossing` while the caller steps the
    batch, and renders the recording in the cassette schema.  Only that env's rows cross PCIe
    (a few hundred bytes per step).

        rec = BatchedTrajectoryRecorder(batched, env_index=17); rec.begin(obs)
        out = batched.step(policy="greedy"); rec.after_step(out)   # ... repeat
        cassette = rec.cassette()
    """

    def __init__(self, env: Any, env_index: int = 0):
        if env.obs is None or env.agent_info is None:
            raise ValueError("recording needs obs_dtype != 'none' and with_info=True")
        self.env, self.k = env, int(env_index)
        cfg = env.config
        self.ids = [f"boarding_{i}" for i in range(cfg.num_boarding_agents)] + [f"exiting_{i}" for i in range(cfg.num_exiting_agents)]
        self.types = ["boarding"] * cfg.num_boarding_agents + ["exiting"] * cfg.num_exiting_agents
        self.traj: dict | None = None
        self._obs: dict = {}
        self._active: list[bool] = []

    def _rows(self, t):
        return t[self.k].detach().cpu().numpy()

    def begin(self, obs=None) -> None:
        """Call after ``reset`` / ``reset_seeded`` / ``set_state`` (``obs``: that call's observation tensor)."""
        rows = self._rows(self.env.observe() if obs is None else obs).astype(np.float32)
        self._obs = {a: rows[i].tolist() for i, a in enumerate(self.ids)}
        self._active = [bool(f & _abi.F_ACTIVE) for f in self._rows(self.env.flags)]
        self.traj = {"config": _config_block(self.env.config), "initial_observations": dict(self._obs),
                     "initial_infos": {a: {"agent_type": self.types[i]} for i, a in enumerate(self.ids)}, "steps": []}

    def after_step(self, out: Any) -> bool:
        """Append the step that produced ``out`` (a ``StepOutput``); returns True when the episode ended."""
        assert self.traj is not None, "call begin() first"
        acts = self._rows(out.actions)
        af, ai, rew = self._rows(out.agent_flags), self._rows(out.agent_info), self._rows(out.reward)
        ef = int(out.env_flags[self.k])
        rows = self._rows(out.obs).astype(np.float32)
        actions = {a: int(acts[i]) for i, a in enumerate(self.ids)}
        step = {"step": len(self.traj["steps"]), "actions": actions,
                "active_actions": {a: v for i, (a, v) in enumerate(actions.items()) if self._active[i]},
                "observations": dict(self._obs), "next_observations": {}, "next_rewards": {}, "next_terminated": {},
                "next_truncated": {}, "next_infos": {}}
        for i, a in enumerate(self.ids):
            bits = int(af[i])
            step["next_terminated"][a] = bool(bits & _abi.O_TERM_VALUE)
            if bits & _abi.O_ALIVE_PREV:
                step["next_rewards"][a] = float(rew[i])
                step["next_truncated"][a] = bool(bits & _abi.O_TRUNC_VALUE)
            if bits & _abi.O_OBS_PRESENT and not (ef & _abi.E_WAS_RESET):
                step["next_observations"][a] = rows[i].tolist()
            if bits & _abi.O_OBS_PRESENT:
                step["next_infos"][a] = {"agent_type": self.types[i], "in_tram_area": bool(ai[i] & _abi.I_IN_TRAM_AREA),
                                         "at_door": bool(ai[i] & _abi.I_AT_DOOR), "active": bool(ai[i] & _abi.I_ACTIVE),
                                         "at_destination": bool(ai[i] & _abi.I_AT_DESTINATION)}
        step["next_terminated"]["__all__"] = bool(ef & _abi.E_TERMINATED_ALL)
        step["next_truncated"]["__all__"] = bool(ef & _abi.E_TRUNCATED_ALL)
        self.traj["steps"].append(step)
        self._obs = dict(step["next_observations"])
        self._active = [bool(b & _abi.O_ACTIVE) for b in af]
        return step["next_terminated"]["__all__"] or step["next_truncated"]["__all__"]

    def cassette(self, path: str | Path | None = None) -> dict:
        assert self.traj is not None
        if path is not None:
            Path(path).write_text(json.dumps(self.traj, indent=2))
        return self.traj
