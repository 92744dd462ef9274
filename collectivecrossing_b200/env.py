"""Single-env facade with the reference's RLlib ``MultiAgentEnv`` dict interface.

``CollectiveCrossingEnv(config)`` is a drop-in for the reference class
(``collectivecrossing.py:30-783``) for everything on the step/reset path: every ``step`` is one
launch of the fused CUDA kernel on a one-env batch (float64 rewards and float32 observations
come back exactly as the reference computes them), ``reset(seed=...)`` runs the numpy-exact
PCG64 placement kernel.  What stays on the host is bookkeeping: the ``_agents`` records tests and
policies read or overwrite (state injection), dict assembly, argument validation.

Rendering: ``render()`` returns an ``rgb_array`` from a numpy rasteriser (``rendering.py``); the
interactive matplotlib ``human`` mode is not provided.  Known divergence: an invalid action raises ``ValueError`` BEFORE anything moves,
while the reference raises midway through its move loop (collectivecrossing.py:197-202) leaving
earlier agents moved and the step counter bumped.
"""

from __future__ import annotations

import secrets
from typing import Any

import numpy as np
import torch

from . import _abi, _spaces
from .batched import BatchedCollectiveCrossing
from .lowering import lower_config
from .types import Agent, AgentType
from .utils.geometry import TramBoundaries, calculate_tram_boundaries

try:  # RLlib base class when ray is installed, so isinstance checks of RLlib pass
    from ray.rllib.env.multi_agent_env import MultiAgentEnv as _Base  # type: ignore
except Exception:  # ray absent (this image)
    class _Base:  # type: ignore[no-redef]
        def __init__(self) -> None:
            pass


_VALID_ACTIONS = (0, 1, 2, 3, 4)


class CollectiveCrossingEnv(_Base):
    """Tram boarding / exiting grid world; see the module docstring."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 4}

    def __init__(self, config: Any, device: Any = "cuda:0"):
        if isinstance(config, dict):  # RLlib passes env_config dicts (examples/training_script.py:28)
            from .configs import CollectiveCrossingConfig

            config = CollectiveCrossingConfig(**config)
        self._config = config
        self._cfg = lower_config(config)  # ValueError for unknown strategy names, like the reference
        self._tram_boundaries = calculate_tram_boundaries(config)
        self._ids = [f"boarding_{i}" for i in range(config.num_boarding_agents)] + [
            f"exiting_{i}" for i in range(config.num_exiting_agents)]
        self._index = {a: k for k, a in enumerate(self._ids)}
        self._step_count = 0
        self._agents: dict[str, Agent] = self._create_dummy_agents()
        self._dev = BatchedCollectiveCrossing(config, 1, device, obs_dtype="float32", reward_dtype="float64",
                                              auto_reset=False, with_info=True)
        self._host = self._dev.make_host_buffers(pinned=False)
        self._host["order"] = torch.zeros((1, len(self._ids)), dtype=torch.int8)
        # the env's generator lives on the device (cc_reset_seeded); `np_random` hands out a numpy view of it (see the property)
        self._np_random: np.random.Generator | None = None
        # the reference's strategy objects (collectivecrossing.py:69-78); their values come from the device
        from .observations import get_observation_function
        from .rewards import get_reward_function
        from .terminateds import get_terminated_function
        from .truncateds import get_truncated_function

        self._observation_function = get_observation_function(config.observation_config)
        self._reward_function = get_reward_function(config.reward_config)
        self._terminated_function = get_terminated_function(config.terminated_config)
        self._truncated_function = get_truncated_function(config.truncated_config)
        self._evaluators: dict = {}
        self._setup_spaces()
        super().__init__()
        self._agents_truncated_or_terminated_this_step: set[str] = set()

    # ---- reference properties (collectivecrossing.py:414-452, 743-783) ------------------------------
    @property
    def config(self):
        return self._config

    @property
    def tram_boundaries(self) -> TramBoundaries:
        return self._tram_boundaries

    tram_door_left = property(lambda self: self._tram_boundaries.tram_door_left)
    tram_door_right = property(lambda self: self._tram_boundaries.tram_door_right)
    tram_left = property(lambda self: self._tram_boundaries.tram_left)
    tram_right = property(lambda self: self._tram_boundaries.tram_right)

    @property
    def action_spaces(self):
        return self._action_spaces

    @property
    def observation_spaces(self):
        return self._observation_spaces

    def get_observation_space(self, agent_id: str):
        return self.observation_space

    def get_action_space(self, agent_id: str):
        return self.action_space

    @property
    def agents(self) -> list[str]:
        return [a for a in self._ids if not self._agents[a].terminated and not self._agents[a].truncated]

    @property
    def possible_agents(self) -> list[str]:
        return list(self._ids)

    @property
    def np_random(self) -> np.random.Generator:
        """gymnasium's ``env.np_random``: in the reference THE generator ``reset()`` places agents with
        (collectivecrossing.py:95,105-106,134-137).  Here that generator lives on the device; this property returns a numpy
        ``Generator(PCG64)`` set to the device generator's current state, and the next unseeded ``reset()`` pushes the state
        back first — so user code that draws from ``env.np_random`` between resets shifts the placement stream exactly as it
        does in the reference."""
        if not getattr(self, "_seeded", False):   # gymnasium seeds from OS entropy on first use; the next reset() takes it over
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
            self._seeded = True
        if self._np_random is None:
            raw = self._dev.get_state()["rng"].cpu().numpy().astype(np.uint64)[0]      # hi, lo, inc hi, inc lo, buffered, has_buffered
            bg = np.random.PCG64()
            bg.state = {"bit_generator": "PCG64", "state": {"state": (int(raw[0]) << 64) | int(raw[1]), "inc": (int(raw[2]) << 64) | int(raw[3])},
                        "has_uint32": int(raw[5]), "uinteger": int(raw[4])}
            self._np_random = np.random.Generator(bg)
        return self._np_random

    def _push_np_random(self) -> None:
        """The host view of the generator (if one was handed out) back to the device, before the device draws from it."""
        if self._np_random is None:
            return
        st = self._np_random.bit_generator.state
        m = (1 << 64) - 1
        raw = np.array([[st["state"]["state"] >> 64, st["state"]["state"] & m, st["state"]["inc"] >> 64, st["state"]["inc"] & m,
                         st["uinteger"], st["has_uint32"]]], dtype=np.uint64)
        state = self._dev.get_state()
        state["rng"] = torch.from_numpy(raw.view(np.int64)).to(self._dev.device)
        self._dev.load_state(state)
        self._np_random = None

    def close(self) -> None:
        self._dev.close()
        for ev in self._evaluators.values():
            ev.close()
        self._evaluators = {}

    def render(self, mode: str | None = None):
        """``rgb_array`` image of the current host view (numpy rasteriser, ``rendering.py``); the
        reference's interactive ``human`` mode needs matplotlib and is not provided."""
        mode = mode or getattr(self._config, "render_mode", None) or "rgb_array"
        if mode != "rgb_array":
            raise NotImplementedError("only render_mode='rgb_array' is available (no matplotlib in this stack)")
        from .rendering import render_env

        return render_env(self)

    # ---- construction helpers ----------------------------------------------------------------------
    def _create_dummy_agents(self) -> dict[str, Agent]:
        """Agents with ids and types but ``[None, None]`` positions, so spaces exist before the
        first reset (collectivecrossing.py:301-343)."""
        return {
            a: Agent(id=a, agent_type=AgentType.BOARDING if a.startswith("boarding") else AgentType.EXITING,
                     position=np.array([None, None]), active=True, terminated=False, truncated=False)
            for a in self._ids
        }

    def _setup_spaces(self) -> None:
        c = self._config
        n_obs = 6 + 4 * len(self._ids)
        self._action_spaces = {a: _spaces.Discrete(5) for a in self._ids}
        # declared bound max(w, h) - 1 although x can equal w: kept as in the reference (observations.py:113-118)
        self._observation_spaces = {
            a: _spaces.Box(low=-1, high=max(c.width, c.height) - 1, shape=(n_obs,), dtype=np.float32) for a in self._ids}
        if self._ids:
            self.action_space = self._action_spaces[self._ids[0]]
            self.observation_space = self._observation_spaces[self._ids[0]]

    # ---- host view <-> device state ----------------------------------------------------------------
    def _push_state(self, dev: Any = None, step_offset: int = 0, allow_unset: bool = False) -> None:
        """Upload the host records (tests and policies may have edited them) to the device.  ``allow_unset``: agents
        without a position yet (before the first reset) go to (0, 0) — the reference's truncation function is callable
        then (tests/collectivecrossing/envs/test_truncateds.py:30-53) and does not look at positions."""
        dev = dev or self._dev
        A = len(self._ids)
        x, y, f = np.zeros((1, A), np.int8), np.zeros((1, A), np.int8), np.zeros((1, A), np.uint8)
        for k, a in enumerate(self._ids):
            ag = self._agents[a]
            if ag.position[0] is None:
                if not allow_unset:
                    raise RuntimeError("call reset() before step()")
                continue
            x[0, k], y[0, k] = int(ag.position[0]), int(ag.position[1])
            f[0, k] = (_abi.F_ACTIVE if ag.active else 0) | (_abi.F_TERMINATED if ag.terminated else 0) | (_abi.F_TRUNCATED if ag.truncated else 0)
        dev.set_state(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(f),
                      torch.tensor([self._step_count + step_offset], dtype=torch.int32))

    def _pull_state(self) -> None:
        x, y, f = self._dev.x.cpu().numpy()[0], self._dev.y.cpu().numpy()[0], self._dev.flags.cpu().numpy()[0]
        self._step_count = int(self._dev.step_count.cpu()[0])
        for k, a in enumerate(self._ids):
            ag = self._agents[a]
            ag.position = np.array([int(x[k]), int(y[k])])
            ag.active = bool(f[k] & _abi.F_ACTIVE)
            ag.terminated = bool(f[k] & _abi.F_TERMINATED)
            ag.truncated = bool(f[k] & _abi.F_TRUNCATED)

    # ---- reset (collectivecrossing.py:91-159) ------------------------------------------------------
    def reset(self, *, seed: int | None = None, options: dict | None = None):
        if seed is None and not getattr(self, "_seeded", False):
            seed = secrets.randbits(63)  # gymnasium would draw OS entropy for an unseeded first reset
        if seed is not None:
            if not 0 <= int(seed) < 2**63:
                raise ValueError("seed must be a non-negative integer below 2**63")
            self._np_random = None      # a fresh generator, like gymnasium's reset(seed=...)
            obs = self._dev.reset_seeded(torch.tensor([int(seed)], dtype=torch.int64, device=self._dev.device))
            self._seeded = True
        else:
            self._push_np_random()      # draws the caller made from env.np_random count, like in the reference
            obs = self._dev.reset_seeded(None)  # keep drawing from the env's generator, like gymnasium
        self._dev.check_error()
        self._agents = self._create_dummy_agents()
        self._pull_state()
        rows = obs.cpu().numpy()[0]
        observations = {a: rows[k].copy() for k, a in enumerate(self._ids)}
        infos = {a: {"agent_type": self._agents[a].agent_type.value} for a in self._ids}
        return observations, infos

    # ---- step (collectivecrossing.py:161-261) ------------------------------------------------------
    def _check_action_and_agent_validity(self, agent_id: str, action: int) -> None:
        if agent_id not in self._agents:
            raise ValueError(
                f"Unknown agent ID: {agent_id} in action_dict. The action_dict keys must be a subset of the agents. "
                f"Current agents: {self._agents.keys()}")
        if isinstance(action, (bool, np.bool_)) or action not in _VALID_ACTIONS:
            raise ValueError(f"Invalid action: {action} for agent {agent_id}. Valid actions are: {list(_VALID_ACTIONS)}")

    def step(self, action_dict: dict[str, int]):
        for agent_id, action in action_dict.items():
            self._check_action_and_agent_validity(agent_id, action)
        self._push_state()
        host = self._host
        host["actions"].fill_(4)
        host["order"].fill_(-1)
        for pos, (agent_id, action) in enumerate(action_dict.items()):  # the dict order is the move order
            k = self._index[agent_id]
            host["actions"][0, k] = int(action)
            host["order"][0, pos] = k
        dev = self._dev
        out = dev.step(host["actions"].to(dev.device), order=host["order"].to(dev.device))
        dev.check_error()
        self._pull_state()
        rows = out.obs.cpu().numpy()[0]
        rew = out.reward.cpu().numpy()[0]
        af = out.agent_flags.cpu().numpy()[0]
        ai = out.agent_info.cpu().numpy()[0]
        ef = int(out.env_flags.cpu()[0])

        observations, rewards, terminateds, truncateds, infos = {}, {}, {}, {}, {}
        self._agents_truncated_or_terminated_this_step = set()
        for k, a in enumerate(self._ids):
            bits = int(af[k])
            terminateds[a] = bool(bits & _abi.O_TERM_VALUE)
            if bits & _abi.O_ALIVE_PREV:
                rewards[a] = float(rew[k])
                truncateds[a] = bool(bits & _abi.O_TRUNC_VALUE)
            if bits & _abi.O_OBS_PRESENT:
                observations[a] = rows[k].copy()
                infos[a] = {
                    "agent_type": self._agents[a].agent_type.value,
                    "in_tram_area": bool(ai[k] & _abi.I_IN_TRAM_AREA), "at_door": bool(ai[k] & _abi.I_AT_DOOR),
                    "active": bool(ai[k] & _abi.I_ACTIVE), "at_destination": bool(ai[k] & _abi.I_AT_DESTINATION),
                }
                if bits & (_abi.O_TERMINATED | _abi.O_TRUNCATED):
                    self._agents_truncated_or_terminated_this_step.add(a)
        terminateds["__all__"] = bool(ef & _abi.E_TERMINATED_ALL)
        truncateds["__all__"] = bool(ef & _abi.E_TRUNCATED_ALL)
        return observations, rewards, terminateds, truncateds, infos

    # ---- strategy values of the current state (collectivecrossing.py:590-633) ----------------------
    def _evaluate(self, reward_config: Any = None, terminated_config: Any = None, truncated_config: Any = None) -> dict:
        """``{"rewards", "terminateds", "truncateds"}`` of the CURRENT host view, as the reference's
        ``calculate_reward / calculate_terminated / calculate_truncated`` return them for every agent
        (absent key = ``None``).  Computed by the step kernel: a scratch one-env handle with the (possibly
        overridden) strategy configs takes ONE step with WAIT actions from this state at step count - 1,
        which evaluates exactly these functions on unchanged positions (collectivecrossing.py:204-241)."""
        key = (repr(reward_config), repr(terminated_config), repr(truncated_config))
        ev = self._evaluators.get(key)
        if ev is None:
            upd = {k: v for k, v in (("reward_config", reward_config), ("terminated_config", terminated_config),
                                     ("truncated_config", truncated_config)) if v is not None}
            cfg = self._config.model_copy(update=upd) if upd else self._config
            ev = self._evaluators[key] = BatchedCollectiveCrossing(cfg, 1, self._dev.device, obs_dtype="none", reward_dtype="float64",
                                                                   auto_reset=False)
        self._push_state(ev, step_offset=-1, allow_unset=True)
        wait = torch.full((1, len(self._ids)), 4, dtype=torch.int8, device=ev.device)
        out = ev.step(wait)
        ev.check_error()
        rew, af = out.reward.cpu().numpy()[0], out.agent_flags.cpu().numpy()[0]
        res: dict = {"rewards": {}, "terminateds": {}, "truncateds": {}}
        for k, a in enumerate(self._ids):
            bits = int(af[k])
            res["terminateds"][a] = bool(bits & _abi.O_TERM_VALUE)
            if bits & _abi.O_ALIVE_PREV:
                res["rewards"][a] = float(rew[k])
                res["truncateds"][a] = bool(bits & _abi.O_TRUNC_VALUE)
        return res

    def _calculate_reward(self, agent_id: str) -> float | None:
        return self._reward_function.calculate_reward(agent_id, self)

    def _calculate_terminated(self, agent_id: str) -> bool | None:
        return self._terminated_function.calculate_terminated(agent_id, self)

    def _calculate_truncated(self, agent_id: str) -> bool | None:
        return self._truncated_function.calculate_truncated(agent_id, self)

    # ---- predicates and accessors used by tests and the baseline policies ---------------------------
    def _get_agent(self, agent_id: str) -> Agent:
        if agent_id not in self._agents:
            raise ValueError(f"Unknown agent ID: {agent_id}")
        return self._agents[agent_id]

    def _get_agent_position(self, agent_id: str) -> np.ndarray:
        return self._get_agent(agent_id).position

    def _get_agents_by_type(self, agent_type: AgentType) -> list[Agent]:
        return [a for a in self._agents.values() if a.agent_type == agent_type]

    def _get_boarding_agents(self) -> list[Agent]:
        return self._get_agents_by_type(AgentType.BOARDING)

    def _get_exiting_agents(self) -> list[Agent]:
        return self._get_agents_by_type(AgentType.EXITING)

    def _get_agent_observation(self, agent_id: str) -> np.ndarray:
        """Observation of the CURRENT host view, computed by the device observe kernel."""
        self._push_state()
        return self._dev.observe().cpu().numpy()[0, self._index[agent_id]].copy()

    def _is_valid_position(self, pos) -> bool:  # collectivecrossing.py:509-534
        x, y, c = pos[0], pos[1], self._config
        if not (0 <= x <= c.width and 0 <= y <= c.height):
            return False
        if y == c.division_y and not (self.tram_door_left < x < self.tram_door_right):
            return False
        if y >= c.division_y and not (self.tram_right > x > self.tram_left):
            return False
        return True

    def _is_position_occupied(self, pos, exclude_agent: str | None = None) -> bool:  # :536-541
        return any(a != exclude_agent and ag.active and np.array_equal(ag.position, pos) for a, ag in self._agents.items())

    def _would_hit_tram_wall(self, current_pos, new_pos) -> bool:  # :565-588
        x, y, d = new_pos[0], new_pos[1], self._config.division_y
        if y == d:
            return not (self.tram_door_left < x < self.tram_door_right)
        return bool(y > d and (x == self.tram_left or x == self.tram_right))

    def _is_move_valid(self, agent_id: str, current_pos, new_pos) -> bool:  # :345-369
        return (self._is_valid_position(new_pos) and not self._is_position_occupied(new_pos, exclude_agent=agent_id)
                and not self._would_hit_tram_wall(current_pos, new_pos))

    def is_in_boarding_destination_area(self, agent_id: str) -> bool:
        return bool(self._get_agent_position(agent_id)[1] == self._config.boarding_destination_area_y)

    def is_in_exiting_destination_area(self, agent_id: str) -> bool:
        return bool(self._get_agent_position(agent_id)[1] == self._config.exiting_destination_area_y)

    def is_in_tram_area(self, agent_id: str) -> bool:
        p = self._get_agent_position(agent_id)
        return bool(p[1] >= self._config.division_y and self.tram_left <= p[0] <= self.tram_right)

    def is_at_tram_door(self, agent_id: str) -> bool:
        p = self._get_agent_position(agent_id)
        return bool(p[1] == self._config.division_y and (p[0] == self.tram_door_left - 1 or p[0] == self.tram_door_right + 1))

    def get_agent_destination_position(self, agent_id: str):
        c = self._config
        return (None, c.boarding_destination_area_y if self._agents[agent_id].is_boarding else c.exiting_destination_area_y)

    def has_agent_reached_destination(self, agent_id: str) -> bool:
        if self._agents[agent_id].is_boarding:
            return self.is_in_boarding_destination_area(agent_id)
        return self.is_in_exiting_destination_area(agent_id)

    def baseline_actions(self, policy: str) -> dict[str, int]:
        """Actions of the on-device greedy / waiting policy (epsilon 0) for every live, active agent
        of the current host view — the loop of scripts/run_greedy_policy_demo.py:71-77."""
        self._push_state()
        acts = self._dev.policy_actions(policy).cpu().numpy()[0]
        return {a: int(acts[k]) for k, a in enumerate(self._ids)
                if self._agents[a].active and not self._agents[a].terminated and not self._agents[a].truncated}
