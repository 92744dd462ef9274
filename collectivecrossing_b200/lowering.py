"""Lower a ``CollectiveCrossingConfig`` to the POD ``cc_config`` the kernels read.

Done once per environment: absolute tram geometry (reference ``utils/geometry.py:33-40``),
strategy registry look-ups (reference ``rewards.py:194-216``, ``terminateds.py:92-114``,
``truncateds.py:106-128``, ``observations.py:127-149`` — unknown names raise ``ValueError`` with
the reference's wording) and the reward parameters as float64.
"""

from __future__ import annotations

from typing import Any

from . import _abi
from .utils.geometry import calculate_tram_boundaries

_REWARD_FIELDS = {
    "default": ("boarding_destination_reward", "tram_door_reward", "tram_area_reward", "distance_penalty_factor"),
    "simple_distance": ("distance_penalty_factor",),
    "binary": ("goal_reward", "no_goal_reward"),
    "constant_negative": ("step_penalty",),
}
_TRUNCATED_FUNCTIONS = ("max_steps", "custom")  # reference truncateds.py:99-102 registers both
_OBSERVATION_FUNCTIONS = ("default",)


def _lookup(name: str, known: Any, noun: str) -> str:
    if name not in known:
        raise ValueError(f"Unknown {noun} function '{name}'. Available: {', '.join(known)}")
    return name


def lower_config(config: Any) -> _abi.CCConfig:
    """``CollectiveCrossingConfig`` -> ``cc_config``.  Raises ``ValueError`` for strategy names
    without a registered function (e.g. the ``custom`` placeholders), like the reference does when
    the env is constructed (collectivecrossing.py:69-78)."""
    _lookup(config.observation_config.get_observation_function_name(), _OBSERVATION_FUNCTIONS, "observation")
    reward = _lookup(config.reward_config.get_reward_function_name(), _abi.REWARD_KINDS, "reward")
    term = _lookup(config.terminated_config.get_terminated_function_name(), _abi.TERMINATED_KINDS, "termination")
    _lookup(config.truncated_config.get_truncated_function_name(), _TRUNCATED_FUNCTIONS, "truncation")

    tb = calculate_tram_boundaries(config)
    out = _abi.CCConfig()
    out.width, out.height, out.division_y = config.width, config.height, config.division_y
    out.tram_left, out.tram_right = tb.tram_left, tb.tram_right
    out.door_left, out.door_right = tb.tram_door_left, tb.tram_door_right
    out.boarding_dest_y = config.boarding_destination_area_y
    out.exiting_dest_y = config.exiting_destination_area_y
    out.num_boarding, out.num_exiting = config.num_boarding_agents, config.num_exiting_agents
    out.max_steps = config.truncated_config.max_steps
    out.reward_kind = _abi.REWARD_KINDS[reward]
    out.terminated_kind = _abi.TERMINATED_KINDS[term]
    params = [float(getattr(config.reward_config, f)) for f in _REWARD_FIELDS[reward]]
    for k in range(4):
        out.reward_params[k] = params[k] if k < len(params) else 0.0
    n_agents = out.num_boarding + out.num_exiting
    if not 1 <= n_agents <= _abi.MAX_AGENTS:
        raise ValueError(f"the B200 kernels support 1..{_abi.MAX_AGENTS} agents per env, got {n_agents}")
    for name in ("width", "height", "division_y", "tram_left", "tram_right", "door_left", "door_right",
                 "boarding_dest_y", "exiting_dest_y"):
        if not -1 <= getattr(out, name) <= 120:
            raise ValueError(f"{name}={getattr(out, name)} is outside the lattice the device tables support (0..120)")
    return out


def describe(cfg: _abi.CCConfig) -> dict:
    """Plain-dict view of a lowered config (logging, bench JSON)."""
    d = {name: getattr(cfg, name) for name, _ in cfg._fields_ if name != "reward_params"}
    d["reward_params"] = list(cfg.reward_params)
    return d
