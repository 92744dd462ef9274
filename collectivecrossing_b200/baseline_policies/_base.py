"""Shared implementation of the two baseline policies."""

from __future__ import annotations

from typing import Any

import numpy as np

from ..actions import ACTION_TO_DIRECTION


class _DevicePolicy:
    """``randomness_factor`` (epsilon) > 0 mixes in uniformly random VALID actions drawn from
    ``np.random.RandomState(seed)`` like the reference (greedy_policy.py:46-60); the greedy /
    waiting decision (epsilon branch not taken) comes from the device kernel ``cc_policy_actions``."""

    kind = "greedy"

    def __init__(self, randomness_factor: float, seed: int) -> None:
        self.randomness_factor = randomness_factor
        self.random_state = np.random.RandomState(seed)

    def _is_valid_action(self, agent_id: str, action: int, env: Any) -> bool:
        if action == 4:
            return True
        cur = env._get_agent_position(agent_id)
        return env._is_move_valid(agent_id, cur, cur + ACTION_TO_DIRECTION[action])

    def get_action(self, agent_id: str, observation: Any, env: Any) -> int:
        if self.randomness_factor > 0.0 and self.random_state.random() < self.randomness_factor:
            valid = [a for a in range(5) if self._is_valid_action(agent_id, a, env)]
            return int(self.random_state.choice(valid)) if valid else 4
        return env.baseline_actions(self.kind).get(agent_id, 4)
