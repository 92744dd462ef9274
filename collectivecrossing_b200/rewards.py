"""Reward functions — the reference's ``rewards`` module surface (rewards.py:16-216).

``calculate_reward(agent_id, env)`` returns what the reference's function returns for the env's
current state — ``None`` for a terminated or truncated agent (rewards.py:65-66), else the float64
reward — computed by the step kernel (``cc_kernels.cuh``: reward tables / float64 path), not on
the host.  A function object carries its own ``reward_config`` and evaluates with it, whatever
the env was configured with.
"""

from __future__ import annotations

from typing import Any

from ._strategies import evaluate, is_done, registry_get


class RewardFunction:
    """Base class (rewards.py:16-38)."""

    def __init__(self, reward_config: Any):
        self.reward_config = reward_config

    def calculate_reward(self, agent_id: str, env: Any) -> float | None:
        if is_done(agent_id, env):   # rewards.py:65-66: no arithmetic, answered without a launch (any env object)
            return None
        return evaluate(env, reward_config=self.reward_config)["rewards"].get(agent_id)


class DefaultRewardFunction(RewardFunction):
    """rewards.py:41-99"""


class SimpleDistanceRewardFunction(RewardFunction):
    """rewards.py:102-129"""


class BinaryRewardFunction(RewardFunction):
    """rewards.py:132-159 (never pays ``goal_reward``: quirk of the reference, reproduced)"""


class ConstantNegativeRewardFunction(RewardFunction):
    """rewards.py:162-182"""


REWARD_FUNCTIONS: dict[str, type[RewardFunction]] = {
    "default": DefaultRewardFunction,
    "simple_distance": SimpleDistanceRewardFunction,
    "binary": BinaryRewardFunction,
    "constant_negative": ConstantNegativeRewardFunction,
}


def get_reward_function(reward_config: Any) -> RewardFunction:
    """rewards.py:194-216"""
    return registry_get(REWARD_FUNCTIONS, reward_config.get_reward_function_name(), "reward", reward_config)
