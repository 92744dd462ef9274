"""Reward configs (reference ``reward_configs.py``)."""

from ._models import (  # noqa: F401
    RewardConfig,
    DefaultRewardConfig,
    SimpleDistanceRewardConfig,
    BinaryRewardConfig,
    ConstantNegativeRewardConfig,
    CustomRewardConfig,
    REWARD_CONFIGS,
    get_reward_config,
)

__all__ = [
    "RewardConfig",
    "DefaultRewardConfig",
    "SimpleDistanceRewardConfig",
    "BinaryRewardConfig",
    "ConstantNegativeRewardConfig",
    "CustomRewardConfig",
    "REWARD_CONFIGS",
    "get_reward_config",
]
