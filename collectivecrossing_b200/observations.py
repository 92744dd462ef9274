"""Observation functions — the reference's ``observations`` module surface (observations.py:15-149);
the vectors come from the observe kernel (``cc_observe``)."""

from __future__ import annotations

from typing import Any

import numpy as np

from . import _spaces
from ._strategies import registry_get


class ObservationFunction:
    """Base class (observations.py:15-37)."""

    def __init__(self, observation_config: Any):
        self.observation_config = observation_config

    def get_agent_observation(self, agent_id: str, env: Any) -> np.ndarray:
        raise NotImplementedError


class DefaultObservationFunction(ObservationFunction):
    """observations.py:40-118: ``[x, y, door centre, division, door left, door right]`` then
    ``[x_j, y_j, type_j, active_j]`` for every agent with the own block masked by -1."""

    def get_agent_observation(self, agent_id: str, env: Any) -> np.ndarray:
        if not hasattr(env, "_get_agent_observation"):
            raise TypeError("observation functions of collectivecrossing_b200 need a collectivecrossing_b200.CollectiveCrossingEnv")
        return env._get_agent_observation(agent_id)

    def return_agent_observation_space(self, agent_id: str, env: Any):
        # declared bound max(w, h) - 1 although x can equal w: kept as in the reference (observations.py:113-118)
        return _spaces.Box(low=-1, high=max(env.config.width, env.config.height) - 1, shape=(2 + 4 + 4 * len(env._agents),), dtype=np.float32)


OBSERVATION_FUNCTIONS: dict[str, type[ObservationFunction]] = {"default": DefaultObservationFunction}


def get_observation_function(observation_config: Any) -> ObservationFunction:
    """observations.py:127-149"""
    return registry_get(OBSERVATION_FUNCTIONS, observation_config.get_observation_function_name(), "observation", observation_config)
