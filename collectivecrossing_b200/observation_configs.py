"""Observation configs (reference ``observation_configs.py``)."""

from ._models import (  # noqa: F401
    ObservationConfig,
    DefaultObservationConfig,
    OBSERVATION_CONFIGS,
    get_observation_config,
)

__all__ = [
    "ObservationConfig",
    "DefaultObservationConfig",
    "OBSERVATION_CONFIGS",
    "get_observation_config",
]
