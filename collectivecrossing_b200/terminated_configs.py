"""Termination configs (reference ``terminated_configs.py``)."""

from ._models import (  # noqa: F401
    TerminatedConfig,
    AllAtDestinationTerminatedConfig,
    IndividualAtDestinationTerminatedConfig,
    CustomTerminatedConfig,
    TERMINATED_CONFIGS,
    get_terminated_config,
)

__all__ = [
    "TerminatedConfig",
    "AllAtDestinationTerminatedConfig",
    "IndividualAtDestinationTerminatedConfig",
    "CustomTerminatedConfig",
    "TERMINATED_CONFIGS",
    "get_terminated_config",
]
