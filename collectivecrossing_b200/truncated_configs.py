"""Truncation configs (reference ``truncated_configs.py``)."""

from ._models import (  # noqa: F401
    TruncatedConfig,
    MaxStepsTruncatedConfig,
    CustomTruncatedConfig,
    TRUNCATED_CONFIGS,
    get_truncated_config,
)

__all__ = [
    "TruncatedConfig",
    "MaxStepsTruncatedConfig",
    "CustomTruncatedConfig",
    "TRUNCATED_CONFIGS",
    "get_truncated_config",
]
