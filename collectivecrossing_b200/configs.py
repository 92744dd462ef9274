"""Environment config (reference ``configs.py``)."""

from ._models import (  # noqa: F401
    CollectiveCrossingConfig,
)

__all__ = [
    "CollectiveCrossingConfig",
]
