"""collectivecrossing_b200 — B200-native batched implementation of the CollectiveCrossing
step/reset hot path (reference: nima-siboni/collectivecrossing, ``CollectiveCrossingEnv``).

Public names follow the reference package so user code can switch by changing the import:

    from collectivecrossing_b200 import CollectiveCrossingEnv            # single-env dict API
    from collectivecrossing_b200.configs import CollectiveCrossingConfig
    from collectivecrossing_b200 import BatchedCollectiveCrossing         # N envs, torch tensors

Importing the package does not touch CUDA; constructing an env loads ``csrc/libccb200.so`` and
raises if it has not been built (there is no fallback path).
"""

from .configs import CollectiveCrossingConfig  # noqa: F401

__version__ = "0.1.0"

_LAZY = {
    "BatchedCollectiveCrossing": ("batched", "BatchedCollectiveCrossing"),
    "StepOutput": ("batched", "StepOutput"),
    "CollectiveCrossingEnv": ("env", "CollectiveCrossingEnv"),
    "ShardedCollectiveCrossing": ("distributed", "ShardedCollectiveCrossing"),
    "VectorCollectiveCrossing": ("vector_env", "VectorCollectiveCrossing"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f".{mod}", __name__), attr)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


try:  # the reference registers its env with gymnasium at import (collectivecrossing/__init__.py:3-10)
    from gymnasium.envs.registration import register as _register  # type: ignore

    _register(id="collectivecrossing_b200/CollectiveCrossing-v0", entry_point="collectivecrossing_b200.env:CollectiveCrossingEnv")
except Exception:  # gymnasium is not part of this stack: nothing to register with
    pass

__all__ = ["CollectiveCrossingConfig", *_LAZY]
