"""Shared machinery of the strategy modules (``rewards``, ``terminateds``, ``truncateds``,
``observations``): the reference's plugin API is "registry name -> class with
``calculate_*(agent_id, env)``" (rewards.py:186-216, terminateds.py:86-114, truncateds.py:99-128,
observations.py:122-149).  Here the classes keep that surface, but the values come from the device:
``env._evaluate(...)`` runs the fused step kernel with WAIT actions on a scratch one-env copy of the
env's current host view and returns what the strategy functions return for every agent.
"""

from __future__ import annotations

from typing import Any


def registry_get(registry: dict, name: str, what: str, config: Any):
    if name not in registry:
        raise ValueError(f"Unknown {what} function '{name}'. Available: {', '.join(registry.keys())}")
    return registry[name](config)


def evaluate(env: Any, **overrides: Any) -> dict:
    if not hasattr(env, "_evaluate"):
        raise TypeError("strategy functions of collectivecrossing_b200 evaluate on the device and need a "
                        "collectivecrossing_b200.CollectiveCrossingEnv (mock envs are not supported)")
    return env._evaluate(**overrides)
