"""Truncation functions — the reference's ``truncateds`` module surface (truncateds.py:12-128);
values come from the step kernel, see ``_strategies``.  ``CustomTruncatedFunction`` is a
placeholder in the reference too; it is not in the registry here (the lowering rejects it)."""

from __future__ import annotations

from typing import Any

from ._strategies import evaluate, is_done, registry_get


class TruncatedFunction:
    """Base class (truncateds.py:12-34)."""

    def __init__(self, truncated_config: Any):
        self.truncated_config = truncated_config

    def calculate_truncated(self, agent_id: str, env: Any) -> bool | None:
        if is_done(agent_id, env):   # truncateds.py:57-58
            return None
        return evaluate(env, truncated_config=self.truncated_config)["truncateds"].get(agent_id)


class MaxStepsTruncatedFunction(TruncatedFunction):
    """truncateds.py:37-61: None for a done agent, else ``step_count >= max_steps``"""


TRUNCATED_FUNCTIONS: dict[str, type[TruncatedFunction]] = {"max_steps": MaxStepsTruncatedFunction}


def get_truncated_function(truncated_config: Any) -> TruncatedFunction:
    """truncateds.py:105-128"""
    return registry_get(TRUNCATED_FUNCTIONS, truncated_config.get_truncated_function_name(), "truncated", truncated_config)
