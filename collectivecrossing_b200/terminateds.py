"""Termination functions — the reference's ``terminateds`` module surface (terminateds.py:12-114);
values come from the step kernel, see ``_strategies``."""

from __future__ import annotations

from typing import Any

from ._strategies import evaluate, registry_get


class TerminatedFunction:
    """Base class (terminateds.py:12-34)."""

    def __init__(self, terminated_config: Any):
        self.terminated_config = terminated_config

    def calculate_terminated(self, agent_id: str, env: Any) -> bool | None:
        return evaluate(env, terminated_config=self.terminated_config)["terminateds"][agent_id]


class AllAtDestinationTerminatedFunction(TerminatedFunction):
    """terminateds.py:37-60"""


class IndividualAtDestinationTerminatedFunction(TerminatedFunction):
    """terminateds.py:63-82"""


TERMINATED_FUNCTIONS: dict[str, type[TerminatedFunction]] = {
    "all_at_destination": AllAtDestinationTerminatedFunction,
    "individual_at_destination": IndividualAtDestinationTerminatedFunction,
}


def get_terminated_function(terminated_config: Any) -> TerminatedFunction:
    """terminateds.py:92-114"""
    return registry_get(TERMINATED_FUNCTIONS, terminated_config.get_terminated_function_name(), "terminated", terminated_config)
