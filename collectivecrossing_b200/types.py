"""Agent record types (reference ``types.py:9-83``)."""

from __future__ import annotations

from dataclasses import dataclass
from enum import Enum

import numpy as np


class AgentType(Enum):
    BOARDING = "boarding"
    EXITING = "exiting"


@dataclass
class Agent:
    """Host-side view of one agent of a single-env facade.

    On the device an agent is three bytes (x, y, flags); this dataclass exists so tests and
    policies can read / inject state the way they do with the reference (``env._agents[id]``).
    The facade pushes the records to the device before every launch.
    """

    id: str
    agent_type: AgentType
    position: np.ndarray
    active: bool
    terminated: bool
    truncated: bool

    def __post_init__(self) -> None:
        if not isinstance(self.position, np.ndarray):
            self.position = np.array(self.position)

    @property
    def x(self) -> int:
        return int(self.position[0])

    @property
    def y(self) -> int:
        return int(self.position[1])

    def update_position(self, new_position: np.ndarray) -> None:
        self.position = np.array(new_position)

    def _flip(self, field: str, value: bool, complaint: str) -> None:
        if getattr(self, field) == value:
            raise ValueError(complaint)
        setattr(self, field, value)

    def deactivate(self) -> None:
        self._flip("active", False, "Agent is already deactivated.")

    def terminate(self) -> None:
        self._flip("terminated", True, "Agent is already terminated.")

    def truncate(self) -> None:
        self._flip("truncated", True, "Agent is already truncated.")

    @property
    def is_boarding(self) -> bool:
        return self.agent_type == AgentType.BOARDING

    @property
    def is_exiting(self) -> bool:
        return self.agent_type == AgentType.EXITING

    @property
    def is_terminated(self) -> bool:
        return self.terminated

    @property
    def is_truncated(self) -> bool:
        return self.truncated
