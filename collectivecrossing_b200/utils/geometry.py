"""Tram geometry (reference ``utils/geometry.py:10-59``)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np


@dataclass
class TramBoundaries:
    """Absolute x-coordinates of the door posts and the tram side walls."""

    tram_door_left: int
    tram_door_right: int
    tram_left: int
    tram_right: int


def calculate_tram_boundaries(config: Any) -> TramBoundaries:
    """Centre the tram in the grid and shift the relative door posts (geometry.py:33-40).

    An odd ``tram_length`` spans ``2 * (tram_length // 2)`` cells, e.g. length 9 on width 12
    gives walls at x = 2 and x = 10.
    """
    half = config.tram_length // 2
    left = config.width // 2 - half
    return TramBoundaries(
        tram_door_left=left + config.tram_door_left,
        tram_door_right=left + config.tram_door_right,
        tram_left=left,
        tram_right=config.width // 2 + half,
    )


def calculate_distance(pos1: tuple, pos2: tuple) -> float:
    """Distance between two positions whose coordinates may be ``None`` (geometry.py:50-59)."""
    if pos1[0] is None or pos2[0] is None:
        return abs(pos1[1] - pos2[1])
    if pos1[1] is None or pos2[1] is None:
        return abs(pos1[0] - pos2[0])
    return float(np.linalg.norm(np.asarray(pos1) - np.asarray(pos2)))
