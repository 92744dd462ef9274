"""Action codes (reference ``actions.py:8-24``)."""

from enum import Enum

import numpy as np


class Actions(Enum):
    """The five moves of an agent; the integer values are part of the public API."""

    right = 0
    up = 1
    left = 2
    down = 3
    wait = 4


_STEPS = ((1, 0), (0, 1), (-1, 0), (0, -1), (0, 0))
ACTION_TO_DIRECTION = {a.value: np.array(_STEPS[a.value]) for a in Actions}
