"""Boarders wait until every exiting agent is done (reference ``baseline_policies/waiting_policy.py``)."""

from ._base import _DevicePolicy


class WaitingPolicy(_DevicePolicy):
    kind = "waiting"


def create_waiting_policy(epsilon: float = 0.1) -> WaitingPolicy:
    return WaitingPolicy(randomness_factor=epsilon, seed=42)
