"""Greedy door-seeking policy (reference ``baseline_policies/greedy_policy.py``)."""

from ._base import _DevicePolicy


class GreedyPolicy(_DevicePolicy):
    kind = "greedy"


def create_greedy_policy(epsilon: float = 0.1) -> GreedyPolicy:
    return GreedyPolicy(randomness_factor=epsilon, seed=42)
