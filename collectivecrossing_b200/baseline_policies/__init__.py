"""Baseline policies with the reference's ``get_action(agent_id, observation, env)`` interface
(reference ``src/baseline_policies``); the decision itself is taken by the CUDA policy kernel."""

from .greedy_policy import GreedyPolicy, create_greedy_policy  # noqa: F401
from .waiting_policy import WaitingPolicy, create_waiting_policy  # noqa: F401

__all__ = ["GreedyPolicy", "WaitingPolicy", "create_greedy_policy", "create_waiting_policy"]
