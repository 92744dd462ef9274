"""Shared machinery of the strategy modules (``rewards``, ``terminateds``, ``truncateds``,
``observations``): the reference's plugin API is "registry name -> class with
``calculate_*(agent_id, env)``" on ANY object that offers a few accessors (rewards.py:186-216,
terminateds.py:86-114, truncateds.py:99-128, observations.py:122-149; its own tests call them on mock
envs, tests/collectivecrossing/envs/test_rewards.py:476-527).  Here the classes keep that surface, but
the values come from the device: the fused step kernel runs with WAIT actions on a scratch one-env
copy of the env's current state and returns what the strategy functions return for every agent.

* a ``collectivecrossing_b200.CollectiveCrossingEnv`` evaluates through its own ``_evaluate``;
* any other env object is read through the accessors the reference's strategy code itself uses:
  ``env._agents[id].terminated / .truncated`` (the ``None`` guards, rewards.py:65-66,
  truncateds.py:57-58 — answered on the host, they involve no arithmetic) and, for a value,
  ``env.config``, the ``_agents`` records (``position``, ``active``, flags) and ``env._step_count``.
"""

from __future__ import annotations

from typing import Any

import numpy as np

_FOREIGN_EVALUATORS: dict = {}


def registry_get(registry: dict, name: str, what: str, config: Any):
    if name not in registry:
        raise ValueError(f"Unknown {what} function '{name}'. Available: {', '.join(registry.keys())}")
    return registry[name](config)


def is_done(agent_id: str, env: Any) -> bool:
    """The guard every reward / truncation function starts with (rewards.py:65-66, truncateds.py:57-58)."""
    agent = env._agents[agent_id]
    return bool(agent.terminated or agent.truncated)


def evaluate(env: Any, **overrides: Any) -> dict:
    if hasattr(env, "_evaluate"):
        return env._evaluate(**overrides)
    return _evaluate_foreign(env, overrides)


def _evaluate_foreign(env: Any, overrides: dict) -> dict:
    """Strategy values for an env object that is not ours: its config and agent records are copied into a
    scratch one-env device batch (cached per config + overrides) and evaluated like ``CollectiveCrossingEnv._evaluate``."""
    import torch

    from . import _abi
    from .batched import BatchedCollectiveCrossing

    config = getattr(env, "config", None) or getattr(env, "_config", None)
    agents = getattr(env, "_agents", None)
    if config is None or agents is None or not hasattr(config, "num_boarding_agents"):
        raise TypeError("strategy functions of collectivecrossing_b200 evaluate on the device: the env object must offer `config` "
                        "(a CollectiveCrossingConfig) and `_agents` records with position / active / terminated / truncated")
    upd = {k: v for k, v in overrides.items() if v is not None}
    cfg = config.model_copy(update=upd) if upd else config
    ids = [f"boarding_{i}" for i in range(cfg.num_boarding_agents)] + [f"exiting_{i}" for i in range(cfg.num_exiting_agents)]
    missing = [a for a in ids if a not in agents]
    if missing:
        raise TypeError(f"env._agents lacks the records of {missing} that env.config declares")
    key = repr(cfg)
    ev = _FOREIGN_EVALUATORS.get(key)
    if ev is None:
        if len(_FOREIGN_EVALUATORS) > 16:
            for old in _FOREIGN_EVALUATORS.values():
                old.close()
            _FOREIGN_EVALUATORS.clear()
        ev = _FOREIGN_EVALUATORS[key] = BatchedCollectiveCrossing(cfg, 1, "cuda:0", obs_dtype="none", reward_dtype="float64", auto_reset=False)
    A = len(ids)
    x, y, f = np.zeros((1, A), np.int8), np.zeros((1, A), np.int8), np.zeros((1, A), np.uint8)
    for k, a in enumerate(ids):
        ag = agents[a]
        x[0, k], y[0, k] = int(ag.position[0]), int(ag.position[1])
        f[0, k] = ((_abi.F_ACTIVE if getattr(ag, "active", True) else 0) | (_abi.F_TERMINATED if ag.terminated else 0)
                   | (_abi.F_TRUNCATED if ag.truncated else 0))
    step = int(getattr(env, "_step_count", 0)) - 1   # the WAIT step below counts it up again (collectivecrossing.py:188)
    ev.set_state(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(f), torch.tensor([step], dtype=torch.int32))
    out = ev.step(torch.full((1, A), 4, dtype=torch.int8, device=ev.device))
    ev.check_error()
    rew, af = out.reward.cpu().numpy()[0], out.agent_flags.cpu().numpy()[0]
    res: dict = {"rewards": {}, "terminateds": {}, "truncateds": {}}
    for k, a in enumerate(ids):
        bits = int(af[k])
        res["terminateds"][a] = bool(bits & _abi.O_TERM_VALUE)
        if bits & _abi.O_ALIVE_PREV:
            res["rewards"][a] = float(rew[k])
            res["truncateds"][a] = bool(bits & _abi.O_TRUNC_VALUE)
    return res
