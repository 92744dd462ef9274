"""Test infrastructure: stand-ins for the parts of ray / RLlib the reference's training script touches
(examples/training_script.py:26-86), and env-runner-shaped sampling loops.

ray is not installed in this image (and cannot be), so what is pinned here is the PROTOCOL an RLlib
``MultiAgentEnvRunner`` speaks with a ``MultiAgentEnv`` (docs/rllib_multiagent_compatibility.md:13-32):

* ``register_env(name, creator)`` / ``creator(env_config)`` builds the env from a config DICT;
* ``reset()`` returns ``(obs, infos)`` keyed by the live agents;
* every step: the runner maps each agent id in the observation dict through ``policy_mapping_fn``, asks that policy's
  module for an action, calls ``step(action_dict)`` and files observations / rewards / flags per agent; agents come
  and go between steps; ``terminateds["__all__"]`` or ``truncateds["__all__"]`` ends the episode and the env is reset.

``install()`` puts the stub modules into ``sys.modules``; ``run_training_script_head(source)`` executes the
register_env + policy_mapping_fn + env_config statements of the reference's script UNCHANGED (AST-extracted).
"""

from __future__ import annotations

import ast
import sys
import types

import numpy as np

ENV_REGISTRY: dict = {}


def install() -> None:
    """ray.tune.registry.register_env + ray.rllib.env.multi_agent_env.MultiAgentEnv (adds to the oracle's gymnasium stubs)."""
    from oracle import refload

    refload._install_stubs()
    ray = sys.modules["ray"]
    tune = types.ModuleType("ray.tune")
    registry = types.ModuleType("ray.tune.registry")

    def register_env(name, creator):
        ENV_REGISTRY[name] = creator

    registry.register_env = register_env
    tune.registry = registry
    ray.tune = tune
    sys.modules.setdefault("ray.tune", tune)
    sys.modules.setdefault("ray.tune.registry", registry)


def run_training_script_head(source: str, extra_globals: dict | None = None) -> dict:
    """Execute, unchanged, the statements of examples/training_script.py that do not need a ray cluster: the
    ``collectivecrossing`` imports, ``register_env(...)`` (:26-29), ``policy_mapping_fn`` (:32-47) and ``env_config`` (:50-64)."""
    tree = ast.parse(source)
    keep = []
    for node in tree.body:
        if isinstance(node, ast.ImportFrom) and node.module and (node.module.startswith("collectivecrossing") or node.module == "ray.tune.registry"):
            keep.append(node)
        elif isinstance(node, ast.Expr) and isinstance(node.value, ast.Call) and getattr(node.value.func, "id", "") == "register_env":
            keep.append(node)
        elif isinstance(node, ast.FunctionDef) and node.name == "policy_mapping_fn":
            keep.append(node)
        elif isinstance(node, ast.Assign) and any(getattr(t, "id", "") == "env_config" for t in node.targets):
            keep.append(node)
    ns: dict = dict(extra_globals or {})
    exec(compile(ast.Module(body=keep, type_ignores=[]), "examples/training_script.py[head]", "exec"), ns)
    return ns


class SeededPolicy:
    """A policy 'module': a deterministic function of (observation, step) -> action, so that two runners fed the same
    observations pick the same actions whatever env implementation produced them."""

    def __init__(self, salt: int):
        self.salt = salt

    def compute_action(self, obs: np.ndarray, t: int) -> int:
        h = int(np.asarray(obs, np.int64).sum()) * 2654435761 + t * 40503 + self.salt * 97
        return (h >> 7) % 5


def sample_episodes(env, policy_mapping_fn, policies: dict, n_steps: int, seed: int):
    """The loop of RLlib's MultiAgentEnvRunner.sample() at protocol level, on ONE dict-API env.  Returns the episodes:
    per agent the lists of observations, actions, rewards, and the per-step flag dicts."""
    episodes, ep = [], None
    obs, infos = env.reset(seed=seed)
    t_ep = 0

    def new_episode(obs, infos):
        return {"agents": {}, "steps": [], "reset_obs": {a: o.copy() for a, o in obs.items()}, "reset_infos": infos}

    ep = new_episode(obs, infos)
    for t in range(n_steps):
        actions = {}
        # only agents with an observation act.  The action dict's order is the move order (collectivecrossing.py:197) and the
        # reference's observation dict is ordered by a SET of ids (hash-seed dependent, SURVEY.md §8 quirk list), so the
        # runner fixes the order itself: possible_agents order.
        for agent_id in env.possible_agents:
            if agent_id not in obs or agent_id not in env.agents:   # done this step: observation is final, no action
                continue
            o = obs[agent_id]
            module = policies[policy_mapping_fn(agent_id)]
            actions[agent_id] = module.compute_action(o, t_ep)
        obs, rewards, terminateds, truncateds, infos = env.step(actions)
        t_ep += 1
        for a in actions:
            ep["agents"].setdefault(a, {"actions": [], "rewards": []})["actions"].append(actions[a])
        for a, r in rewards.items():
            ep["agents"].setdefault(a, {"actions": [], "rewards": []})["rewards"].append(float(r))
        ep["steps"].append({"obs": {a: o.copy() for a, o in obs.items()}, "rewards": dict(rewards), "terminateds": dict(terminateds),
                            "truncateds": dict(truncateds), "infos": infos})
        if terminateds["__all__"] or truncateds["__all__"]:
            episodes.append(ep)
            obs, infos = env.reset()
            t_ep = 0
            ep = new_episode(obs, infos)
    episodes.append(ep)
    return episodes


def assert_same_episodes(got, want, what=""):
    assert len(got) == len(want), f"{what}: {len(got)} episodes vs {len(want)}"
    for e, (g, w) in enumerate(zip(got, want)):
        assert g["reset_obs"].keys() == w["reset_obs"].keys() and all(np.array_equal(g["reset_obs"][a], w["reset_obs"][a]) for a in g["reset_obs"]), f"{what}: reset obs of episode {e}"
        assert len(g["steps"]) == len(w["steps"]), f"{what}: length of episode {e}"
        for t, (gs, ws) in enumerate(zip(g["steps"], w["steps"])):
            for key in ("rewards", "terminateds", "truncateds", "infos"):
                assert gs[key] == ws[key], f"{what}: episode {e} step {t}: {key}: {gs[key]} vs {ws[key]}"
            assert gs["obs"].keys() == ws["obs"].keys(), f"{what}: episode {e} step {t}: observation keys"
            for a in gs["obs"]:
                assert gs["obs"][a].dtype == np.float32 and np.array_equal(gs["obs"][a], ws["obs"][a]), f"{what}: episode {e} step {t}: obs[{a}]"
        assert g["agents"] == w["agents"], f"{what}: per-agent actions / rewards of episode {e}"
