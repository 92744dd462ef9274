"""Import the UNMODIFIED reference: from ``/root/reference/src`` in the build container, else from the archive
``oracle/_ref/reference.zip`` that ``oracle/stage_ref.py`` stages there (git-ignored, travels with gpurun).

gymnasium, ray and matplotlib are not installed and the hot path needs almost nothing from them
(SURVEY.md §8c), so tiny stand-ins are injected into ``sys.modules`` first:

* ``gymnasium.Env.reset(seed)`` builds ``np.random.Generator(np.random.PCG64(SeedSequence(seed)))``
  — exactly gymnasium's ``utils.seeding.np_random``;
* ``gymnasium.spaces.Discrete/Box`` keep their constructor arguments and ``sample()``;
* ``gymnasium.envs.registration.register`` is a no-op;
* ``ray.rllib.env.multi_agent_env.MultiAgentEnv`` is an ``Env`` with an empty ``__init__``;
* ``matplotlib.pyplot`` only needs an ``Axes`` attribute (evaluated in a signature).

Where neither the mount nor the archive exists ``available()`` is False and every caller skips.
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

ARCHIVE = Path(__file__).resolve().parent / "_ref" / "reference.zip"
_MOUNT = Path("/root/reference")
if (_MOUNT / "src" / "collectivecrossing" / "collectivecrossing.py").exists():
    REFERENCE_SRC: Path | None = _MOUNT / "src"
    REFERENCE_TESTS: Path | None = _MOUNT / "tests"
    SOURCE = "mount"
elif ARCHIVE.exists():
    REFERENCE_SRC = ARCHIVE / "src"   # zipimport: a directory inside an archive is a valid sys.path entry
    REFERENCE_TESTS = None            # extract_tests() unpacks them on demand
    SOURCE = "archive"
else:
    REFERENCE_SRC = REFERENCE_TESTS = None
    SOURCE = None


def available() -> bool:
    return REFERENCE_SRC is not None


def extract(dest: Path, prefixes=("tests/", "src/", "scripts/", "examples/")) -> Path:
    """Unpack the reference (mount or archive) under ``dest``; returns ``dest``."""
    import shutil
    import zipfile

    dest = Path(dest)
    if SOURCE == "mount":
        for pre in prefixes:
            src = _MOUNT / pre.rstrip("/")
            if src.exists():
                shutil.copytree(src, dest / pre.rstrip("/"), ignore=shutil.ignore_patterns("__pycache__"), dirs_exist_ok=True)
    elif SOURCE == "archive":
        with zipfile.ZipFile(ARCHIVE) as z:
            z.extractall(dest, [n for n in z.namelist() if n.startswith(tuple(prefixes))])
    else:
        raise RuntimeError("reference neither mounted nor staged")
    return dest


def _install_stubs() -> None:
    if "gymnasium" in sys.modules and not getattr(sys.modules["gymnasium"], "__ccb200_stub__", False):
        return  # a real gymnasium is installed: use it

    gym = types.ModuleType("gymnasium")
    gym.__ccb200_stub__ = True

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape, self.dtype = shape, dtype

    class Discrete(Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)
            self._rng = np.random.default_rng()

        def sample(self):
            return int(self._rng.integers(0, self.n))

        def contains(self, v):
            return 0 <= int(v) < self.n

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            super().__init__(tuple(shape), dtype)
            self.low, self.high = low, high

    class Env:
        metadata: dict = {}
        _np_random = None

        def reset(self, *, seed=None, options=None):
            if seed is not None or self._np_random is None:
                self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
            return self._np_random

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Space, spaces.Discrete, spaces.Box = Space, Discrete, Box
    gym.Env, gym.Space, gym.spaces = Env, Space, spaces
    envs = types.ModuleType("gymnasium.envs")
    registration = types.ModuleType("gymnasium.envs.registration")
    registration.register = lambda *a, **k: None
    envs.registration = registration
    gym.envs = envs

    ray = types.ModuleType("ray")
    rllib = types.ModuleType("ray.rllib")
    rllib_env = types.ModuleType("ray.rllib.env")
    mae = types.ModuleType("ray.rllib.env.multi_agent_env")

    class MultiAgentEnv(Env):
        def __init__(self):
            pass

    mae.MultiAgentEnv = MultiAgentEnv
    ray.rllib, rllib.env, rllib_env.multi_agent_env = rllib, rllib_env, mae

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.Axes = object
    mpl.pyplot = plt

    for name, mod in {
        "gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.envs": envs,
        "gymnasium.envs.registration": registration,
        "ray": ray, "ray.rllib": rllib, "ray.rllib.env": rllib_env, "ray.rllib.env.multi_agent_env": mae,
        "matplotlib": mpl, "matplotlib.pyplot": plt,
    }.items():
        sys.modules.setdefault(name, mod)


_ref = None


def load():
    """Returns a namespace with the reference's public names (imports it on first use)."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference sources are neither mounted at /root/reference nor staged at oracle/_ref/reference.zip")
    _install_stubs()
    if str(REFERENCE_SRC) not in sys.path:
        sys.path.insert(0, str(REFERENCE_SRC))
    import collectivecrossing as cc  # noqa: F401  (the reference package)
    from baseline_policies.greedy_policy import GreedyPolicy
    from baseline_policies.waiting_policy import WaitingPolicy
    from collectivecrossing import configs, observation_configs, reward_configs, terminated_configs, truncated_configs
    from collectivecrossing.collectivecrossing import CollectiveCrossingEnv

    ns = types.SimpleNamespace(
        CollectiveCrossingEnv=CollectiveCrossingEnv,
        CollectiveCrossingConfig=configs.CollectiveCrossingConfig,
        reward_configs=reward_configs,
        terminated_configs=terminated_configs,
        truncated_configs=truncated_configs,
        observation_configs=observation_configs,
        GreedyPolicy=GreedyPolicy,
        WaitingPolicy=WaitingPolicy,
    )
    _ref = ns
    return ns


def to_reference_config(cfg, validate: bool = True):
    """Rebuild one of OUR config objects as the reference's class (same field names)."""
    ref = load()
    d = cfg.model_dump()
    # dump the strategy configs from the instances: the env config declares them by their base
    # class, so a plain model_dump() of the parent drops the subclass fields (e.g. max_steps)
    for k in ("reward_config", "terminated_config", "truncated_config", "observation_config"):
        d.pop(k)
    rc, tc, uc, oc = (getattr(cfg, k).model_dump() for k in
                      ("reward_config", "terminated_config", "truncated_config", "observation_config"))
    sub = dict(
        reward_config=ref.reward_configs.REWARD_CONFIGS[rc["reward_function"]](**rc),
        terminated_config=ref.terminated_configs.TERMINATED_CONFIGS[tc["terminated_function"]](**tc),
        truncated_config=ref.truncated_configs.TRUNCATED_CONFIGS[uc["truncated_function"]](**uc),
        observation_config=ref.observation_configs.OBSERVATION_CONFIGS[oc["observation_function"]](**oc),
    )
    if validate:
        return ref.CollectiveCrossingConfig(**d, **sub)
    return ref.CollectiveCrossingConfig.model_construct(**d, **sub)
