"""CPU oracle for the CollectiveCrossing step/reset hot path — TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``collectivecrossing_b200``) never does.

* ``cc_oracle.c``  — plain-C restatement of the reference algorithm (each function cites the
  reference file:line), bound here through ctypes on numpy arrays;
* ``pyport.py``    — single-env pure-Python port with the reference's dict API and float64
  rewards (used for the facade parity tests and as the Python-speed CPU baseline);
* ``refload.py``   — imports the UNMODIFIED reference from ``/root/reference`` behind stub
  ``gymnasium`` / ``ray`` / ``matplotlib`` modules (build container only) to pin both.

Parity status: pinned (see ``cc_oracle.c`` header).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from collectivecrossing_b200 import _abi

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libcc_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _HERE / "cc_oracle.c"
    hdr = _HERE.parent / "include" / "ccb200.h"
    stale = (not _LIB_PATH.exists()) or _LIB_PATH.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime)
    if force or stale:
        subprocess.run(["make", "-s", "-B", "-C", str(_HERE)], check=True)
    return _LIB_PATH


_P, _I32, _I64, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
_EXPORTS = {
    "cc_oracle_step": (C.c_int, [C.POINTER(_abi.CCConfig), _I64, _I64, _U64, _U64, _P, _P, _P, _P, _P,
                                 C.POINTER(_abi.CCStepIO), C.POINTER(_abi.CCStats)]),
    "cc_oracle_reset": (C.c_int, [C.POINTER(_abi.CCConfig), _I64, _I64, _U64, _U64, _P, _P, _P, _P, _P, _P, _P, _I32]),
    "cc_oracle_policy_actions": (C.c_int, [C.POINTER(_abi.CCConfig), _I64, _I64, _U64, _U64, _I32, _P, _P, _P, _P, _P]),
    "cc_oracle_observe": (C.c_int, [C.POINTER(_abi.CCConfig), _I64, _P, _P, _P, _P, _P, _I32]),
    "cc_oracle_reset_seeded": (C.c_int, [C.POINTER(_abi.CCConfig), _I64, _P, _P, _P, _P, _P, _P, _P, _P, _I32]),
    "cc_oracle_pcg64_integers": (None, [_U64, _I64, _I64, _I64, _P]),
    "cc_oracle_philox4x32_10": (None, [_P, _P, _P]),
    "cc_oracle_abi_version": (C.c_int, []),
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = _abi.bind(C.CDLL(str(_LIB_PATH)), _EXPORTS)
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_OBS_NP = {_abi.OBS_INT8: np.int8, _abi.OBS_FP32: np.float32}
_REW_NP = {_abi.REWARD_F32: np.float32, _abi.REWARD_F64: np.float64}


class OracleEnvs:
    """N independent envs stepped by the C oracle; same state layout and outputs as the device
    class ``BatchedCollectiveCrossing`` so tests can diff them field by field."""

    def __init__(self, cfg: _abi.CCConfig, num_envs: int, seed: int = 0, global_env_offset: int = 0,
                 threads: int | None = None):
        if threads is not None:
            os.environ["OMP_NUM_THREADS"] = str(threads)
        self.cfg = cfg
        self.n = int(num_envs)
        self.a = cfg.num_agents
        self.obs_len = cfg.obs_len
        self.seed = int(seed)
        self.offset = int(global_env_offset)
        self.t = 0
        self.x = np.zeros((self.n, self.a), np.int8)
        self.y = np.zeros((self.n, self.a), np.int8)
        self.flags = np.zeros((self.n, self.a), np.uint8)
        self.step_count = np.zeros(self.n, np.int32)
        self.episode_return = np.zeros(self.n, np.float32)
        self.stats = _abi.CCStats()
        self._out = {}

    # ---- state -----------------------------------------------------------------------------
    def set_state(self, x, y, flags, step):
        self.x[...] = np.asarray(x, np.int8).reshape(self.n, self.a)
        self.y[...] = np.asarray(y, np.int8).reshape(self.n, self.a)
        self.flags[...] = np.asarray(flags, np.uint8).reshape(self.n, self.a)
        self.step_count[...] = np.asarray(step, np.int32).reshape(self.n)
        self.episode_return[...] = 0

    def get_state(self):
        return self.x.copy(), self.y.copy(), self.flags.copy(), self.step_count.copy()

    def _bufs(self, obs_dtype, reward_dtype):
        key = (obs_dtype, reward_dtype)
        if key not in self._out:
            n, a = self.n, self.a
            self._out[key] = dict(
                obs=None if obs_dtype == _abi.OBS_NONE else (np.zeros((n, a, 4), np.int8) if obs_dtype == _abi.OBS_TABLE
                                                             else np.zeros((n, a, self.obs_len), _OBS_NP[obs_dtype])),
                reward=np.zeros((n, a), _REW_NP[reward_dtype]),
                agent_flags=np.zeros((n, a), np.uint8),
                agent_info=np.zeros((n, a), np.uint8),
                env_flags=np.zeros(n, np.uint8),
                actions_out=np.zeros((n, a), np.int8),
            )
        return self._out[key]

    # ---- hot path --------------------------------------------------------------------------
    def step(self, actions=None, *, order=None, policy="external", auto_reset=False,
             obs_dtype=_abi.OBS_INT8, reward_dtype=_abi.REWARD_F32, check=True):
        b = self._bufs(obs_dtype, reward_dtype)
        io = _abi.CCStepIO()
        if actions is not None:
            actions = np.ascontiguousarray(actions, np.int8).reshape(self.n, self.a)
        if order is not None:
            order = np.ascontiguousarray(order, np.int8).reshape(self.n, self.a)
        io.actions, io.order = _ptr(actions), _ptr(order)
        io.actions_out = _ptr(b["actions_out"])
        io.obs, io.reward = _ptr(b["obs"]), _ptr(b["reward"])
        io.agent_flags, io.agent_info, io.env_flags = _ptr(b["agent_flags"]), _ptr(b["agent_info"]), _ptr(b["env_flags"])
        io.obs_dtype, io.reward_dtype = obs_dtype, reward_dtype
        io.policy = _abi.POLICIES[policy] if isinstance(policy, str) else int(policy)
        io.auto_reset = int(bool(auto_reset))
        rc = lib().cc_oracle_step(C.byref(self.cfg), self.n, self.offset, self.seed, self.t, _ptr(self.x), _ptr(self.y),
                                  _ptr(self.flags), _ptr(self.step_count), _ptr(self.episode_return), C.byref(io),
                                  C.byref(self.stats))
        self.t += 1
        if check and rc != _abi.OK:
            raise ValueError(f"oracle step failed with status {rc}")
        out = dict(b)
        out["status"] = rc
        return out

    def reset(self, mask=None, obs_dtype=_abi.OBS_INT8):
        b = self._bufs(obs_dtype, _abi.REWARD_F32)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        rc = lib().cc_oracle_reset(C.byref(self.cfg), self.n, self.offset, self.seed, self.t, _ptr(mask), _ptr(self.x),
                                   _ptr(self.y), _ptr(self.flags), _ptr(self.step_count), _ptr(self.episode_return),
                                   _ptr(b["obs"]), obs_dtype)
        self.t += 1
        if rc != _abi.OK:
            raise RuntimeError(f"oracle reset failed with status {rc}")
        return b["obs"]

    def reset_seeded(self, seeds, obs_dtype=_abi.OBS_INT8):
        """seeds: int64 per env (reset(seed=s)); None continues the stored generators (reset())."""
        b = self._bufs(obs_dtype, _abi.REWARD_F32)
        if not hasattr(self, "gen"):
            self.gen = np.zeros((self.n, 6), np.uint64)
        if seeds is not None:
            seeds = np.ascontiguousarray(seeds, np.int64).reshape(self.n)
        rc = lib().cc_oracle_reset_seeded(C.byref(self.cfg), self.n, _ptr(seeds), _ptr(self.gen), _ptr(self.x), _ptr(self.y),
                                          _ptr(self.flags), _ptr(self.step_count), _ptr(self.episode_return),
                                          _ptr(b["obs"]), obs_dtype)
        if rc != _abi.OK:
            raise RuntimeError(f"oracle reset_seeded failed with status {rc}")
        return b["obs"]

    def policy_actions(self, policy):
        out = np.zeros((self.n, self.a), np.int8)
        pol = _abi.POLICIES[policy] if isinstance(policy, str) else int(policy)
        rc = lib().cc_oracle_policy_actions(C.byref(self.cfg), self.n, self.offset, self.seed, self.t, pol, _ptr(self.x),
                                            _ptr(self.y), _ptr(self.flags), _ptr(self.step_count), _ptr(out))
        if rc != _abi.OK:
            raise RuntimeError(f"oracle policy failed with status {rc}")
        return out

    def observe(self, obs_dtype=_abi.OBS_INT8):
        obs = np.zeros((self.n, self.a, 4), np.int8) if obs_dtype == _abi.OBS_TABLE else np.zeros((self.n, self.a, self.obs_len), _OBS_NP[obs_dtype])
        lib().cc_oracle_observe(C.byref(self.cfg), self.n, _ptr(self.x), _ptr(self.y), _ptr(self.flags),
                                _ptr(self.step_count), _ptr(obs), obs_dtype)
        return obs


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, np.uint32)
    k = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().cc_oracle_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def pcg64_integers(seed: int, low: int, high: int, count: int):
    out = np.zeros(count, np.int64)
    lib().cc_oracle_pcg64_integers(seed, low, high, count, _ptr(out))
    return out
