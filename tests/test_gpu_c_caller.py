"""The C ABI driven by a plain C program (examples/host_caller.c: no CUDA headers, no Python, no torch in the process) against the
oracle: cc_create -> cc_reset -> K x cc_step_host (greedy policy in the kernel, auto-reset, float32 rows into malloc'ed memory)
-> cc_get_state_host / cc_stats_read.  Everything the program dumps must equal the oracle's run of the same seed bit for bit."""

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
from cases import readme_config

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "collectivecrossing_b200" / "csrc"


def _compile(tmp_path):
    exe = tmp_path / "host_caller"
    cmd = ["gcc", "-std=c99", "-D_POSIX_C_SOURCE=199309L", "-O2", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}",
           str(ROOT / "examples" / "host_caller.c"), f"-L{CSRC}", "-lccb200", f"-Wl,-rpath,{CSRC}", "-o", str(exe)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_c_example_compiles_against_the_header_as_c99(tmp_path):
    """(no GPU) include/ccb200.h is a C header: the example builds with -std=c99 -pedantic -Werror and links every symbol it uses."""
    from collectivecrossing_b200 import _native

    _native.library()   # builds the library if it is not there yet
    assert _compile(tmp_path).exists()


@pytest.mark.gpu
@pytest.mark.parametrize("n,steps,seed", [(4099, 130, 11), (70000, 40, 5)])
def test_c_caller_matches_the_oracle(tmp_path, n, steps, seed):
    import oracle
    from collectivecrossing_b200 import _abi
    from collectivecrossing_b200.lowering import lower_config

    exe = _compile(tmp_path)
    dump = tmp_path / "dump.bin"
    out = subprocess.run([str(exe), str(n), str(steps), str(seed), "--dump", str(dump)], check=True, capture_output=True, text=True, timeout=300)
    assert "agent-steps/s through host buffers" in out.stdout and "cc_step_tpe2_kernel" in out.stdout, out.stdout

    low = lower_config(readme_config())
    A, L = 8, 38
    orc = oracle.OracleEnvs(low, n, seed=seed)
    orc.reset()
    for _ in range(steps):
        res = orc.step(policy="greedy", auto_reset=True, obs_dtype=_abi.OBS_FP32)
    raw = dump.read_bytes()
    pos = 0

    def take(dtype, shape):
        nonlocal pos
        cnt = int(np.prod(shape))
        a = np.frombuffer(raw, dtype=dtype, count=cnt, offset=pos).reshape(shape)
        pos += cnt * np.dtype(dtype).itemsize
        return a

    x, y, flags, step = take(np.int8, (n, A)), take(np.int8, (n, A)), take(np.uint8, (n, A)), take(np.int32, (n,))
    obs, reward = take(np.float32, (n, A, L)), take(np.float32, (n, A))
    agent_flags, env_flags, applied = take(np.uint8, (n, A)), take(np.uint8, (n,)), take(np.int8, (n, A))
    st = _abi.CCStats.from_buffer_copy(raw[pos:pos + C.sizeof(_abi.CCStats)])
    assert pos + C.sizeof(_abi.CCStats) == len(raw)
    assert np.array_equal(x, orc.x) and np.array_equal(y, orc.y) and np.array_equal(flags, orc.flags) and np.array_equal(step, orc.step_count)
    assert np.array_equal(obs, res["obs"]) and np.array_equal(reward, res["reward"])
    assert np.array_equal(agent_flags, res["agent_flags"]) and np.array_equal(env_flags, res["env_flags"]) and np.array_equal(applied, res["actions_out"])
    assert (st.env_steps, st.episodes, st.terminated_all, st.truncated_all, st.arrivals, st.episode_length_sum) == \
           (orc.stats.env_steps, orc.stats.episodes, orc.stats.terminated_all, orc.stats.truncated_all, orc.stats.arrivals, orc.stats.episode_length_sum)
    assert st.episodes > 0
    np.testing.assert_allclose(st.reward_sum, orc.stats.reward_sum, rtol=1e-9)
