"""The host-thread crew of the host-buffer path (csrc/cc_workers.h) under stress, on the CPU: thousands of rounds of varying size,
every index exactly once per round, no early return — and the same under ThreadSanitizer where g++ has it."""

import subprocess
from pathlib import Path

import pytest

SRC = Path(__file__).resolve().parent / "native" / "workers_stress.cpp"


def _build(tmp_path, flags, name):
    exe = tmp_path / name
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", *flags, str(SRC), "-o", str(exe)], capture_output=True, text=True)
    return exe if r.returncode == 0 else None


def test_worker_pool_runs_every_index_exactly_once(tmp_path):
    exe = _build(tmp_path, [], "workers_stress")
    assert exe is not None
    out = subprocess.run([str(exe), "4000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_worker_pool_is_clean_under_thread_sanitizer(tmp_path):
    exe = _build(tmp_path, ["-fsanitize=thread"], "workers_stress_tsan")
    if exe is None:
        pytest.skip("g++ has no ThreadSanitizer runtime here")
    out = subprocess.run([str(exe), "600"], capture_output=True, text=True, timeout=600)
    if "FATAL: ThreadSanitizer" in out.stderr and "unexpected memory mapping" in out.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this container")
    assert out.returncode == 0 and "WARNING: ThreadSanitizer" not in out.stderr, out.stdout + out.stderr[-3000:]


def test_row_expansion_is_clean_under_address_and_ub_sanitizers(tmp_path):
    """cc_expand.cpp (staging writer, byte-shuffle path, ragged heads and tails) against a naive restatement, with guard bytes around
    the destination, built with -fsanitize=address,undefined."""
    src = Path(__file__).resolve().parent / "native" / "expand_sanitized.cpp"
    exe = tmp_path / "expand_sanitized"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-g", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    if r.returncode != 0:
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-g", "-pthread", str(src), "-o", str(exe)], capture_output=True, text=True)   # no sanitizer runtime: plain build
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok"), out.stdout[-2000:] + out.stderr[-3000:]
