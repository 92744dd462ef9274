"""The threading contract of include/ccb200.h: a handle is driven by one thread at a time, DISTINCT handles may be driven from
distinct threads concurrently (the only shared mutable state is the thread-local error string and the mutex-protected table of
kernel attributes).  Two threads step their own envs at the same time — device path on their own streams, host-buffer path with
their own crews of host threads — and must end where a sequential run of the same envs ends."""

import threading

import pytest
from cases import readme_config, readme_crew

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _drive(cfg, n, seed, obs, result, key, errors):
    try:
        from collectivecrossing_b200 import BatchedCollectiveCrossing

        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            env = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=seed, obs_dtype=obs, auto_reset=True)
            env.reset()
            host = env.make_host_buffers()
            for t in range(40):
                if t % 3 == 2:
                    env.step_host(host, policy="waiting")
                elif t % 3 == 1:
                    env.rollout_trajectory(3, policy="greedy")
                else:
                    env.step(policy="greedy")
            stream.synchronize()
            env.check_error()
            result[key] = (env.x.cpu(), env.y.cpu(), env.flags.cpu(), env.step_count.cpu(), env.stats(), host["reward"].clone())
            env.close()
    except Exception as e:  # noqa: BLE001 - reported by the main thread
        errors.append((key, repr(e)))


def test_two_threads_drive_two_handles_concurrently():
    jobs = [(readme_config(max_steps=30), 70_000, 1, "float32"), (readme_crew(3, 2, max_steps=25), 50_000, 2, "int8"),
            (readme_config(max_steps=20), 33_000, 3, "table")]
    sequential, concurrent, errors = {}, {}, []
    for k, job in enumerate(jobs):
        _drive(*job, sequential, k, errors)
    assert not errors, errors
    threads = [threading.Thread(target=_drive, args=(*job, concurrent, k, errors)) for k, job in enumerate(jobs)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for k in range(len(jobs)):
        a, b = sequential[k], concurrent[k]
        assert all(torch.equal(u, v) for u, v in zip(a[:4], b[:4])), f"job {k}: state"
        assert a[4] == b[4], f"job {k}: statistics"
        assert torch.equal(a[5], b[5]), f"job {k}: last rewards in host memory"
