"""world_size-2 gloo tests of the sharding layer (CPU; the rank-local env is the C oracle, which
tests may use as a stand-in for the CUDA env)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from cases import readme_config

from collectivecrossing_b200.distributed import ShardedCollectiveCrossing, derived_stats, reduce_stats, shard_range
from collectivecrossing_b200.lowering import lower_config


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 1000, 16 * 2**20):
        for world_size in (1, 2, 3, 4, 8):
            blocks = [shard_range(total, r, world_size) for r in range(world_size)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (o0, c0), (o1, _) in zip(blocks, blocks[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


class _OracleEnv:
    """Adapter giving OracleEnvs the surface ShardedCollectiveCrossing forwards to."""

    def __init__(self, cfg, n, global_env_offset=0, seed=0):
        import oracle

        self.o = oracle.OracleEnvs(lower_config(cfg), n, seed=seed, global_env_offset=global_env_offset)

    def reset(self):
        return self.o.reset()

    def step(self, policy):
        return self.o.step(policy=policy, auto_reset=True)

    def stats(self):
        return self.o.stats.as_dict()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, total, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        sh = ShardedCollectiveCrossing(readme_config(max_steps=20), total, seed=5, env_factory=_OracleEnv)
        sh.reset()
        for _ in range(steps):
            sh.step("waiting")
        g = sh.global_stats()
        out[rank] = dict(offset=sh.offset, count=sh.count, x=sh.env.o.x.copy(), stats=g, local=sh.local_stats())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_run_equals_single_rank_run():
    total, steps = 301, 45
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), total, steps, out), nprocs=2, join=True)
    assert out[0]["offset"] == 0 and out[0]["count"] == 151 and out[1]["offset"] == 151 and out[1]["count"] == 150
    single = _OracleEnv(readme_config(max_steps=20), total, seed=5)
    single.reset()
    for _ in range(steps):
        single.step("waiting")
    assert np.array_equal(np.concatenate([out[0]["x"], out[1]["x"]]), single.o.x)  # env-by-env identical
    want = single.stats()
    for r in (0, 1):
        g = out[r]["stats"]
        for k in ("env_steps", "episodes", "terminated_all", "truncated_all", "arrivals", "episode_length_sum"):
            assert g[k] == want[k] == out[0]["local"][k] + out[1]["local"][k], k
        assert g["reward_sum"] == pytest.approx(want["reward_sum"], rel=1e-12)
        assert g["episode_len_mean"] == pytest.approx(want["episode_length_sum"] / want["episodes"])
    assert want["episodes"] > 0


def test_reduce_stats_without_process_group_is_identity():
    st = dict(env_steps=10, episodes=2, terminated_all=1, truncated_all=1, arrivals=5, episode_length_sum=40,
              episode_return_sum=-3.5, reward_sum=-4.0)
    assert reduce_stats(st) == st
    d = derived_stats(st)
    assert d["episode_len_mean"] == 20 and d["episode_return_mean"] == -1.75 and d["terminated_fraction"] == 0.5


def test_host_threads_are_shared_between_the_ranks_of_a_node(monkeypatch):
    """Ranks of one node split the node's host threads for the host-buffer path (cc_set_host_expand); a lone process keeps the
    library's automatic choice."""
    import os

    from collectivecrossing_b200.distributed import host_threads_per_rank

    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    assert host_threads_per_rank() == 0
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert host_threads_per_rank() == 0
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert host_threads_per_rank() == max(2, (os.cpu_count() or 1) // 8)
