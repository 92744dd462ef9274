"""GPU: the parallel resolution of the ordered moves (lane-group kernel, one env per warp: crews above 16) on states built to
exercise its rules — queues that move as a chain, several agents asking for one cell, swaps and cycles, targets held by agents that
move later / never / are ghosts, stacked active agents (the sequential fallback) — each against the oracle's sequential loop
(collectivecrossing.py:197-202), for many steps."""

import numpy as np
import pytest
from cases import unchecked

from collectivecrossing_b200 import _abi
from collectivecrossing_b200.lowering import lower_config
from collectivecrossing_b200.reward_configs import SimpleDistanceRewardConfig
from collectivecrossing_b200.terminated_configs import AllAtDestinationTerminatedConfig
from collectivecrossing_b200.truncated_configs import MaxStepsTruncatedConfig

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
R, U, L, D, W = 0, 1, 2, 3, 4


def corridor_config(boarding, exiting, width=30, height=10):
    """A wide waiting area (rows 0-3) under a tram (rows 4-10): crews of 17-40 agents, one env per warp."""
    return unchecked(width=width, height=height, division_y=4, tram_door_left=8, tram_door_right=20, tram_length=26,
                     num_boarding_agents=boarding, num_exiting_agents=exiting, exiting_destination_area_y=0, boarding_destination_area_y=height,
                     reward_config=SimpleDistanceRewardConfig(), terminated_config=AllAtDestinationTerminatedConfig(),
                     truncated_config=MaxStepsTruncatedConfig(max_steps=1000))


def run_case(cfg, x, y, flags, action_seq, tag):
    import oracle
    from collectivecrossing_b200 import BatchedCollectiveCrossing

    low = lower_config(cfg)
    n, A = x.shape
    env = BatchedCollectiveCrossing(cfg, n, "cuda:0", obs_dtype="table", auto_reset=False, with_info=True)
    orc = oracle.OracleEnvs(low, n)
    s = np.zeros(n, np.int32)
    env.set_state(*(torch.from_numpy(v).cuda() for v in (x, y, flags, s)))
    orc.set_state(x, y, flags, s)
    for t, acts in enumerate(action_seq):
        out = env.step(torch.from_numpy(acts).cuda())
        res = orc.step(acts, obs_dtype=_abi.OBS_TABLE)
        assert np.array_equal(env.x.cpu().numpy(), orc.x) and np.array_equal(env.y.cpu().numpy(), orc.y), f"{tag}: positions after step {t}"
        assert np.array_equal(env.flags.cpu().numpy(), orc.flags), f"{tag}: flags after step {t}"
        assert np.array_equal(out.obs.cpu().numpy(), res["obs"]) and np.array_equal(out.reward.cpu().numpy(), res["reward"]), f"{tag}: outputs of step {t}"
    assert env.last_kernel_name.startswith("ccb::cc_kernel<32,")
    env.check_error()
    env.close()


def test_queues_move_as_chains_in_agent_order_only():
    """A row of 20 agents that all step RIGHT: with ascending agent indices from right to left the whole queue advances
    (each agent's target was vacated earlier in the same step); with descending indices only the head moves."""
    cfg = corridor_config(20, 4)
    A, n = 24, 2
    x, y, f = np.zeros((n, A), np.int8), np.zeros((n, A), np.int8), np.ones((n, A), np.uint8)
    x[0, :20] = np.arange(19, -1, -1) + 2      # agent 0 at the head (x = 21): everyone can follow
    x[1, :20] = np.arange(0, 20) + 2           # agent 19 at the head: agent k finds agent k+1 still in place
    y[:, :20] = 1
    x[:, 20:] = np.array([5, 7, 9, 11]); y[:, 20:] = 7
    acts = np.full((n, A), W, np.int8)
    acts[:, :20] = R
    run_case(cfg, x, y, f, [acts.copy() for _ in range(6)], "queues")


def test_contenders_for_one_cell_and_targets_of_every_kind():
    """Four agents around one free cell all step into it (the first in agent order wins); agents stepping onto a ghost
    (free), onto an agent that never moves (blocked), onto one that moves later (blocked) and onto one that moved earlier (free)."""
    cfg = corridor_config(18, 2)
    A, n = 20, 1
    x, y, f = np.zeros((n, A), np.int8), np.zeros((n, A), np.int8), np.ones((n, A), np.uint8)
    pos = {0: (10, 2), 1: (9, 1), 2: (11, 1), 3: (10, 0),        # all want (10, 1)
           4: (15, 1), 5: (16, 1),                               # 4 steps onto ghost 5
           6: (20, 1), 7: (21, 1),                               # 6 steps onto 7, who waits
           8: (24, 1), 9: (25, 1),                               # 8 steps onto 9, who moves away LATER in the order
           11: (3, 2), 10: (4, 2),                               # 11 steps onto 10's cell, vacated EARLIER in the order
           12: (1, 0), 13: (2, 0),                               # 12 and 13 swap: neither can
           14: (27, 0), 15: (28, 0), 16: (28, 1), 17: (27, 1)}   # a 4-cycle: nobody can
    for k, (px, py) in pos.items():
        x[0, k], y[0, k] = px, py
    x[0, 18:], y[0, 18:] = [12, 14], [7, 7]
    f[0, 5] = 0                                                   # ghost: inactive, not done
    acts = np.full((n, A), W, np.int8)
    acts[0, [0, 1, 2, 3]] = [D, R, L, U]
    acts[0, 4], acts[0, 6], acts[0, 8], acts[0, 9] = R, R, R, U
    acts[0, 10], acts[0, 11] = R, R
    acts[0, 12], acts[0, 13] = R, L
    acts[0, [14, 15, 16, 17]] = [R, U, L, D]
    run_case(cfg, x, y, f, [acts, acts, acts], "contenders")


def test_stacked_active_agents_use_the_sequential_turns():
    """Two ACTIVE agents on one cell cannot arise from play but can be injected: the parallel maps cannot represent them, the
    kernel falls back to the reference's sequential turns for that env and still agrees with the oracle."""
    cfg = corridor_config(18, 2)
    A, n = 20, 3
    rng = np.random.default_rng(3)
    x, y, f = np.zeros((n, A), np.int8), np.zeros((n, A), np.int8), np.ones((n, A), np.uint8)
    for e in range(n):
        cells = rng.permutation(30 * 4)[:18]
        x[e, :18], y[e, :18] = cells % 30, cells // 30
        x[e, 18:], y[e, 18:] = [12, 14], [7, 7]
    x[1, 3], y[1, 3] = x[1, 2], y[1, 2]                      # env 1: agents 2 and 3 stacked
    x[2, 7], y[2, 7] = x[2, 0], y[2, 0]                      # env 2: agents 0 and 7 stacked
    seq = [rng.integers(0, 5, size=(n, A)).astype(np.int8) for _ in range(12)]
    run_case(cfg, x, y, f, seq, "stacked")


@pytest.mark.parametrize("crew", [(17, 0), (24, 8), (30, 10)])
def test_dense_random_crowds(crew):
    """Dense crowds (a third of the waiting area occupied) under random actions: conflicts and chains in almost every step."""
    b, e = crew
    cfg = corridor_config(b, e)
    A, n = b + e, 64
    rng = np.random.default_rng(b)
    x, y, f = np.zeros((n, A), np.int8), np.zeros((n, A), np.int8), np.ones((n, A), np.uint8)
    for k in range(n):
        cells = rng.permutation(12 * 4)[:b] if b <= 48 else None
        x[k, :b], y[k, :b] = cells % 12 + 6, cells // 12     # packed into a 12 x 4 block below the door
        tc = rng.permutation(11 * 5)[:e]
        x[k, b:], y[k, b:] = tc % 11 + 9, tc // 11 + 5       # inside the tram
    seq = [rng.integers(0, 5, size=(n, A)).astype(np.int8) for _ in range(40)]
    run_case(cfg, x, y, f, seq, f"crowd {crew}")
