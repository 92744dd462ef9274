"""The C oracle pinned directly against the UNMODIFIED reference (imported from /root/reference
behind stub modules).  Skipped where the reference is not mounted (the GPU box)."""

import json

import numpy as np
import pytest
from cases import TERMS, REWARDS, cassette_config, large_config, readme_config

from oracle import refload, refrun

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference sources not mounted")


@pytest.mark.parametrize("source", ["random", "greedy", "waiting"])
@pytest.mark.parametrize("term", list(TERMS))
@pytest.mark.parametrize("reward", list(REWARDS))
def test_strategy_matrix(source, term, reward):
    cfg = readme_config(reward, term, max_steps=30)
    rec = refrun.record(cfg, range(3), 40, source=source, stream_seed=7, shuffle_order=source == "random",
                        drop_prob=0.2 if source == "random" else 0.0)
    refrun.compare(rec, refrun.replay_with_oracle(cfg, rec))
    if source != "random":
        refrun.compare(rec, refrun.replay_with_oracle(cfg, rec, policy=source), what="oracle-policy", policy_actions=True)


def test_reference_cassettes_replay_through_oracle():
    """golden_basic_trajectory.json / regression_test.json, straight from the reference tree."""
    from golden.make_golden import convert_cassette
    from helpers import CASSETTE_BITS, assert_same, replay_oracle

    gold = refload.REFERENCE_TESTS / "fixtures" / "trajectories" / "golden"
    for name in ("golden_basic_trajectory.json", "regression_test.json"):
        cas = json.loads((gold / name).read_text())
        assert cas["config"]["truncated_config"]["max_steps"] == 50
        rec = convert_cassette(gold / name)
        assert_same(rec, replay_oracle(cassette_config(), rec), name, flag_mask=CASSETTE_BITS)


def test_large_config_and_seeded_reset():
    cfg = large_config(20)
    rec = refrun.record(cfg, [5], 24, source="greedy", validate=False)
    refrun.compare(rec, refrun.replay_with_oracle(cfg, rec, policy="greedy"), policy_actions=True)


def test_random_geometries():
    from collectivecrossing_b200.configs import CollectiveCrossingConfig
    from collectivecrossing_b200.lowering import lower_config
    from collectivecrossing_b200.truncated_configs import MaxStepsTruncatedConfig

    rng = np.random.default_rng(123)
    done = 0
    while done < 12:
        W, H = int(rng.integers(3, 30)), int(rng.integers(3, 20))
        D, L = int(rng.integers(1, H)), int(rng.integers(1, W + 1))
        dl = int(rng.integers(0, L))
        dr = int(rng.integers(dl, L))
        B, E = int(rng.integers(0, 6)), int(rng.integers(0, 5))
        try:
            cfg = CollectiveCrossingConfig(
                width=W, height=H, division_y=D, tram_door_left=dl, tram_door_right=dr, tram_length=L,
                num_boarding_agents=B, num_exiting_agents=E, exiting_destination_area_y=int(rng.integers(0, D)),
                boarding_destination_area_y=int(rng.integers(D, H + 1)),
                truncated_config=MaxStepsTruncatedConfig(max_steps=int(rng.integers(1, 40))),
                terminated_config=list(TERMS.values())[int(rng.integers(0, 2))](),
            )
        except ValueError:
            continue
        low = lower_config(cfg) if B + E else None
        if low is None:
            continue
        free = max(0, low.tram_right - low.tram_left - 1) * (H - D - 1) + max(0, low.door_right - low.door_left - 1)
        if E > free:
            continue  # reference reset() would spin forever
        for source in ("random", "greedy", "waiting"):
            rec = refrun.record(cfg, range(2), 30, source=source, stream_seed=done, shuffle_order=True, drop_prob=0.15)
            refrun.compare(rec, refrun.replay_with_oracle(cfg, rec))
            if source != "random":
                refrun.compare(rec, refrun.replay_with_oracle(cfg, rec, policy=source), policy_actions=True)
        done += 1
