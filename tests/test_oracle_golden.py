"""The C oracle against the committed golden vectors (generated from the unmodified reference by
tests/golden/make_golden.py) — runs everywhere, including the GPU box."""

import numpy as np
import pytest
from cases import GOLDEN_CASES, cassette_config
from helpers import CASSETTE_BITS, assert_same, load_golden, replay_oracle

from collectivecrossing_b200.lowering import lower_config

POLICY_CASES = [n for n, (_, kw) in GOLDEN_CASES.items() if kw["source"] in ("greedy", "waiting")]


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_oracle_replays_golden(name):
    cfg = GOLDEN_CASES[name][0]()
    rec = load_golden(name)
    assert_same(rec, replay_oracle(cfg, rec), f"oracle/{name}")


@pytest.mark.parametrize("name", POLICY_CASES)
def test_oracle_policies_match_reference_policies(name):
    """With the policy evaluated by the oracle itself (no action tensor), the whole trajectory —
    including every action the reference's Greedy/WaitingPolicy chose — is reproduced."""
    cfg, kw = GOLDEN_CASES[name][0](), GOLDEN_CASES[name][1]
    rec = load_golden(name)
    assert_same(rec, replay_oracle(cfg, rec, policy=kw["source"]), f"oracle-policy/{name}", policy_actions=True)


@pytest.mark.parametrize("name", ["cassette_basic", "cassette_regression"])
def test_oracle_replays_reference_cassettes(name):
    """The reference's own golden cassettes (converted, see make_golden.py): exact float64 rewards
    (-0.30000000000000004 ...), infos, x == width reachable, wall-blocked exiter."""
    rec = load_golden(name)
    got = replay_oracle(cassette_config(), rec)
    assert_same(rec, got, f"oracle/{name}", flag_mask=CASSETTE_BITS)
    if name == "cassette_basic":
        assert rec["obs"][9, 0, 0, 0] == 10  # boarding_0 stands on x == width at step 9
        assert -0.30000000000000004 in rec["reward"]


def test_oracle_seeded_reset_matches_golden_initial_states():
    """reset(seed) restated with PCG64/SeedSequence/Lemire reproduces the reference's placements."""
    import oracle

    for name, (make_cfg, kw) in GOLDEN_CASES.items():
        rec = load_golden(name)
        o = oracle.OracleEnvs(lower_config(make_cfg()), len(rec["seeds"]))
        obs = o.reset_seeded(rec["seeds"])
        assert np.array_equal(o.x, rec["init_x"]) and np.array_equal(o.y, rec["init_y"]), name
        assert np.array_equal(obs, rec["init_obs"]), name
    o = oracle.OracleEnvs(lower_config(cassette_config()), 1)
    o.reset_seeded([42])
    assert o.x.tolist() == [[0, 6, 4]] and o.y.tolist() == [[2, 1, 5]]  # SURVEY.md §4 cassette row
