"""Known-answer tests of the two generators the oracle restates."""

import numpy as np
import oracle
import pytest


def test_philox4x32_10_random123_known_answers():
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert oracle.philox4x32_10(ctr, key).tolist() == want


@pytest.mark.parametrize("seed", [0, 1, 42, 2**31, 2**32 - 1, 2**32, 2**40 + 7, 2**63 - 1])
def test_pcg64_integers_match_numpy(seed):
    """gymnasium's np_random(seed) == Generator(PCG64(SeedSequence(seed))); integers(low, high) on
    ranges below 2**32 is Lemire's method over buffered 32-bit halves."""
    for lo, hi in [(0, 12), (0, 4), (2, 11), (4, 8), (0, 1), (0, 100), (-5, 5), (0, 3)]:
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        want = np.array([g.integers(lo, hi) for _ in range(300)])
        assert np.array_equal(oracle.pcg64_integers(seed, lo, hi, 300), want), (seed, lo, hi)


def test_pcg64_mixed_ranges_match_numpy():
    """reset() alternates ranges, some of them of width 1 (no draw consumed)."""
    from cases import readme_config

    from collectivecrossing_b200.lowering import lower_config

    low = lower_config(readme_config())
    for seed in range(50):
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        # replay the reference's reset loop with numpy directly (collectivecrossing.py:101-150)
        placed = []
        for i in range(low.num_agents):
            while True:
                if i < low.num_boarding:
                    x, y = int(g.integers(0, low.width)), int(g.integers(0, low.division_y))
                    ok = not (low.door_left <= x <= low.door_right and y == low.division_y - 1)
                else:
                    x, y = int(g.integers(low.tram_left, low.tram_right + 1)), int(g.integers(low.division_y, low.height))
                    ok = low.tram_left < x < low.tram_right and (y != low.division_y or low.door_left < x < low.door_right)
                if ok and (x, y) not in placed:
                    placed.append((x, y))
                    break
        o = oracle.OracleEnvs(low, 1)
        o.reset_seeded([seed])
        assert list(zip(o.x[0].tolist(), o.y[0].tolist())) == placed, seed
