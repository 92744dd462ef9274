"""Named configurations shared by the tests, the golden generator and the bench."""

from __future__ import annotations

from collectivecrossing_b200.configs import CollectiveCrossingConfig
from collectivecrossing_b200.observation_configs import DefaultObservationConfig
from collectivecrossing_b200.reward_configs import (
    BinaryRewardConfig,
    ConstantNegativeRewardConfig,
    DefaultRewardConfig,
    SimpleDistanceRewardConfig,
)
from collectivecrossing_b200.terminated_configs import (
    AllAtDestinationTerminatedConfig,
    IndividualAtDestinationTerminatedConfig,
)
from collectivecrossing_b200.truncated_configs import MaxStepsTruncatedConfig

REWARDS = {
    "default": DefaultRewardConfig,
    "simple_distance": SimpleDistanceRewardConfig,
    "binary": BinaryRewardConfig,
    "constant_negative": ConstantNegativeRewardConfig,
}
TERMS = {"individual": IndividualAtDestinationTerminatedConfig, "all": AllAtDestinationTerminatedConfig}


def readme_config(reward="default", term="individual", max_steps=100, **reward_kw):
    """BASELINE configs 1/2/4/5: the README's 12x8 grid, door 5-7, 5 boarding + 3 exiting."""
    return CollectiveCrossingConfig(
        width=12, height=8, division_y=4, tram_door_left=5, tram_door_right=7, tram_length=9,
        num_boarding_agents=5, num_exiting_agents=3, exiting_destination_area_y=0, boarding_destination_area_y=8,
        reward_config=REWARDS[reward](**reward_kw), terminated_config=TERMS[term](),
        truncated_config=MaxStepsTruncatedConfig(max_steps=max_steps),
    )


def readme_crew(boarding, exiting, reward="default", term="individual", max_steps=40):
    """The README geometry (small lattice: the policies use per-thread bitmaps) with another crew."""
    return CollectiveCrossingConfig(
        width=12, height=8, division_y=4, tram_door_left=5, tram_door_right=7, tram_length=9,
        num_boarding_agents=boarding, num_exiting_agents=exiting, exiting_destination_area_y=0, boarding_destination_area_y=8,
        reward_config=REWARDS[reward](), terminated_config=TERMS[term](), truncated_config=MaxStepsTruncatedConfig(max_steps=max_steps),
    )


def cassette_config():
    """The env of the reference's golden cassettes (tests/.../test_trajectory_vcr.py:323-340)."""
    return CollectiveCrossingConfig(
        width=10, height=6, division_y=3, tram_door_left=3, tram_door_right=5, tram_length=8,
        num_boarding_agents=2, num_exiting_agents=1, exiting_destination_area_y=0, boarding_destination_area_y=5,
        truncated_config=MaxStepsTruncatedConfig(max_steps=50), reward_config=DefaultRewardConfig(),
        terminated_config=IndividualAtDestinationTerminatedConfig(),
    )


def unchecked(**fields):
    """A config that bypasses the agent-count cap (reference configs.py:161-171), as BASELINE
    config 3 requires; every other field is still what the validated class would hold."""
    fields.setdefault("render_mode", None)
    fields.setdefault("observation_config", DefaultObservationConfig())
    fields.setdefault("reward_config", DefaultRewardConfig())
    fields.setdefault("terminated_config", IndividualAtDestinationTerminatedConfig())
    fields.setdefault("truncated_config", MaxStepsTruncatedConfig())
    return CollectiveCrossingConfig.model_construct(**fields)


def large_config(max_steps=512):
    """BASELINE config 3: 64x32, 48 boarding + 16 exiting, SimpleDistance, AllAtDestination
    (division / door / tram length as proposed in SURVEY.md §8)."""
    return unchecked(
        width=64, height=32, division_y=16, tram_door_left=24, tram_door_right=32, tram_length=56,
        num_boarding_agents=48, num_exiting_agents=16, exiting_destination_area_y=0, boarding_destination_area_y=32,
        reward_config=SimpleDistanceRewardConfig(), terminated_config=AllAtDestinationTerminatedConfig(),
        truncated_config=MaxStepsTruncatedConfig(max_steps=max_steps),
    )


def crew_config(boarding, exiting, width=40, height=24, max_steps=60, reward="default", term="individual"):
    """Mid-size grid with an arbitrary crew: exercises every lanes-per-env / agents-per-lane
    instantiation of the kernel (A<=4, 8, 16, 32, 64, 128)."""
    return unchecked(
        width=width, height=height, division_y=height // 2, tram_door_left=10, tram_door_right=16, tram_length=width - 6,
        num_boarding_agents=boarding, num_exiting_agents=exiting, exiting_destination_area_y=0,
        boarding_destination_area_y=height, reward_config=REWARDS[reward](), terminated_config=TERMS[term](),
        truncated_config=MaxStepsTruncatedConfig(max_steps=max_steps),
    )


# name -> (config factory, kwargs for oracle.refrun.record)
GOLDEN_CASES = {
    "readme_random": (lambda: readme_config(), dict(seeds=range(24), n_steps=110, source="random", stream_seed=11, shuffle_order=True, drop_prob=0.1)),
    "readme_greedy": (lambda: readme_config(), dict(seeds=range(100, 124), n_steps=110, source="greedy")),
    "readme_waiting": (lambda: readme_config(), dict(seeds=range(200, 224), n_steps=60, source="waiting")),
    "readme_binary_all": (lambda: readme_config("binary", "all", 40, goal_reward=1.0, no_goal_reward=-1.0), dict(seeds=range(8), n_steps=50, source="waiting")),
    "readme_constneg_ind": (lambda: readme_config("constant_negative", "individual", 40, step_penalty=-2.5), dict(seeds=range(8), n_steps=50, source="random", stream_seed=5)),
    "readme_simple_all": (lambda: readme_config("simple_distance", "all", 40, distance_penalty_factor=0.3), dict(seeds=range(8), n_steps=50, source="greedy")),
    "large_random": (lambda: large_config(30), dict(seeds=range(2), n_steps=36, source="random", stream_seed=3, validate=False)),
    "large_waiting": (lambda: large_config(512), dict(seeds=range(2, 4), n_steps=40, source="waiting", validate=False)),
    "crew_1_0": (lambda: crew_config(1, 0, max_steps=30), dict(seeds=range(4), n_steps=40, source="greedy", validate=False)),
    "crew_3_2": (lambda: crew_config(3, 2, max_steps=50), dict(seeds=range(4), n_steps=60, source="waiting", validate=False)),
    "crew_7_5": (lambda: crew_config(7, 5, max_steps=50, term="all"), dict(seeds=range(4), n_steps=60, source="random", stream_seed=8, shuffle_order=True, validate=False)),
    "crew_13_7": (lambda: crew_config(13, 7, max_steps=50), dict(seeds=range(3), n_steps=60, source="greedy", validate=False)),
    "crew_25_15": (lambda: crew_config(25, 15, max_steps=50, reward="simple_distance"), dict(seeds=range(2), n_steps=56, source="waiting", validate=False)),
    "crew_60_40": (lambda: crew_config(60, 40, max_steps=30), dict(seeds=range(1), n_steps=34, source="random", stream_seed=9, shuffle_order=True, validate=False)),
}
