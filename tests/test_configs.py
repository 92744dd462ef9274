"""API surface of the config classes (mirrors what the reference's tests rely on)."""

import pytest
from cases import readme_config
from pydantic import ValidationError

from collectivecrossing_b200 import _abi
from collectivecrossing_b200.configs import CollectiveCrossingConfig
from collectivecrossing_b200.lowering import lower_config
from collectivecrossing_b200.observation_configs import get_observation_config
from collectivecrossing_b200.reward_configs import REWARD_CONFIGS, CustomRewardConfig, get_reward_config
from collectivecrossing_b200.terminated_configs import CustomTerminatedConfig, get_terminated_config
from collectivecrossing_b200.truncated_configs import MaxStepsTruncatedConfig, get_truncated_config
from collectivecrossing_b200.utils.geometry import calculate_distance, calculate_tram_boundaries


def test_readme_geometry_lowers_to_surveyed_constants():
    low = lower_config(readme_config())
    assert (low.tram_left, low.tram_right, low.door_left, low.door_right) == (2, 10, 7, 9)
    assert (low.num_agents, low.obs_len, low.max_steps) == (8, 38, 100)
    assert list(low.reward_params) == [15.0, 10.0, 5.0, 0.1]
    tb = calculate_tram_boundaries(readme_config())
    assert (tb.tram_left, tb.tram_door_right) == (2, 9)


def test_frozen_forbid_extra_and_model_copy():
    cfg = readme_config()
    with pytest.raises(ValidationError):
        cfg.width = 3
    with pytest.raises(ValidationError):
        CollectiveCrossingConfig(**{**cfg.model_dump(exclude={"reward_config", "terminated_config", "truncated_config", "observation_config"}), "bogus": 1})
    c2 = cfg.model_copy(update={"truncated_config": MaxStepsTruncatedConfig(max_steps=1)})
    assert lower_config(c2).max_steps == 1 and lower_config(cfg).max_steps == 100


@pytest.mark.parametrize("field,value,fragment", [
    ("tram_length", 13, "cannot exceed grid width"),
    ("tram_door_left", 9, "must be within tram boundaries"),
    ("tram_door_right", 4, "cannot be greater than"),
    ("exiting_destination_area_y", 4, "must be within waiting area"),
    ("boarding_destination_area_y", 3, "must be within tram area"),
    ("division_y", 8, "must be less than environment height"),
    ("num_boarding_agents", 40, "exceeds reasonable limit"),
    ("render_mode", "ansi", "Invalid render_mode"),
])
def test_validators_reject_like_the_reference(field, value, fragment):
    base = readme_config().model_dump(exclude={"reward_config", "terminated_config", "truncated_config", "observation_config"})
    base[field] = value
    with pytest.raises(ValueError, match=fragment):
        CollectiveCrossingConfig(**base)
    loose = CollectiveCrossingConfig.model_construct(**base)
    assert not loose.is_valid() and any(fragment in e for e in loose.get_validation_errors())


def test_factories_and_registries():
    assert get_reward_config("binary", goal_reward=2.0).goal_reward == 2.0
    assert get_reward_config("constant_negative", step_penalty=-2.5, reward_function="ignored").step_penalty == -2.5
    assert get_terminated_config("all_at_destination").get_terminated_function_name() == "all_at_destination"
    assert get_truncated_config("max_steps", max_steps=5).max_steps == 5
    assert get_observation_config("default").get_observation_function_name() == "default"
    assert set(REWARD_CONFIGS) == {"default", "simple_distance", "binary", "constant_negative", "custom"}
    with pytest.raises(ValueError, match="Unknown reward function 'nope'"):
        get_reward_config("nope")
    with pytest.raises(ValidationError):
        get_reward_config("constant_negative", step_penalty=1.0)
    with pytest.raises(ValidationError):
        get_truncated_config("max_steps", max_steps=0)


def test_custom_strategy_names_raise_at_lowering_like_env_construction():
    """reference: get_reward_function raises ValueError for unregistered names (rewards.py:210-214,
    tests/.../test_rewards.py:210-219, test_terminateds.py:135-145)."""
    base = readme_config()
    with pytest.raises(ValueError, match="Unknown reward function 'my_reward'"):
        lower_config(base.model_copy(update={"reward_config": CustomRewardConfig(reward_function="my_reward")}))
    with pytest.raises(ValueError, match="Unknown termination function"):
        lower_config(base.model_copy(update={"terminated_config": CustomTerminatedConfig(terminated_function="mine")}))


def test_lowering_limits():
    with pytest.raises(ValueError, match="agents per env"):
        lower_config(CollectiveCrossingConfig.model_construct(**{**readme_config().__dict__, "num_boarding_agents": 100, "num_exiting_agents": 100}))
    assert _abi.REWARD_KINDS["binary"] == 2 and _abi.TERMINATED_KINDS["all_at_destination"] == 1


def test_calculate_distance_none_aware():
    assert calculate_distance((3, 4), (None, 9)) == 5
    assert calculate_distance((3, 4), (7, None)) == 4
    assert calculate_distance((0, 0), (3, 4)) == 5.0
