"""Pydantic models of the CollectiveCrossing API surface (one module, re-exported under the
reference's module names by ``configs.py``, ``reward_configs.py``, ...).

Field names, defaults, bounds and validation outcomes follow the reference so that user code
constructing configs keeps working unchanged:

* ``ConfigClass``                      -> reference ``utils/pydantic.py:6-30``
* ``CollectiveCrossingConfig``         -> ``configs.py:15-241``
* reward configs + registry/factory    -> ``reward_configs.py:12-197``
* terminated configs                   -> ``terminated_configs.py:11-114``
* truncated configs                    -> ``truncated_configs.py:11-109``
* observation configs                  -> ``observation_configs.py:11-74``

The strategy *functions* of the reference (rewards.py, terminateds.py, ...) have no Python
counterpart here: a config is lowered once to the POD ``cc_config`` (``lowering.py``) and the
strategy is evaluated inside the fused CUDA step kernel.
"""

from __future__ import annotations

from typing import Any, Callable

from pydantic import BaseModel, ConfigDict, Field, model_validator


class ConfigClass(BaseModel):
    """Frozen, strict base model (reference ``utils/pydantic.py:6-30``)."""

    model_config = ConfigDict(
        extra="forbid",
        frozen=True,
        validate_assignment=True,
        validate_default=True,
        arbitrary_types_allowed=False,
        use_enum_values=True,
        populate_by_name=True,
        validate_by_name=True,
        loc_by_alias=True,
    )


def _bounded(default: Any, lo: float, hi: float, doc: str) -> Any:
    return Field(default=default, ge=lo, le=hi, description=doc)


# --------------------------------------------------------------------------------------------
# rewards  (reference reward_configs.py)
# --------------------------------------------------------------------------------------------
class RewardConfig(ConfigClass):
    """Base of the reward configs; ``reward_function`` selects the strategy by registry name."""

    reward_function: str = Field(description="registry name of the reward strategy")

    def get_reward_function_name(self) -> str:
        return self.reward_function


class DefaultRewardConfig(RewardConfig):
    reward_function: str = Field(default="default", description="default reward strategy")
    boarding_destination_reward: float = _bounded(15.0, -100.0, 100.0, "paid on arrival (both agent types)")
    tram_door_reward: float = _bounded(10.0, -100.0, 100.0, "paid next to the door cells")
    tram_area_reward: float = _bounded(5.0, -100.0, 100.0, "paid inside (boarding) / outside (exiting) the tram")
    distance_penalty_factor: float = _bounded(0.1, 0.0, 10.0, "scale of the Manhattan door distance term")

    def get_reward_function_name(self) -> str:
        return "default"


class SimpleDistanceRewardConfig(RewardConfig):
    reward_function: str = Field(default="simple_distance", description="distance-to-goal reward")
    distance_penalty_factor: float = _bounded(0.1, 0.0, 10.0, "scale of the |y - y_goal| term")

    def get_reward_function_name(self) -> str:
        return "simple_distance"


class BinaryRewardConfig(RewardConfig):
    reward_function: str = Field(default="binary", description="goal / no-goal reward")
    goal_reward: float = _bounded(1.0, 0.0, 100.0, "reward at the goal")
    no_goal_reward: float = _bounded(0.0, -100.0, 100.0, "reward elsewhere")

    def get_reward_function_name(self) -> str:
        return "binary"


class ConstantNegativeRewardConfig(RewardConfig):
    reward_function: str = Field(default="constant_negative", description="fixed penalty per step")
    step_penalty: float = _bounded(-1.0, -100.0, 0.0, "reward handed out every step")

    def get_reward_function_name(self) -> str:
        return "constant_negative"


class CustomRewardConfig(RewardConfig):
    """Placeholder kept for API parity; no strategy is registered under a custom name, so an env
    built with it raises ``ValueError`` exactly like the reference (rewards.py:210-214)."""

    reward_function: str = Field(description="name of a user strategy")
    time_penalty: float = _bounded(0.0, -10.0, 0.0, "penalty per step")
    goal_bonus: float = _bounded(0.0, 0.0, 100.0, "bonus at the goal")
    collision_penalty: float = _bounded(0.0, -100.0, 0.0, "penalty per collision")
    efficiency_bonus: float = _bounded(0.0, 0.0, 100.0, "bonus for short paths")


# --------------------------------------------------------------------------------------------
# termination  (reference terminated_configs.py)
# --------------------------------------------------------------------------------------------
class TerminatedConfig(ConfigClass):
    terminated_function: str = Field(description="registry name of the termination strategy")

    def get_terminated_function_name(self) -> str:
        return self.terminated_function


class AllAtDestinationTerminatedConfig(TerminatedConfig):
    terminated_function: str = Field(default="all_at_destination", description="nobody terminates before everybody arrived")

    def get_terminated_function_name(self) -> str:
        return "all_at_destination"


class IndividualAtDestinationTerminatedConfig(TerminatedConfig):
    terminated_function: str = Field(default="individual_at_destination", description="an agent terminates on its own arrival")

    def get_terminated_function_name(self) -> str:
        return "individual_at_destination"


class CustomTerminatedConfig(TerminatedConfig):
    terminated_function: str = Field(description="name of a user strategy")
    max_steps_per_agent: int = _bounded(1000, 1, 10000, "per-agent step budget")
    require_all_completion: bool = Field(default=False, description="wait for everybody")
    timeout_penalty: bool = Field(default=False, description="penalise unfinished agents")


# --------------------------------------------------------------------------------------------
# truncation  (reference truncated_configs.py)
# --------------------------------------------------------------------------------------------
class TruncatedConfig(ConfigClass):
    truncated_function: str = Field(description="registry name of the truncation strategy")

    def get_truncated_function_name(self) -> str:
        return self.truncated_function


class MaxStepsTruncatedConfig(TruncatedConfig):
    truncated_function: str = Field(default="max_steps", description="truncate at max_steps")
    max_steps: int = _bounded(1000, 1, 100000, "episode length limit")

    def get_truncated_function_name(self) -> str:
        return "max_steps"


class CustomTruncatedConfig(TruncatedConfig):
    truncated_function: str = Field(description="name of a user strategy")
    max_steps: int = _bounded(1000, 1, 100000, "episode length limit")
    early_truncation_threshold: float = _bounded(0.0, 0.0, 1.0, "0 disables early truncation")
    require_all_agents_active: bool = Field(default=False, description="only truncate full crews")


# --------------------------------------------------------------------------------------------
# observation  (reference observation_configs.py)
# --------------------------------------------------------------------------------------------
class ObservationConfig(ConfigClass):
    observation_function: str = Field(description="registry name of the observation strategy")

    def get_observation_function_name(self) -> str:
        return self.observation_function


class DefaultObservationConfig(ObservationConfig):
    observation_function: str = Field(default="default", description="own position, door info, every agent's (x, y, type, active)")

    def get_observation_function_name(self) -> str:
        return "default"


# --------------------------------------------------------------------------------------------
# registries + factories (reference: REWARD_CONFIGS / get_reward_config etc.)
# --------------------------------------------------------------------------------------------
REWARD_CONFIGS: dict[str, type[RewardConfig]] = {
    "default": DefaultRewardConfig,
    "simple_distance": SimpleDistanceRewardConfig,
    "binary": BinaryRewardConfig,
    "constant_negative": ConstantNegativeRewardConfig,
    "custom": CustomRewardConfig,
}
TERMINATED_CONFIGS: dict[str, type[TerminatedConfig]] = {
    "all_at_destination": AllAtDestinationTerminatedConfig,
    "individual_at_destination": IndividualAtDestinationTerminatedConfig,
    "custom": CustomTerminatedConfig,
}
TRUNCATED_CONFIGS: dict[str, type[TruncatedConfig]] = {
    "max_steps": MaxStepsTruncatedConfig,
    "custom": CustomTruncatedConfig,
}
OBSERVATION_CONFIGS: dict[str, type[ObservationConfig]] = {"default": DefaultObservationConfig}


def _factory(registry: dict[str, type], key: str, noun: str) -> Callable[..., Any]:
    def make(name: str, **kwargs: Any) -> Any:
        if name not in registry:
            raise ValueError(f"Unknown {noun} function '{name}'. Available: {', '.join(registry)}")
        kwargs.pop(key, None)
        return registry[name](**{key: name}, **kwargs)

    make.__doc__ = f"Build the {noun} config registered under ``name`` (extra kwargs are its fields)."
    return make


get_reward_config = _factory(REWARD_CONFIGS, "reward_function", "reward")
get_terminated_config = _factory(TERMINATED_CONFIGS, "terminated_function", "termination")
get_truncated_config = _factory(TRUNCATED_CONFIGS, "truncated_function", "truncation")
get_observation_config = _factory(OBSERVATION_CONFIGS, "observation_function", "observation")


# --------------------------------------------------------------------------------------------
# environment config  (reference configs.py)
# --------------------------------------------------------------------------------------------
_ENV_CHECKS = (
    ("Tram parameter error", "_validate_tram_parameters"),
    ("Destination area error", "_validate_destination_areas"),
    ("Environment bounds error", "_validate_environment_bounds"),
    ("Agent count error", "_validate_agent_counts"),
    ("Render mode error", "_validate_render_mode"),
)


def _dim(doc: str, lo: int = 1) -> Any:
    return Field(description=doc, ge=lo, le=100)


class CollectiveCrossingConfig(ConfigClass):
    """Geometry, crew and strategy selection of one environment (reference configs.py:15-77).

    Door coordinates are RELATIVE to the tram's left edge; absolute coordinates are derived by
    ``utils.geometry.calculate_tram_boundaries``.  Use ``model_construct`` to bypass the agent
    count cap for large crews (BASELINE config 3), as one would with the reference.
    """

    width: int = _dim("grid width")
    height: int = _dim("grid height")
    division_y: int = _dim("y of the wall between waiting area and tram")
    tram_door_left: int = _dim("left door post, relative to the tram", 0)
    tram_door_right: int = _dim("right door post, relative to the tram", 0)
    tram_length: int = _dim("horizontal extent of the tram")
    num_boarding_agents: int = _dim("agents that want to get on", 0)
    num_exiting_agents: int = _dim("agents that want to get off", 0)
    render_mode: str | None = Field(default=None, description="'human', 'rgb_array' or None")
    exiting_destination_area_y: int = Field(description="row exiting agents walk to")
    boarding_destination_area_y: int = Field(description="row boarding agents walk to")
    observation_config: ObservationConfig = Field(default_factory=DefaultObservationConfig)
    reward_config: RewardConfig = Field(default_factory=DefaultRewardConfig)
    terminated_config: TerminatedConfig = Field(default_factory=IndividualAtDestinationTerminatedConfig)
    truncated_config: TruncatedConfig = Field(default_factory=MaxStepsTruncatedConfig)

    # Each checker raises ValueError with a message naming the offending values
    # (reference configs.py:89-195).
    def _validate_tram_parameters(self) -> None:
        if self.tram_length > self.width:
            raise ValueError(f"Tram length ({self.tram_length}) cannot exceed grid width ({self.width})")
        for side, v in (("left", self.tram_door_left), ("right", self.tram_door_right)):
            if not 0 <= v < self.tram_length:
                raise ValueError(
                    f"Tram door {side} boundary ({v}) must be within tram boundaries (0 to {self.tram_length - 1})"
                )
        if self.tram_door_left > self.tram_door_right:
            raise ValueError(
                f"Tram door left boundary ({self.tram_door_left}) cannot be greater than "
                f"right boundary ({self.tram_door_right})"
            )

    def _validate_destination_areas(self) -> None:
        if not 0 <= self.exiting_destination_area_y < self.division_y:
            raise ValueError(
                f"Exiting destination area y-coordinate ({self.exiting_destination_area_y}) must be "
                f"within waiting area (0 to {self.division_y - 1})"
            )
        if not self.division_y <= self.boarding_destination_area_y <= self.height:
            raise ValueError(
                f"Boarding destination area y-coordinate ({self.boarding_destination_area_y}) must be "
                f"within tram area ({self.division_y} to {self.height})"
            )

    def _validate_environment_bounds(self) -> None:
        if self.division_y >= self.height:
            raise ValueError(
                f"Division line y-coordinate ({self.division_y}) must be less than environment height ({self.height})"
            )
        for side, v in (("left", self.tram_door_left), ("right", self.tram_door_right)):
            if v >= self.width:
                raise ValueError(f"Tram door {side} boundary ({v}) must be less than environment width ({self.width})")

    def _validate_agent_counts(self) -> None:
        crew = self.num_boarding_agents + self.num_exiting_agents
        cap = min(self.width * self.height // 4, 50)
        if crew > cap:
            raise ValueError(
                f"Total number of agents ({crew}) exceeds reasonable limit ({cap}) for environment size "
                f"{self.width}x{self.height}"
            )
        waiting_cells = self.width * self.division_y
        tram_cells = self.width * (self.height - self.division_y)
        if self.num_exiting_agents > waiting_cells // 2:
            raise ValueError(
                f"Number of exiting agents ({self.num_exiting_agents}) may be too high for waiting area size ({waiting_cells})"
            )
        if self.num_boarding_agents > tram_cells // 2:
            raise ValueError(
                f"Number of boarding agents ({self.num_boarding_agents}) may be too high for tram area size ({tram_cells})"
            )

    def _validate_render_mode(self) -> None:
        allowed = ["human", "rgb_array", None]
        if self.render_mode not in allowed:
            raise ValueError(f"Invalid render_mode: {self.render_mode}. Valid modes are: {allowed}")

    @model_validator(mode="after")
    def validate_config(self) -> "CollectiveCrossingConfig":
        for _, name in _ENV_CHECKS:
            getattr(self, name)()
        return self

    def get_validation_errors(self) -> list[str]:
        """All validation messages without raising (only useful on ``model_construct``-ed objects)."""
        found = []
        for label, name in _ENV_CHECKS:
            try:
                getattr(self, name)()
            except ValueError as exc:
                found.append(f"{label}: {exc}")
        return found

    def is_valid(self) -> bool:
        return not self.get_validation_errors()
