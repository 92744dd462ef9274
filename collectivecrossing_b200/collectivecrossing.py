"""Import-path compatibility: the reference exposes the env as
``collectivecrossing.collectivecrossing.CollectiveCrossingEnv``."""

from .env import CollectiveCrossingEnv  # noqa: F401

__all__ = ["CollectiveCrossingEnv"]
