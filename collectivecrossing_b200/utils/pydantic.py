"""Reference ``utils/pydantic.py``: the strict frozen base model."""

from .._models import ConfigClass  # noqa: F401

__all__ = ["ConfigClass"]
