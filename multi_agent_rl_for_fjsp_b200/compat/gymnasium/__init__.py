"""Stand-in for `gymnasium` (absent from this image): only `gymnasium.spaces` is provided."""
from . import spaces  # noqa: F401
from .spaces import Space  # noqa: F401

__version__ = "0.0-fjsp-b200-standin"
