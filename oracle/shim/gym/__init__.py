"""Test-only stand-in for legacy `gym`; distinct classes so isinstance tuples in pufferlib.spaces work."""
from . import spaces
from gymnasium import Env, Wrapper  # noqa: F401
