"""Test-only stand-in for `pettingzoo` (only the base-class names the reference subclasses)."""


class ParallelEnv:
    pass


class AECEnv:
    pass
