"""Test-only stand-in for `gymnasium` so the unmodified reference imports here (see spaces.py)."""
from . import spaces


class Env:
    metadata = {}
    render_mode = None

    def reset(self, seed=None, options=None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self):
        return None

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = env.observation_space
        self.action_space = env.action_space

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()

    @property
    def render_mode(self):
        return self.env.render_mode

    @property
    def unwrapped(self):
        return getattr(self.env, 'unwrapped', self.env)
