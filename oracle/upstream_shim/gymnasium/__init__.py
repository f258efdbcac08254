"""Restated sliver of the public `gymnasium` API (TEST INFRASTRUCTURE ONLY).

Just enough for the unmodified reference `environment.py` wrappers and the
restated MiniGridEnv to run headless: Env (np_random + seeding), Wrapper,
ObservationWrapper, spaces.{Box,Dict,Discrete}, logger.  See ../README.md.
"""
import numpy as np

from . import spaces  # noqa: F401

__shim__ = True


class _Logger:
    DEBUG, INFO, WARN, ERROR, DISABLED = 10, 20, 30, 40, 50
    min_level = 30


logger = _Logger()


class Env:
    _np_random = None
    render_mode = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.Generator(np.random.PCG64())
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    @property
    def unwrapped(self):
        return self

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.Generator(np.random.PCG64(seed))

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self._observation_space = None
        self._action_space = None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def observation_space(self):
        if self._observation_space is None:
            return self.env.observation_space
        return self._observation_space

    @observation_space.setter
    def observation_space(self, space):
        self._observation_space = space

    @property
    def action_space(self):
        if self._action_space is None:
            return self.env.action_space
        return self._action_space

    @action_space.setter
    def action_space(self, space):
        self._action_space = space

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, action):
        return self.env.step(action)

    def close(self):
        return self.env.close()


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, observation):
        raise NotImplementedError


def make(*_a, **_k):
    raise NotImplementedError("only env_name == 'custom' is on the hot path (environment.py:11-12)")
