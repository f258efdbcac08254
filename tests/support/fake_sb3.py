"""Stand-ins for the two pieces of Stable-Baselines3 that consume the vec-env (TEST INFRASTRUCTURE; SB3 is not installable
here).  Restated from the public SB3 2.x behaviour, [UPSTREAM] and unpinned like oracle/sb3_oracle.py:

* `VecEnv`: the abstract base class (`stable_baselines3.common.vec_env.base_vec_env.VecEnv`): what its constructor sets and
  checks (`render_mode` through `get_attr`), `step = step_async + step_wait`, `seed`, the abstract method set.
* `collect_rollouts`: the loop of `OnPolicyAlgorithm.collect_rollouts` as far as it touches the environment: observation
  dict -> arrays of the spaces' shapes and dtypes (`obs_as_tensor` / `DictRolloutBuffer.add`), `env.step(actions)`, the
  info scan of `_update_info_buffer`, and the truncation bootstrap
  `rewards[idx] += gamma * V(infos[idx]["terminal_observation"])` for `TimeLimit.truncated`.
"""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod

import numpy as np


class VecEnv(ABC):
    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.reset_infos = [{} for _ in range(num_envs)]
        self._seeds = [None for _ in range(num_envs)]
        self._options = [{} for _ in range(num_envs)]
        try:
            render_modes = self.get_attr("render_mode")
        except AttributeError:
            warnings.warn("The `render_mode` attribute is not defined in your environment.")
            render_modes = [None for _ in range(num_envs)]
        assert all(m == render_modes[0] for m in render_modes), "render_mode mode should be the same for all environments"
        self.render_mode = render_modes[0]
        self.metadata = {"render_modes": [] if self.render_mode is None else [self.render_mode]}

    @abstractmethod
    def reset(self): ...

    @abstractmethod
    def step_async(self, actions): ...

    @abstractmethod
    def step_wait(self): ...

    @abstractmethod
    def close(self): ...

    @abstractmethod
    def get_attr(self, attr_name, indices=None): ...

    @abstractmethod
    def set_attr(self, attr_name, value, indices=None): ...

    @abstractmethod
    def env_method(self, method_name, *method_args, indices=None, **method_kwargs): ...

    @abstractmethod
    def env_is_wrapped(self, wrapper_class, indices=None): ...

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def seed(self, seed=None):
        if seed is None:
            seed = int(np.random.randint(0, np.iinfo(np.uint32).max, dtype=np.uint32))
        self._seeds = [seed + idx for idx in range(self.num_envs)]
        return self._seeds

    @property
    def unwrapped(self):
        return self


def obs_shapes(space):
    """get_obs_shape for a Dict of Box / Discrete spaces"""
    return {k: (tuple(s.shape) if len(tuple(s.shape)) else (1,)) for k, s in space.spaces.items()}


def collect_rollouts(env, n_steps, act, value_of, gamma, seed=0):
    """Returns the rollout buffer (dict of [n_steps, n_envs, ...] arrays), the episode info buffer and the last obs."""
    assert isinstance(env, VecEnv), "SB3's _wrap_env would wrap a non-VecEnv in DummyVecEnv"
    n = env.num_envs
    shapes = obs_shapes(env.observation_space)
    buf = {k: np.zeros((n_steps, n) + shp, env.observation_space.spaces[k].dtype) for k, shp in shapes.items()}
    buf.update(actions=np.zeros((n_steps, n), np.int64), rewards=np.zeros((n_steps, n), np.float32),
               episode_starts=np.zeros((n_steps, n), np.float32))
    ep_info = []
    env.seed(seed)
    last_obs = env.reset()
    last_starts = np.ones(n, dtype=bool)
    for t in range(n_steps):
        for k in shapes:
            arr = np.asarray(last_obs[k])
            assert arr.shape == (n,) + tuple(env.observation_space.spaces[k].shape), (k, arr.shape)
            assert arr.dtype == env.observation_space.spaces[k].dtype, (k, arr.dtype)
            buf[k][t] = arr.reshape((n,) + shapes[k])                  # DictRolloutBuffer.add copies
        actions = act(t, last_obs)
        new_obs, rewards, dones, infos = env.step(actions)
        rewards = np.array(rewards, np.float32, copy=True)
        for idx, info in enumerate(infos):                                # _update_info_buffer
            maybe = info.get("episode")
            if maybe is not None:
                ep_info.append((t, idx, maybe["r"], maybe["l"]))
        for idx, done in enumerate(dones):                                # truncation bootstrap
            if done and infos[idx].get("terminal_observation") is not None and infos[idx].get("TimeLimit.truncated", False):
                rewards[idx] += np.float32(gamma) * np.float32(value_of(infos[idx]["terminal_observation"]))
        buf["actions"][t], buf["rewards"][t], buf["episode_starts"][t] = actions, rewards, last_starts
        last_obs, last_starts = new_obs, dones
    return buf, ep_info, last_obs
