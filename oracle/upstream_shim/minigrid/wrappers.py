"""Wrappers of the public minigrid API that the reference imports (restated).

FullyObsWrapper is the one `experts_test.py:29` uses; RGBImgObsWrapper is only
imported by `environment.py:3` (never applied on the PPO path) and is a stub.
"""
import gymnasium as gym
from gymnasium import spaces

from .core.constants import COLOR_TO_IDX, OBJECT_TO_IDX


class FullyObsWrapper(gym.ObservationWrapper):
    def __init__(self, env):
        gym.ObservationWrapper.__init__(self, env)
        base = env
        while hasattr(base, "env"):
            base = base.env
        self._base = base
        new_image_space = spaces.Box(low=0, high=255,
                                     shape=(base.width, base.height, 3), dtype="uint8")
        self.observation_space = spaces.Dict(
            {**self.observation_space.spaces, "image": new_image_space})

    def observation(self, obs):
        env = self._base
        full_grid = env.grid.encode()
        full_grid[env.agent_pos[0]][env.agent_pos[1]] = (
            OBJECT_TO_IDX["agent"], COLOR_TO_IDX["red"], env.agent_dir)
        return {**obs, "image": full_grid}


class RGBImgObsWrapper(gym.ObservationWrapper):
    def __init__(self, env, tile_size=8):
        raise NotImplementedError("rendering is outside the hot path (SURVEY.md §2 row 3)")
