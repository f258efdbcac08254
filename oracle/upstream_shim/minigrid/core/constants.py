"""Constants of the public minigrid API (restated)."""
import numpy as np

COLOR_TO_IDX = {"red": 0, "green": 1, "blue": 2, "purple": 3, "yellow": 4, "grey": 5}
IDX_TO_COLOR = {v: k for k, v in COLOR_TO_IDX.items()}
# upstream sorts the colour names; `choice(COLOR_NAMES)` in custom_env.py:635 depends on it
COLOR_NAMES = sorted(COLOR_TO_IDX.keys())

OBJECT_TO_IDX = {
    "unseen": 0, "empty": 1, "wall": 2, "floor": 3, "door": 4, "key": 5,
    "ball": 6, "box": 7, "goal": 8, "lava": 9, "agent": 10,
}
IDX_TO_OBJECT = {v: k for k, v in OBJECT_TO_IDX.items()}

STATE_TO_IDX = {"open": 0, "closed": 1, "locked": 2}

# 0 east, 1 south, 2 west, 3 north (y grows downwards)
DIR_TO_VEC = [np.array((1, 0)), np.array((0, 1)), np.array((-1, 0)), np.array((0, -1))]
