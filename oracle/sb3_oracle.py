"""numpy restatement of the SB3 vec-env wrappers on the reference's path (TEST INFRASTRUCTURE).

[UPSTREAM] stable_baselines3 (not vendored by the reference; call sites ppo.py:118-126):
  VecTransposeImage       image (N,7,7,3) -> (N,3,7,7)
  VecFrameStack(4,'first') via StackedObservations.update: roll by one frame along axis 1,
                          zero the stack of finished envs, append the newest frame; the
                          terminal observation gets the three previous frames prepended.
plus environment.py:91-112 (mission tokens) and :144-149 (direction one-hot).
"""
import numpy as np

VOCAB = [" ", "\n", "-", ":", ",", "."] + [chr(c) for c in range(ord("a"), ord("z") + 1)]


def tokenize(mission: str) -> np.ndarray:       # environment.py:91-101
    out = np.zeros(32, np.int64)
    for i, ch in enumerate(mission.lower()):
        out[i] = VOCAB.index(ch)
    return out


def one_hot_dir(d: np.ndarray) -> np.ndarray:   # environment.py:144-149
    out = np.zeros((d.shape[0], 4), np.uint8)
    out[np.arange(d.shape[0]), d] = 1
    return out


class FrameStack:
    """StackedObservations for one key, channels_order='first' (stack along axis 1)."""

    def __init__(self, n, frame_shape, dtype, n_stack=4):
        self.c = frame_shape[0]
        self.stacked = np.zeros((n, self.c * n_stack) + tuple(frame_shape[1:]), dtype)

    def reset(self, obs):
        self.stacked[...] = 0
        self.stacked[:, -self.c:] = obs
        return self.stacked.copy()

    def update(self, obs, dones, terminal):
        """returns (stacked, {env: stacked terminal observation})"""
        self.stacked = np.roll(self.stacked, -self.c, axis=1)
        term = {}
        for i in np.flatnonzero(dones):
            term[int(i)] = np.concatenate([self.stacked[i, :-self.c], terminal[i]], axis=0)
            self.stacked[i] = 0
        self.stacked[:, -self.c:] = obs
        return self.stacked.copy(), term
