"""B200-native batched MiniGrid simulator + PPO rollout engine.

Host-side mirror of the vec-env surface that Idokorro/MiniGrid-RL's ppo.py / policies.py /
experts.py consume (SB3 `VecEnv` after `VecTransposeImage` + `VecFrameStack(4,'first')`),
over a C-ABI CUDA library of hand-written sm_100a kernels (include/mgrl.h).  There is no
CPU implementation in this package: every op raises if the CUDA library or a GPU is missing.
"""
from .config import EnvConfig
from .missions import MISSIONS, expert_token_table, mission_id, mission_string, token_table, tokenize
from ._native import NativeError, lib, build_library, library_path
from .vec_env import B200VecEnv, DeviceEnv

__all__ = [
    "EnvConfig", "MISSIONS", "expert_token_table", "mission_id", "mission_string", "token_table", "tokenize",
    "NativeError", "lib", "build_library", "library_path", "B200VecEnv", "DeviceEnv",
]
from .policy import Policy  # noqa: E402
from .ppo import PPOConfig, RolloutEngine, Updater  # noqa: E402

__all__ += ["Policy", "PPOConfig", "RolloutEngine", "Updater"]
