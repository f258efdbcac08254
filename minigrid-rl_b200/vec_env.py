"""Vector-environment front-ends over the C-ABI CUDA library.

`B200VecEnv` is the Stable-Baselines3 `VecEnv` that the reference builds at
/root/reference/src/ppo.py:118-126 (make_vec_env(make_env) -> VecTransposeImage ->
VecFrameStack(4, channels_order='first')) and drives at ppo.py:134,145,159,161,210,242,265,292: numpy in,
numpy out, same observation dict (`image` (N,12,7,7) u8, `direction` (N,16) u8, `mission` (N,128)
i64), float32 rewards, bool dones and SB3-style infos (`terminal_observation`,
`TimeLimit.truncated`, `episode`).  When `stable_baselines3` / `gymnasium` are importable the class
derives from SB3's `VecEnv` and its spaces are `gymnasium.spaces` objects (SB3's `_wrap_env`,
`get_obs_shape`, `preprocess_obs` and `DictRolloutBuffer` check with isinstance); without them the
same class stands on `object` with the duck-typed spaces below, so the package needs neither.
`obs_mode="full"` is the experts' surface (experts_test.py:27-47: FullyObsWrapper + tokens).

`DeviceEnv` is the device-resident fast path used by the rollout engine and the benchmark:
torch CUDA tensors in and out, nothing crosses PCIe.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _native as nat
from .config import EnvConfig, PROBLEMS
from .missions import MISSIONS, expert_token_table, token_table

OBS_BYTES = 147
STATE_BYTES = 140
FRAMES = 4

# numpy view of the 140-byte state record (include/mgrl.h)
STATE_DTYPE = np.dtype([
    ("grid", np.uint8, (121,)),
    ("agent_x", np.uint8), ("agent_y", np.uint8), ("agent_dir", np.uint8),
    ("carrying", np.uint8), ("step_count", np.uint8),
    ("target_x", np.uint8), ("target_y", np.uint8), ("target_action", np.uint8),
    ("mission_id", np.uint8), ("mission_done", np.uint8), ("latch_step", np.uint8),
    ("episode", np.uint32), ("reset_draws", np.uint16), ("error", np.uint8), ("pad", np.uint8),
])


LAYOUTS = {"hwc": (0, 147), "chw": (1, 147), "hwc148": (2, 148)}


def _native_config(cfg: EnvConfig, num_envs: int, env_id_base: int, layout) -> nat.Config:
    cfg.validate()
    if isinstance(layout, bool):
        layout = "chw" if layout else "hwc"
    return nat.Config(cfg.size, cfg.num_objects, PROBLEMS[cfg.problem],
                      -1 if cfg.mission is None else int(cfg.mission), int(cfg.all_doors_open),
                      int(cfg.see_through_walls), cfg.max_steps, cfg.num_obstacles, int(num_envs),
                      LAYOUTS[layout][0], int(env_id_base))


class _Handle:
    """Owns one mgrl_env."""

    def __init__(self, cfg: EnvConfig, num_envs: int, device: int, env_id_base: int, chw: bool):
        self.lib = nat.lib()
        self.ncfg = _native_config(cfg, num_envs, env_id_base, chw)
        self.ptr = C.c_void_p()
        nat.check(self.lib.mgrl_create(C.byref(self.ncfg), int(device), C.byref(self.ptr)), "mgrl_create")

    def close(self):
        if self.ptr:
            self.lib.mgrl_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------- optional upstream base classes
def gym_spaces():
    """`gymnasium.spaces` when importable (looked up on every construction: cheap, and test-friendly), else None."""
    try:
        from gymnasium import spaces
        return spaces
    except Exception:
        return None


def sb3_vecenv_base():
    """SB3's abstract `VecEnv` when importable, else `object`."""
    try:
        from stable_baselines3.common.vec_env import VecEnv
        return VecEnv
    except Exception:
        return object


def make_spaces(obs_mode: str = "stacked", size: int = 11):
    """(observation_space, action_space) of the reference's env stack: gymnasium objects when available.
    stacked: ppo.py:118-126 (after VecTransposeImage + VecFrameStack(4,'first')); full: experts_test.py:27-30."""
    sp = gym_spaces()
    box = (lambda lo, hi, shape, dt: sp.Box(low=lo, high=hi, shape=shape, dtype=dt)) if sp else Box
    if obs_mode == "full":
        spaces = {"direction": (sp.Discrete(4) if sp else Discrete(4)), "image": box(0, 255, (size, size, 3), np.uint8),
                  "mission": box(0, 32, (32,), np.int64)}
    else:
        spaces = {"direction": box(0, 1, (4 * FRAMES,), np.uint8), "image": box(0, 255, (3 * FRAMES, 7, 7), np.uint8),
                  "mission": box(0, 32, (32 * FRAMES,), np.int64)}
    return (sp.Dict(spaces) if sp else DictSpace(spaces)), (sp.Discrete(7) if sp else Discrete(7))


# ----------------------------------------------------------------------------- spaces (duck types, no gymnasium)
class Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class Discrete:
    def __init__(self, n):
        self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def __repr__(self):
        return f"Discrete({self.n})"


class DictSpace:
    def __init__(self, spaces):
        self.spaces = dict(sorted(spaces.items()))   # gymnasium sorts plain-dict keys

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def __repr__(self):
        return f"Dict({self.spaces})"


class _Pinned:
    """numpy array over pinned host memory from mgrl_host_alloc."""

    def __init__(self, lib, shape, dtype):
        self.lib = lib
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        self.ptr = C.c_void_p()
        nat.check(lib.mgrl_host_alloc(C.byref(self.ptr), nbytes), "mgrl_host_alloc")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self.array[...] = 0

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.mgrl_host_free(self.ptr)
            self.ptr = C.c_void_p()


class _View:
    """typed window into a _Pinned slab"""

    def __init__(self, slab, offset, shape, dtype):
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        self.ptr = C.c_void_p(slab.ptr.value + offset)
        self.array = slab.array[offset:offset + count * dtype.itemsize].view(dtype).reshape(shape)

    def free(self):
        self.array = None


import types  # noqa: E402

_EMPTY_INFO = types.MappingProxyType({})      # read-only info of an environment that did not finish (large batches)
_FRESH_INFO_LIMIT = 4096                      # up to this many environments every info is its own (mutable) dict


class _B200VecEnvImpl:
    """Drop-in for the reference's stacked SB3 VecEnv (host numpy surface).  `B200VecEnv` below is this class over
    SB3's `VecEnv` when that is importable, over `object` otherwise (`bind_vecenv_base`)."""

    metadata = {"render_modes": []}
    _vecenv_base = object

    def __init__(self, cfg: EnvConfig | None = None, num_envs: int = 16, seed: int | None = None,
                 device: int = 0, env_id_base: int = 0, n_frames_stack: int = 4, layout: str = "chw",
                 obs_mode: str = "stacked", token_vocab: str = "reference", host_stack: str | None = None):
        if n_frames_stack != FRAMES:
            raise ValueError("only n_frames_stack == 4 is built (hydra_configs/algorithm/ppo.yaml:5)")
        if obs_mode not in ("stacked", "full"):
            raise ValueError("obs_mode must be 'stacked' (ppo.py:118-126) or 'full' (experts_test.py:27-30)")
        self.cfg = cfg or EnvConfig()
        self.obs_mode = obs_mode
        obs_space, act_space = make_spaces(obs_mode, self.cfg.size)
        if self._vecenv_base is not object:
            # SB3's VecEnv.__init__: num_envs, spaces, reset_infos, _seeds, _options, render_mode (via get_attr)
            self._vecenv_base.__init__(self, int(num_envs), obs_space, act_space)
        else:
            self.num_envs, self.observation_space, self.action_space = int(num_envs), obs_space, act_space
            self.reset_infos = [dict() for _ in range(int(num_envs))]
            self._seeds = [None] * int(num_envs)
            self._options = [dict() for _ in range(int(num_envs))]
            self.render_mode = None
        self._seed = 0 if seed is None else int(seed)
        self.layout, self.pitch = layout, LAYOUTS[layout][1]     # "hwc148" serves step_frames only (fast records)
        self._h = _Handle(self.cfg, self.num_envs, device, env_id_base, layout)
        lib = self._h.lib
        self._table = np.ascontiguousarray(token_table())
        # full mode: tokens in the reference wrapper's vocabulary (environment.py:74-80) or in the one Expert.decode_missions
        # decodes with (experts.py:181-182; the two differ in the reference)
        self._full_table = np.ascontiguousarray(expert_token_table() if token_vocab == "expert" else token_table())
        nat.check(lib.mgrl_set_token_table(self._h.ptr, self._table.ctypes.data_as(C.c_void_p)), "set_token_table")
        n = self.num_envs
        # Default (in place): ONE stacked observation in pinned memory, updated by every step the way SB3's VecFrameStack
        # updates `stacked_obs` (the arrays returned by step t are the arrays step t+1 rewrites; SB3's rollout buffer copies
        # them on add); the stacked terminal observations of finished environments land in a second set of arrays.
        # MGRL_WIRE=0: the device-side stack, double-buffered on the host (arrays stay valid until step t+2).
        if host_stack not in (None, "inplace", "device"):
            raise ValueError("host_stack must be 'inplace' (mgrl_vec_step_stacked_host) or 'device' (mgrl_vec_step_host)")
        self._inplace = host_stack == "inplace" if host_stack else os.environ.get("MGRL_WIRE", "1") != "0"
        self._obs_bufs = [{
            "image": _Pinned(lib, (n, 3 * FRAMES, 7, 7), np.uint8),
            "direction": _Pinned(lib, (n, 4 * FRAMES), np.uint8),
            "mission": _Pinned(lib, (n, 32 * FRAMES), np.int64),
        } for _ in range(2)]
        self._cur = 0
        # the small per-step outputs share one pinned slab laid out like the library's device slab
        # (reward | dir | mission | term | trunc | ep_len | term_dir), so step_frames gets them in one copy
        self._slab = _Pinned(lib, (10 * n,), np.uint8)
        self._p = {
            "actions": _Pinned(lib, (n,), np.uint8),
            "reward": _View(self._slab, 0, (n,), np.float32),
            "f_dir": _View(self._slab, 4 * n, (n,), np.uint8),
            "f_mission": _View(self._slab, 5 * n, (n,), np.uint8),
            "term": _View(self._slab, 6 * n, (n,), np.uint8),
            "trunc": _View(self._slab, 7 * n, (n,), np.uint8),
            "ep_len": _View(self._slab, 8 * n, (n,), np.uint8),
            "term_dir": _View(self._slab, 9 * n, (n,), np.uint8),
            "term_image": _Pinned(lib, (n, 3, 7, 7), np.uint8),
        }
        self._frames = None
        self._full = None
        self._actions = None
        self._t0 = time.time()

    # ------------------------------------------------------------------ VecEnv protocol
    def seed(self, seed=None):
        """SB3 semantics: env i is seeded with seed + i at its next reset (ppo.py:134 -> set_random_seed -> env.seed).
        Here the one seed keys the Philox streams of all environments by (seed, global env id, episode)."""
        if seed is not None:
            self._seed = int(seed)
        self._seeds = [self._seed + i for i in range(self.num_envs)]
        return list(self._seeds)

    def set_options(self, options=None):
        self._options = [dict() for _ in range(self.num_envs)]      # PlaygroundEnv.reset takes no options

    def _obs(self):
        b = self._obs_bufs[self._cur]
        return {"direction": b["direction"].array, "image": b["image"].array, "mission": b["mission"].array}

    def reset(self):
        self.reset_infos = [dict() for _ in range(self.num_envs)]
        if self.obs_mode == "full":
            img, d, m = self.reset_frames()
            return self._full_obs(d, m)
        b = self._obs_bufs[self._cur]
        nat.check(self._h.lib.mgrl_vec_reset_host(
            self._h.ptr, self._seed, b["image"].ptr, b["direction"].ptr, b["mission"].ptr, None), "vec_reset")
        return self._obs()

    def _full_obs(self, d, m):
        """experts_test.py:27-30: FullyObsWrapper image (S,S,3) with the agent cell (10, 0, dir), integer direction,
        mission tokens."""
        if self._full is None:
            S = self.cfg.size
            self._full = _Pinned(self._h.lib, (self.num_envs, S, S, 3), np.uint8)
        nat.check(self._h.lib.mgrl_full_obs_host(self._h.ptr, self._full.ptr, None), "full_obs_host")
        return {"direction": d.astype(np.int64), "image": self._full.array, "mission": self._full_table[m]}

    def step_async(self, actions):
        a = np.asarray(actions)
        if a.shape != (self.num_envs,):
            raise ValueError(f"actions must have shape ({self.num_envs},), got {a.shape}")
        if a.size and (a.min() < 0 or a.max() > 6):
            raise ValueError(f"Unknown action: {int(a.max() if a.max() > 6 else a.min())}")  # upstream step raises
        self._p["actions"].array[:] = a
        self._actions = a

    def step_wait(self):
        if self.obs_mode == "full":
            return self._step_wait_full()
        p = self._p
        if self._inplace:
            self._step_inplace()
            dones = (p["term"].array | p["trunc"].array).astype(bool)
            return self._obs(), p["reward"].array, dones, self._infos(dones, self._term_obs())
        prev = self._obs()           # terminal_observation needs the three frames before the terminal one
        self._cur ^= 1
        b = self._obs_bufs[self._cur]
        nat.check(self._h.lib.mgrl_vec_step_host(
            self._h.ptr, p["actions"].ptr, b["image"].ptr, b["direction"].ptr, b["mission"].ptr,
            p["reward"].ptr, p["term"].ptr, p["trunc"].ptr, p["ep_len"].ptr, p["term_image"].ptr,
            p["term_dir"].ptr, None), "vec_step")
        term, trunc = p["term"].array, p["trunc"].array
        dones = (term | trunc).astype(bool)
        return self._obs(), p["reward"].array, dones, self._infos(dones, prev)

    def stacked_terminal_obs(self):
        """in-place mode: the observation dict whose rows of finished environments hold info['terminal_observation']"""
        return self._term_obs()

    def _term_obs(self):
        b = self._obs_bufs[1]        # in-place mode: the second set holds the stacked terminal observations
        return {"direction": b["direction"].array, "image": b["image"].array, "mission": b["mission"].array}

    def _step_inplace(self):
        """mgrl_vec_step_stacked_host: 64-byte records over PCIe, the observation dict updated in place by the library's
        host threads, stacked terminal observations of finished environments into the second buffer set."""
        p, b, t = self._p, self._obs_bufs[0], self._obs_bufs[1]
        nat.check(self._h.lib.mgrl_vec_step_stacked_host(
            self._h.ptr, p["actions"].ptr, b["image"].ptr, b["direction"].ptr, b["mission"].ptr,
            p["reward"].ptr, p["term"].ptr, p["trunc"].ptr, p["ep_len"].ptr, t["image"].ptr, t["direction"].ptr,
            t["mission"].ptr, None), "vec_step_stacked")

    def _infos(self, dones, prev):
        """SB3 infos: `terminal_observation` (the stacked observation the finished episode ended on), `TimeLimit.truncated`
        and Monitor's `episode` for finished environments; gathered for all of them at once, one dict each."""
        p, n = self._p, self.num_envs
        infos = [dict() for _ in range(n)] if n <= _FRESH_INFO_LIMIT else [_EMPTY_INFO] * n
        idx = np.flatnonzero(dones)
        if idx.size == 0:
            return infos
        t = round(time.time() - self._t0, 6)
        trunc_only = (p["trunc"].array[idx] != 0) & (p["term"].array[idx] == 0)
        rew, length = p["reward"].array[idx].tolist(), p["ep_len"].array[idx].tolist()
        trunc_l = trunc_only.tolist()
        if prev is None:             # full mode: no terminal observation (the full grid of the finished episode is gone)
            for i, tr, r, l in zip(idx.tolist(), trunc_l, rew, length):
                infos[i] = {"TimeLimit.truncated": tr, "episode": {"r": r, "l": l, "t": t}}
            return infos
        if self._inplace:            # the library wrote the stacked terminal observations (rows of finished environments)
            t_dir, t_img, t_mis = prev["direction"][idx], prev["image"][idx], prev["mission"][idx]
        else:
            tdir = np.zeros((idx.size, 4), np.uint8)
            tdir[np.arange(idx.size), p["term_dir"].array[idx]] = 1
            t_dir = np.concatenate([prev["direction"][idx, 4:], tdir], axis=1)
            t_img = np.concatenate([prev["image"][idx, 3:], p["term_image"].array[idx]], axis=1)
            t_mis = np.concatenate([prev["mission"][idx, 32:], prev["mission"][idx, 96:]], axis=1)
        # (iterating the arrays yields the row views a third faster than indexing them one by one: this loop is the SB3
        #  protocol's per-environment cost, 9 000 dicts per vector step of 65 536 uniform-random environments)
        for i, d, im, mi, tr, r, l in zip(idx.tolist(), t_dir, t_img, t_mis, trunc_l, rew, length):
            infos[i] = {"terminal_observation": {"direction": d, "image": im, "mission": mi},
                        "TimeLimit.truncated": tr, "episode": {"r": r, "l": l, "t": t}}
        return infos

    def _step_wait_full(self):
        out = self.step_frames(self._p["actions"].array)
        _, d, m, rew, term, trunc, _ = out
        dones = (term | trunc).astype(bool)
        return self._full_obs(d, m), rew, dones, self._infos(dones, None)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def step_arrays(self, actions):
        """`step` without the per-env Python info dicts: returns (obs, rewards, term, trunc, ep_len,
        term_image, term_dir) as pinned numpy arrays straight from the library; term_image [N,3,7,7] is the last frame of
        the terminal observation (rows of finished environments).  In place (default): mgrl_vec_step_stacked_host, `obs` is
        the persistent stack, term_dir is that frame's one-hot direction [N,4] and `stacked_terminal_obs()` gives the whole
        stacked terminal observations; MGRL_WIRE=0: mgrl_vec_step_host, term_dir [N] is the direction index."""
        p = self._p
        p["actions"].array[:] = actions
        if self._inplace:
            self._step_inplace()
            return (self._obs(), p["reward"].array, p["term"].array, p["trunc"].array, p["ep_len"].array,
                    self._obs_bufs[1]["image"].array[:, 9:], self._obs_bufs[1]["direction"].array[:, 12:])
        self._cur ^= 1
        b = self._obs_bufs[self._cur]
        nat.check(self._h.lib.mgrl_vec_step_host(
            self._h.ptr, p["actions"].ptr, b["image"].ptr, b["direction"].ptr, b["mission"].ptr,
            p["reward"].ptr, p["term"].ptr, p["trunc"].ptr, p["ep_len"].ptr, p["term_image"].ptr,
            p["term_dir"].ptr, None), "vec_step")
        return (self._obs(), p["reward"].array, p["term"].array, p["trunc"].array, p["ep_len"].array,
                p["term_image"].array, p["term_dir"].array)

    # ------------------------------------------------------------------ un-stacked host path
    def _frame_bufs(self):
        if self._frames is None:
            lib, n = self._h.lib, self.num_envs
            self._frames = {"image": _Pinned(lib, (n, self.pitch), np.uint8), "dir": self._p["f_dir"],
                            "mission": self._p["f_mission"], "term_image": _Pinned(lib, (n, self.pitch), np.uint8)}
        return self._frames

    def reset_frames(self):
        """reset() without the SB3 wrapper stack: (image [N,pitch] u8 in this env's layout, dir [N], mission id [N])."""
        f = self._frame_bufs()
        nat.check(self._h.lib.mgrl_vec_reset_frames_host(self._h.ptr, self._seed, f["image"].ptr, f["dir"].ptr,
                                                         f["mission"].ptr, None), "vec_reset_frames")
        return f["image"].array, f["dir"].array, f["mission"].array

    def step_frames(self, actions, want_terminal: bool = False):
        """step() without the SB3 wrapper stack (same outputs as the CPU oracle's vector step): pinned numpy arrays
        (image, dir, mission id, reward, term, trunc, ep_len[, term_image, term_dir]) via mgrl_vec_step_frames_host."""
        p, f = self._p, self._frame_bufs()
        p["actions"].array[:] = actions
        nat.check(self._h.lib.mgrl_vec_step_frames_host(
            self._h.ptr, p["actions"].ptr, f["image"].ptr, f["dir"].ptr, f["mission"].ptr, p["reward"].ptr, p["term"].ptr,
            p["trunc"].ptr, p["ep_len"].ptr, f["term_image"].ptr if want_terminal else None,
            p["term_dir"].ptr if want_terminal else None, None), "vec_step_frames")
        out = (f["image"].array, f["dir"].array, f["mission"].array, p["reward"].array, p["term"].array, p["trunc"].array,
               p["ep_len"].array)
        return out + ((f["term_image"].array, p["term_dir"].array) if want_terminal else ())

    def close(self):
        if self._full is not None:
            self._full.free()
        if self._frames is not None:
            for v in self._frames.values():
                v.free()
        for v in self._p.values():
            v.free()
        self._slab.free()
        for b in self._obs_bufs:
            for v in b.values():
                v.free()
        self._h.close()

    def render(self, mode=None):
        """ppo.py:265 calls vec_env.render() when cfg.render is set.  There is no renderer on the device path (rendering is
        out of scope); like an SB3 VecEnv whose render_mode is None this returns None."""
        return None

    def get_images(self):
        return [None for _ in range(self.num_envs)]

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        return [int(i) for i in np.atleast_1d(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def get_attr(self, name, indices=None):
        idx = self._indices(indices)
        if name == "render_mode":        # (SB3's VecEnv.__init__ asks before the device handle exists)
            return [None for _ in idx]
        if name == "llm_description":    # environment.py:190
            return [None for _ in idx]
        if name in ("size", "width", "height", "max_steps"):
            v = self.cfg.max_steps if name == "max_steps" else self.cfg.size
            return [v for _ in idx]
        if name == "mission":
            st = self.get_state()
            return [MISSIONS[int(st["mission_id"][i])] for i in idx]
        if name in ("agent_pos", "agent_dir", "step_count", "mission_done"):
            st = self.get_state()
            if name == "agent_pos":
                return [(int(st["agent_x"][i]), int(st["agent_y"][i])) for i in idx]
            return [int(st[name][i]) for i in idx]
        raise AttributeError(name)

    def set_attr(self, name, value, indices=None):
        raise AttributeError(f"attribute {name} cannot be set on device-resident environments")

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        """The per-environment methods SB3 and the reference reach through the vec-env: `render`/`close`/`seed` are
        tolerated as no-ops (one simulator owns all environments); anything else has no per-environment object."""
        idx = self._indices(indices)
        if method_name in ("render", "close", "seed", "get_wrapper_attr"):
            return [None for _ in idx]
        raise AttributeError(f"env_method({method_name}) is not available on device-resident environments")

    # ------------------------------------------------------------------ extras
    def get_state(self) -> np.ndarray:
        """Canonical [N] structured state dump (copied to the host)."""
        out = np.zeros(self.num_envs, STATE_DTYPE)
        nat.check(self._h.lib.mgrl_get_state_host(self._h.ptr, out.ctypes.data_as(C.c_void_p), out.nbytes, None),
                  "get_state_host")
        return out

    def set_state(self, state: np.ndarray, seed: int | None = None) -> None:
        """Restore a canonical state dump (the frame-stack history is left as is)."""
        st = np.ascontiguousarray(state)
        assert st.nbytes == self.num_envs * STATE_BYTES
        if seed is not None:
            self._seed = int(seed)
        nat.check(self._h.lib.mgrl_set_state_host(self._h.ptr, st.ctypes.data_as(C.c_void_p), st.nbytes,
                                                  self._seed, None), "set_state_host")

    def mission_strings(self):
        return self.get_attr("mission")


def bind_vecenv_base(base=None):
    """The drop-in class over `base` (SB3's `VecEnv` when importable): what `B200VecEnv` is.  Exposed so that the
    binding can be exercised against a stand-in base class in tests."""
    base = sb3_vecenv_base() if base is None else base
    if base is object:
        return type("B200VecEnv", (_B200VecEnvImpl,), {"__doc__": _B200VecEnvImpl.__doc__})
    return type("B200VecEnv", (_B200VecEnvImpl, base), {"_vecenv_base": base, "__doc__": _B200VecEnvImpl.__doc__})


B200VecEnv = bind_vecenv_base()


# ================================================================================= device path
class DeviceEnv:
    """Device-resident environments: torch CUDA tensors in/out, no host copies.

    Observations are un-stacked frames ([N,147] u8 in CHW or HWC order); the frame stack of
    the policy input is gathered on the fly by the rollout engine."""

    def __init__(self, cfg: EnvConfig | None = None, num_envs: int = 65536, seed: int = 0,
                 device: int | None = None, env_id_base: int = 0, chw: bool = True, layout: str | None = None):
        import torch
        if not torch.cuda.is_available():
            raise nat.NativeError("DeviceEnv needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.cfg = cfg or EnvConfig()
        self.num_envs = int(num_envs)
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.seed = int(seed)
        self.layout = layout or ("chw" if chw else "hwc")
        self.pitch = LAYOUTS[self.layout][1]
        self._h = _Handle(self.cfg, self.num_envs, self.device_index, env_id_base, self.layout)
        n = self.num_envs
        u8 = dict(dtype=torch.uint8, device=self.device)
        self.image = torch.zeros((n, self.pitch), **u8)
        self.dir = torch.zeros(n, **u8)
        self.mission = torch.zeros(n, **u8)
        self.reward = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.term = torch.zeros(n, **u8)
        self.trunc = torch.zeros(n, **u8)
        self.ep_len = torch.zeros(n, **u8)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def reset(self, seed: int | None = None):
        if seed is not None:
            self.seed = int(seed)
        nat.check(self._h.lib.mgrl_reset(self._h.ptr, self.seed, self._ptr(self.image), self._ptr(self.dir),
                                         self._ptr(self.mission), self._stream()), "mgrl_reset")
        return self.image, self.dir, self.mission

    def step(self, actions, image=None, dir=None, mission=None, reward=None, term=None, trunc=None,
             ep_len=None, term_image=None, term_dir=None):
        """One vector step.  Output tensors default to this object's buffers."""
        image = self.image if image is None else image
        dir = self.dir if dir is None else dir
        mission = self.mission if mission is None else mission
        reward = self.reward if reward is None else reward
        term = self.term if term is None else term
        trunc = self.trunc if trunc is None else trunc
        ep_len = self.ep_len if ep_len is None else ep_len
        assert actions.dtype == self.torch.uint8 and actions.is_cuda and actions.numel() == self.num_envs
        nat.check(self._h.lib.mgrl_step(
            self._h.ptr, self._ptr(actions), self._ptr(image), self._ptr(dir), self._ptr(mission),
            self._ptr(reward), self._ptr(term), self._ptr(trunc), self._ptr(ep_len), self._ptr(term_image),
            self._ptr(term_dir), self._stream()), "mgrl_step")
        return image, reward, term, trunc

    def step_many(self, actions, image=None, dir=None, mission=None, reward=None, term=None, trunc=None,
                  ep_len=None):
        """T steps in one launch; actions [T,N] u8; outputs [T,N,...] (image/dir/mission/ep_len optional)."""
        T = int(actions.shape[0])
        assert actions.dtype == self.torch.uint8 and actions.is_cuda and actions.shape[1] == self.num_envs
        assert actions.is_contiguous()
        nat.check(self._h.lib.mgrl_step_many(
            self._h.ptr, T, self._ptr(actions), self._ptr(image), self._ptr(dir), self._ptr(mission),
            self._ptr(reward), self._ptr(term), self._ptr(trunc), self._ptr(ep_len), self._stream()),
            "mgrl_step_many")

    def observe(self):
        nat.check(self._h.lib.mgrl_observe(self._h.ptr, self._ptr(self.image), self._ptr(self.dir),
                                           self._ptr(self.mission), self._stream()), "mgrl_observe")
        return self.image, self.dir, self.mission

    def full_obs(self):
        S = self.cfg.size
        out = self.torch.zeros((self.num_envs, S, S, 3), dtype=self.torch.uint8, device=self.device)
        nat.check(self._h.lib.mgrl_full_obs(self._h.ptr, self._ptr(out), self._stream()), "mgrl_full_obs")
        return out

    def get_state(self):
        out = self.torch.empty((self.num_envs, STATE_BYTES), dtype=self.torch.uint8, device=self.device)
        nat.check(self._h.lib.mgrl_get_state(self._h.ptr, self._ptr(out), out.numel(), self._stream()), "get_state")
        return out

    def set_state(self, state, seed: int | None = None):
        if seed is not None:
            self.seed = int(seed)
        assert state.is_cuda and state.dtype == self.torch.uint8 and state.numel() == self.num_envs * STATE_BYTES
        state = state.contiguous()
        nat.check(self._h.lib.mgrl_set_state(self._h.ptr, self._ptr(state), state.numel(), self.seed,
                                             self._stream()), "set_state")

    def get_state_numpy(self) -> np.ndarray:
        return self.get_state().cpu().numpy().view(STATE_DTYPE).reshape(-1)

    def set_state_numpy(self, state: np.ndarray, seed: int | None = None):
        t = self.torch.from_numpy(np.ascontiguousarray(state).view(np.uint8).reshape(-1, STATE_BYTES).copy())
        self.set_state(t.to(self.device), seed)

    def error_flags(self) -> int:
        flags = C.c_int(0)
        nat.check(self._h.lib.mgrl_error_flags(self._h.ptr, C.byref(flags), self._stream()), "error_flags")
        return flags.value

    def close(self):
        self._h.close()


def gae(rewards, values, episode_starts, last_values, last_dones, gamma: float, gae_lambda: float,
        advantages=None, returns=None):
    """SB3 RolloutBuffer.compute_returns_and_advantage on device tensors ([T,N] f32 / u8)."""
    import torch
    T, N = rewards.shape
    for t in (rewards, values, episode_starts, last_values, last_dones):
        assert t.is_cuda and t.is_contiguous()
    assert rewards.dtype == torch.float32 and values.dtype == torch.float32
    assert episode_starts.dtype == torch.uint8 and last_dones.dtype == torch.uint8
    advantages = torch.empty_like(rewards) if advantages is None else advantages
    returns = torch.empty_like(rewards) if returns is None else returns
    s = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    nat.check(nat.lib().mgrl_gae(p(rewards), p(values), p(episode_starts), p(last_values), p(last_dones),
                                 float(gamma), float(gae_lambda), T, N, p(advantages), p(returns), s), "mgrl_gae")
    return advantages, returns
