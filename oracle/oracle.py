"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, NOT THE PRODUCT).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module (see mg_oracle.h).  It builds oracle/_build/libmg_oracle.so
on first use with gcc.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmg_oracle.so")

OBS_BYTES = 147
GRID_CELLS = 121
NONE = 0xFF

# numpy mirror of mg_state (140 bytes)
STATE_DTYPE = np.dtype([
    ("grid", np.uint8, (GRID_CELLS,)),
    ("agent_x", np.uint8), ("agent_y", np.uint8), ("agent_dir", np.uint8),
    ("carrying", np.uint8), ("step_count", np.uint8),
    ("target_x", np.uint8), ("target_y", np.uint8), ("target_action", np.uint8),
    ("mission_id", np.uint8), ("mission_done", np.uint8), ("latch_step", np.uint8),
    ("episode", np.uint32), ("reset_draws", np.uint16), ("error", np.uint8), ("pad", np.uint8),
], align=False)
assert STATE_DTYPE.itemsize == 140

PROBLEMS = {"multi": 0, "gto": 1, "gtg": 2, "opn": 3, "pkp": 4, "drp": 5, "mov": 6, "full": 7}


class Config(C.Structure):
    _fields_ = [("size", C.c_int32), ("num_objects", C.c_int32), ("problem", C.c_int32),
                ("mission", C.c_int32), ("all_doors_open", C.c_int32),
                ("see_through_walls", C.c_int32), ("max_steps", C.c_int32),
                ("num_obstacles", C.c_int32)]


def make_config(problem="multi", mission=5, size=11, num_objects=4, all_doors_open=False,
                see_through_walls=True, obstacles=False, percent_obstacles=0.05) -> Config:
    from math import floor
    n_obst = floor((size - 2) ** 2 * percent_obstacles) if obstacles else 0
    return Config(size, num_objects, PROBLEMS[problem], -1 if mission is None else int(mission),
                  int(bool(all_doors_open)), int(bool(see_through_walls)), size * size, n_obst)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mg_oracle.c")
    hdr = os.path.join(_HERE, "mg_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or (os.path.exists(src) and os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        u8p, f32p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        vp = C.c_void_p
        _lib.mg_philox4x32_10.argtypes = [u32p, u32p, u32p]
        _lib.mg_draw_below.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.mg_draw_below.restype = C.c_uint32
        _lib.mg_reward_lut.argtypes = [C.c_int, f32p]
        _lib.mg_kind_encode.argtypes = [C.c_uint8, u8p]
        _lib.mg_generate.argtypes = [C.POINTER(Config), C.c_uint64, C.c_uint64, vp]
        _lib.mg_generate.restype = C.c_int
        _lib.mg_reset_env.argtypes = [C.POINTER(Config), C.c_uint64, C.c_uint64, vp]
        _lib.mg_reset_env.restype = C.c_int
        _lib.mg_step_env.argtypes = [C.POINTER(Config), f32p, vp, C.c_int, f32p, u8p, u8p, u8p]
        _lib.mg_gen_obs.argtypes = [C.POINTER(Config), vp, C.c_uint8, vp]
        _lib.mg_full_obs.argtypes = [C.POINTER(Config), vp, vp]
        _lib.mg_vec_reset.argtypes = [C.POINTER(Config), C.c_uint64, C.c_uint64, C.c_int, vp, vp, vp, vp, C.c_int]
        _lib.mg_vec_step.argtypes = [C.POINTER(Config), C.c_uint64, C.c_uint64, C.c_int, vp, vp, vp, vp, vp,
                                     vp, vp, vp, vp, vp, C.c_int]
        _lib.mg_gae.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().mg_philox4x32_10(c, k, o)
    return list(o)


def reward_lut(max_steps: int) -> np.ndarray:
    lut = np.zeros(max_steps + 1, np.float32)
    lib().mg_reward_lut(max_steps, lut.ctypes.data_as(C.POINTER(C.c_float)))
    return lut


def kind_encode(kind: int):
    o = (C.c_uint8 * 3)()
    lib().mg_kind_encode(kind, o)
    return tuple(o)


class OracleVecEnv:
    """N oracle environments with DummyVecEnv-style auto-reset (mg_vec_reset / mg_vec_step)."""

    def __init__(self, cfg: Config, n: int, seed: int = 0, env_id_base: int = 0, nthreads: int = 1):
        self.cfg, self.n, self.seed, self.base, self.nthreads = cfg, n, seed, env_id_base, nthreads
        self.states = np.zeros(n, STATE_DTYPE)
        self.obs = np.zeros((n, 7, 7, 3), np.uint8)
        self.dir = np.zeros(n, np.uint8)
        self.mission = np.zeros(n, np.uint8)
        self.reward = np.zeros(n, np.float32)
        self.term = np.zeros(n, np.uint8)
        self.trunc = np.zeros(n, np.uint8)
        self.ep_len = np.zeros(n, np.uint8)
        self.term_obs = np.zeros((n, 7, 7, 3), np.uint8)

    def reset(self):
        lib().mg_vec_reset(C.byref(self.cfg), self.seed, self.base, self.n, _p(self.states),
                           _p(self.obs), _p(self.dir), _p(self.mission), self.nthreads)
        return self.obs

    def step(self, actions: np.ndarray, want_term_obs: bool = True):
        a = np.ascontiguousarray(actions, np.uint8)
        assert a.shape == (self.n,)
        lib().mg_vec_step(C.byref(self.cfg), self.seed, self.base, self.n, _p(self.states), _p(a),
                          _p(self.obs), _p(self.dir), _p(self.mission), _p(self.reward), _p(self.term),
                          _p(self.trunc), _p(self.ep_len), _p(self.term_obs) if want_term_obs else None,
                          self.nthreads)
        return self.obs, self.reward, self.term, self.trunc

    def full_obs(self) -> np.ndarray:
        S = self.cfg.size
        out = np.zeros((self.n, S, S, 3), np.uint8)
        for i in range(self.n):
            lib().mg_full_obs(C.byref(self.cfg), _p(self.states[i:i + 1]), _p(out[i]))
        return out


def step_one(cfg: Config, lut: np.ndarray, state: np.ndarray, action: int):
    """Single-env PlaygroundEnv.step on a length-1 state array; returns (reward, term, trunc, carry_obs)."""
    r = C.c_float()
    te, tr, co = C.c_uint8(), C.c_uint8(), C.c_uint8()
    lib().mg_step_env(C.byref(cfg), lut.ctypes.data_as(C.POINTER(C.c_float)), _p(state), int(action),
                      C.byref(r), C.byref(te), C.byref(tr), C.byref(co))
    return np.float32(r.value), te.value, tr.value, co.value


def gen_obs(cfg: Config, state: np.ndarray, carrying=None) -> np.ndarray:
    out = np.zeros((7, 7, 3), np.uint8)
    c = int(state["carrying"][0]) if carrying is None else int(carrying)
    lib().mg_gen_obs(C.byref(cfg), _p(state), c, _p(out))
    return out


def generate(cfg: Config, seed: int, env_id: int, state: np.ndarray) -> int:
    return lib().mg_generate(C.byref(cfg), seed, env_id, _p(state))


def gae(rewards, values, episode_starts, last_values, last_dones, gamma, gae_lambda):
    rewards = np.ascontiguousarray(rewards, np.float32)
    values = np.ascontiguousarray(values, np.float32)
    episode_starts = np.ascontiguousarray(episode_starts, np.uint8)
    last_values = np.ascontiguousarray(last_values, np.float32)
    last_dones = np.ascontiguousarray(last_dones, np.uint8)
    T, N = rewards.shape
    adv = np.zeros((T, N), np.float32)
    ret = np.zeros((T, N), np.float32)
    lib().mg_gae(_p(rewards), _p(values), _p(episode_starts), _p(last_values), _p(last_dones),
                 float(gamma), float(gae_lambda), T, N, _p(adv), _p(ret))
    return adv, ret
