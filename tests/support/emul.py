"""Host build of the product's per-environment device functions (TEST INFRASTRUCTURE).

Compiles minigrid-rl_b200/csrc/mgrl_core.cuh with g++ (tests/support/host_emul.cpp) so the
CPU test tier can check the kernel *logic* against the oracle.  Not a product path."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "tests", "support", "host_emul.cpp")
CORE = os.path.join(ROOT, "minigrid-rl_b200", "csrc", "mgrl_core.cuh")
OUT = os.path.join(ROOT, "tests", "_build", "libhost_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if (not os.path.exists(OUT)
                or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(CORE))):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                                   "-Wno-unknown-pragmas", "-I", os.path.dirname(CORE), "-o", OUT, SRC])
        _lib = C.CDLL(OUT)
        _lib.emul_kind_encode.restype = C.c_uint32
        vp = C.c_void_p
        _lib.emul_vec_step.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int] + [vp] * 11
        _lib.emul_generate.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
        _lib.emul_obs.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        _lib.emul_full_obs.argtypes = [vp, vp, vp]
    return _lib


def p(a):
    return a.ctypes.data_as(C.c_void_p)


class EmulVecEnv:
    def __init__(self, cfg, n, seed=0, env_id_base=0):
        from oracle import oracle as orc
        self.cfg, self.n, self.seed, self.base = cfg, n, seed, env_id_base
        self.lut = orc.reward_lut(cfg.max_steps)
        self.states = np.zeros(n, orc.STATE_DTYPE)
        self.obs = np.zeros((n, 7, 7, 3), np.uint8)
        self.term_obs = np.zeros((n, 7, 7, 3), np.uint8)
        self.dir = np.zeros(n, np.uint8)
        self.mission = np.zeros(n, np.uint8)
        self.reward = np.zeros(n, np.float32)
        self.term = np.zeros(n, np.uint8)
        self.trunc = np.zeros(n, np.uint8)
        self.ep_len = np.zeros(n, np.uint8)

    def reset(self):
        self.states[:] = 0
        for i in range(self.n):
            lib().emul_generate(C.byref(self.cfg), self.seed, self.base + i, p(self.states[i:i + 1]))
            lib().emul_obs(C.byref(self.cfg), p(self.states[i:i + 1]), 0, 0, p(self.obs[i]))
        self.dir[:] = self.states["agent_dir"]
        self.mission[:] = self.states["mission_id"]
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.uint8)
        lib().emul_vec_step(C.byref(self.cfg), self.seed, self.base, self.n, p(self.lut), p(self.states), p(a),
                            p(self.obs), p(self.dir), p(self.mission), p(self.reward), p(self.term),
                            p(self.trunc), p(self.ep_len), p(self.term_obs))
        return self.obs, self.reward, self.term, self.trunc
