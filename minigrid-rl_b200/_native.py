"""ctypes binding of the C-ABI CUDA library (include/mgrl.h).  Fails loudly: there is no
Python/CPU fallback for any entry point."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_lib", "libmgrl.so")
ABI_VERSION = 1


class NativeError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("size", C.c_int32), ("num_objects", C.c_int32), ("problem", C.c_int32),
                ("mission", C.c_int32), ("all_doors_open", C.c_int32), ("see_through_walls", C.c_int32),
                ("max_steps", C.c_int32), ("num_obstacles", C.c_int32), ("num_envs", C.c_int32),
                ("obs_layout", C.c_int32), ("env_id_base", C.c_uint64)]


class RolloutView(C.Structure):          # mgrl_rollout_view
    _fields_ = [(n, C.c_void_p) for n in ("frames", "dirs", "mission", "age", "actions", "values", "logp", "adv", "ret")] + [
        ("num_envs", C.c_int)]


class PPOHyper(C.Structure):             # mgrl_ppo_hyper
    _fields_ = [("clip_range", C.c_float), ("clip_range_vf", C.c_float), ("ent_coef", C.c_float), ("vf_coef", C.c_float),
                ("normalize_advantage", C.c_int), ("strict_fp32", C.c_int), ("use_tcgen05", C.c_int)]


def library_path() -> str:
    return _LIB


def build_library(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["sh", os.path.join(_HERE, "csrc", "build.sh")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise NativeError(f"nvcc build failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stdout.strip())
    return _LIB


_lib = None
vp = C.c_void_p

_SIGNATURES = {
    "mgrl_abi_version": (C.c_int, []),
    "mgrl_last_error": (C.c_char_p, []),
    "mgrl_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(vp)]),
    "mgrl_destroy": (C.c_int, [vp]),
    "mgrl_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
    "mgrl_host_free": (C.c_int, [vp]),
    "mgrl_reset": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp]),
    "mgrl_step": (C.c_int, [vp] * 12),
    "mgrl_step_many": (C.c_int, [vp, C.c_int] + [vp] * 9),
    "mgrl_get_state": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "mgrl_set_state": (C.c_int, [vp, vp, C.c_size_t, C.c_uint64, vp]),
    "mgrl_get_state_host": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "mgrl_set_state_host": (C.c_int, [vp, vp, C.c_size_t, C.c_uint64, vp]),
    "mgrl_state_ptr": (C.c_int, [vp, C.POINTER(vp)]),
    "mgrl_observe": (C.c_int, [vp, vp, vp, vp, vp]),
    "mgrl_full_obs": (C.c_int, [vp, vp, vp]),
    "mgrl_full_obs_host": (C.c_int, [vp, vp, vp]),
    "mgrl_error_flags": (C.c_int, [vp, C.POINTER(C.c_int), vp]),
    "mgrl_stack_push": (C.c_int, [C.c_int] + [vp] * 9),
    "mgrl_gae": (C.c_int, [vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp, vp]),
    "mgrl_vec_reset_host": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp]),
    "mgrl_vec_step_host": (C.c_int, [vp] * 12),
    "mgrl_vec_step_stacked_host": (C.c_int, [vp] * 13),
    "mgrl_debug_tc5_shift_probe": (C.c_int, [vp, vp]),
    "mgrl_set_token_table": (C.c_int, [vp, vp]),
    "mgrl_vec_reset_frames_host": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp]),
    "mgrl_vec_step_frames_host": (C.c_int, [vp] * 12),
    "mgrl_policy_forward": (C.c_int, [vp] * 12 + [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, vp]),
    "mgrl_policy_pack_fragments": (C.c_int, [vp, vp]),
    "mgrl_policy_last_error": (C.c_char_p, []),
    "mgrl_conv1_pool_forward": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp]),
    "mgrl_conv1_pool_backward": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp]),
    "mgrl_lut_grad": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
    "mgrl_colsum": (C.c_int, [vp, C.c_longlong, C.c_int, vp, vp]),
    "mgrl_patch2x2_forward": (C.c_int, [vp, C.c_int, vp, vp]),
    "mgrl_patch2x2_backward": (C.c_int, [vp, C.c_int, vp, vp]),
    "mgrl_ppo_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "mgrl_ppo_destroy": (C.c_int, [vp]),
    "mgrl_ppo_bind": (C.c_int, [vp] * 6),
    "mgrl_ppo_mission_table": (C.c_int, [vp, vp, vp]),
    "mgrl_ppo_moments": (C.c_int, [vp, vp, C.c_int, C.c_longlong, vp, vp]),
    "mgrl_ppo_gradients": (C.c_int, [vp, C.POINTER(RolloutView), vp, C.c_int, vp, C.POINTER(PPOHyper), vp, vp, vp, vp]),
    "mgrl_ppo_apply": (C.c_int, [vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, vp, vp]),
    "mgrl_ppo_debug_buffer": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "mgrl_ppo_debug_copy": (C.c_int, [vp, C.c_char_p, vp, C.c_longlong, vp]),
}
EXPORTS = sorted(_SIGNATURES)


def lib():
    """The loaded library; raises NativeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise NativeError(
                f"{_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the simulator has no CPU fallback)")
        handle = C.CDLL(_LIB)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        if handle.mgrl_abi_version() != ABI_VERSION:
            raise NativeError("libmgrl.so ABI version mismatch: rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().mgrl_last_error().decode(errors="replace")
        raise NativeError(f"{what or 'mgrl call'} failed ({rc}): {msg}")
