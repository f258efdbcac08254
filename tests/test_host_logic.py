"""CPU tier: host-side logic and the C-ABI library surface (no compute without a GPU)."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

import minigrid_rl_b200 as mg
from minigrid_rl_b200 import _native as nat
from oracle import sb3_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mgrl.h")).read()
    declared = sorted(set(re.findall(r"\b(mgrl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    handle = C.CDLL(mg.library_path())
    for name in declared:
        assert hasattr(handle, name), f"libmgrl.so does not export {name}"
    assert sorted(nat.EXPORTS) == declared, "python binding table out of sync with include/mgrl.h"
    assert mg.lib().mgrl_abi_version() == 1


def test_create_fails_loudly_without_gpu_or_with_bad_config():
    import torch
    lib = mg.lib()
    ptr = C.c_void_p()
    bad = nat.Config(11, 4, 0, 3, 0, 1, 121, 0, 16, 1, 0)      # mission 3 ('drop' on multi) is invalid
    rc = lib.mgrl_create(C.byref(bad), 0, C.byref(ptr))
    assert rc < 0 and lib.mgrl_last_error()
    if not torch.cuda.is_available():
        ok = nat.Config(11, 4, 0, 5, 0, 1, 121, 0, 16, 1, 0)
        assert lib.mgrl_create(C.byref(ok), 0, C.byref(ptr)) == -3   # MGRL_ERR_NO_DEVICE
        with pytest.raises(mg.NativeError):
            mg.B200VecEnv(num_envs=4)


def test_mission_table_matches_reference_tokens():
    # tokens recorded from the reference's TokenizeVocabWrapper (oracle/gen_golden.py)
    table = mg.token_table()
    seen = set()
    for path in sorted(glob.glob(os.path.join(GOLDEN, "trace_*.npz"))):
        z = np.load(path)
        ids = z["state"]["mission_id"]
        assert np.array_equal(table[ids].astype(np.int8), z["tokens"])
        assert np.array_equal(table[z["init_state"]["mission_id"]].astype(np.int8), z["init_tokens"])
        seen |= set(np.unique(ids).tolist())
        for s in z["missions"]:
            assert mg.mission_string(mg.mission_id(str(s))) == str(s)
    assert len(seen) >= 56      # every mission the five tasks can produce shows up in the fixtures
    # SURVEY App. C-12
    assert mg.tokenize("go to goal")[:11].tolist() == [12, 20, 0, 25, 20, 0, 12, 20, 6, 17, 0]
    assert np.array_equal(sb3_oracle.tokenize("pick up purple ball"), mg.tokenize("pick up purple ball"))


def test_config_mirrors_reference_keys():
    cfg = mg.EnvConfig.from_cfg(dict(problem="multi", mission=None, all_doors_open=False, size=11,
                                     num_objects=4, see_through_walls=True, obstacles=True,
                                     percent_obstacles=0.05))
    assert cfg.max_steps == 121 and cfg.num_obstacles == 4
    assert mg.EnvConfig.for_task("TGL").mission == 1 and mg.EnvConfig.for_task("ALL").mission is None
    mg.EnvConfig(problem="full").validate()                      # custom_env.py:134-152: every problem of the reference
    mg.EnvConfig(problem="mov").validate()
    with pytest.raises(ValueError):                              # ... and its error for anything else (:151-152)
        mg.EnvConfig(problem="maze").validate()
    assert [mg.MISSIONS[i] for i in (24, 25, 26, 27, 72, 73)] == ["move left", "move right", "move up", "move down", "go to goal", "drop"]


def test_frame_stack_oracle_known_answers():
    # SURVEY App. C-14
    fs = sb3_oracle.FrameStack(2, (1,), np.int64)
    assert fs.reset(np.array([[1], [10]])).tolist() == [[0, 0, 0, 1], [0, 0, 0, 10]]
    s, _ = fs.update(np.array([[2], [11]]), np.array([False, False]), None)
    assert s.tolist() == [[0, 0, 1, 2], [0, 0, 10, 11]]
    s, term = fs.update(np.array([[3], [50]]), np.array([False, True]), np.array([[0], [12]]))
    assert s.tolist() == [[0, 1, 2, 3], [0, 0, 0, 50]]
    assert term[1].tolist() == [0, 10, 11, 12]


def test_spaces_are_gymnasium_objects_when_gymnasium_is_importable():
    """ppo.py:134: SB3 isinstance-checks the spaces against gymnasium.spaces; without gymnasium the duck types stand in."""
    import os
    import sys
    from minigrid_rl_b200 import vec_env as ve
    obs, act = ve.make_spaces()
    if ve.gym_spaces() is None:
        assert isinstance(obs, ve.DictSpace) and isinstance(act, ve.Discrete)
    assert list(obs.keys()) == ["direction", "image", "mission"] and act.n == 7
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "upstream_shim")
    sys.path.insert(0, shim)
    try:
        import gymnasium
        obs, act = ve.make_spaces()
        assert isinstance(obs, gymnasium.spaces.Dict) and isinstance(obs["image"], gymnasium.spaces.Box)
        assert isinstance(act, gymnasium.spaces.Discrete) and act.n == 7
        assert tuple(obs["image"].shape) == (12, 7, 7) and obs["mission"].dtype == np.int64 and tuple(obs["direction"].shape) == (16,)
        full, _ = ve.make_spaces("full", 9)
        assert tuple(full["image"].shape) == (9, 9, 3) and isinstance(full["direction"], gymnasium.spaces.Discrete)
    finally:
        sys.path.remove(shim)
        for k in [k for k in sys.modules if k == "gymnasium" or k.startswith("gymnasium.")]:
            if getattr(sys.modules[k], "__file__", "") and shim in (sys.modules[k].__file__ or ""):
                del sys.modules[k]


def test_vecenv_class_binds_over_the_sb3_base_class():
    from tests.support import fake_sb3
    from minigrid_rl_b200 import vec_env as ve
    cls = ve.bind_vecenv_base(fake_sb3.VecEnv)
    assert issubclass(cls, fake_sb3.VecEnv) and cls._vecenv_base is fake_sb3.VecEnv
    assert not getattr(cls, "__abstractmethods__", None)          # every abstract method of the base is implemented
    plain = ve.bind_vecenv_base(object)
    assert plain.__mro__[1] is ve._B200VecEnvImpl and fake_sb3.VecEnv not in plain.__mro__


def test_expert_vocabulary_round_trip():
    import minigrid_rl_b200 as mg
    tab = mg.expert_token_table()
    vocab = [" "] + [chr(c) for c in range(ord("a"), ord("z") + 1)]
    for i, text in enumerate(mg.MISSIONS):
        assert "".join(vocab[k] for k in tab[i]).rstrip() == text


def test_wire_format_host_expander_matches_the_code_table():
    """csrc/mgrl_wire_host.cpp (SSSE3 shuffles) against the definition of the PCIe wire code: 128 | state << 3 | colour for
    doors (type 4), type << 3 | colour otherwise.  Host code only: runs without a GPU."""
    import ctypes
    lib = ctypes.CDLL(mg.library_path())
    if not lib.mgrl_wire_have_ssse3():
        pytest.skip("CPU without SSSE3: the library uses its scalar table loop")
    rs = np.random.RandomState(0)
    valid = [c for c in range(256) if (c < 128 and (c >> 3) <= 10 and (c & 7) < 6) or (128 <= c < 160 and ((c >> 3) & 3) < 3 and (c & 7) < 6)]
    for trial in range(100):
        rec = np.zeros(64, np.uint8)
        rec[:49] = rs.choice(valid, 49)
        rec[49:] = rs.randint(0, 256, 15)                    # scalars / tag / pad: must not leak into the image
        for pad in (0, 1):
            out = np.full(160, 0xAA, np.uint8)
            lib.mgrl_wire_expand_hwc_ssse3(rec.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), pad)
            want = np.array([(4, c & 7, (c >> 3) & 3) if c >= 128 else (c >> 3, c & 7, 0) for c in rec[:49].tolist()], np.uint8)
            assert np.array_equal(out[:147], want.reshape(-1)), trial
            assert out[147] == (0 if pad else 0xAA) and (out[148:] == 0xAA).all()


def test_wire_record_expansion_matches_the_format_on_every_instruction_set():
    """Host half of the PCIe wire format (csrc/mgrl_wire_host.cpp): 64-byte records (49 cell codes, code = 128 | state << 3 |
    colour for doors, type << 3 | colour otherwise) -> HWC observation records, SSSE3 and - when the CPU has it - AVX-512
    VBMI, for both pitches, against a numpy decode of the format.  No GPU involved."""
    import ctypes as C
    from minigrid_rl_b200 import _native as nat
    lib = nat.lib()
    lib.mgrl_wire_expand_block_hwc_ssse3.restype = C.c_int
    lib.mgrl_wire_expand_block_hwc_avx512.restype = C.c_int
    if not lib.mgrl_wire_have_ssse3():
        pytest.skip("no SSSE3")
    rs = np.random.RandomState(5)
    n, tag = 1536, 9
    t = rs.randint(0, 11, size=(n, 49)); c = rs.randint(0, 6, size=(n, 49)); s = rs.randint(0, 3, size=(n, 49)) * (t == 4)
    rec = np.zeros((n, 64), np.uint8)
    rec[:, :49] = np.where(t == 4, 128 | (s << 3) | c, (t << 3) | c)
    rec[:, 49:55] = rs.randint(0, 256, size=(n, 6)); rec[:, 55] = tag; rec[:, 56:60] = rs.randint(0, 256, size=(n, 4))
    want = np.stack([t, c, s], axis=2).reshape(n, 147).astype(np.uint8)
    abort = C.c_int(0)
    fns = [lib.mgrl_wire_expand_block_hwc_ssse3] + ([lib.mgrl_wire_expand_block_hwc_avx512] if lib.mgrl_wire_have_avx512vbmi() else [])
    for fn in fns:
        for pitch in (147, 148):
            raw = np.full(n * pitch + 128, 0xEE, np.uint8)
            off = (-raw.ctypes.data) % 64                      # the block routines want an aligned destination
            out = raw[off:off + n * pitch]
            got = fn(C.c_void_p(rec.ctypes.data), n, C.c_void_p(out.ctypes.data), pitch, tag, 55, C.byref(abort), None)
            assert got == n
            img = out.reshape(n, pitch)
            assert np.array_equal(img[:, :147], want), (fn, pitch)
            if pitch == 148:
                assert not img[:, 147].any()
            assert raw[off + n * pitch] == 0xEE               # nothing written past the block
    if lib.mgrl_wire_have_avx512vbmi():
        # groups of 16 records of the 148-byte pitch: images realigned in registers, the step's scalars 16 at a time; a ragged
        # count stops at the last whole group
        lib.mgrl_wire_expand_groups_hwc148_avx512.restype = C.c_int
        for count in (n, n - 5, 16, 15):
            raw = np.full(n * 148 + 128, 0xEE, np.uint8)
            off = (-raw.ctypes.data) % 64
            out = raw[off:off + n * 148]
            sc = [np.full(n, 0xEE, np.uint8) for _ in range(6)]
            rew = np.full(n, 0xEE, np.uint8).repeat(4).view(np.float32)
            got = lib.mgrl_wire_expand_groups_hwc148_avx512(
                C.c_void_p(rec.ctypes.data), count, C.c_void_p(out.ctypes.data), tag, 55, C.byref(abort), None,
                *[C.c_void_p(a.ctypes.data) for a in sc], C.c_void_p(rew.ctypes.data))
            done = count & ~15
            assert got == done
            img = out.reshape(n, 148)
            assert np.array_equal(img[:done, :147], want[:done]) and not img[:done, 147].any()
            assert (raw[off + done * 148:] == 0xEE).all()      # nothing written past the last whole group
            for a, col in zip(sc, (49, 50, 51, 52, 53, 54)):
                assert np.array_equal(a[:done], rec[:done, col]) and (a[done:] == 0xEE).all(), col
            assert np.array_equal(rew.view(np.uint8).reshape(n, 4)[:done], rec[:done, 56:60])
            assert (rew.view(np.uint8).reshape(n, 4)[done:] == 0xEE).all()
