"""Product device functions (mgrl_core.cuh, compiled for the host by tests/support) against
the CPU oracle: bit-exact states/observations/rewards/flags on seeded random traces and on
the golden traces recorded from the reference.  CPU tier: no GPU, no compute through the
C-ABI library."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.support import emul

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TRACES = sorted(glob.glob(os.path.join(GOLDEN, "trace_*.npz")))

CONFIGS = {
    "gtg": dict(problem="multi", mission=5),
    "gto": dict(problem="multi", mission=0),
    "pkp": dict(problem="multi", mission=2),
    "tgl": dict(problem="multi", mission=1),
    "all": dict(problem="multi", mission=None),
    "all_vis": dict(problem="multi", mission=None, see_through_walls=False),
    "all_open_n6": dict(problem="multi", mission=None, all_doors_open=True, num_objects=6),
    "all_s9": dict(problem="multi", mission=None, size=9),
    "lava": dict(problem="multi", mission=None, obstacles=True),
    "single_gtg_obst": dict(problem="gtg", mission=None, num_objects=6, obstacles=True, see_through_walls=False),
    "single_opn": dict(problem="opn", mission=None),
    "single_pkp": dict(problem="pkp", mission=None),
    "single_gto": dict(problem="gto", mission=None),
    "single_drp": dict(problem="drp", mission=None),
    "single_mov": dict(problem="mov", mission=None, num_objects=6),
    "single_full": dict(problem="full", mission=None),
    "single_full_obst_s10": dict(problem="full", mission=None, obstacles=True, percent_obstacles=0.08, size=10),
}


def assert_same(a, b, ctx):
    for name in orc.STATE_DTYPE.names:
        assert np.array_equal(a.states[name], b.states[name]), f"{ctx}: state.{name}"
    assert np.array_equal(a.obs, b.obs), ctx
    assert np.array_equal(a.dir, b.dir), ctx
    assert np.array_equal(a.mission, b.mission), ctx


def biased_actions(rs, n, t):
    # uniform actions end episodes after ~7 steps; mix in a 'never done' population so that
    # doors, keys, boxes, truncation and the latch leak all get exercised
    a = rs.randint(0, 7, size=n).astype(np.uint8)
    long_lived = (np.arange(n) % 3) != 0
    redo = long_lived & (a == 6) & (rs.rand(n) < 0.97)
    a[redo] = rs.choice([2, 2, 2, 0, 1, 3, 5, 5, 4], size=int(redo.sum())).astype(np.uint8)
    return a


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_random_traces_match_oracle(name):
    cfg = orc.make_config(**CONFIGS[name])
    n, T = 192, 260
    o = orc.OracleVecEnv(cfg, n, seed=2024, env_id_base=1000)
    e = emul.EmulVecEnv(cfg, n, seed=2024, env_id_base=1000)
    o.reset(); e.reset()
    assert_same(o, e, f"{name} reset")
    rs = np.random.RandomState(5)
    n_done = 0
    for t in range(T):
        a = biased_actions(rs, n, t)
        o.step(a); e.step(a)
        ctx = f"{name} t={t}"
        assert_same(o, e, ctx)
        assert np.array_equal(o.reward.view(np.uint32), e.reward.view(np.uint32)), ctx
        assert np.array_equal(o.term, e.term) and np.array_equal(o.trunc, e.trunc), ctx
        assert np.array_equal(o.ep_len, e.ep_len), ctx
        d = (o.term | o.trunc).astype(bool)
        assert np.array_equal(o.term_obs[d], e.term_obs[d]), ctx
        n_done += int(d.sum())
    assert n_done > n
    assert int(o.states["error"].max()) == 0


@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p)[6:-4] for p in TRACES])
def test_golden_traces_through_product_core(path):
    z = np.load(path)
    cfg = orc.make_config(**json.loads(bytes(z["cfg_json"]).decode()))
    E, T = z["init_state"].shape[0], z["actions"].shape[0]
    e = emul.EmulVecEnv(cfg, E, seed=int(z["seed"]))
    obs = e.reset()
    assert np.array_equal(obs, z["init_obs"])
    for name in orc.STATE_DTYPE.names:
        assert np.array_equal(e.states[name], z["init_state"][name]), name
    for t in range(T):
        obs, rew, term, trunc = e.step(z["actions"][t])
        for name in orc.STATE_DTYPE.names:
            assert np.array_equal(e.states[name], z["state"][t][name]), (t, name)
        assert np.array_equal(obs, z["obs"][t]), t
        assert np.array_equal(rew.view(np.uint32), z["reward"][t].view(np.uint32)), t
        assert np.array_equal(term, z["term"][t]) and np.array_equal(trunc, z["trunc"][t]), t


def test_chw_layout_is_transpose_of_hwc():
    import ctypes as C
    o = None
    for see in (False, True):
      cfg = orc.make_config(problem="multi", mission=None, see_through_walls=see)
      o = orc.OracleVecEnv(cfg, 64, seed=3)
      o.reset()
      for i in range(64):
        hwc = np.zeros((7, 7, 3), np.uint8)
        chw = np.zeros((3, 7, 7), np.uint8)
        emul.lib().emul_obs(C.byref(cfg), emul.p(o.states[i:i + 1]), 9, 0, emul.p(hwc))
        emul.lib().emul_obs(C.byref(cfg), emul.p(o.states[i:i + 1]), 9, 1, emul.p(chw))
        assert np.array_equal(chw, hwc.transpose(2, 0, 1))
        packed = np.full(148, 255, np.uint8)      # 148-byte-pitch packed record == HWC + one zero pad byte
        emul.lib().emul_obs(C.byref(cfg), emul.p(o.states[i:i + 1]), 9, 2, emul.p(packed))
        assert np.array_equal(packed[:147].reshape(7, 7, 3), hwc) and packed[147] == 0
        assert tuple(hwc[3, 6]) == (5, 1, 0)  # carried green key shown on the agent's cell
        assert np.array_equal(hwc, orc.gen_obs(cfg, o.states[i:i + 1], carrying=9))


def test_kind_encode_table():
    for k in list(range(4)) + list(range(8, 14)) + list(range(16, 22)) + \
            [24 + 8 * s + c for s in range(3) for c in range(6)] + [64 + 8 * m + c for m in range(7) for c in range(6)]:
        e = emul.lib().emul_kind_encode(k)
        assert (e & 255, (e >> 8) & 255, (e >> 16) & 255) == orc.kind_encode(k), k


def test_full_obs_matches_oracle():
    import ctypes as C
    cfg = orc.make_config(problem="gtg", mission=None, num_objects=6, obstacles=True)
    o = orc.OracleVecEnv(cfg, 32, seed=11)
    o.reset()
    want = o.full_obs()
    for i in range(32):
        got = np.zeros((11, 11, 3), np.uint8)
        emul.lib().emul_full_obs(C.byref(cfg), emul.p(o.states[i:i + 1]), emul.p(got))
        assert np.array_equal(got, want[i])
