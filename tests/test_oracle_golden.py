"""The C oracle (oracle/mg_oracle.c) against golden traces recorded from the UNMODIFIED
reference PlaygroundEnv + wrappers (oracle/gen_golden.py).  Bit-exact on every byte:
full state after every reset/step, observations, direction, float32 reward bits,
terminated/truncated, episode lengths and terminal observations."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TRACES = sorted(glob.glob(os.path.join(GOLDEN, "trace_*.npz")))
LAYOUTS = sorted(glob.glob(os.path.join(GOLDEN, "layouts_*.npz")))


def _cfg(z):
    return orc.make_config(**json.loads(bytes(z["cfg_json"]).decode()))


def _assert_states_equal(got, want, ctx):
    for name in want.dtype.names:
        if not np.array_equal(got[name], want[name]):
            bad = np.argwhere(got[name] != want[name])[0]
            raise AssertionError(f"{ctx}: state field {name} differs at {bad}: got {got[name][tuple(bad)]}, "
                                 f"want {want[name][tuple(bad)]}")


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 with 10 rounds
    assert orc.philox((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    m = 0xFFFFFFFF
    assert orc.philox((m, m, m, m), (m, m)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_reward_lut_matches_readme_values():
    # /root/reference/README.md:84-95 prints these float32 rewards
    lut = orc.reward_lut(121)
    want = {1: "0.9925619960", 6: "0.9553719163", 7: "0.9479339123", 18: "0.8661156893",
            23: "0.8289256096", 34: "0.7471074462", 36: "0.7322313786", 37: "0.7247933745",
            121: "0.1000000015"}
    for k, txt in want.items():
        assert "%.10f" % float(lut[k]) == txt
    ref = np.array([np.float32(1 - 0.9 * (k / 121)) for k in range(122)], np.float32)
    assert np.array_equal(lut.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p)[6:-4] for p in TRACES])
def test_trace_replay_bit_exact(path):
    z = np.load(path)
    cfg = _cfg(z)
    E = z["init_state"].shape[0]
    T = z["actions"].shape[0]
    env = orc.OracleVecEnv(cfg, E, seed=int(z["seed"]))
    obs = env.reset()
    _assert_states_equal(env.states, z["init_state"], "reset")
    assert np.array_equal(obs, z["init_obs"])
    assert np.array_equal(env.dir, z["init_dir"])
    for t in range(T):
        obs, rew, term, trunc = env.step(z["actions"][t])
        ctx = f"{os.path.basename(path)} t={t}"
        _assert_states_equal(env.states, z["state"][t], ctx)
        assert np.array_equal(obs, z["obs"][t]), ctx
        assert np.array_equal(env.dir, z["dir"][t]), ctx
        assert np.array_equal(rew.view(np.uint32), z["reward"][t].view(np.uint32)), ctx
        assert np.array_equal(term, z["term"][t]), ctx
        assert np.array_equal(trunc, z["trunc"][t]), ctx
        assert np.array_equal(env.ep_len, z["ep_len"][t]), ctx
        done = (term | trunc).astype(bool)
        assert np.array_equal(env.term_obs[done], z["term_obs"][t][done]), ctx
    assert int(env.states["error"].max()) == 0


@pytest.mark.parametrize("path", LAYOUTS, ids=[os.path.basename(p)[:-4] for p in LAYOUTS])
def test_layouts_bit_exact(path):
    z = np.load(path)
    cfg = _cfg(z)
    want = z["states"]
    n_envs, n_eps = want.shape
    for i in range(n_envs):
        for ep in range(n_eps):
            st = np.zeros(1, orc.STATE_DTYPE)
            st["episode"] = ep
            nd = orc.generate(cfg, int(z["seed"]), i, st)
            assert nd == int(want[i, ep]["reset_draws"])
            _assert_states_equal(st[0], want[i, ep], f"layout env={i} ep={ep}")
            assert np.array_equal(orc.gen_obs(cfg, st), z["obs"][i, ep])


EXPERTS = sorted(glob.glob(os.path.join(GOLDEN, "expert_*.npz")))


@pytest.mark.parametrize("path", EXPERTS, ids=[os.path.basename(p)[7:-4] for p in EXPERTS])
def test_full_observation_matches_the_reference_expert_fixture(path):
    """`mg_full_obs` (FullyObsWrapper as the reference runs it, experts_test.py:27-30) and the step dynamics under the
    actions the UNMODIFIED reference Expert took (oracle/gen_expert_golden.py): every full-grid image, direction, token row,
    reward and episode end."""
    from minigrid_rl_b200 import missions
    z = np.load(path)
    env = orc.OracleVecEnv(_cfg(z), 1, seed=int(z["seed"]))
    env.reset()
    table, table_x = missions.token_table(), missions.expert_token_table()
    for t in range(len(z["action"])):
        assert np.array_equal(env.full_obs()[0], z["image"][t]), t
        assert int(env.dir[0]) == int(z["dir"][t]), t
        assert np.array_equal(table[env.mission[0]], z["tokens"][t]) and np.array_equal(table_x[env.mission[0]], z["tokens_expert"][t]), t
        env.step(np.array([z["action"][t]], np.uint8))
        assert env.reward[0] == z["reward"][t] and env.term[0] == z["term"][t] and env.trunc[0] == z["trunc"][t], t
    assert (z["returns"] > 0).sum() >= 3          # the expert solves episodes (README.md:84-95 style rewards)
