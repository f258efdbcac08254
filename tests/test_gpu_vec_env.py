"""GPU tier: the drop-in host surface (B200VecEnv: numpy in / numpy out, SB3 VecEnv protocol after
VecTransposeImage + VecFrameStack(4,'first')) against the oracle env + numpy SB3-wrapper oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import minigrid_rl_b200 as mg  # noqa: E402
from oracle import oracle as orc, sb3_oracle  # noqa: E402


def biased_actions(rs, n):
    a = rs.randint(0, 7, size=n)
    long_lived = (np.arange(n) % 3) != 0
    redo = long_lived & (a == 6) & (rs.rand(n) < 0.97)
    a[redo] = rs.choice([2, 2, 2, 0, 1, 3, 5, 5, 4], size=int(redo.sum()))
    return a


@pytest.mark.parametrize("kw", [dict(problem="multi", mission=5), dict(problem="multi", mission=None),
                                dict(problem="multi", mission=1, see_through_walls=False)],
                         ids=["gtg", "all", "tgl_vis"])
@pytest.mark.parametrize("host_stack,n,T", [("inplace", 48, 400), ("inplace", 1100, 40), ("device", 48, 200)],
                         ids=["inplace", "inplace_blocks", "device_stack"])
def test_vec_env_matches_sb3_stack_semantics(kw, host_stack, n, T):
    """`inplace`: mgrl_vec_step_stacked_host (64-byte wire records, the observation dict maintained in place by the library's
    host threads; 1100 environments = three work blocks); `device_stack`: mgrl_vec_step_host (stack on the device, full copy)."""
    env = mg.B200VecEnv(mg.EnvConfig(**kw), num_envs=n, seed=42, host_stack=host_stack)
    assert env.observation_space["image"].shape == (12, 7, 7)
    assert env.observation_space["direction"].shape == (16,)
    assert env.observation_space["mission"].shape == (128,) and env.action_space.n == 7
    assert list(env.observation_space.keys()) == ["direction", "image", "mission"]
    o = orc.OracleVecEnv(orc.make_config(**kw), n, seed=42)
    table = mg.token_table()
    fs_img = sb3_oracle.FrameStack(n, (3, 7, 7), np.uint8)
    fs_dir = sb3_oracle.FrameStack(n, (4,), np.uint8)
    fs_mis = sb3_oracle.FrameStack(n, (32,), np.int64)

    obs = env.reset()
    o.reset()
    assert np.array_equal(obs["image"], fs_img.reset(o.obs.transpose(0, 3, 1, 2)))
    assert np.array_equal(obs["direction"], fs_dir.reset(sb3_oracle.one_hot_dir(o.dir)))
    assert np.array_equal(obs["mission"], fs_mis.reset(table[o.mission]))
    assert obs["mission"].dtype == np.int64 and obs["image"].dtype == np.uint8
    rs = np.random.RandomState(0)
    lut = orc.reward_lut(121)
    n_done = n_trunc = 0
    for t in range(T):
        a = biased_actions(rs, n)
        # terminal direction / mission from the oracle, before its auto-reset
        pre_mission = o.mission.copy()
        tdir = np.zeros(n, np.uint8)
        for i in range(n):
            s = o.states[i:i + 1].copy()
            orc.step_one(o.cfg, lut, s, int(a[i]))
            tdir[i] = s["agent_dir"][0]
        obs, rew, dones, infos = env.step(a)
        o.step(a.astype(np.uint8))
        want_done = (o.term | o.trunc).astype(bool)
        want_img, term_img = fs_img.update(o.obs.transpose(0, 3, 1, 2), want_done, o.term_obs.transpose(0, 3, 1, 2))
        want_dir, term_dir = fs_dir.update(sb3_oracle.one_hot_dir(o.dir), want_done, sb3_oracle.one_hot_dir(tdir))
        want_mis, term_mis = fs_mis.update(table[o.mission], want_done, table[pre_mission])
        assert np.array_equal(obs["image"], want_img), t
        assert np.array_equal(obs["direction"], want_dir), t
        assert np.array_equal(obs["mission"], want_mis), t
        assert rew.dtype == np.float32 and np.array_equal(rew.view(np.uint32), o.reward.view(np.uint32)), t
        assert dones.dtype == bool and np.array_equal(dones, want_done), t
        assert len(infos) == n
        for i in range(n):
            if want_done[i]:
                info = infos[i]
                assert np.array_equal(info["terminal_observation"]["image"], term_img[i]), (t, i)
                assert np.array_equal(info["terminal_observation"]["direction"], term_dir[i]), (t, i)
                assert np.array_equal(info["terminal_observation"]["mission"], term_mis[i]), (t, i)
                assert info["TimeLimit.truncated"] == bool(o.trunc[i] and not o.term[i])
                assert info["episode"]["l"] == int(o.ep_len[i])
                assert np.float32(info["episode"]["r"]) == o.reward[i]
                n_trunc += int(info["TimeLimit.truncated"])
            else:
                assert "terminal_observation" not in infos[i]
        n_done += int(want_done.sum())
    assert n_done > 50 and (n_trunc > 0 or T < 121)
    st = env.get_state()
    for name in orc.STATE_DTYPE.names:
        assert np.array_equal(st[name], o.states[name]), name
    assert env.get_attr("mission") == [mg.MISSIONS[i] for i in o.mission]
    with pytest.raises(ValueError):
        env.step(np.full(n, 7))
    env.close()


def test_step_frames_host_path_matches_oracle():
    """mgrl_vec_step_frames_host: the un-stacked host path (like for like with the CPU arm of bench.py)."""
    kw = dict(problem="multi", mission=None)
    n = 700
    for layout in ("hwc148", "hwc"):
        env = mg.B200VecEnv(mg.EnvConfig(**kw), num_envs=n, seed=9, env_id_base=64, layout=layout)
        o = orc.OracleVecEnv(orc.make_config(**kw), n, seed=9, env_id_base=64)
        img, d, m = env.reset_frames()
        o.reset()
        assert np.array_equal(img[:, :147].reshape(n, 7, 7, 3), o.obs) and np.array_equal(d, o.dir) and np.array_equal(m, o.mission)
        rs = np.random.RandomState(3)
        for t in range(60):
            a = biased_actions(rs, n).astype(np.uint8)
            img, d, m, rew, term, trunc, ep_len, timg, tdir = env.step_frames(a, want_terminal=True)
            o.step(a)
            assert np.array_equal(img[:, :147].reshape(n, 7, 7, 3), o.obs), t
            assert np.array_equal(d, o.dir) and np.array_equal(m, o.mission), t
            assert np.array_equal(rew.view(np.uint32), o.reward.view(np.uint32)), t
            assert np.array_equal(term, o.term) and np.array_equal(trunc, o.trunc) and np.array_equal(ep_len, o.ep_len), t
            done = (o.term | o.trunc).astype(bool)
            assert np.array_equal(timg[done][:, :147].reshape(-1, 7, 7, 3), o.term_obs[done]), t
        env.close()


@pytest.mark.parametrize("direct", ["auto", "0", "0.4", "1"])
@pytest.mark.parametrize("layout,see", [("hwc148", True), ("chw", False), ("hwc", False)])
def test_step_frames_wire_format_with_chunks_and_host_threads(layout, see, direct, monkeypatch):
    """The host path's PCIe wire format (csrc/mgrl_wire.cu: one code byte per view cell, 64-byte records, chunked copies
    expanded by the handle's host threads) at a size that uses several chunks and threads: every output array bit-exact
    against the oracle, including occluded views (unseen cells = (0,0,0)) and the terminal observations.  `direct`: the
    share of the batch whose images cross PCIe as they are (MGRL_WIRE_DIRECT; auto = the split follows the host threads'
    idle share, so it moves while the test runs); steps with terminal outputs always go as records."""
    monkeypatch.setenv("MGRL_WIRE_DIRECT", direct)
    kw = dict(problem="multi", mission=None, see_through_walls=see)
    n = 20000
    env = mg.B200VecEnv(mg.EnvConfig(**kw), num_envs=n, seed=21, env_id_base=5, layout=layout)
    o = orc.OracleVecEnv(orc.make_config(**kw), n, seed=21, env_id_base=5, nthreads=8)
    img, d, m = env.reset_frames()
    o.reset()

    def view(x):
        x = x[:, :147]
        return x.reshape(-1, 3, 7, 7).transpose(0, 2, 3, 1) if layout == "chw" else x.reshape(-1, 7, 7, 3)

    assert np.array_equal(view(img), o.obs)
    rs = np.random.RandomState(8)
    for t in range(24):
        a = biased_actions(rs, n).astype(np.uint8)
        img, d, m, rew, term, trunc, ep_len, timg, tdir = env.step_frames(a, want_terminal=True)
        o.step(a)
        assert np.array_equal(view(img), o.obs), t
        assert np.array_equal(d, o.dir) and np.array_equal(m, o.mission), t
        assert np.array_equal(rew.view(np.uint32), o.reward.view(np.uint32)), t
        assert np.array_equal(term, o.term) and np.array_equal(trunc, o.trunc) and np.array_equal(ep_len, o.ep_len), t
        done = (o.term | o.trunc).astype(bool)
        assert np.array_equal(view(timg[done]), o.term_obs[done]), t
        for _ in range(3 if t % 2 else 0):             # and without the terminal outputs (one job; hybrid transfer)
            a = biased_actions(rs, n).astype(np.uint8)
            img, d, m, rew, term, trunc, ep_len = env.step_frames(a)[:7]
            o.step(a)
            assert np.array_equal(view(img), o.obs) and np.array_equal(rew.view(np.uint32), o.reward.view(np.uint32)), t
            assert np.array_equal(d, o.dir) and np.array_equal(m, o.mission), t
            assert np.array_equal(term, o.term) and np.array_equal(trunc, o.trunc) and np.array_equal(ep_len, o.ep_len), t
    env.close()


def test_sb3_vecenv_binding_drives_collect_rollouts():
    """ppo.py:134,159: the class bound over SB3's VecEnv base (a faithful stand-in here, tests/support/fake_sb3.py) with
    gymnasium-style spaces goes through OnPolicyAlgorithm.collect_rollouts' use of the environment - spaces -> buffers,
    step, infos, truncation bootstrap - and fills the same buffers as the oracle env behind the numpy SB3-wrapper oracle."""
    import os
    import sys
    from tests.support import fake_sb3
    from minigrid_rl_b200 import vec_env as ve
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "upstream_shim")
    sys.path.insert(0, shim)
    try:
        import gymnasium
        cls = ve.bind_vecenv_base(fake_sb3.VecEnv)
        kw = dict(problem="multi", mission=5)
        n, T, gamma = 96, 160, 0.81
        env = cls(mg.EnvConfig(**kw), num_envs=n, seed=5)
    finally:
        sys.path.remove(shim)
    assert isinstance(env, fake_sb3.VecEnv) and isinstance(env.observation_space, gymnasium.spaces.Dict)
    assert isinstance(env.observation_space["image"], gymnasium.spaces.Box) and isinstance(env.action_space, gymnasium.spaces.Discrete)
    assert env.render_mode is None and env.render() is None and env.env_method("render") == [None] * n
    assert env.env_is_wrapped(object) == [False] * n and len(env.reset_infos) == n

    rs = np.random.RandomState(11)
    plan = np.stack([biased_actions(rs, n) for _ in range(T)])
    plan[:, ::2] = np.where(plan[:, ::2] == 6, 2, plan[:, ::2])           # half of the envs never say done: truncations happen
    value_of = lambda obs: float(obs["image"].astype(np.float64).sum() % 97) / 97.0    # noqa: E731  a deterministic critic
    buf, ep_info, last = fake_sb3.collect_rollouts(env, T, lambda t, obs: plan[t], value_of, gamma, seed=5)

    # the same loop on the oracle env + numpy wrapper stack
    o = orc.OracleVecEnv(orc.make_config(**kw), n, seed=5)
    table = mg.token_table()
    fs = (sb3_oracle.FrameStack(n, (3, 7, 7), np.uint8), sb3_oracle.FrameStack(n, (4,), np.uint8),
          sb3_oracle.FrameStack(n, (32,), np.int64))
    o.reset()
    cur = (fs[0].reset(o.obs.transpose(0, 3, 1, 2)), fs[1].reset(sb3_oracle.one_hot_dir(o.dir)), fs[2].reset(table[o.mission]))
    lut = orc.reward_lut(121)
    want_eps, n_boot = [], 0
    starts = np.ones(n, bool)
    for t in range(T):
        assert np.array_equal(buf["image"][t], cur[0]) and np.array_equal(buf["direction"][t], cur[1]), t
        assert np.array_equal(buf["mission"][t], cur[2]) and np.array_equal(buf["episode_starts"][t], starts.astype(np.float32)), t
        a = plan[t]
        pre_mission = o.mission.copy()
        tdir = np.zeros(n, np.uint8)
        for i in range(n):
            s = o.states[i:i + 1].copy()
            orc.step_one(o.cfg, lut, s, int(a[i]))
            tdir[i] = s["agent_dir"][0]
        o.step(a.astype(np.uint8))
        done = (o.term | o.trunc).astype(bool)
        img, timg = fs[0].update(o.obs.transpose(0, 3, 1, 2), done, o.term_obs.transpose(0, 3, 1, 2))
        d, tdirs = fs[1].update(sb3_oracle.one_hot_dir(o.dir), done, sb3_oracle.one_hot_dir(tdir))
        m, tmis = fs[2].update(table[o.mission], done, table[pre_mission])
        r = o.reward.copy()
        for i in np.flatnonzero(done):
            want_eps.append((t, int(i), float(o.reward[i]), int(o.ep_len[i])))
            if o.trunc[i] and not o.term[i]:
                r[i] += np.float32(gamma) * np.float32(value_of({"image": timg[int(i)], "direction": tdirs[int(i)], "mission": tmis[int(i)]}))
                n_boot += 1
        assert np.array_equal(buf["rewards"][t].view(np.uint32), r.view(np.uint32)), t
        cur, starts = (img, d, m), done
    assert ep_info == want_eps and n_boot > 0 and len(want_eps) > 100
    env.close()


@pytest.mark.parametrize("name", ["single_gtg_obst", "multi_gtg", "multi_gto", "multi_pkp"])
def test_full_obs_mode_replays_the_reference_expert(name):
    """experts_test.py:27-47 on the device path: `obs_mode="full"` must show the unmodified reference Expert exactly the
    observations it saw on the reference env (tests/golden/expert_*.npz, oracle/gen_expert_golden.py), so that it takes
    the same actions; rewards and episode ends follow."""
    import json
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"expert_{name}.npz"))
    kw = json.loads(bytes(z["cfg_json"]).decode())
    for vocab, key in (("expert", "tokens_expert"), ("reference", "tokens")):
        env = mg.B200VecEnv(mg.EnvConfig(**kw), num_envs=1, seed=int(z["seed"]), obs_mode="full", token_vocab=vocab)
        S = int(z["size"])
        assert tuple(env.observation_space["image"].shape) == (S, S, 3)
        obs = env.reset()
        for t in range(len(z["action"])):
            assert np.array_equal(obs["image"][0], z["image"][t]), (name, t)
            assert int(obs["direction"][0]) == int(z["dir"][t]) and np.array_equal(obs["mission"][0], z[key][t]), (name, t)
            obs, rew, dones, infos = env.step(np.array([z["action"][t]]))
            assert np.float32(rew[0]) == z["reward"][t] and bool(dones[0]) == bool(z["term"][t] or z["trunc"][t]), (name, t)
            if dones[0]:
                assert infos[0]["episode"]["r"] == float(z["reward"][t])
        env.close()
