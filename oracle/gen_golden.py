#!/usr/bin/env python
"""Record golden traces from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs in the build container only (needs /root/reference).  Imports the reference's own
`custom_env.PlaygroundEnv`, `environment.TokenizeVocabWrapper` and
`environment.Discrete2BoxWrapper` on top of oracle/upstream_shim (the restated public
minigrid/gymnasium API — see its README), replaces the reference's two RNG sources
(Python `random.choice/randint` imported at custom_env.py:3-4 and gymnasium `np_random`)
by this repo's Philox stream *draw for draw*, and records, for every scenario:

  * the canonical state dump (oracle/mg_oracle.h `mg_state`, 140 B) after every reset/step,
  * observations, direction, mission tokens, float32 reward, terminated, truncated,
    terminal observations, under DummyVecEnv-style auto-reset.

tests/test_oracle_golden.py replays the same seeds/actions through oracle/mg_oracle.c and
requires every byte to match, which pins the C oracle to the reference's code (the
[UPSTREAM] shim itself stays "unpinned").

Usage:  python oracle/gen_golden.py            # rewrites tests/golden/*.npz
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE_SRC = "/root/reference/src"
GOLDEN_DIR = os.environ.get("MGRL_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden"))   # (override: regenerate elsewhere and compare)

# ----------------------------------------------------------------------------- Philox (independent of mg_oracle.c)
M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return [c0, c1, c2, c3]


class Stream:
    """Per-(seed, env, episode) draw stream: draw d = word d&3 of Philox block d>>2."""

    def __init__(self, seed, env_id, episode):
        self.key = (seed & M32, (seed >> 32) & M32)
        self.tail = (episode & M32, env_id & M32, (env_id >> 32) & M32)
        self.ndraw = 0
        self.buf = None

    def below(self, n):
        w = self.ndraw & 3
        if w == 0:
            self.buf = philox4x32_10((self.ndraw >> 2,) + self.tail, self.key)
        self.ndraw += 1
        return (self.buf[w] * int(n)) >> 32


CURRENT: Stream | None = None


def _choice(seq):
    return seq[CURRENT.below(len(seq))]


def _randint(a, b):
    return a + CURRENT.below(b - a + 1)


class _NPRandom:
    def integers(self, low, high=None):
        if high is None:
            low, high = 0, low
        return low + CURRENT.below(high - low)

    def choice(self, seq):
        return seq[CURRENT.below(len(seq))]


# ----------------------------------------------------------------------------- reference import
def import_reference():
    sys.path.insert(0, os.path.join(HERE, "upstream_shim"))
    sys.path.insert(0, REFERENCE_SRC)
    import custom_env
    import environment
    assert getattr(sys.modules["minigrid"], "__shim__", False)
    custom_env.choice = _choice      # `from random import choice, randint` (custom_env.py:4)
    custom_env.randint = _randint
    return custom_env, environment


class NS(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def make_cfg(problem="multi", mission=5, size=11, num_objects=4, all_doors_open=False,
             see_through_walls=True, obstacles=False, percent_obstacles=0.05, seed=42):
    return NS(seed=seed,
              env=NS(problem=problem, mission=mission, all_doors_open=all_doors_open, size=size,
                     num_objects=num_objects, see_through_walls=see_through_walls,
                     obstacles=obstacles, percent_obstacles=percent_obstacles),
              algorithm=NS(n_frames_stack=4, recurrent=False))


# ----------------------------------------------------------------------------- canonical dump of a reference env
COLOR_TO_IDX = {"red": 0, "green": 1, "blue": 2, "purple": 3, "yellow": 4, "grey": 5}
TYPE4 = {"key": 0, "ball": 1, "box": 2, "door": 3}

STATE_DTYPE = np.dtype([
    ("grid", np.uint8, (121,)),
    ("agent_x", np.uint8), ("agent_y", np.uint8), ("agent_dir", np.uint8),
    ("carrying", np.uint8), ("step_count", np.uint8),
    ("target_x", np.uint8), ("target_y", np.uint8), ("target_action", np.uint8),
    ("mission_id", np.uint8), ("mission_done", np.uint8), ("latch_step", np.uint8),
    ("episode", np.uint32), ("reset_draws", np.uint16), ("error", np.uint8), ("pad", np.uint8),
])


def kind_of(obj):
    if obj is None:
        return 0
    c = COLOR_TO_IDX[obj.color]
    t = obj.type
    if t == "wall":
        assert obj.color == "grey"
        return 1
    if t == "goal":
        return 2
    if t == "lava":
        return 3
    if t == "key":
        return 8 + c
    if t == "ball":
        return 16 + c
    if t == "door":
        s = 0 if obj.is_open else (2 if obj.is_locked else 1)
        return 24 + 8 * s + c
    if t == "box":
        m = 0
        if obj.contains is not None:
            assert obj.contains.type == "key"
            m = 1 + COLOR_TO_IDX[obj.contains.color]
        return 64 + 8 * m + c
    raise ValueError(t)


def mission_id_of(mission: str) -> int:
    if mission == "go to goal":
        return 72
    if mission == "drop":
        return 73
    if mission.startswith("move "):      # ids of the never-generated 'toggle <colour> key' (mg_oracle.h)
        return 24 + ("left", "right", "up", "down").index(mission[5:])
    for g, prefix in enumerate(("go to ", "toggle ", "pick up ")):
        if mission.startswith(prefix):
            colour, typ = mission[len(prefix):].split(" ")
            return g * 24 + TYPE4[typ] * 6 + COLOR_TO_IDX[colour]
    raise ValueError(mission)


def dump_state(base, episode, reset_draws):
    """PlaygroundEnv -> mg_state record (public attributes only: SURVEY.md §7 'canonical state dump')."""
    S = base.width
    st = np.zeros((), STATE_DTYPE)
    for y in range(S):
        for x in range(S):
            st["grid"][y * S + x] = kind_of(base.grid.get(x, y))
    st["agent_x"], st["agent_y"] = int(base.agent_pos[0]), int(base.agent_pos[1])
    st["agent_dir"] = int(base.agent_dir)
    st["carrying"] = kind_of(base.carrying)
    st["step_count"] = base.step_count
    if base.target_pos is None:
        st["target_x"] = st["target_y"] = 0xFF
    else:
        st["target_x"], st["target_y"] = int(base.target_pos[0]), int(base.target_pos[1])
    st["target_action"] = 0 if base.target_action is None else int(base.target_action)
    st["mission_id"] = mission_id_of(base.mission)
    if base.mission.startswith("move "):
        # target_range (custom_env.py:216-256) as decimal digits over the four bytes target_x, target_y, target_action, pad:
        # digit k = the coordinate of the cell in row y = k+1 (left / right) or column x = k+1 (up / down), 0 = none
        assert base.target_pos is None and base.target_action is None
        horizontal = base.mission[5:] in ("left", "right")
        v, seen = 0, set()
        for (x, y) in base.target_range:
            k, c = (y, x) if horizontal else (x, y)
            assert k not in seen and 1 <= c <= S - 2 and 1 <= k <= S - 2
            seen.add(k)
            v += int(c) * 10 ** (k - 1)
        st["target_x"], st["target_y"], st["target_action"], st["pad"] = v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF, v >> 24
    else:
        assert not base.target_range
    st["mission_done"] = int(bool(base.mission_done))
    if base.reward is None:
        st["latch_step"] = 0
    else:
        ks = [k for k in range(base.max_steps + 1) if 1 - 0.9 * (k / base.max_steps) == base.reward]
        assert len(ks) == 1, (base.reward, ks)
        st["latch_step"] = ks[0]
    assert bool(base.mission_done) == (base.reward is not None)
    st["episode"] = episode
    st["reset_draws"] = reset_draws
    return st


# ----------------------------------------------------------------------------- action policies
def front_kind(base):
    fx, fy = base.front_pos
    return kind_of(base.grid.get(int(fx), int(fy)))


def pick_action(policy, base, rs):
    if policy == 0:                      # uniform random over the 7 actions
        return int(rs.randint(7))
    k = front_kind(base)
    # policy 2 only says done right after a reset that inherited a latched reward from a
    # truncated episode: that is the stale-reward path of SURVEY App. B Q1
    p_done = 0.02 if policy == 1 else (0.5 if (base.mission_done and base.step_count < 4) else 0.0)
    u = rs.rand()
    if u < p_done:
        return 6
    if k >= 24 and (k < 48 or k >= 64) and rs.rand() < 0.6:   # door or box ahead
        return 5 if rs.rand() < 0.7 else 3
    if (8 <= k < 24) and rs.rand() < 0.6:                     # key / ball ahead
        return 3
    if base.carrying is not None and k == 0 and rs.rand() < 0.08:
        return 4
    u = rs.rand()
    if u < 0.55:
        return 2
    if u < 0.75:
        return 0
    if u < 0.95:
        return 1
    return int(rs.randint(3, 6))


# ----------------------------------------------------------------------------- scenario runner
def run_scenario(mods, name, cfg_kwargs, n_envs, n_steps, seed, stats):
    global CURRENT
    custom_env, environment = mods
    cfg = make_cfg(**cfg_kwargs)
    envs, bases, episodes = [], [], [0] * n_envs
    for _ in range(n_envs):
        e = environment.make_env("custom", "rgb_array", cfg=cfg)
        b = e.unwrapped
        b._np_random = _NPRandom()
        envs.append(e)
        bases.append(b)

    def do_reset(i):
        global CURRENT
        CURRENT = Stream(seed, i, episodes[i])
        obs, _ = envs[i].reset()
        nd = CURRENT.ndraw
        CURRENT = None
        episodes[i] += 1
        return obs, nd

    E, T = n_envs, n_steps
    out = dict(
        init_state=np.zeros(E, STATE_DTYPE), init_obs=np.zeros((E, 7, 7, 3), np.uint8),
        init_dir=np.zeros(E, np.uint8), init_tokens=np.zeros((E, 32), np.int8),
        actions=np.zeros((T, E), np.uint8), obs=np.zeros((T, E, 7, 7, 3), np.uint8),
        dir=np.zeros((T, E), np.uint8), tokens=np.zeros((T, E, 32), np.int8),
        reward=np.zeros((T, E), np.float32), term=np.zeros((T, E), np.uint8),
        trunc=np.zeros((T, E), np.uint8), ep_len=np.zeros((T, E), np.uint8),
        term_obs=np.zeros((T, E, 7, 7, 3), np.uint8), state=np.zeros((T, E), STATE_DTYPE),
        carry_obs=np.zeros((T, E), np.uint8),
    )
    missions = set()
    for i in range(E):
        obs, nd = do_reset(i)
        out["init_state"][i] = dump_state(bases[i], episodes[i], nd)
        out["init_obs"][i] = obs["image"]
        assert obs["direction"].sum() == 1
        out["init_dir"][i] = int(np.argmax(obs["direction"]))
        out["init_tokens"][i] = obs["mission"]
        missions.add(bases[i].mission)

    rs = np.random.RandomState(seed * 7919 + 13)
    last_nd = [int(s["reset_draws"]) for s in out["init_state"]]
    for t in range(T):
        for i in range(E):
            b = bases[i]
            a = pick_action(i % 3, b, rs)
            carrying_before = kind_of(b.carrying)
            fk = front_kind(b)
            leaked = bool(b.mission_done) and b.step_count < 4 and i % 3 == 2
            obs, r, term, trunc, _ = envs[i].step(a)
            out["actions"][t, i] = a
            out["reward"][t, i] = np.float32(r)       # VecEnv float32 reward buffer [UPSTREAM]
            out["term"][t, i], out["trunc"][t, i] = term, trunc
            out["carry_obs"][t, i] = tuple(obs["image"][3, 6]) != (1, 0, 0)
            # coverage counters
            if a == 5 and 40 <= fk < 48 and 24 <= front_kind(b) < 32:
                stats["unlock"] += 1
            if a == 5 and fk >= 64:
                stats["box_open_key" if fk >= 72 else "box_open_empty"] += 1
            if a == 5 and carrying_before and b.carrying is None:
                stats["consumed"] += 1
            if a == 3 and carrying_before == 0 and b.carrying is not None:
                stats["pickup"] += 1
            if a == 4 and carrying_before and b.carrying is None:
                stats["drop"] += 1
            if term and r > 0:
                stats["success"] += 1
            if trunc:
                stats["truncated"] += 1
            if leaked and a == 6 and r > 0:
                stats["q1_stale_reward"] += 1
            if term or trunc:
                out["ep_len"][t, i] = b.step_count
                out["term_obs"][t, i] = obs["image"]
                obs, nd = do_reset(i)
                last_nd[i] = nd
                missions.add(b.mission)
                stats["episodes"] += 1
            out["obs"][t, i] = obs["image"]
            assert obs["direction"].sum() == 1
            out["dir"][t, i] = int(np.argmax(obs["direction"]))
            out["tokens"][t, i] = obs["mission"]
            out["state"][t, i] = dump_state(b, episodes[i], last_nd[i])
    out["cfg_json"] = np.frombuffer(json.dumps(cfg_kwargs).encode(), np.uint8)
    out["seed"] = np.array(seed, np.uint64)
    out["missions"] = np.array(sorted(missions))
    return out


def run_layouts(mods, cfg_kwargs, n_envs, n_episodes, seed):
    """Layout-only fixture: many (env, episode) keys -> state after reset."""
    global CURRENT
    custom_env, environment = mods
    cfg = make_cfg(**cfg_kwargs)
    e = environment.make_env("custom", "rgb_array", cfg=cfg)
    b = e.unwrapped
    b._np_random = _NPRandom()
    states = np.zeros((n_envs, n_episodes), STATE_DTYPE)
    obs0 = np.zeros((n_envs, n_episodes, 7, 7, 3), np.uint8)
    for i in range(n_envs):
        for ep in range(n_episodes):
            CURRENT = Stream(seed, i, ep)
            b.mission_done, b.reward = False, None
            obs, _ = e.reset()
            states[i, ep] = dump_state(b, ep + 1, CURRENT.ndraw)
            obs0[i, ep] = obs["image"]
            CURRENT = None
    return dict(states=states, obs=obs0, seed=np.array(seed, np.uint64),
                cfg_json=np.frombuffer(json.dumps(cfg_kwargs).encode(), np.uint8))


SCENARIOS = {
    # name: (cfg kwargs, n_envs, n_steps, seed)     BASELINE.json configs 1-5 first
    "multi_gtg": (dict(problem="multi", mission=5), 9, 400, 42),
    "multi_gto": (dict(problem="multi", mission=0), 9, 400, 43),
    "multi_pkp": (dict(problem="multi", mission=2), 9, 500, 44),
    "multi_tgl": (dict(problem="multi", mission=1), 9, 500, 45),
    "multi_all": (dict(problem="multi", mission=None), 12, 600, 46),
    "multi_all_vis": (dict(problem="multi", mission=None, see_through_walls=False), 9, 400, 47),
    "multi_all_open": (dict(problem="multi", mission=None, all_doors_open=True), 9, 300, 48),
    "multi_all_n6_s9": (dict(problem="multi", mission=None, num_objects=6, size=9), 6, 300, 49),
    # (size 8 is not used: its 2x2 rooms can leave the reference's unbounded rejection
    #  loops without any admissible cell -- SURVEY App. B Q7 -- and the reference hangs)
    "multi_all_s10_n2": (dict(problem="multi", mission=None, num_objects=2, size=10), 6, 200, 50),
    "multi_gtg_lava": (dict(problem="multi", mission=5, obstacles=True), 6, 300, 51),
    "single_gto": (dict(problem="gto", mission=None), 6, 200, 52),
    "single_gtg_obst": (dict(problem="gtg", mission=None, num_objects=6, obstacles=True,
                             see_through_walls=False), 6, 300, 53),
    "single_opn": (dict(problem="opn", mission=None), 6, 200, 54),
    "single_pkp": (dict(problem="pkp", mission=None), 6, 200, 55),
    "single_drp": (dict(problem="drp", mission=None), 6, 200, 56),
    "single_mov": (dict(problem="mov", mission=None, num_objects=6), 8, 250, 57),
    "single_full": (dict(problem="full", mission=None), 12, 300, 58),
    "single_full_obst": (dict(problem="full", mission=None, obstacles=True, percent_obstacles=0.08, size=10), 6, 200, 59),
}

LAYOUTS = {
    "layouts_multi_all": (dict(problem="multi", mission=None), 64, 8, 1234),
    "layouts_multi_all_open": (dict(problem="multi", mission=None, all_doors_open=True), 32, 4, 99),
    "layouts_multi_tgl_n6": (dict(problem="multi", mission=1, num_objects=6), 32, 4, 7),
    "layouts_full": (dict(problem="full", mission=None), 48, 6, 77),
    "layouts_mov_n8_s9": (dict(problem="mov", mission=None, num_objects=8, size=9), 32, 4, 78),
}


def main():
    mods = import_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    # Philox known-answer vectors (Random123 kat_vectors, philox4x32 10 rounds)
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox4x32_10((M32,) * 4, (M32, M32)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    import signal
    signal.signal(signal.SIGALRM, lambda *_: (_ for _ in ()).throw(TimeoutError("reference hung (Q7)")))
    only = set(sys.argv[1:])             # optional: names of the fixtures to (re)generate
    for name, (kw, n_envs, n_steps, seed) in SCENARIOS.items():
        if only and name not in only:
            continue
        signal.alarm(120)
        stats = dict(unlock=0, box_open_key=0, box_open_empty=0, consumed=0, pickup=0, drop=0,
                     success=0, truncated=0, q1_stale_reward=0, episodes=0)
        out = run_scenario(mods, name, kw, n_envs, n_steps, seed, stats)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"trace_{name}.npz"), **out)
        print(f"{name:18s} {stats}  missions={len(out['missions'])}")
    for name, (kw, n_envs, n_eps, seed) in LAYOUTS.items():
        if only and name not in only:
            continue
        signal.alarm(120)
        out = run_layouts(mods, kw, n_envs, n_eps, seed)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **out)
        print(f"{name:18s} {out['states'].shape}")
    signal.alarm(0)


if __name__ == "__main__":
    main()
