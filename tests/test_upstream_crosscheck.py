"""Cross-check of the [UPSTREAM] half of the oracle against the real Farama `minigrid` package, for the day it is
installable (oracle/upstream_shim/README.md).  Skipped in this image: `minigrid` and `gymnasium` are
not installed and cannot be (no network), which is why the upstream semantics are pinned only through the reference's
own call sites (tests/golden/, oracle/gen_golden.py) and stay "[UPSTREAM] - unpinned".

What it does when the package is there: oracle states (layouts of the C oracle's generator) are rebuilt as real
`minigrid` grids, then `MiniGridEnv.gen_obs()` and `MiniGridEnv.step()` of the real package are compared with
`mg_gen_obs` / `mg_step_one` for every action."""
import numpy as np
import pytest

minigrid = pytest.importorskip("minigrid")
if getattr(minigrid, "__shim__", False):
    pytest.skip("only the restated shim is importable, not the real package", allow_module_level=True)

from minigrid.core.grid import Grid  # noqa: E402
from minigrid.core.mission import MissionSpace  # noqa: E402
from minigrid.core.world_object import Ball, Box, Door, Goal, Key, Lava, Wall  # noqa: E402
from minigrid.minigrid_env import MiniGridEnv  # noqa: E402

from oracle import oracle as orc  # noqa: E402

COLOURS = ["red", "green", "blue", "purple", "yellow", "grey"]


def world_object(kind):
    """kind byte (include/mgrl.h) -> minigrid WorldObj"""
    if kind == 0:
        return None
    if kind == 1:
        return Wall()
    if kind == 2:
        return Goal()
    if kind == 3:
        return Lava()
    c = COLOURS[kind & 7]
    if kind < 16:
        return Key(c)
    if kind < 24:
        return Ball(c)
    if kind < 48:
        state = (kind - 24) >> 3
        return Door(c, is_open=state == 0, is_locked=state == 2)
    m = (kind - 64) >> 3
    return Box(c, contains=Key(COLOURS[m - 1]) if m else None)


class Replay(MiniGridEnv):
    def __init__(self, size, see_through_walls, max_steps):
        super().__init__(mission_space=MissionSpace(mission_func=lambda: "x"), grid_size=size, max_steps=max_steps,
                         see_through_walls=see_through_walls, agent_view_size=7)

    def _gen_grid(self, width, height):
        self.grid = Grid(width, height)
        self.agent_pos, self.agent_dir, self.mission = (1, 1), 0, "x"

    def load(self, st, S):
        self.reset()
        for y in range(S):
            for x in range(S):
                self.grid.set(x, y, world_object(int(st["grid"][y * S + x])))
        self.agent_pos = (int(st["agent_x"]), int(st["agent_y"]))
        self.agent_dir = int(st["agent_dir"])
        self.carrying = world_object(int(st["carrying"]))
        self.step_count = int(st["step_count"])


@pytest.mark.parametrize("see", [True, False])
def test_gen_obs_and_step_match_the_real_minigrid(see):
    cfg = orc.make_config(problem="multi", mission=None, see_through_walls=see)
    S = 11
    env = orc.OracleVecEnv(cfg, 64, seed=3)
    env.reset()
    rs = np.random.RandomState(0)
    real = Replay(S, see, 121)
    lut = orc.reward_lut(121)
    for _ in range(30):
        env.step(rs.randint(0, 6, size=64).astype(np.uint8))
        for i in range(0, 64, 7):
            st = env.states[i:i + 1].copy()
            real.load(st[0], S)
            assert np.array_equal(real.gen_obs()["image"], orc.gen_obs(cfg, st)[0])
            for a in range(6):
                s2 = st.copy()
                r, term, trunc, _ = orc.step_one(cfg, lut, s2, a)[:4]
                real.load(st[0], S)
                _, _, rterm, rtrunc, _ = real.step(a)
                assert (int(real.agent_pos[0]), int(real.agent_pos[1]), int(real.agent_dir)) == \
                    (int(s2["agent_x"][0]), int(s2["agent_y"][0]), int(s2["agent_dir"][0]))
                assert bool(rtrunc) == bool(trunc)
