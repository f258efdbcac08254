"""Environment configuration: same keys as `cfg.env` in the reference's hydra configs
(/root/reference/src/hydra_configs/single.yaml:20-28), consumed at custom_env.py:75-80,
110-114, 134, 155-156, 601, 637."""
from __future__ import annotations

from dataclasses import dataclass
from math import floor

PROBLEMS = {"multi": 0, "gto": 1, "gtg": 2, "opn": 3, "pkp": 4, "drp": 5, "mov": 6, "full": 7}
# README task names -> cfg.env.mission on the `multi` map (single.yaml:22)
TASKS = {"GTG": 5, "GTO": 0, "PKP": 2, "TGL": 1, "ALL": None}


@dataclass
class EnvConfig:
    problem: str = "multi"
    mission: int | None = 5
    all_doors_open: bool = False
    size: int = 11
    num_objects: int = 4
    see_through_walls: bool = True
    obstacles: bool = False
    percent_obstacles: float = 0.05

    @classmethod
    def for_task(cls, task: str, **kw) -> "EnvConfig":
        return cls(problem="multi", mission=TASKS[task.upper()], **kw)

    @classmethod
    def from_cfg(cls, env_cfg) -> "EnvConfig":
        """Build from the reference's `cfg.env` mapping/namespace (hydra DictConfig or dict)."""
        get = (lambda k, d: env_cfg.get(k, d)) if hasattr(env_cfg, "get") else (lambda k, d: getattr(env_cfg, k, d))
        return cls(problem=get("problem", "multi"), mission=get("mission", None),
                   all_doors_open=bool(get("all_doors_open", False)), size=int(get("size", 11)),
                   num_objects=int(get("num_objects", 4)),
                   see_through_walls=bool(get("see_through_walls", True)),
                   obstacles=bool(get("obstacles", False)),
                   percent_obstacles=float(get("percent_obstacles", 0.05)))

    @property
    def max_steps(self) -> int:
        return self.size * self.size  # custom_env.py:114

    @property
    def num_obstacles(self) -> int:
        return floor((self.size - 2) ** 2 * self.percent_obstacles) if self.obstacles else 0  # custom_env.py:156

    def validate(self) -> None:
        if self.problem not in PROBLEMS:
            raise ValueError(f"Invalid problem type given: {self.problem} (supported: {sorted(PROBLEMS)})")
        if self.mission not in (None, 0, 1, 2, 5):
            raise ValueError("mission must be one of 0 ('go to'), 1 ('toggle'), 2 ('pick up'), 5 ('go to goal') or None")
        if not 5 <= self.size <= 11:
            raise ValueError("size must be in 5..11")
