#!/usr/bin/env python
"""Record what the UNMODIFIED reference `Expert` (/root/reference/src/experts.py:128-178) does on full-grid observations
(TEST INFRASTRUCTURE; runs in the build container only, like gen_golden.py whose machinery it reuses).

Pipeline of /root/reference/src/experts_test.py:27-47: PlaygroundEnv -> FullyObsWrapper -> tokens, one environment,
`Expert(cfg)(obs, None, False)` until the episode ends.  One deviation, forced by the reference itself: experts_test.py
tokenises the mission with TokenizeVocabWrapper (32 symbols, environment.py:74-80) while Expert.decode_missions
(experts.py:181-182) decodes with its own 27-symbol vocabulary (' ' + a..z), so the unmodified test script raises
IndexError on its first step (verified here).  The fixture therefore feeds the expert tokens in ITS vocabulary, built
from the environment's mission string; the product exposes both encodings (`B200VecEnv(obs_mode="full",
token_vocab="expert")`).

Recorded per scenario (tests/golden/expert_*.npz): for every step the full-grid image (S, S, 3), the direction, the
mission tokens in both vocabularies, the expert's action, float32 reward, terminated, truncated and the episode returns.
tests/test_oracle_golden.py replays the actions through the C oracle's full observation (pinning `mg_full_obs` to the
FullyObsWrapper semantics as the reference runs them); tests/test_gpu_vec_env.py does the same through the CUDA path.

Usage:  python oracle/gen_expert_golden.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402

EXPERT_VOCAB = [" "] + [chr(c) for c in range(ord("a"), ord("z") + 1)]           # experts.py:181-182


def expert_tokens(mission: str) -> np.ndarray:
    out = np.zeros(32, np.int64)
    for i, ch in enumerate(mission[:32]):
        out[i] = EXPERT_VOCAB.index(ch)
    return out


SCENARIOS = {
    # name: (cfg kwargs, episodes, seed)
    "single_gtg_obst": (dict(problem="gtg", mission=None, num_objects=6, obstacles=True), 10, 1337),   # experts_test.py:24
    "multi_gtg": (dict(problem="multi", mission=5), 10, 7),
    "multi_gto": (dict(problem="multi", mission=0), 10, 8),
    "multi_pkp": (dict(problem="multi", mission=2), 10, 9),
}


def run(mods, kw, n_episodes, seed):
    custom_env, environment = mods
    import experts
    from minigrid.wrappers import FullyObsWrapper
    cfg = gg.make_cfg(**kw)
    cfg["algo"] = "test"
    base = custom_env.PlaygroundEnv(render_mode="rgb_array", cfg=cfg, manual=False)
    base._np_random = gg._NPRandom()
    env = environment.TokenizeVocabWrapper(FullyObsWrapper(base))
    S = cfg.env.size
    rec = dict(image=[], dir=[], tokens=[], tokens_expert=[], action=[], reward=[], term=[], trunc=[], first=[])
    returns = []
    for ep in range(n_episodes):
        gg.CURRENT = gg.Stream(seed, 0, ep)
        obs, _ = env.reset()
        gg.CURRENT = None
        expert = experts.Expert(cfg)                           # experts_test.py:40-41: a new expert per episode
        done, first = False, 1
        while not done:
            batch = {"image": np.asarray(obs["image"])[None], "direction": np.asarray([obs["direction"]]),
                     "mission": expert_tokens(base.mission)[None]}
            action, _ = expert(batch, None, False)
            a = int(action[0])
            rec["image"].append(np.asarray(obs["image"], np.uint8)); rec["dir"].append(int(obs["direction"]))
            rec["tokens"].append(np.asarray(obs["mission"], np.int64)); rec["tokens_expert"].append(batch["mission"][0])
            rec["action"].append(a); rec["first"].append(first)
            first = 0
            obs, r, term, trunc, _ = env.step(a)
            rec["reward"].append(np.float32(r)); rec["term"].append(int(term)); rec["trunc"].append(int(trunc))
            done = term or trunc
        returns.append(float(np.float32(r)))
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["returns"] = np.asarray(returns, np.float32)
    out["cfg_json"] = np.frombuffer(json.dumps(kw).encode(), np.uint8)
    out["seed"] = np.array(seed, np.uint64)
    out["size"] = np.array(S)
    return out


def main():
    mods = gg.import_reference()
    for name, (kw, n_ep, seed) in SCENARIOS.items():
        out = run(mods, kw, n_ep, seed)
        np.savez_compressed(os.path.join(gg.GOLDEN_DIR, f"expert_{name}.npz"), **out)
        print(f"expert_{name:16s} steps={len(out['action'])} returns={np.round(out['returns'], 3).tolist()}")


if __name__ == "__main__":
    main()
