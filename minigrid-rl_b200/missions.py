"""Mission strings and their token encoding.

Mission text is produced at /root/reference/src/custom_env.py:185,197,208,213,259 and
tokenised by TokenizeVocabWrapper (environment.py:69-112): vocabulary
[' ', '\\n', '-', ':', ',', '.', 'a'..'z'] -> indices 0..31, zero padded to 32 tokens.
The simulator stores a one-byte mission id; this module is the id <-> text <-> tokens map.
"""
from __future__ import annotations

import numpy as np

COLOURS = ["red", "green", "blue", "purple", "yellow", "grey"]      # upstream COLOR_TO_IDX order
TYPES = ["key", "ball", "box", "door"]
COMMANDS = ["go to", "toggle", "pick up"]                            # msn_commands[0..2], custom_env.py:87-94
N_MISSIONS = 74
MISSION_GOAL, MISSION_DROP = 72, 73
# 'move <direction>' (problems mov / full, custom_env.py:216-256) takes the ids of 'toggle <colour> key', which the reference
# never generates ('toggle' only picks boxes and doors, :197-201): the table - and the policy's 74 x 4 mission GRU rows - keep
# their size
MISSION_MOVE0 = 24
DIRECTIONS = ["left", "right", "up", "down"]                         # msn_directions, custom_env.py:96-101
MSN_LEN = 32
VOCAB = [" ", "\n", "-", ":", ",", "."] + [chr(c) for c in range(ord("a"), ord("z") + 1)]


def mission_string(mid: int) -> str:
    if mid == MISSION_GOAL:
        return "go to goal"
    if mid == MISSION_DROP:
        return "drop"
    if MISSION_MOVE0 <= mid < MISSION_MOVE0 + 4:
        return f"move {DIRECTIONS[mid - MISSION_MOVE0]}"
    if not 0 <= mid < 72:
        raise ValueError(f"bad mission id {mid}")
    group, rem = divmod(mid, 24)
    typ, colour = divmod(rem, 6)
    return f"{COMMANDS[group]} {COLOURS[colour]} {TYPES[typ]}"


MISSIONS = [mission_string(i) for i in range(N_MISSIONS)]
_IDS = {s: i for i, s in enumerate(MISSIONS)}


def mission_id(text: str) -> int:
    return _IDS[text]


def tokenize(text: str) -> np.ndarray:
    out = np.zeros(MSN_LEN, np.int64)
    for i, ch in enumerate(text.lower()):
        out[i] = VOCAB.index(ch)
    return out


def token_table() -> np.ndarray:
    """[74, 32] int64: row = mission id."""
    return np.stack([tokenize(s) for s in MISSIONS])


EXPERT_VOCAB = [" "] + [chr(c) for c in range(ord("a"), ord("z") + 1)]     # /root/reference/src/experts.py:181-182


def expert_token_table() -> np.ndarray:
    """[74, 32] int64 in the 27-symbol vocabulary that Expert.decode_missions decodes with (it differs from
    TokenizeVocabWrapper's 32-symbol one, environment.py:74-80: a reference inconsistency)."""
    out = np.zeros((N_MISSIONS, MSN_LEN), np.int64)
    for i, text in enumerate(MISSIONS):
        for k, ch in enumerate(text):
            out[i, k] = EXPERT_VOCAB.index(ch)
    return out
