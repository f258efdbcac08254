"""MissionSpace stand-in: the reference only constructs it (custom_env.py:103-106)."""


class MissionSpace:
    def __init__(self, mission_func, ordered_placeholders=None, seed=None):
        self.mission_func = mission_func
        self.ordered_placeholders = ordered_placeholders
        self.shape = None
        self.dtype = str
