"""Sampled-trajectory container indexed by floor(t / dt) — counterpart of ref:ctrl/Trajectory.py:4-15.

The reference module cannot be imported (`from ctrl.conf import dt`, :1, names a module attribute that does
not exist) and get_frame indexes `self.traj[t]` with the float time (:13).  Here dt is a constructor argument
defaulting to RobotConfig.dt and the bound check uses the sample index, which is what :9-14 intend.
"""
import math

from .conf import RobotConfig


class Trajectory:
    def __init__(self, dt: float = RobotConfig.dt):
        self.dt = dt
        self.traj = []

    def get_frame(self, t, diff):
        k = math.floor(t / self.dt)
        if k < 0 or k >= len(self.traj):
            raise IndexError("Time index out of bounds")
        if diff < 0 or diff >= len(self.traj[k]):
            raise IndexError("Difference index out of bounds")
        return self.traj[k][diff]
