"""Linear inverted pendulum CoM reference — counterpart of ref:ctrl/LIPM.py:5-49 (same class, same methods,
same semi-implicit Euler update), plus a batched form over numpy/torch arrays for per-env references.
"""
import math

import numpy as np

from .Trajectory import Trajectory as Traj


class LIPM:
    def __init__(self, h0, dt: float = None):
        """pos/vel/acc/dcm/zmp are [x, y] pairs (ref:ctrl/LIPM.py:7-13)."""
        self.w = np.sqrt(9.80665 / h0)  # ref:ctrl/LIPM.py:15 (note: 9.80665 here, 9.81 in the dynamics)
        kw = {} if dt is None else {"dt": dt}
        self.x = Traj(**kw)
        self.y = Traj(**kw)

    def pos(self, t):
        return np.array([self.x.get_frame(t, 0), self.y.get_frame(t, 0)])

    def vel(self, t):
        return np.array([self.x.get_frame(t, 1), self.y.get_frame(t, 1)])

    def acc(self, t):
        return np.array([self.x.get_frame(t, 2), self.y.get_frame(t, 2)])

    def dcm(self, t):
        return self.pos(t) + self.vel(t) / self.w

    def zmp(self, t):
        return self.pos(t) - self.acc(t) / self.w**2

    def make_trajectory(self, t, dt, pos0, vel0, acc0, zmp):
        """ref:ctrl/LIPM.py:34-49.  Like the reference, pos0 and vel0 are updated IN PLACE (pos = pos0 aliases
        the caller's array and `+=` mutates it), so a caller can chain segments by passing the same arrays."""
        duration = t[1] - t[0]
        pos = pos0
        vel = vel0
        acc = acc0
        for _ in range(math.floor(duration / dt)):
            acc = (zmp - pos) * self.w**2
            vel += acc * dt
            pos += vel * dt
            self.x.traj.append([pos[0], vel[0], acc[0]])
            self.y.traj.append([pos[1], vel[1], acc[1]])

    # ---- batched form (an addition): one Euler step for N pendulums at once -----------------------------
    @staticmethod
    def step(pos, vel, zmp, w, dt):
        """pos, vel, zmp: [..., 2] numpy arrays or torch tensors; returns (pos, vel, acc) after one step of
        acc = (zmp - pos) w^2; vel += acc dt; pos += vel dt  (the loop body of make_trajectory)."""
        acc = (zmp - pos) * w**2
        vel = vel + acc * dt
        pos = pos + vel * dt
        return pos, vel, acc
