"""numpy restatement of the device gait phase machine (tsid_control_b200/csrc/tsidb_gait.cuh) — TEST
INFRASTRUCTURE.  It follows the reference pieces the device code cites: contact switching
(ref:legacy/biped.py:168-212), swing-foot shape (ref:ctrl/Foot_Trajectory.py:8-19, rise_ratio 0.5), LIPM Euler
step (ref:ctrl/LIPM.py:15,44-47), and the gait cycle of the benchmark workload.  The swing
polynomial is checked against this repo's FootTrajectory (itself pinned to the reference's scipy splines by
tests/golden/planners.npz) in tests/test_gait.py."""
from __future__ import annotations

import numpy as np


def mask_of(phi):
    return np.where(phi < 0.2, 3, np.where(phi < 0.6, 1, 2)).astype(np.uint8)


class GaitRef:
    def __init__(self, n, dt, step_duration, step_length, step_height, com_height, defaults, phase0=None, vcmd=None,
                 steps=None, n_steps=None, rise_ratio=0.5):
        """steps [n,S,4] / n_steps [n]: a footstep plan (FootstepPlanner.plan) -> planned mode: every swing is this repo's
        FootTrajectory (pinned to the reference's scipy splines by tests/golden/planners.npz) from the lift-off placement
        to the env's next footstep of that side."""
        self.steps, self.n_steps, self.rr = steps, n_steps, rise_ratio
        self.step_idx = np.full(n, 2, np.int64)
        self.swing = [[None, None] for _ in range(n)]
        self.n, self.dt, self.T, self.L, self.h = n, dt, step_duration, step_length, step_height
        self.w2 = 9.80665 / com_height
        self.com_z = float(defaults["com"][2])
        self.phi = np.zeros(n) if phase0 is None else np.array(phase0, dtype=np.float64)
        self.mask = mask_of(self.phi)
        self.vcmd = np.zeros((n, 2)) if vcmd is None else np.array(vcmd, dtype=np.float64)
        self.pos = np.tile(defaults["com"][:2], (n, 1)).astype(np.float64)
        self.vel = self.vcmd.copy()
        self.com = np.tile(defaults["com"], (n, 1)).astype(np.float64)
        self.com[:, 3:5] = self.vcmd
        self.foot = [np.tile(defaults["foot_lf"], (n, 1)).astype(np.float64), np.tile(defaults["foot_rf"], (n, 1)).astype(np.float64)]
        self.contact = [np.tile(defaults["contact_lf"], (n, 1)).astype(np.float64), np.tile(defaults["contact_rf"], (n, 1)).astype(np.float64)]
        self.origin = [self.foot[0][:, :12].copy(), self.foot[1][:, :12].copy()]
        self.fails = np.zeros(n, np.int32)
        if steps is not None:  # a plan handed over while a foot is in the air: that swing starts from the standing placement
            for f in (1, 0):
                up = (self.mask & (1 << f)) == 0
                self._begin(f, up, self.origin[f])

    def _begin(self, f, lift, now):
        from tsid_control_b200.ctrl.Foot_Trajectory import FootTrajectory

        for e in np.where(lift)[0]:
            idx = int(self.step_idx[e])
            while idx < self.n_steps[e] and int(self.steps[e, idx, 3]) != f:
                idx += 1
            yaw = np.arctan2(now[e, 4], now[e, 3])
            start = np.array([now[e, 0], now[e, 1], now[e, 2], yaw])
            if idx < self.n_steps[e]:
                target = np.array([self.steps[e, idx, 0], self.steps[e, idx, 1], now[e, 2], self.steps[e, idx, 2]])
                idx += 1
            else:
                target = start.copy()
            self.step_idx[e] = idx
            self.swing[e][f] = (start, FootTrajectory([0.0, self.T], start, target, self.h, self.rr))

    def _planned(self, f, lift, sw, s, now):
        self._begin(f, lift, now)
        fr, org = self.foot[f], self.origin[f]
        for e in np.where(sw)[0]:
            start, tj = self.swing[e][f]
            t = s[e] * self.T
            fr[e, :3] = tj.get_position(t)
            dy = tj.yaw(t) - start[3]
            R0 = org[e, 3:12].reshape(3, 3).T  # column-major
            Rz = np.array([[np.cos(dy), -np.sin(dy), 0], [np.sin(dy), np.cos(dy), 0], [0, 0, 1]])
            fr[e, 3:12] = (Rz @ R0).T.ravel()
            fr[e, 12:15] = tj.velocity(t)
            fr[e, 15:18] = [0.0, 0.0, tj.yaw(t, 1)]
            fr[e, 18:21] = tj.acceleration(t)
            fr[e, 21:24] = [0.0, 0.0, tj.yaw(t, 2)]

    def refs(self):
        return {"com": self.com, "foot_lf": self.foot[0], "foot_rf": self.foot[1], "contact_lf": self.contact[0],
                "contact_rf": self.contact[1]}

    def step(self, foot_now_lf, foot_now_rf, status=None):
        phi = self.phi + 0.4 * self.dt / self.T
        phi = np.where(phi >= 1.0, phi - 1.0, phi)
        old, nm = self.mask, mask_of(phi)
        L = np.where(self.vcmd[:, 0] >= 0.0, self.L, -self.L)
        T, h = self.T, self.h
        for f, now in ((0, foot_now_lf), (1, foot_now_rf)):
            bit = 1 << f
            lift = ((old & bit) != 0) & ((nm & bit) == 0)
            land = ((old & bit) == 0) & ((nm & bit) != 0)
            self.origin[f][lift] = now[lift]
            self.contact[f][land] = now[land]
            self.foot[f][land, :12] = now[land]
            self.foot[f][land, 12:] = 0.0
            sw = (nm & bit) == 0
            s = (phi - (0.2 if f == 1 else 0.6)) / 0.4
            fr, org = self.foot[f], self.origin[f]
            if self.steps is not None:
                self._planned(f, lift, sw, s, now)
                continue
            fr[sw, 0] = org[sw, 0] + L[sw] * s[sw]
            fr[sw, 1] = org[sw, 1]
            fr[sw, 2] = org[sw, 2] + 4.0 * h * s[sw] * (1.0 - s[sw])
            fr[sw, 3:12] = org[sw, 3:12]
            fr[sw, 12:] = 0.0
            fr[sw, 12] = L[sw] / T
            fr[sw, 14] = 4.0 * h * (1.0 - 2.0 * s[sw]) / T
            fr[sw, 20] = -8.0 * h / (T * T)
        cl, cr = self.contact[0][:, :2], self.contact[1][:, :2]
        zmp = np.where((nm == 3)[:, None], 0.5 * (cl + cr), np.where((nm == 1)[:, None], cl, cr))
        acc = (zmp - self.pos) * self.w2
        self.vel = self.vel + acc * self.dt
        self.pos = self.pos + self.vel * self.dt
        self.com[:, 0:2] = self.pos
        self.com[:, 2] = self.com_z
        self.com[:, 3:5] = self.vel
        self.com[:, 5] = 0.0
        self.com[:, 6:8] = acc
        self.com[:, 8] = 0.0
        self.phi, self.mask = phi, nm
        if status is not None:
            self.fails += (np.asarray(status) != 0).astype(np.int32)
