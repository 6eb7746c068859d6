"""Swing-foot trajectory — counterpart of ref:ctrl/Foot_Trajectory.py:5-43.

The reference builds scipy CubicSplines through 2 knots (x, y, yaw), 3 knots (z when rise_ratio == 0.5) or
4 knots (z otherwise).  With scipy's default not-a-knot boundary those are single polynomials of degree 1, 2
and 3, so they are evaluated here in closed form (Newton divided differences) — no scipy at run time, and the
same code runs on numpy arrays and torch tensors for per-env batches.

Accessor semantics are the reference's, including its derivative orders: get_velocity returns the SECOND
derivative and get_acceleration the THIRD (ref:ctrl/Foot_Trajectory.py:35,43 pass 2 and 3 to the spline).
velocity()/acceleration() give the first and second derivative, which is what a TSID foot reference needs.
"""
from typing import List

import numpy as np


class _Poly:
    """Interpolating polynomial through (ts, ys) in Newton form; derivative-aware evaluation."""

    def __init__(self, ts, ys):
        ts = [float(t) for t in ts]
        self.ts = ts
        n = len(ts)
        dd = [y for y in ys]
        coef = [dd[0]]
        for j in range(1, n):
            dd = [(dd[i + 1] - dd[i]) / (ts[i + j] - ts[i]) for i in range(n - j)]
            coef.append(dd[0])
        # expand to monomials in (t - ts[0]) for easy differentiation
        mono = [0.0 * coef[0]] * n
        basis = [1.0]  # coefficients of prod_{k<j} (t - ts[k]) in powers of s = t - ts[0]
        for j in range(n):
            for p, b in enumerate(basis):
                mono[p] = mono[p] + coef[j] * b
            if j + 1 < n:
                shift = ts[j] - ts[0]
                nxt = [0.0] * (len(basis) + 1)
                for p, b in enumerate(basis):
                    nxt[p + 1] += b
                    nxt[p] -= shift * b
                basis = nxt
        self.mono = mono

    def __call__(self, t, nu: int = 0):
        s = t - self.ts[0]
        n = len(self.mono)
        out = 0.0 * s
        for p in range(n - 1, nu - 1, -1):
            f = 1.0
            for k in range(nu):
                f *= p - k
            out = out * s + f * self.mono[p]
        return out


class FootTrajectory:
    def __init__(self, t: List, start, target, step_height: float, rise_ratio: float = 0.5):
        self.t = t
        self.x = _Poly(t, [start[0], target[0]])
        self.y = _Poly(t, [start[1], target[1]])
        self.z = None
        self.yaw = _Poly(t, [start[3], target[3]]) if len(start) > 3 else None
        duration = t[1] - t[0]
        if rise_ratio != 0.5:
            rise_time = duration * rise_ratio
            new_t = [t[0], t[0] + rise_time, t[1] - rise_time, t[1]]
            self.z = _Poly(new_t, [start[2], start[2] + step_height, target[2] + step_height, target[2]])
        else:
            self.z = _Poly([t[0], t[0] + duration * rise_ratio, t[1]], [start[2], start[2] + step_height, target[2]])

    def _stack(self, t, nu):
        vals = [self.x(t, nu), self.y(t, nu), self.z(t, nu)]
        try:
            import torch

            if any(isinstance(v, torch.Tensor) for v in vals):
                return torch.stack([torch.as_tensor(v) for v in vals], dim=-1)
        except ImportError:
            pass
        return np.stack([np.asarray(v, dtype=np.float64) for v in vals], axis=-1)

    def get_position(self, t):
        return self._stack(t, 0)

    def get_velocity(self, t):
        return self._stack(t, 2)  # sic: the reference asks the spline for derivative order 2 (:35)

    def get_acceleration(self, t):
        return self._stack(t, 3)  # sic: order 3 (:43)

    # first and second derivatives (additions)
    def velocity(self, t):
        return self._stack(t, 1)

    def acceleration(self, t):
        return self._stack(t, 2)
