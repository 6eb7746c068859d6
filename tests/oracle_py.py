"""ctypes wrapper of oracle/liboracle*.so — the CPU checker (test infrastructure only).

Nothing under tsid_control_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Sequence

import numpy as np

from tsid_control_b200._capi import MAX_BODIES, MAX_NA, MAX_NV, TsidbConf, TsidbModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

NMAX = MAX_NV + 24
NINMAX = 2 * (34 + MAX_NA + MAX_NV)

CI_FORCE_LF, CI_FORCE_RF, CI_ACTUATION, CI_JOINT_BOUNDS = 0, 1, 2, 3
T_FORCEREG_LF, T_FORCEREG_RF, T_FOOT_LF, T_FOOT_RF, T_COM, T_POSTURE, T_AM = range(7)

dp = C.POINTER(C.c_double)


class OracleProblem(C.Structure):
    _fields_ = [
        ("q", dp), ("v", dp),
        ("nc", C.c_int32), ("contact_order", C.c_int32 * 2),
        ("n_ci_blocks", C.c_int32), ("ci_order", C.c_int32 * 4),
        ("n_cost", C.c_int32), ("cost_order", C.c_int32 * 8),
        ("ref_com", dp), ("ref_foot", dp * 2), ("ref_contact", dp * 2), ("ref_posture", dp),
    ]


class OracleResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("iters", C.c_int32), ("n", C.c_int32), ("n_active", C.c_int32),
        ("active", C.c_int32 * (NMAX + 1)),
        ("tau", C.c_double * MAX_NA), ("dv", C.c_double * MAX_NV), ("f", C.c_double * 24),
        ("x", C.c_double * NMAX), ("lam", C.c_double * (NMAX + 19)),
        ("com", C.c_double * 9), ("foot", (C.c_double * 12) * 2),
    ]


class OracleDump(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("neq", C.c_int32), ("nin", C.c_int32), ("nv", C.c_int32),
        ("M", C.c_double * (MAX_NV * MAX_NV)), ("nle", C.c_double * MAX_NV),
        ("JF", (C.c_double * (6 * MAX_NV)) * 2), ("vF", (C.c_double * 6) * 2), ("aF", (C.c_double * 6) * 2),
        ("Jcom", C.c_double * (3 * MAX_NV)), ("Ag", C.c_double * (6 * MAX_NV)), ("dAg_v_ang", C.c_double * 3),
        ("H", C.c_double * (NMAX * NMAX)), ("g", C.c_double * NMAX),
        ("CE", C.c_double * (18 * NMAX)), ("ce0", C.c_double * 18),
        ("CI", C.c_double * (NINMAX * NMAX)), ("ci0", C.c_double * NINMAX),
        ("oMi_R", (C.c_double * 9) * MAX_BODIES), ("oMi_p", (C.c_double * 3) * MAX_BODIES),
    ]


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_LIBS: Dict[str, C.CDLL] = {}


def _lib(name: str) -> C.CDLL:
    if name not in _LIBS:
        path = os.path.join(ORACLE_DIR, name)
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.oracle_tick.argtypes = [C.POINTER(TsidbModel), C.POINTER(TsidbConf), C.POINTER(OracleProblem),
                                    C.POINTER(OracleResult), C.POINTER(OracleDump)]
        lib.oracle_tick.restype = C.c_int
        lib.oracle_tick_batch.argtypes = [C.POINTER(TsidbModel), C.POINTER(TsidbConf), C.POINTER(OracleProblem),
                                          C.POINTER(OracleResult), C.c_int, C.c_int]
        lib.oracle_tick_batch.restype = C.c_int
        lib.oracle_integrate.argtypes = [C.POINTER(TsidbModel), dp, dp, dp, C.c_double]
        lib.oracle_integrate.restype = C.c_int
        lib.oracle_real_bytes.restype = C.c_int
        _LIBS[name] = lib
    return _LIBS[name]


def walkcontroller_orders(mask: int):
    """Level-0 inequality and level-1 cost insertion order of a freshly constructed
    WalkController (ref:ctrl/WalkController.py:83-184) restricted to the active contacts."""
    contacts = [f for f in (0, 1) if mask & (1 << f)]
    ci = [b for b in (CI_FORCE_LF, CI_FORCE_RF) if mask & (1 << b)] + [CI_ACTUATION, CI_JOINT_BOUNDS]
    cost = []
    if mask & 1:
        cost.append(T_FORCEREG_LF)
    cost.append(T_FOOT_LF)
    if mask & 2:
        cost.append(T_FORCEREG_RF)
    cost += [T_FOOT_RF, T_COM, T_POSTURE]
    return contacts, ci, cost


def biped_orders(mask: int):
    """Same for legacy Biped (ref:legacy/biped.py:35-130): RF contact first, AM/CoM/posture,
    then LF/RF foot tasks; actuation bounds last; no joint-bounds task (w_joint_bounds = 0)."""
    contacts = [f for f in (1, 0) if mask & (1 << f)]
    ci = [b for b in (CI_FORCE_RF, CI_FORCE_LF) if mask & (1 << b)] + [CI_ACTUATION]
    cost = [t for t, f in ((T_FORCEREG_RF, 1), (T_FORCEREG_LF, 0)) if mask & (1 << f)]
    cost += [T_AM, T_COM, T_POSTURE, T_FOOT_LF, T_FOOT_RF]
    return contacts, ci, cost


class Oracle:
    def __init__(self, cmodel: TsidbModel, cconf: TsidbConf, variant: str = "liboracle.so"):
        self.lib = _lib(variant)
        self.cm, self.cc = cmodel, cconf
        self.nb = cmodel.nb
        self.na, self.nv, self.nq = self.nb - 1, self.nb + 5, self.nb + 6
        self.legacy = cconf.w_am > 0.0

    def orders(self, mask: int):
        o = biped_orders(mask) if self.legacy else walkcontroller_orders(mask)
        contacts, ci, cost = o
        if not self.cc.use_torque_bounds:
            ci = [b for b in ci if b != CI_ACTUATION]
        if not self.cc.use_joint_bounds:
            ci = [b for b in ci if b != CI_JOINT_BOUNDS]
        return contacts, ci, cost

    def _problem(self, keep, q, v, mask, refs, orders=None):
        q = np.ascontiguousarray(q, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        contacts, ci, cost = orders if orders is not None else self.orders(mask)
        pb = OracleProblem()
        arrs = {k: np.ascontiguousarray(refs[k], dtype=np.float64) for k in
                ("com", "foot_lf", "foot_rf", "contact_lf", "contact_rf", "posture")}
        keep.extend([q, v, arrs])
        pb.q = q.ctypes.data_as(dp)
        pb.v = v.ctypes.data_as(dp)
        pb.nc = len(contacts)
        for i, f in enumerate(contacts):
            pb.contact_order[i] = f
        pb.n_ci_blocks = len(ci)
        for i, b in enumerate(ci):
            pb.ci_order[i] = b
        pb.n_cost = len(cost)
        for i, t in enumerate(cost):
            pb.cost_order[i] = t
        pb.ref_com = arrs["com"].ctypes.data_as(dp)
        pb.ref_foot[0] = arrs["foot_lf"].ctypes.data_as(dp)
        pb.ref_foot[1] = arrs["foot_rf"].ctypes.data_as(dp)
        pb.ref_contact[0] = arrs["contact_lf"].ctypes.data_as(dp)
        pb.ref_contact[1] = arrs["contact_rf"].ctypes.data_as(dp)
        pb.ref_posture = arrs["posture"].ctypes.data_as(dp)
        return pb

    def tick(self, q, v, mask: int, refs: dict, dump: bool = False, orders=None) -> dict:
        keep: list = []
        pb = self._problem(keep, q, v, mask, refs, orders)
        res = OracleResult()
        dmp = OracleDump() if dump else None
        rc = self.lib.oracle_tick(C.byref(self.cm), C.byref(self.cc), C.byref(pb), C.byref(res),
                                  C.byref(dmp) if dump else None)
        assert rc == 0
        out = self._unpack(res)
        if dump:
            out["dump"] = self._unpack_dump(dmp)
        return out

    def _unpack(self, res: OracleResult) -> dict:
        na, nv = self.na, self.nv
        return {
            "status": res.status, "iters": res.iters, "n": res.n,
            "active": np.array(res.active[: res.n_active], dtype=np.int64),
            "tau": np.array(res.tau[:na]), "dv": np.array(res.dv[:nv]), "f": np.array(res.f[:24]),
            "x": np.array(res.x[: res.n]), "lam": np.array(res.lam[: NMAX + 19]),
            "com": np.array(res.com[:9]),
            "foot": np.array([list(res.foot[0]), list(res.foot[1])]),
        }

    def _unpack_dump(self, d: OracleDump) -> dict:
        n, neq, nin, nv = d.n, d.neq, d.nin, d.nv
        return {
            "n": n, "neq": neq, "nin": nin,
            "M": np.array(d.M[: nv * nv]).reshape(nv, nv), "nle": np.array(d.nle[:nv]),
            "JF": np.array([np.array(d.JF[s][: 6 * nv]).reshape(6, nv) for s in range(2)]),
            "vF": np.array([list(d.vF[0]), list(d.vF[1])]), "aF": np.array([list(d.aF[0]), list(d.aF[1])]),
            "Jcom": np.array(d.Jcom[: 3 * nv]).reshape(3, nv), "Ag": np.array(d.Ag[: 6 * nv]).reshape(6, nv),
            "dAg_v_ang": np.array(d.dAg_v_ang[:3]),
            "H": np.array(d.H[: n * n]).reshape(n, n), "g": np.array(d.g[:n]),
            "CE": np.array(d.CE[: neq * n]).reshape(neq, n), "ce0": np.array(d.ce0[:neq]),
            "CI": np.array(d.CI[: nin * n]).reshape(nin, n), "ci0": np.array(d.ci0[:nin]),
            "oMi_R": np.array([list(d.oMi_R[b]) for b in range(self.nb)]).reshape(self.nb, 3, 3),
            "oMi_p": np.array([list(d.oMi_p[b]) for b in range(self.nb)]),
        }

    def batch(self, q, v, mask, refs: dict, n_threads: int = 1) -> dict:
        """q [N,nq], v [N,nv], mask [N] uint8, refs: dict of [N,k] arrays (or [k] broadcast)."""
        N = q.shape[0]
        keep: list = []
        pbs = (OracleProblem * N)()
        for i in range(N):
            r = {k: (a[i] if np.ndim(a) == 2 else a) for k, a in refs.items()}
            pbs[i] = self._problem(keep, q[i], v[i], int(mask[i]), r)
        res = (OracleResult * N)()
        rc = self.lib.oracle_tick_batch(C.byref(self.cm), C.byref(self.cc), pbs, res, N, n_threads)
        assert rc == 0
        outs = [self._unpack(res[i]) for i in range(N)]
        return {
            "status": np.array([o["status"] for o in outs]), "iters": np.array([o["iters"] for o in outs]),
            "tau": np.array([o["tau"] for o in outs]), "dv": np.array([o["dv"] for o in outs]),
            "f": np.array([o["f"] for o in outs]), "active": [o["active"] for o in outs],
            "com": np.array([o["com"] for o in outs]), "foot": np.array([o["foot"] for o in outs]),
            "lam": [o["lam"] for o in outs],  # sol.lambda: [equalities (6 + 6 nc); active inequalities in working-set order]
        }

    def timed_batch(self, q, v, mask, refs: dict, n_threads: int):
        """Build the problem array once, then return a zero-argument callable that runs the batch
        (so that bench.py times the C code, not the Python marshalling)."""
        N = q.shape[0]
        keep: list = []
        pbs = (OracleProblem * N)()
        for i in range(N):
            r = {k: (a[i] if np.ndim(a) == 2 else a) for k, a in refs.items()}
            pbs[i] = self._problem(keep, q[i], v[i], int(mask[i]), r)
        res = (OracleResult * N)()

        def run():
            rc = self.lib.oracle_tick_batch(C.byref(self.cm), C.byref(self.cc), pbs, res, N, n_threads)
            assert rc == 0
            return res

        run._keep = (keep, pbs, res)
        return run

    def integrate(self, q, v, dv, dt: float):
        q = np.array(q, dtype=np.float64)
        v = np.array(v, dtype=np.float64)
        dv = np.ascontiguousarray(dv, dtype=np.float64)
        self.lib.oracle_integrate(C.byref(self.cm), q.ctypes.data_as(dp), v.ctypes.data_as(dp),
                                  dv.ctypes.data_as(dp), float(dt))
        return q, v

    # reference CI row index (as SolverHQuadProgFast stacks it) -> (block id, side, i)
    def ci_rows(self, mask: int, orders=None):
        _, ci, _ = orders if orders is not None else self.orders(mask)
        rows = []
        for b in ci:
            nr = 17 if b in (CI_FORCE_LF, CI_FORCE_RF) else (self.na if b == CI_ACTUATION else self.nv)
            for side in (0, 1):
                for i in range(nr):
                    rows.append((b, side, i))
        return rows
