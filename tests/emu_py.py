"""ctypes driver of the host warp emulator build of the CUDA kernel header (tests/emu).
Debug/test infrastructure only: it executes tsidb_kernels.cuh's tick_env() with 32 lock-step
fibers per env so the kernel logic can be compared with the oracle on a box without a GPU."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from tsid_control_b200._capi import TsidbConf, TsidbModel

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
SRC_DIR = os.path.join(os.path.dirname(HERE), "tsid_control_b200", "csrc")


class TickArgs(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("layout", C.c_int32), ("pad_", C.c_int32),
        ("q", C.c_void_p), ("v", C.c_void_p), ("mask", C.c_void_p),
        ("r_com", C.c_void_p), ("r_foot", C.c_void_p * 2), ("r_contact", C.c_void_p * 2), ("r_posture", C.c_void_p),
        ("tau", C.c_void_p), ("ddq", C.c_void_p), ("f", C.c_void_p),
        ("status", C.c_void_p), ("iters", C.c_void_p), ("active", C.c_void_p),
        ("o_com", C.c_void_p), ("o_foot", C.c_void_p * 2), ("o_wrench", C.c_void_p), ("o_lambda", C.c_void_p), ("o_lambda_row", C.c_void_p),
        ("counter", C.c_void_p), ("ws", C.c_void_p), ("ws3", C.c_void_p), ("perm", C.c_void_p), ("kin_only", C.c_int32), ("slot", C.c_int32),
        ("pred", C.c_void_p), ("tables", C.c_void_p),
    ]


def build_emu() -> str:
    so = os.path.join(EMU_DIR, "libtsidb_emu.so")
    srcs = [os.path.join(EMU_DIR, "emu_main.cpp"), os.path.join(EMU_DIR, "emu_cuda.h")] + [
        os.path.join(SRC_DIR, f) for f in ("tsidb_kernels.cuh", "tsidb_gait.cuh", "tsidb_const.h", "tsidb_host_const.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        "-o", so, os.path.join(EMU_DIR, "emu_main.cpp")], check=True)
    return so


class Emu:
    def __init__(self, cm: TsidbModel, cc: TsidbConf, refs: dict):
        self.lib = C.CDLL(build_emu())
        self.lib.emu_tick.argtypes = [C.POINTER(TickArgs), C.c_void_p]
        self.lib.emu_fill_const.argtypes = [C.POINTER(TsidbModel), C.POINTER(TsidbConf), C.c_void_p]
        self.nb = cm.nb
        self.na, self.nv, self.nq = self.nb - 1, self.nb + 5, self.nb + 6
        post = np.zeros(23)
        post[: self.na] = refs["posture"]
        flat = np.concatenate([refs["com"], refs["foot_lf"], refs["foot_rf"], refs["contact_lf"], refs["contact_rf"], post])
        flat = np.ascontiguousarray(flat, dtype=np.float64)
        assert self.lib.emu_fill_const(C.byref(cm), C.byref(cc), flat.ctypes.data) == 0
        self.sm_per_env = self.lib.emu_sm_per_env()

    def tick(self, q, v, mask, refs: dict | None = None, layout: int = 0, kin_only: bool = False, aux: bool = True):
        """q [N,nq], v [N,nv], mask [N]; refs: optional per-env arrays [N,k] (else handle defaults)."""
        N = q.shape[0]
        na, nv = self.na, self.nv

        def prep(a, nd):
            a = np.ascontiguousarray(a, dtype=np.float64).reshape(N, nd)
            return np.ascontiguousarray(a.T) if layout else a

        qd, vd = prep(q, self.nq), prep(v, nv)
        md = np.ascontiguousarray(mask, dtype=np.uint8)
        shape = (lambda nd: (nd, N)) if layout else (lambda nd: (N, nd))
        out = {"tau": np.zeros(shape(na)), "ddq": np.zeros(shape(nv)), "f": np.zeros(shape(24)),
               "status": np.full(N, -7, np.int32), "iters": np.zeros(N, np.int32), "active": np.zeros((3, N), np.uint64),
               "com": np.zeros(shape(9)), "foot_lf": np.zeros(shape(12)), "foot_rf": np.zeros(shape(12)),
               "wrench": np.zeros(shape(12)), "lambda": np.zeros((N, 32)), "lambda_row": np.full((N, 32), -1, np.int32)}
        a = TickArgs()
        a.n_envs, a.layout, a.kin_only, a.slot = N, layout, int(kin_only), 0
        a.q, a.v, a.mask = qd.ctypes.data, vd.ctypes.data, md.ctypes.data
        keep = []
        if refs is not None:
            for key, nd, field in (("com", 9, "r_com"), ("posture", na, "r_posture")):
                arr = prep(refs[key], nd); keep.append(arr); setattr(a, field, arr.ctypes.data)
            for i, key in enumerate(("foot_lf", "foot_rf")):
                arr = prep(refs[key], 24); keep.append(arr); a.r_foot[i] = arr.ctypes.data
            for i, key in enumerate(("contact_lf", "contact_rf")):
                arr = prep(refs[key], 12); keep.append(arr); a.r_contact[i] = arr.ctypes.data
        a.tau, a.ddq, a.f = out["tau"].ctypes.data, out["ddq"].ctypes.data, out["f"].ctypes.data
        a.status, a.iters, a.active = out["status"].ctypes.data, out["iters"].ctypes.data, out["active"].ctypes.data
        if aux:
            a.o_com, a.o_wrench = out["com"].ctypes.data, out["wrench"].ctypes.data
            a.o_lambda, a.o_lambda_row = out["lambda"].ctypes.data, out["lambda_row"].ctypes.data
            a.o_foot[0], a.o_foot[1] = out["foot_lf"].ctypes.data, out["foot_rf"].ctypes.data
        sm = np.zeros(self.sm_per_env)
        rc = self.lib.emu_tick(C.byref(a), sm.ctypes.data)
        assert rc == 0, rc
        if layout:
            for k in ("tau", "ddq", "f", "com", "foot_lf", "foot_rf", "wrench"):
                out[k] = np.ascontiguousarray(out[k].T)
        out["sm"] = sm
        return out


def active_bits(words) -> list:
    """[3] uint64 words -> sorted list of set bit indices (tsidb_ci_row numbering)."""
    bits = []
    for w in range(3):
        val = int(words[w])
        for b in range(64):
            if (val >> b) & 1:
                bits.append(64 * w + b)
    return bits


class EmuGait:
    """Host build of the device gait phase machine (tsidb_gait.cuh via tests/emu): same state arrays as the
    library keeps on the device."""

    def __init__(self, n, dt, step_duration, step_length, step_height, com_height, defaults, phase0=None, vcmd=None,
                 steps=None, n_steps=None, rise_ratio=0.5):
        self.lib = C.CDLL(build_emu())
        self.steps = None if steps is None else np.ascontiguousarray(steps, np.float64)
        self.n_steps = None if n_steps is None else np.ascontiguousarray(n_steps, np.int32)
        self.step_idx = np.full(n, 2, np.int32)
        self.swing = np.zeros((n, 2, 8))
        self.rise_ratio = float(rise_ratio)
        self.n = n
        self.g = np.array([dt, step_duration, step_length, step_height, 9.80665 / com_height, defaults["com"][2]], np.float64)
        z = lambda *s: np.zeros(s, np.float64)
        self.phi, self.mask, self.vcmd, self.lipm, self.origin = z(n), np.zeros(n, np.uint8), z(n, 2), z(n, 4), z(n, 24)
        self.com, self.foot, self.contact = z(n, 9), [z(n, 24), z(n, 24)], [z(n, 12), z(n, 12)]
        self.fails = np.zeros(n, np.int32)
        d = np.concatenate([defaults["com"], defaults["foot_lf"], defaults["foot_rf"], defaults["contact_lf"], defaults["contact_rf"]]).astype(np.float64)
        p0 = None if phase0 is None else np.ascontiguousarray(phase0, np.float64)
        vc = None if vcmd is None else np.ascontiguousarray(vcmd, np.float64)
        self._call(d, p0, vc, None, None, None)

    def _call(self, d, p0, vc, fl, fr, st):
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self.lib.emu_gait.argtypes = [C.c_int] + [C.c_void_p] * 22 + [C.c_int, C.c_double]
        rc = self.lib.emu_gait(self.n, ptr(self.g), ptr(self.phi), ptr(self.mask), ptr(self.vcmd), ptr(self.lipm), ptr(self.origin),
                               ptr(self.com), ptr(self.foot[0]), ptr(self.foot[1]), ptr(self.contact[0]), ptr(self.contact[1]),
                               ptr(self.fails), ptr(d), ptr(p0), ptr(vc), ptr(fl), ptr(fr), ptr(st), ptr(self.steps), ptr(self.n_steps),
                               ptr(self.step_idx), ptr(self.swing), 0 if self.steps is None else self.steps.shape[1], self.rise_ratio)
        assert rc == 0

    def step(self, foot_now_lf, foot_now_rf, status=None):
        fl = np.ascontiguousarray(foot_now_lf, np.float64)
        fr = np.ascontiguousarray(foot_now_rf, np.float64)
        st = None if status is None else np.ascontiguousarray(status, np.int32)
        self._call(None, None, None, fl, fr, st)
