"""Shared host logic of WalkController (ref:ctrl/WalkController.py) and Biped (ref:legacy/biped.py):
per-env references and contact phases held as CUDA tensors, the batched tick, and the bookkeeping
that mirrors how the reference's formulation orders contacts and constraint blocks.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import REF_KEYS, TickOutput, TsidEngine
from .model_compiler import load_model
from .tsid_mirror import (SE3, Contact, FormulationMirror, ModelMirror, RobotWrapperMirror, SolverMirror, Task,
                          TrajectoryEuclidianConstant, TrajectorySE3Constant, TrajectorySample)

BLK_FORCE_LF, BLK_FORCE_RF, BLK_ACT, BLK_JB = 0, 1, 2, 3


class BatchedController:
    def __init__(self, conf, urdf: str, srdf: str, lf_frame: str, rf_frame: str, legacy: bool, n_envs: int,
                 contact_order: Sequence[int]):
        self.conf = conf
        self.n_envs = int(n_envs)
        compiled = load_model(urdf, srdf)
        self.compiled = compiled
        self.engine = TsidEngine(compiled, conf, lf_frame, rf_frame, legacy=legacy,
                                 max_envs=max(self.n_envs, int(getattr(conf, "max_envs", 1))),
                                 device=int(getattr(conf, "device", 0)))
        self.device = self.engine.device
        self.model = ModelMirror(compiled, [lf_frame, rf_frame])
        self.robot = RobotWrapperMirror(self)
        self.formulation = FormulationMirror(self)
        self.solver = SolverMirror(self)
        self.LF_frame, self.RF_frame = 0, 1
        # formulation bookkeeping of the single-robot view (env 0): x order and level-0 block order
        self._contact_order: List[int] = list(contact_order)
        e = self.engine
        self._ci_order: List[int] = [BLK_FORCE_LF + f for f in contact_order]
        if e.cc.use_torque_bounds:
            self._ci_order.append(BLK_ACT)
        if e.cc.use_joint_bounds:
            self._ci_order.append(BLK_JB)
        self._contact_names = {0: "contact_lfoot", 1: "contact_rfoot"}
        self.refs: Dict[str, torch.Tensor] = {}
        self.contact_mask = torch.full((self.n_envs,), 3, dtype=torch.uint8, device=self.device)
        self.last: Optional[TickOutput] = None

    # ------------------------------------------------------------------ construction helpers
    def _standing(self) -> np.ndarray:
        return self.model.referenceConfigurations["standing"].copy()

    def _kin1(self, q: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        qd = torch.as_tensor(q.reshape(1, -1), device=self.device)
        com, lf, rf = self.engine.kinematics(qd)
        return com[0].cpu().numpy(), lf[0].cpu().numpy(), rf[0].cpu().numpy()

    def _init_refs(self, com: np.ndarray, lf12: np.ndarray, rf12: np.ndarray, q: np.ndarray) -> None:
        z12 = np.zeros(12)
        com9 = np.zeros(9)
        com9[:3] = com[:3]
        d = {"com": com9, "foot_lf": np.r_[lf12, z12], "foot_rf": np.r_[rf12, z12], "contact_lf": lf12.copy(),
             "contact_rf": rf12.copy(), "posture": q[7:].copy()}
        self.default_refs = d
        self.engine.set_default_refs(d)
        self.refs = {k: torch.as_tensor(np.tile(v, (self.n_envs, 1)), device=self.device).contiguous() for k, v in d.items()}

    # ------------------------------------------------------------------ the tick
    def _tick(self, q: torch.Tensor, v: torch.Tensor, env_slice: Optional[slice] = None, aux: bool = False) -> TickOutput:
        if env_slice is None:
            refs, mask = self.refs, self.contact_mask
        else:
            refs = {k: t[env_slice].contiguous() for k, t in self.refs.items()}
            mask = self.contact_mask[env_slice].contiguous()
        self.last = self.engine.compute(q, v, mask, refs, aux=aux)
        return self.last

    def compute(self, q: torch.Tensor, v: torch.Tensor, t: float = 0.0):
        """The batched drop-in of ref:main.py:119-127: (tau [N,na], ddq [N,nv], f_contact [N,24]).
        f_contact holds the LF corner forces in 0..11 and the RF ones in 12..23 (zeros for a foot not in
        contact); per-env solver status / iterations / active set are in ``self.last``."""
        if q.shape[0] != self.n_envs:
            raise ValueError(f"compute: controller was built for {self.n_envs} envs, got {q.shape[0]}")
        # aux outputs on: CoM and sole placements of this tick are what the reference reads from formulation.data()
        # after the solve (ref:main.py:135-142) and what update_tasks / set_contact_phase need at a contact switch
        out = self._tick(q, v, aux=True)
        return out.tau, out.ddq, out.f

    def rollout(self, q: torch.Tensor, v: torch.Tensor, n_steps: int, phase0: Optional[torch.Tensor] = None,
                vcmd: Optional[torch.Tensor] = None, restart: bool = False, use_graph: bool = True):
        """Closed-loop batched form of the reference's loop (ref:main.py:110-128) with the walking references it
        never wires in (ref:main.py:117): n_steps x {tick, integrate_dv, gait phase machine} on the device, no host
        round trip.  q [N,nq] and v [N,nv] are advanced in place; returns (tau, ddq, f_contact) of the last step.
        The gait is (re)started from the controller's default references on the first call or with restart=True,
        with per-env phase0 [N] in [0,1) and velocity command vcmd [N,2]; gait parameters come from conf
        (dt, step_duration, step_length, step_height; LIPM height = reference CoM height)."""
        if restart or not getattr(self, "_gait_started", False):
            c = self.conf
            self.engine.gait_reset(self.n_envs, float(c.dt), float(getattr(c, "step_duration", 0.5)),
                                   float(getattr(c, "step_length", 0.1)), float(getattr(c, "step_height", 0.05)),
                                   float(self.default_refs["com"][2]), phase0, vcmd)
            self._gait_started = True
        self.last = self.engine.rollout(q, v, n_steps, use_graph=use_graph)
        return self.last.tau, self.last.ddq, self.last.f

    # ------------------------------------------------------------------ references
    def _set_task_reference(self, key: str, sample) -> None:
        if key == "am":
            return  # TaskAMEquality reference is the zero sample in the reference (ref:legacy/biped.py:86-87)
        pos, vel, acc = sample.value(), sample.derivative(), sample.second_derivative()

        def dev(x, nd):
            t = torch.as_tensor(x, dtype=torch.float64, device=self.device)
            if t.dim() == 1:
                t = t.unsqueeze(0).expand(self.n_envs, nd)
            return t

        if key in ("foot_lf", "foot_rf"):
            self.refs[key][:, :12] = dev(pos, 12)
            self.refs[key][:, 12:18] = dev(vel, 6)
            self.refs[key][:, 18:24] = dev(acc, 6)
        elif key == "com":
            self.refs[key][:, 0:3] = dev(pos, 3)
            self.refs[key][:, 3:6] = dev(vel, 3)
            self.refs[key][:, 6:9] = dev(acc, 3)
        elif key in ("contact_lf", "contact_rf"):
            self.refs[key][:] = dev(pos, 12)
        elif key == "posture":
            self.refs[key][:] = dev(pos, self.engine.na)
        else:
            raise KeyError(key)

    # ------------------------------------------------------------------ contacts
    def _contact_foot(self, name: str) -> int:
        for f, n in self._contact_names.items():
            if n == name:
                return f
        raise KeyError(name)

    def _remove_contact_by_name(self, name: str) -> bool:
        f = self._contact_foot(name)
        if f in self._contact_order:
            self._contact_order.remove(f)
            self._ci_order.remove(BLK_FORCE_LF + f)
        self.contact_mask &= ~torch.tensor(1 << f, dtype=torch.uint8, device=self.device)
        return True

    def _add_contact_by_name(self, name: str) -> bool:
        f = self._contact_foot(name)
        if f not in self._contact_order:
            # [UPSTREAM addRigidContact] a re-added contact goes to the END of x and of level 0
            self._contact_order.append(f)
            self._ci_order.append(BLK_FORCE_LF + f)
        self.contact_mask |= torch.tensor(1 << f, dtype=torch.uint8, device=self.device)
        return True

    def set_contact_phase(self, contact_LF: torch.Tensor, contact_RF: torch.Tensor, foot_lf_now: Optional[torch.Tensor] = None,
                          foot_rf_now: Optional[torch.Tensor] = None) -> None:
        """Per-env contact switching with the legacy semantics (ref:legacy/biped.py:168-212): on lift-off the
        foot task reference becomes the current placement; on touch-down the contact reference becomes the
        current placement.  contact_* are bool [N] tensors; foot_*_now are [N,12] current placements
        (taken from the last tick's aux outputs when omitted)."""
        if foot_lf_now is None or foot_rf_now is None:
            if self.last is None or self.last.foot_lf is None:
                raise RuntimeError("set_contact_phase needs current foot placements: run a tick with aux or pass them")
            foot_lf_now, foot_rf_now = self.last.foot_lf, self.last.foot_rf
        for f, want, now, ck, fk in ((0, contact_LF, foot_lf_now, "contact_lf", "foot_lf"),
                                     (1, contact_RF, foot_rf_now, "contact_rf", "foot_rf")):
            bit = 1 << f
            have = (self.contact_mask & bit) != 0
            want = want.to(torch.bool)
            lift = have & ~want
            land = ~have & want
            if bool(lift.any()):
                self.refs[fk][lift, :12] = now[lift]
                self.refs[fk][lift, 12:] = 0.0
            if bool(land.any()):
                self.refs[ck][land] = now[land]
            self.contact_mask = torch.where(want, self.contact_mask | bit, self.contact_mask & ~torch.tensor(bit, dtype=torch.uint8, device=self.device)).to(torch.uint8)

    # ------------------------------------------------------------------ reference-order bookkeeping
    def _block_rows(self, blk: int) -> int:
        e = self.engine
        return 17 if blk in (BLK_FORCE_LF, BLK_FORCE_RF) else (e.na if blk == BLK_ACT else e.nv)

    def _bits_to_reference_rows(self, bits: Sequence[int]) -> List[int]:
        """Canonical active-set bits (tsidb_ci_row numbering) -> indices into the reference's stacked CI for the
        current level-0 block order of the single-robot view."""
        e = self.engine
        canon = {}
        for blk in range(4):
            for side in (0, 1):
                for i in range(self._block_rows(blk)):
                    canon[e.ci_row(blk, side, i)] = (blk, side, i)
        ref_off, off = {}, 0
        for blk in self._ci_order:
            ref_off[blk] = off
            off += 2 * self._block_rows(blk)
        out = []
        for b in bits:
            blk, side, i = canon[b]
            out.append(ref_off[blk] + side * self._block_rows(blk) + i)
        return sorted(out)

    # ------------------------------------------------------------------ integrate / display / CoP
    def display(self, q) -> None:
        pass  # the reference's viewer is disabled (conf.visualizer = None, ref:ctrl/conf.py:74-75)

    def integrate_dv(self, q, v, dv, dt):
        """ref:ctrl/WalkController.py:291-295 / ref:legacy/biped.py:236-240.  numpy [nq]/[nv] for one robot (v is
        updated in place like the reference does) or CUDA tensors [N,*] (both updated in place)."""
        if isinstance(q, torch.Tensor):
            self.engine.integrate(q, v, dv, dt)
            return q, v
        qd = torch.as_tensor(np.asarray(q, dtype=np.float64).reshape(1, -1), device=self.device).clone()
        vd = torch.as_tensor(np.asarray(v, dtype=np.float64).reshape(1, -1), device=self.device).clone()
        dvd = torch.as_tensor(np.asarray(dv, dtype=np.float64).reshape(1, -1), device=self.device).contiguous()
        self.engine.integrate(qd, vd, dvd, dt)
        v[:] = vd[0].cpu().numpy()
        return qd[0].cpu().numpy(), v

    def cop_batch(self, out: Optional[TickOutput] = None) -> torch.Tensor:
        """Batched centre of pressure [N,3] with the formulas of ref:ctrl/WalkController.py:255-289:
        cop_local = (w[4]/w[2], w[3]/w[2], 0) from each foot's wrench w = T f when w[2] > 1e-3, mapped to
        the world by the sole placement and force-weighted over the feet in contact.  (The reference raises
        with a single contact because it reads f_rf unconditionally; here a single contact returns that
        foot's CoP.)"""
        out = out or self.last
        if out is None or out.wrench is None:
            raise RuntimeError("cop_batch needs a tick computed with aux outputs")
        N = out.wrench.shape[0]
        num = torch.zeros((N, 2), dtype=torch.float64, device=self.device)
        den = torch.zeros((N,), dtype=torch.float64, device=self.device)
        for f, foot in ((0, out.foot_lf), (1, out.foot_rf)):
            w = out.wrench[:, 6 * f:6 * f + 6]
            on = ((self.contact_mask >> f) & 1).to(torch.bool)
            fz = w[:, 2]
            ok = fz > 1e-3
            safe = torch.where(ok, fz, torch.ones_like(fz))
            loc = torch.stack([torch.where(ok, w[:, 4] / safe, torch.zeros_like(fz)),
                               torch.where(ok, w[:, 3] / safe, torch.zeros_like(fz)), torch.zeros_like(fz)], dim=1)
            R = foot[:, 3:].reshape(N, 3, 3).transpose(1, 2)
            world = torch.einsum("nij,nj->ni", R, loc) + foot[:, :3]
            wgt = torch.where(on, fz, torch.zeros_like(fz))
            num += world[:, :2] * wgt[:, None]
            den += wgt
        cop = num / torch.where(den != 0, den, torch.ones_like(den))[:, None]
        return torch.cat([cop, torch.zeros((N, 1), dtype=torch.float64, device=self.device)], dim=1)
