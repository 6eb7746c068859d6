"""WalkController — batched, CUDA-backed counterpart of ref:ctrl/WalkController.py.

Same constructor argument (``RobotConfig``), same attribute and method names; the TSID problem it
describes (2 x Contact6d, 2 x TaskSE3Equality, CoM, posture, actuation and joint bounds,
ref:ctrl/WalkController.py:55-187) is solved for ``n_envs`` robots per call by libtsidb.so.
``compute(q, v, t) -> (tau, ddq, f_contact)`` is the batched form of the reference's tick
(ref:main.py:119-127).
"""
from __future__ import annotations

import numpy as np
import torch

from ..controller_base import BatchedController
from ..tsid_mirror import SE3, Contact, Task, TrajectoryEuclidianConstant, TrajectorySE3Constant
from .conf import RobotConfig


class WalkController(BatchedController):
    def __init__(self, conf: RobotConfig, n_envs: int = 1):
        # ref:ctrl/WalkController.py:13-25 — model, formulation, standing configuration
        super().__init__(conf, conf.urdf, conf.srdf, conf.lf_fixed_joint, conf.rf_fixed_joint, legacy=False,
                         n_envs=n_envs, contact_order=(0, 1))
        self.q0 = self.q = self._standing()  # aliased on purpose, as in the reference (:23)
        self.v = np.zeros(self.robot.nv)
        self.formulation.computeProblemData(0.0, self.q, self.v)
        self.data = self.formulation.data()

        # ref:ctrl/WalkController.py:72-79 — put the left sole on z = 0, then take the references
        H_lf_ref = self.robot.framePosition(self.data, self.LF_frame)
        self.q[2] -= H_lf_ref.translation[2]
        self.formulation.computeProblemData(0.0, self.q, self.v)
        self.data = self.formulation.data()
        H_lf_ref = self.robot.framePosition(self.data, self.LF_frame)
        H_rf_ref = self.robot.framePosition(self.data, self.RF_frame)
        self._init_refs(self.data.com, H_lf_ref.to_vec12(), H_rf_ref.to_vec12(), self.q)

        T = self._force_generator()
        # contacts (ref:ctrl/WalkController.py:59-86, 107-128)
        self.contactLF = Contact(self, "contact_lfoot", "contact_lf", 0, T)
        self.contactRF = Contact(self, "contact_rfoot", "contact_rf", 1, T)
        self.contactLF_active = True
        self.contactRF_active = True
        # foot tasks (ref:ctrl/WalkController.py:90-104, 131-144)
        self.task_LF = Task(self, "task_lfoot", "foot_lf")
        self.task_RF = Task(self, "task_rfoot", "foot_rf")
        self.traj_LF = TrajectorySE3Constant("traj_lfoot", H_lf_ref)
        self.traj_RF = TrajectorySE3Constant("traj_rfoot", H_rf_ref)
        # the reference's remove_contact() names these two (ref:ctrl/WalkController.py:222,229)
        self.leftFootTask, self.rightFootTask = self.task_LF, self.task_RF
        # CoM and posture (ref:ctrl/WalkController.py:147-165)
        self.comTask = Task(self, "task-com", "com")
        self.traj_COM = TrajectoryEuclidianConstant("traj-com", self.robot.com(self.data))
        self.postureTask = Task(self, "task-posture", "posture")
        self.traj_posture = TrajectoryEuclidianConstant("traj-posture", self.q[7:])
        # bounds (ref:ctrl/WalkController.py:167-184)
        self.tau_max = conf.tau_max_scaling * self.model.effortLimit[-self.robot.na:]
        self.tau_min = -self.tau_max
        self.v_max = conf.v_max_scaling * self.model.velocityLimit[-self.robot.na:]
        self.v_min = -self.v_max
        self.actuationBoundsTask = Task(self, "task-actuation-bounds", "none")
        self.jointBoundsTask = Task(self, "task-joint-bounds", "none")
        self.solver.resize(self.formulation.nVar, self.formulation.nEq, self.formulation.nIn)

    def _force_generator(self) -> np.ndarray:
        c = self.engine.cc
        T = np.zeros((6, 12))
        for i in range(4):
            p = np.array([c.contact_points[0][i], c.contact_points[1][i], c.contact_points[2][i]])
            T[:3, 3 * i:3 * i + 3] = np.eye(3)
            T[3:, 3 * i:3 * i + 3] = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
        return T

    # ref:ctrl/WalkController.py:189-206
    def update_tasks(self, sampleLF, sampleRF, contact_LF, contact_RF):
        self.task_LF.setReference(sampleLF)
        self.task_RF.setReference(sampleRF)
        if isinstance(contact_LF, torch.Tensor) or isinstance(contact_RF, torch.Tensor):
            cl = contact_LF if isinstance(contact_LF, torch.Tensor) else torch.full((self.n_envs,), bool(contact_LF), device=self.device)
            cr = contact_RF if isinstance(contact_RF, torch.Tensor) else torch.full((self.n_envs,), bool(contact_RF), device=self.device)
            self.set_contact_phase(cl, cr)
            return
        if contact_LF and not self.contactLF_active:
            self.add_contact(left_foot=True, right_foot=False)
        elif not contact_LF and self.contactLF_active:
            self.remove_contact(left_foot=True, right_foot=False)
        if contact_RF and not self.contactRF_active:
            self.add_contact(left_foot=False, right_foot=True)
        elif not contact_RF and self.contactRF_active:
            self.remove_contact(left_foot=False, right_foot=True)

    # ref:ctrl/WalkController.py:215-232 with the working semantics of ref:legacy/biped.py:168-184
    # (the reference method names attributes that do not exist and would raise on first use)
    def remove_contact(self, left_foot=True, right_foot=True):
        if left_foot and self.contactLF_active:
            T_lf = self.robot.framePosition(self.formulation.data(), self.LF_frame)
            self.traj_LF.setReference(T_lf)
            self.leftFootTask.setReference(self.traj_LF.computeNext())
            self.formulation.removeRigidContact(self.contactLF.name)
            self.contactLF_active = False
        if right_foot and self.contactRF_active:
            T_rf = self.robot.framePosition(self.formulation.data(), self.RF_frame)
            self.traj_RF.setReference(T_rf)
            self.rightFootTask.setReference(self.traj_RF.computeNext())
            self.formulation.removeRigidContact(self.contactRF.name)
            self.contactRF_active = False

    # ref:ctrl/WalkController.py:234-253 / ref:legacy/biped.py:186-212
    def add_contact(self, left_foot=True, right_foot=True):
        if left_foot and not self.contactLF_active:
            T_lf = self.robot.framePosition(self.formulation.data(), self.LF_frame)
            self.contactLF.setReference(T_lf)
            self.formulation.addRigidContact(self.contactLF, self.conf.w_forceRef)
            self.contactLF_active = True
        if right_foot and not self.contactRF_active:
            T_rf = self.robot.framePosition(self.formulation.data(), self.RF_frame)
            self.contactRF.setReference(T_rf)
            self.formulation.addRigidContact(self.contactRF, self.conf.w_forceRef)
            self.contactRF_active = True

    # ref:ctrl/WalkController.py:255-289 (single-robot view)
    def get_cop(self, sol):
        data = self.formulation.data()
        tot, acc = 0.0, np.zeros(2)
        if not (self.contactLF_active and self.contactRF_active):
            return None  # the reference only returns a CoP in double support (:284)
        for active, contact, frame in ((self.contactLF_active, self.contactLF, self.LF_frame),
                                       (self.contactRF_active, self.contactRF, self.RF_frame)):
            w = contact.getForceGeneratorMatrix.dot(self.formulation.getContactForce(contact.name, sol))
            cop = np.zeros(3)
            if w[2] > 1e-3:
                cop = np.array([w[4] / w[2], w[3] / w[2], 0.0])
            world = self.robot.framePosition(data, frame).act(cop)
            acc += world[:2] * w[2]
            tot += w[2]
        return np.append(acc / tot, 0.0)
