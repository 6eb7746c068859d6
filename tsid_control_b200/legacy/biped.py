"""Biped — batched, CUDA-backed counterpart of ref:legacy/biped.py (the OP3-era controller used with
``op3_conf``).  Same problem as WalkController plus the angular-momentum task (ref:legacy/biped.py:82-87),
a zero-weight CoP force task (:79-80, w_cop = 0 contributes nothing), the RIGHT foot contact inserted
first (:35-77) and no joint-bounds task (w_joint_bounds = 0, :129).  Contact switching follows
:168-212.
"""
from __future__ import annotations

import numpy as np

from ..controller_base import BatchedController
from ..tsid_mirror import Contact, Task, TrajectoryEuclidianConstant, TrajectorySE3Constant


class Biped(BatchedController):
    def __init__(self, conf, n_envs: int = 1):
        super().__init__(conf, conf.urdf, conf.srdf, conf.lf_frame_name, conf.rf_frame_name, legacy=True,
                         n_envs=n_envs, contact_order=(1, 0))
        self.q0 = q = self._standing()
        v = np.zeros(self.robot.nv)
        self.formulation.computeProblemData(0.0, q, v)
        data = self.formulation.data()
        self.LF, self.RF = self.LF_frame, self.RF_frame
        H_rf_ref = self.robot.framePosition(data, self.RF)
        H_lf_ref = self.robot.framePosition(data, self.LF)
        self._init_refs(data.com, H_lf_ref.to_vec12(), H_rf_ref.to_vec12(), q)
        from ..ctrl.WalkController import WalkController

        T = WalkController._force_generator(self)
        self.contactRF = Contact(self, "contact_rfoot", "contact_rf", 1, T)
        self.contactLF = Contact(self, "contact_lfoot", "contact_lf", 0, T)
        self.copTask = Task(self, "task-cop", "none")
        self.amTask = Task(self, "task-am", "am")
        self.comTask = Task(self, "task-com", "com")
        self.postureTask = Task(self, "task-posture", "posture")
        self.leftFootTask = Task(self, "task-left-foot", "foot_lf")
        self.rightFootTask = Task(self, "task-right-foot", "foot_rf")
        self.trajLF = TrajectorySE3Constant("traj-left-foot", H_lf_ref)
        self.trajRF = TrajectorySE3Constant("traj-right-foot", H_rf_ref)
        self.tau_max = conf.tau_max_scaling * self.model.effortLimit[-self.robot.na:]
        self.tau_min = -self.tau_max
        self.v_max = conf.v_max_scaling * self.model.velocityLimit[-self.robot.na:]
        self.v_min = -self.v_max
        self.actuationBoundsTask = Task(self, "task-actuation-bounds", "none")
        self.jointBoundsTask = Task(self, "task-joint-bounds", "none")
        self.trajCom = TrajectoryEuclidianConstant("traj_com", self.robot.com(data))
        self.sample_com = self.trajCom.computeNext()
        self.trajPosture = TrajectoryEuclidianConstant("traj_joint", q[7:])
        self.sampleLF = self.trajLF.computeNext()
        self.sample_LF_pos, self.sample_LF_vel, self.sample_LF_acc = (
            self.sampleLF.value(), self.sampleLF.derivative(), self.sampleLF.second_derivative())
        self.sampleRF = self.trajRF.computeNext()
        self.sample_RF_pos, self.sample_RF_vel, self.sample_RF_acc = (
            self.sampleRF.value(), self.sampleRF.derivative(), self.sampleRF.second_derivative())
        self.solver.resize(self.formulation.nVar, self.formulation.nEq, self.formulation.nIn)
        self.q, self.v = q, v
        self.contact_LF_active = True
        self.contact_RF_active = True

    # ref:legacy/biped.py:168-184
    def removeLeftFootContact(self):
        if self.contact_LF_active:
            H = self.robot.framePosition(self.formulation.data(), self.LF)
            self.trajLF.setReference(H)
            self.leftFootTask.setReference(self.trajLF.computeNext())
            self.formulation.removeRigidContact(self.contactLF.name)
            self.contact_LF_active = False

    def removeRightFootContact(self):
        if self.contact_RF_active:
            H = self.robot.framePosition(self.formulation.data(), self.RF)
            self.trajRF.setReference(H)
            self.rightFootTask.setReference(self.trajRF.computeNext())
            self.formulation.removeRigidContact(self.contactRF.name)
            self.contact_RF_active = False

    # ref:legacy/biped.py:186-212
    def addLeftFootContact(self):
        if not self.contact_LF_active:
            H = self.robot.framePosition(self.formulation.data(), self.LF)
            self.contactLF.setReference(H)
            self.formulation.addRigidContact(self.contactLF, self.conf.w_forceRef)
            self.contact_LF_active = True

    def addRightFootContact(self):
        if not self.contact_RF_active:
            H = self.robot.framePosition(self.formulation.data(), self.RF)
            self.contactRF.setReference(H)
            self.formulation.addRigidContact(self.contactRF, self.conf.w_forceRef)
            self.contact_RF_active = True

    # ref:legacy/biped.py:214-222 (the reference builds the array and returns nothing; here it is returned)
    def gen_footstep(self, pos, r_foot, steps, height):
        pos0 = self.robot.framePosition(self.formulation.data(), self.RF if r_foot else self.LF).translation
        traj = np.zeros((steps, 3))
        traj[:, 0] = np.linspace(pos0[0], pos[0], steps)
        traj[:, 1] = np.linspace(pos0[1], pos[1], steps)
        traj[:, 2] = [4 * height * (i / steps) * (1 - i / steps) for i in range(steps)]
        return traj

    # ref:legacy/biped.py:224-227
    def compute_capture_point(self, com, dcom, w):
        cp = com + dcom / w
        cp[2] = 0
        return cp

    # ref:legacy/biped.py:229-234
    def compute_support_polygon(self):
        data = self.formulation.data()
        rf = self.robot.framePosition(data, self.RF).translation
        lf = self.robot.framePosition(data, self.LF).translation
        return np.array([lf[:2], rf[:2]])
