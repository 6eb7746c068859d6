"""Controller configuration for robot/v1 — same attribute names and values as the
reference's ``RobotConfig`` (ref:ctrl/conf.py:5-75) so that a script written against
``conf.<name>`` keeps working.  Differences: no pinocchio / gepetto imports (neither
exists on the GPU path) and model paths resolve to the tables compiled by
``tsid_control_b200.model_compiler`` when the URDF file itself is not present.
"""
import numpy as np


class RobotConfig:
    # ---- model files (ref:ctrl/conf.py:9-15) --------------------------------------
    robot_path = "./robot/v1"
    root_urdf = f"{robot_path}/urdf"
    urdf = f"{robot_path}/urdf/robot_mod.urdf"
    pin_urdf = f"{root_urdf}/robot_mod.urdf"
    mjcf = f"{robot_path}/mujoco/scene.xml"
    srdf = f"{root_urdf}/robot.srdf"

    # sole frames: fixed joints of the URDF (ref:ctrl/conf.py:17-18)
    lf_fixed_joint = "left_sole_joint_fixed"
    rf_fixed_joint = "right_sole_joint_fixed"

    # ---- timing (ref:ctrl/conf.py:21) ---------------------------------------------
    dt = 0.002

    # ---- gait (ref:ctrl/conf.py:24-28) --------------------------------------------
    step_height = 0.2
    step_width = 0.2
    step_length = 0.3
    step_duration = 0.5
    rise_ratio = 0.5

    # ---- foot rectangle in the sole frame (ref:ctrl/conf.py:31-35) ----------------
    lxn = 0.055
    lyn = 0.0275
    lxp = 0.055
    lyp = 0.0275
    lz = 0.0

    # ---- Contact6d (ref:ctrl/conf.py:38-44) ---------------------------------------
    mu = 0.5
    fMin = 10.0
    fMax = 1000.0
    contactNormal = np.array([0.0, 0.0, 1.0])
    w_contact = -1.0  # < 0: contact motion is a hard constraint (2-argument addRigidContact)
    w_forceRef = 1e-5
    kp_contact = 10.0

    # ---- swing / stance foot SE3 task (ref:ctrl/conf.py:47-48) --------------------
    w_foot = 1e-1
    kp_foot = 10.0

    # ---- centre of mass (ref:ctrl/conf.py:51-52) ----------------------------------
    w_com = 1e-1
    kp_com = 10.0

    # ---- posture (ref:ctrl/conf.py:55-66) -----------------------------------------
    w_posture = 1e-1
    kp_posture = 10.0
    gain_vector = np.array(
        [100.0, 100.0]  # head yaw, pitch
        + [10.0, 5.0, 5.0, 1.0, 1.0, 1.0]  # left hip yaw/roll/pitch, knee, ankle pitch/roll
        + [10.0, 10.0, 10.0]  # left shoulder pitch/roll, elbow
        + [10.0, 5.0, 5.0, 1.0, 1.0, 1.0]  # right leg
        + [10.0, 10.0, 10.0]  # right arm
    )
    masks_posture = np.ones(20)

    # ---- bounds (ref:ctrl/conf.py:69-72) ------------------------------------------
    tau_max_scaling = 5.0
    v_max_scaling = 10.0
    w_torque_bounds = 1e-2
    w_joint_bounds = 1e-2

    # the reference marks the Gepetto viewer as not working (ref:ctrl/conf.py:74-75)
    visualizer = None

    # ---- additions of this implementation (not in the reference) ------------------
    device = 0  # CUDA device of the handle
    max_envs = 65536  # workspace size of the handle
