"""Legacy OP3-era configuration — same module-level names and values as
ref:legacy/op3_conf.py:1-62 (consumed by ``Biped(conf)``, ref:legacy/biped.py:7).
The stale ``./robot/robot.urdf`` paths of the reference are kept; the model loader maps
them to the robot/v0 tables (18 joints, ``leg_*_sole_joint_fixed`` frames), which is the
model these constants were written for.
"""
import numpy as np

N_SIMULATION = 500
dt = 0.002
g = 9.81
z0 = 0.4

# gait (ref:legacy/op3_conf.py:9-12)
step_length = 0.1
step_height = 0.05
step_width = 0.1275
step_time = 0.7

# task weights (ref:legacy/op3_conf.py:14-22)
w_com = 1.0
w_am = 1e-3
w_foot = 1e-1
w_contact = -1.0
w_posture = 1e-1
w_forceRef = 1e-5
w_cop = 0.0
w_torque_bounds = 1e-1
w_joint_bounds = 0.0

# foot rectangle (ref:legacy/op3_conf.py:24-28) and contact model (:29-31)
lyp = 0.055
lyn = 0.055
lxp = 0.0275
lxn = 0.0275
lz = 0.0
mu = 0.5
fMin = 0.0
fMax = 1000.0

tau_max_scaling = 3.0
v_max_scaling = 10.0

# gains (ref:legacy/op3_conf.py:36-40)
kp_contact = 10.0
kp_foot = 10.0
kp_com = 10.0
kp_am = 10.0
kp_posture = 1.0

masks_posture = np.ones(18)
gain_vector = np.array(
    [100.0, 100.0]  # head
    + [10.0, 5.0, 5.0, 1.0, 1.0]  # left hip pitch/roll/yaw, knee, ankle pitch
    + [10.0, 10.0, 10.0]  # left arm
    + [10.0, 5.0, 5.0, 1.0, 1.0]  # right leg
    + [10.0, 10.0, 10.0]  # right arm
)

contactNormal = np.array([0.0, 0.0, 1.0])

rf_frame_name = "leg_right_sole_joint_fixed"
lf_frame_name = "leg_left_sole_joint_fixed"

urdf = "./robot/robot.urdf"
srdf = "./robot/robot.srdf"
path_to_urdf = "./robot"
mujoco_model_path = "./robot/robot.xml"

# additions of this implementation
device = 0
max_envs = 65536
