"""Oracle validation (CPU).  The reference has no tests or golden vectors for its tick
(parity unpinned), so the restatement in oracle/ is pinned by
  * the data-file KATs of SURVEY.md §4,
  * physical identities checked by finite differences on independent quantities,
  * a KKT optimality check of every QP solution (convex QP: KKT <=> optimal),
  * agreement between the fp64 build and the 80-bit long-double build.
"""
import numpy as np
import pytest

from common import SOLE_DUMMY, assert_whole_body_matches_mjcf, mjcf_case_q, mjcf_golden, setup
from tsid_control_b200 import synth

KINDS = ["v1", "v0"]


def _plus(orc, q, dv_dir, eps):
    """q (+) eps*dv_dir through the oracle's own integrate (v=0, dv = dir, dt chosen so that
    dt*v_mean = eps*dir): integrate_dv uses v_mean = v + dt/2 dv."""
    # v = dir, dv = 0, dt = eps  ->  q (+) eps*dir
    qn, _ = orc.integrate(q, dv_dir, np.zeros_like(dv_dir), eps)
    return qn


def _log3(R):
    tr = np.trace(R)
    th = np.arccos(np.clip((tr - 1) / 2, -1, 1))
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return w * (0.5 if th < 1e-6 else th / (2 * np.sin(th)))


# ------------------------------------------------------------------ KATs
def test_kat_joint_order_and_sizes():
    s1, s0 = setup("v1"), setup("v0")
    m1, m0 = s1["model"], s0["model"]
    assert (m1.nq, m1.nv, m1.na) == (27, 26, 20)
    assert (m0.nq, m0.nv, m0.na) == (25, 24, 18)
    # ref:main.py:14-42: q[7..8] head, q[9..14] left leg, q[15..17] left arm, q[18..23] right leg, q[24..26] right arm
    names = m1.joint_names
    assert names[0:2] == ["head_yaw", "head_pitch"]
    assert names[2:8] == ["left_hip_yaw", "left_hip_roll", "left_hip_pitch", "left_knee", "left_ankle_pitch", "left_ankle_roll"]
    assert names[8:11] == ["left_shoulder_pitch", "left_shoulder_roll", "left_elbow"]
    assert names[11:17] == ["right_hip_yaw", "right_hip_roll", "right_hip_pitch", "right_knee", "right_ankle_pitch", "right_ankle_roll"]
    assert names[17:20] == ["right_shoulder_pitch", "right_shoulder_roll", "right_elbow"]
    assert m0.joint_names[2:7] == ["left_hip_pitch", "left_hip_roll", "left_hip_yaw", "left_knee", "left_ankle_pitch"]


def test_kat_mass_and_limits():
    assert abs(setup("v1")["model"].total_mass - 2.893639) < 1e-6
    assert abs(setup("v0")["model"].total_mass - 2.784987) < 1e-6
    assert np.all(setup("v1")["model"].effort == 10) and np.all(setup("v1")["model"].velocity == 10)
    assert np.all(setup("v0")["model"].effort == 3) and np.all(setup("v0")["model"].velocity == 4)
    cc = setup("v1")["cc"]
    assert cc.tau_max[0] == 50.0 and cc.v_max[0] == 100.0  # ref:ctrl/conf.py:69-70 x URDF limits
    assert setup("v0")["cc"].tau_max[0] == 9.0


def test_kat_standing_sole_height_and_orientation():
    s = setup("v1")
    m, orc = s["model"], s["oracle"]
    q = m.q_ref["standing"].copy()
    r = orc.tick(q, np.zeros(m.nv), 3, s["refs"])
    # SRDF z = 0.331699 puts both soles at -2.687e-4 before the z-shift of ref:ctrl/WalkController.py:74
    assert abs(r["foot"][0][2] + 2.687e-4) < 1e-6 and abs(r["foot"][1][2] + 2.687e-4) < 1e-6
    R = r["foot"][0][3:].reshape(3, 3).T
    assert np.abs(R - np.eye(3)).max() < 2e-5 and np.abs(R - np.eye(3)).max() > 1e-7  # truncated literals kept
    r = orc.tick(s["q0"], np.zeros(m.nv), 3, s["refs"])
    assert abs(r["foot"][0][2]) < 1e-15
    s0 = setup("v0")
    r0 = s0["oracle"].tick(s0["q0"], np.zeros(24), 3, s0["refs"])
    assert np.abs(r0["foot"][:, 2]).max() < 1e-15
    assert np.abs(r0["foot"][0][3:].reshape(3, 3).T - np.diag([-1.0, -1.0, 1.0])).max() < 1e-10


def test_problem_sizes_match_survey_table():
    # n, m_e, one-sided m_i
    s = setup("v1")
    q, v = synth.random_states(s["q0"], 1, 1)
    for mask, (n, me, mi) in {3: (50, 18, 160), 1: (38, 12, 126), 2: (38, 12, 126), 0: (26, 6, 92)}.items():
        d = s["oracle"].tick(q[0], v[0], mask, s["refs"], dump=True)["dump"]
        assert (d["n"], d["neq"], d["nin"]) == (n, me, mi)
    s = setup("v0")
    q, v = synth.random_states(s["q0"], 1, 1)
    for mask, (n, me, mi) in {3: (48, 18, 104), 1: (36, 12, 70)}.items():
        d = s["oracle"].tick(q[0], v[0], mask, s["refs"], dump=True)["dump"]
        assert (d["n"], d["neq"], d["nin"]) == (n, me, mi)


# ------------------------------------------------------------------ dynamics identities
@pytest.mark.parametrize("kind", KINDS)
def test_mass_matrix_properties(kind):
    s = setup(kind)
    m, orc = s["model"], s["oracle"]
    q, v = synth.random_states(s["q0"], 4, 11)
    for i in range(4):
        d = orc.tick(q[i], v[i], 3, s["refs"], dump=True)["dump"]
        M = d["M"]
        assert np.abs(M - M.T).max() == 0.0
        assert np.linalg.eigvalsh(M).min() > 0
        assert np.abs(M[:3, :3] - m.total_mass * np.eye(3)).max() < 1e-12
        # linear rows of the centroidal map are m * Jcom
        assert np.abs(d["Ag"][:3] - m.total_mass * d["Jcom"]).max() < 1e-12
        # different branches do not couple: left leg x right leg block is zero
        names = m.joint_names
        li = [6 + k for k, nme in enumerate(names) if nme.startswith("left_hip") or nme.startswith("left_knee")]
        ri = [6 + k for k, nme in enumerate(names) if nme.startswith("right_hip") or nme.startswith("right_knee")]
        assert np.all(M[np.ix_(li, ri)] == 0.0)


@pytest.mark.parametrize("kind", KINDS)
def test_kinetic_energy_from_body_velocities(kind):
    """1/2 v^T M v == sum_i 1/2 m |c_i_dot|^2 + 1/2 w_i^T I_i w_i with body velocities taken by
    central finite differences of the world placements (independent of CRBA)."""
    s = setup(kind)
    m, orc = s["model"], s["oracle"]
    q, v = synth.random_states(s["q0"], 2, 12)
    eps = 1e-6
    for i in range(2):
        d0 = orc.tick(q[i], v[i], 3, s["refs"], dump=True)["dump"]
        dp_ = orc.tick(_plus(orc, q[i], v[i], eps), v[i], 3, s["refs"], dump=True)["dump"]
        dm_ = orc.tick(_plus(orc, q[i], v[i], -eps), v[i], 3, s["refs"], dump=True)["dump"]
        T = 0.0
        for b in range(m.nb):
            cp = dp_["oMi_p"][b] + dp_["oMi_R"][b] @ m.com[b]
            cm_ = dm_["oMi_p"][b] + dm_["oMi_R"][b] @ m.com[b]
            cdot = (cp - cm_) / (2 * eps)
            Rdot = (dp_["oMi_R"][b] - dm_["oMi_R"][b]) / (2 * eps)
            W = Rdot @ d0["oMi_R"][b].T
            w = np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) / 2
            Iw = d0["oMi_R"][b] @ m.inertia[b] @ d0["oMi_R"][b].T
            T += 0.5 * m.mass[b] * cdot @ cdot + 0.5 * w @ Iw @ w
        assert abs(0.5 * v[i] @ d0["M"] @ v[i] - T) < 1e-8 * max(1.0, T)


@pytest.mark.parametrize("kind", KINDS)
def test_gravity_term_is_potential_gradient(kind):
    s = setup(kind)
    m, orc = s["model"], s["oracle"]
    q, _ = synth.random_states(s["q0"], 2, 13)
    eps = 1e-6
    z = np.zeros(m.nv)
    for i in range(2):
        G = orc.tick(q[i], z, 3, s["refs"], dump=True)["dump"]["nle"]
        for k in range(m.nv):
            e = np.zeros(m.nv)
            e[k] = 1.0
            Up = m.total_mass * 9.81 * orc.tick(_plus(orc, q[i], e, eps), z, 3, s["refs"])["com"][2]
            Um = m.total_mass * 9.81 * orc.tick(_plus(orc, q[i], e, -eps), z, 3, s["refs"])["com"][2]
            assert abs((Up - Um) / (2 * eps) - G[k]) < 2e-7, (k, (Up - Um) / (2 * eps), G[k])


@pytest.mark.parametrize("kind", KINDS)
def test_energy_conservation_in_free_fall(kind):
    """With no contact and zero torque, dv = -M^-1 h; T + U must be conserved: checks the Coriolis
    part of nle against M (d/dt(1/2 v'Mv) + G'v = 0 iff C is consistent with M)."""
    s = setup(kind)
    m, orc = s["model"], s["oracle"]
    q, v = synth.random_states(s["q0"], 1, 14)
    q, v = q[0], v[0] * 3.0

    def energy(q, v):
        r = orc.tick(q, v, 0, s["refs"], dump=True)
        return 0.5 * v @ r["dump"]["M"] @ v + m.total_mass * 9.81 * r["com"][2], r["dump"]

    def deriv(q, v):
        # dE/dt along the passive flow, by central differences with a tiny step
        h = 1e-6
        _, d = energy(q, v)
        dv = -np.linalg.solve(d["M"], d["nle"])
        qp, vp = orc.integrate(q, v.copy(), dv, h)
        qm, vm = orc.integrate(q, v.copy(), dv, -h)
        return (energy(qp, vp)[0] - energy(qm, vm)[0]) / (2 * h), dv

    dE, dv = deriv(q, v)
    scale = abs(v @ (orc.tick(q, v, 0, s["refs"], dump=True)["dump"]["nle"]))
    assert abs(dE) < 1e-6 * max(1.0, scale), (dE, scale)


@pytest.mark.parametrize("kind", KINDS)
def test_jacobians_by_finite_differences(kind):
    s = setup(kind)
    m, orc = s["model"], s["oracle"]
    q, v = synth.random_states(s["q0"], 1, 15)
    q, v = q[0], v[0]
    eps = 1e-6
    r0 = orc.tick(q, v, 3, s["refs"], dump=True)
    d0 = r0["dump"]
    for k in range(m.nv):
        e = np.zeros(m.nv)
        e[k] = 1.0
        rp = orc.tick(_plus(orc, q, e, eps), v, 3, s["refs"])
        rm = orc.tick(_plus(orc, q, e, -eps), v, 3, s["refs"])
        assert np.abs((rp["com"][:3] - rm["com"][:3]) / (2 * eps) - d0["Jcom"][:, k]).max() < 1e-7
        for f in range(2):
            R0 = r0["foot"][f][3:].reshape(3, 3).T
            Rp = rp["foot"][f][3:].reshape(3, 3).T
            Rm = rm["foot"][f][3:].reshape(3, 3).T
            lin = R0.T @ (rp["foot"][f][:3] - rm["foot"][f][:3]) / (2 * eps)
            ang = _log3(Rm.T @ Rp) / (2 * eps)
            # log3(Rm^T Rp) is expressed in the frame at q-eps; to first order equal to the local twist at q
            assert np.abs(lin - d0["JF"][f][:3, k]).max() < 1e-6
            assert np.abs(ang - d0["JF"][f][3:, k]).max() < 1e-6
    # velocities: J v
    assert np.abs(d0["Jcom"] @ v - r0["com"][3:6]).max() < 1e-12
    for f in range(2):
        assert np.abs(d0["JF"][f] @ v - d0["vF"][f]).max() < 1e-12
    # momentum: Ag v angular part == sum of body angular momenta about the CoM is covered by the AM drift test


@pytest.mark.parametrize("kind", KINDS)
def test_drift_terms_by_finite_differences(kind):
    """Zero-joint-acceleration drifts: with dv = 0, d/dt of (JF v), (Jcom v), (Ag v)."""
    s = setup(kind)
    m, orc = s["model"], s["oracle"]
    q, v = synth.random_states(s["q0"], 1, 16)
    q, v = q[0], v[0]
    eps = 1e-6
    r0 = orc.tick(q, v, 3, s["refs"], dump=True)
    rp = orc.tick(_plus(orc, q, v, eps), v, 3, s["refs"], dump=True)
    rm = orc.tick(_plus(orc, q, v, -eps), v, 3, s["refs"], dump=True)
    assert np.abs((rp["com"][3:6] - rm["com"][3:6]) / (2 * eps) - r0["com"][6:9]).max() < 1e-7
    for f in range(2):
        spatial = (rp["dump"]["vF"][f] - rm["dump"]["vF"][f]) / (2 * eps)
        vF = r0["dump"]["vF"][f]
        classic = spatial.copy()
        classic[:3] += np.cross(vF[3:], vF[:3])
        assert np.abs(classic - r0["dump"]["aF"][f]).max() < 1e-6
    Lp = rp["dump"]["Ag"][3:] @ v
    Lm = rm["dump"]["Ag"][3:] @ v
    assert np.abs((Lp - Lm) / (2 * eps) - r0["dump"]["dAg_v_ang"]).max() < 1e-7


# ------------------------------------------------------------------ QP: KKT optimality
def _kkt_check(d, r, tol_feas=1e-5):
    H, g, CE, ce0, CI, ci0 = d["H"], d["g"], d["CE"], d["ce0"], d["CI"], d["ci0"]
    x = r["x"]
    assert np.abs(CE @ x + ce0).max() < 1e-7
    s = CI @ x + ci0
    # eiquadprog stops when the summed violation is below nIn*eps*c1*c2*100 (~1e-6 here)
    assert s.min() > -tol_feas
    act = r["active"]
    assert len(set(act.tolist())) == len(act)
    assert np.abs(s[act]).max(initial=0.0) < tol_feas
    # stationarity H x + g = CE^T le + CI_W^T u, u >= 0.  Solve in the H^-1 metric for conditioning.
    N = np.vstack([CE, CI[act]]).T
    L = np.linalg.cholesky(H)
    rhs = np.linalg.solve(L, H @ x + g)
    B = np.linalg.solve(L, N)
    lam, *_ = np.linalg.lstsq(B, rhs, rcond=None)
    res = B @ lam - rhs
    assert np.abs(res).max() < 1e-6 * max(1.0, np.abs(rhs).max())
    u = lam[CE.shape[0]:]
    assert u.min(initial=0.0) > -1e-7 * max(1.0, np.abs(u).max(initial=0.0))
    return lam


@pytest.mark.parametrize("kind,mask", [("v1", 3), ("v1", 1), ("v1", 2), ("v1", 0), ("v0", 3), ("v0", 2)])
def test_qp_solutions_satisfy_kkt(kind, mask):
    s = setup(kind)
    orc = s["oracle"]
    q, v = synth.random_states(s["q0"], 24, 2)
    n_act = 0
    for i in range(24):
        r = orc.tick(q[i], v[i], mask, s["refs"], dump=True)
        assert r["status"] == 0
        lam = _kkt_check(r["dump"], r)
        n_act += len(r["active"])
        # the oracle's own multipliers agree with the independent least-squares ones
        k = r["dump"]["neq"] + len(r["active"])
        assert np.abs(lam - r["lam"][:k]).max() < 1e-5 * max(1.0, np.abs(lam).max())
    if mask != 0:
        assert n_act > 0  # the sample really exercises the active-set iterations


def test_decode_matches_dynamics():
    """tau = h_a + M_a dv - J_a^T f and the base rows of the dynamics hold (ref:main.py:126-127)."""
    s = setup("v1")
    orc = s["oracle"]
    q, v = synth.random_states(s["q0"], 6, 3)
    T = np.zeros((6, 12))
    pts = np.array(s["cc"].contact_points)
    for c in range(4):
        T[:3, 3 * c:3 * c + 3] = np.eye(3)
        p = pts[:, c]
        T[3:, 3 * c:3 * c + 3] = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
    for i in range(6):
        r = orc.tick(q[i], v[i], 3, s["refs"], dump=True)
        d = r["dump"]
        gen = d["M"] @ r["dv"] + d["nle"]
        for f in range(2):
            gen -= d["JF"][f].T @ (T @ r["f"][12 * f:12 * f + 12])
        assert np.abs(gen[:6]).max() < 1e-8
        assert np.abs(gen[6:] - r["tau"]).max() < 1e-9
        # contact motion constraint holds: J dv = b
        assert np.abs(np.abs(r["tau"]).max()) <= 50.0 + 1e-6


def test_infeasible_and_status_codes():
    """fMin above what the robot weighs with torque limits shrunk -> not OPTIMAL, never a crash
    (SURVEY.md §5 failure handling; ref:main.py:122-124)."""
    import copy
    import ctypes as C

    from oracle_py import Oracle
    from tsid_control_b200._capi import TsidbConf

    s = setup("v1")
    cc = TsidbConf()
    C.memmove(C.byref(cc), C.byref(s["cc"]), C.sizeof(TsidbConf))
    cc.fmin = 900.0
    cc.fmax = 1000.0
    for i in range(20):
        cc.tau_max[i] = 0.05
        cc.tau_min[i] = -0.05
    orc = Oracle(s["cm"], cc)
    r = orc.tick(s["q0"], np.zeros(26), 3, s["refs"])
    assert r["status"] in (1, 3, 4)
    assert np.all(r["tau"] == 0) or r["status"] == 3


# ------------------------------------------------------------------ precision: fp64 vs 80-bit
@pytest.mark.parametrize("kind,mask", [("v1", 3), ("v1", 1), ("v0", 3)])
def test_fp64_against_long_double_truth(kind, mask):
    s64, s80 = setup(kind), setup(kind, "liboracle_ld.so")
    assert s80["oracle"].lib.oracle_real_bytes() == 16
    q, v = synth.random_states(s64["q0"], 32, 4)
    from common import canonical_active

    worst = 0.0
    worst_f = 0.0
    same_exact = same_canon = 0
    T = np.zeros((6, 12))
    pts = np.array(s64["cc"].contact_points)
    for c in range(4):
        T[:3, 3 * c:3 * c + 3] = np.eye(3)
        p = pts[:, c]
        T[3:, 3 * c:3 * c + 3] = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
    for i in range(32):
        a = s64["oracle"].tick(q[i], v[i], mask, s64["refs"])
        b = s80["oracle"].tick(q[i], v[i], mask, s64["refs"])
        assert a["status"] == b["status"] == 0
        rows = s64["oracle"].ci_rows(mask)
        ra, rb = [rows[k] for k in a["active"]], [rows[k] for k in b["active"]]
        same_exact += set(ra) == set(rb)
        same_canon += canonical_active(ra) == canonical_active(rb)
        wa = np.r_[T @ a["f"][:12], T @ a["f"][12:]]
        wb = np.r_[T @ b["f"][:12], T @ b["f"][12:]]
        for x, y in ((a["tau"], b["tau"]), (a["dv"], b["dv"]), (wa, wb)):
            worst = max(worst, (np.abs(x - y) / (1e-10 / 1e-8 + np.abs(y))).max())
        worst_f = max(worst_f, (np.abs(a["f"] - b["f"]) / (1e-10 / 1e-8 + np.abs(b["f"]))).max())
    # Pivoting is identical up to the choice of 3-of-4 rows at unloaded corners (rounding decides).
    assert same_canon >= 31, (same_exact, same_canon)
    # tau, dv and the 6-D contact wrenches of the fp64 build sit at about 1e-8 rel / 1e-10 abs of the
    # truth in the worst env (3e-9 for v1, 1.4e-8 for v0 on these seeds): that is the noise floor of
    # the reference algorithm in fp64, so no fp64 implementation can be closer to it than that.
    # The 12 corner forces per foot are only determined through the 1e-8 Hessian regulariser and
    # carry ~1e-7 of fp64 noise (SURVEY.md §7 "Conditioning").
    assert worst < 5e-8, worst
    assert worst_f < 1e-5, worst_f


# ------------------------------------------------------------------ integrate
def test_integrate_matches_se3_exponential():
    s = setup("v1")
    orc = s["oracle"]
    q, v = synth.random_states(s["q0"], 3, 5)
    rng = np.random.default_rng(0)
    for i in range(3):
        dv = rng.uniform(-5, 5, 26)
        dt = 0.002
        qn, vn = orc.integrate(q[i], v[i], dv, dt)
        assert np.allclose(vn, v[i] + dt * dv, rtol=0, atol=1e-15)
        vm = dt * (v[i] + 0.5 * dt * dv)
        assert np.allclose(qn[7:], q[i][7:] + vm[6:], rtol=0, atol=1e-15)
        # independent SE3 exp (Rodrigues + V matrix)
        w, u = vm[3:6], vm[:3]
        th = np.linalg.norm(w)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        R1 = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * K + (th - np.sin(th)) / th**3 * K @ K
        x, y, z, ww = q[i][3:7]
        R0 = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
                       [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
                       [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(qn[:3], q[i][:3] + R0 @ (V @ u), atol=1e-14)
        x, y, z, ww = qn[3:7]
        Rn = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
                       [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
                       [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(Rn, R0 @ R1, atol=1e-13)
        assert abs(np.linalg.norm(qn[3:7]) - 1) < 1e-12


# ------------------------------------------------------------------ reference-held model data (robot/v1/mujoco/robot.xml)
def test_kat_model_tables_match_the_references_mujoco_export():
    """Body by body: joint placement, mass, lever and inertia tensor of the tables compiled from the URDF against the
    reference's independent MuJoCo export of the same CAD (tests/golden/make_mjcf_golden.py).  Tolerances are the
    print precision of the two files (6 significant digits; 1.5708 / 3.14159 literals in the URDF)."""
    g, m = mjcf_golden(), setup("v1")["model"]
    assert sorted(g["joint_names"]) == sorted(m.joint_names)
    body_of = {n: b + 1 for b, n in enumerate(m.joint_names)}
    body_of[None] = 0
    for gb in g["bodies"]:
        b = body_of[gb["joint"]]
        assert m.parent[b] == (body_of[gb["parent_joint"]] if b else -1), gb["name"]
        if b:
            assert np.abs(m.jp[b] - np.array(gb["pos"])).max() < 1e-8, gb["name"]
            assert np.abs(m.jR[b] - np.array(gb["R"])).max() < 2e-5, gb["name"]
        mass, com, I = m.mass[b], m.com[b], m.inertia[b]
        if gb["joint"] in SOLE_DUMMY:  # take the merged dummy link out again
            fr = m.frames[SOLE_DUMMY[gb["joint"]]]
            assert fr["body"] == b
            m2 = mass - 0.01
            c2 = (mass * com - 0.01 * fr["p"]) / m2
            d1, d2 = fr["p"] - com, c2 - com
            I = I - 1e-4 * np.eye(3) - 0.01 * (d1 @ d1 * np.eye(3) - np.outer(d1, d1)) - m2 * (d2 @ d2 * np.eye(3) - np.outer(d2, d2))
            mass, com = m2, c2
        assert abs(mass - gb["mass"]) < 1e-9, gb["name"]
        assert np.abs(com - np.array(gb["com"])).max() < 1e-8, gb["name"]
        assert np.abs(I - np.array(gb["inertia"])).max() < 1e-9, gb["name"]


def test_kat_whole_body_quantities_match_the_references_mujoco_export():
    """The oracle's kinematics and CRBA (C code, tables from the URDF) against whole-body quantities evaluated from the
    MuJoCo file by an independent tree walk: total mass, CoM and rotational inertia about the CoM in the torso frame at
    12 joint configurations — joint order and sign, placements, levers and inertias all enter."""
    g, s = mjcf_golden(), setup("v1")
    m, orc = s["model"], s["oracle"]
    for case in g["cases"]:
        q = mjcf_case_q(m, case)
        r = orc.tick(q, np.zeros(m.nv), 3, s["refs"], dump=True)
        assert_whole_body_matches_mjcf(m, case, r["dump"]["M"], r["com"][:3])
