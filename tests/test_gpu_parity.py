"""GPU parity tests: the CUDA tick, called through the C ABI (libtsidb.so via ctypes), against the CPU
oracle on the same seeded inputs.  Tolerances (BASELINE.json north_star): tau, ddq and contact wrenches
within 1e-8 relative / 1e-10 absolute, i.e. |a-b| <= 1e-10 + 1e-8 |b|, written below as
err = |a-b| / (1e-2 + |b|) <= 1e-8.
"""
import os

import numpy as np
import pytest

import parity_log
from common import TOL, assert_parity, bits_to_rows, canonical_active, compare_outputs, make_conf, setup
from tsid_control_b200 import synth

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _err(a, b):
    return float((np.abs(a - b) / (1e-2 + np.abs(b))).max())


def _T(cc):
    T = np.zeros((6, 12))
    pts = np.array(cc.contact_points)
    for c in range(4):
        T[:3, 3 * c:3 * c + 3] = np.eye(3)
        p = pts[:, c]
        T[3:, 3 * c:3 * c + 3] = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
    return T


def _wrench(T, f):
    return np.concatenate([f[:, :12] @ T.T, f[:, 12:] @ T.T], axis=1)


_LIVE = []


@pytest.fixture(autouse=True)
def _close_handles():
    yield
    while _LIVE:
        _LIVE.pop().engine.close()


def _controller(kind, n, overrides=()):
    c = _make_controller(kind, n, overrides)
    _LIVE.append(c)
    return c


def _make_controller(kind, n, overrides=()):
    conf = make_conf(kind, overrides)
    conf.max_envs = n
    if kind == "v1":
        from tsid_control_b200.ctrl.WalkController import WalkController

        return WalkController(conf, n_envs=n)
    from tsid_control_b200.legacy.biped import Biped

    return Biped(conf, n_envs=n)


def _bits(words):
    out = []
    for w in range(3):
        val = int(words[w]) & ((1 << 64) - 1)
        out += [64 * w + b for b in range(64) if (val >> b) & 1]
    return out


def _run(ctrl, q, v, mask, refs=None):
    dev = ctrl.device
    ctrl.contact_mask = torch.as_tensor(mask, device=dev)
    if refs is not None:
        ctrl.refs = {k: torch.as_tensor(np.ascontiguousarray(a), device=dev) for k, a in refs.items()}
    ctrl._tick(torch.as_tensor(q, device=dev), torch.as_tensor(v, device=dev), aux=True)
    torch.cuda.synchronize()
    o = ctrl.last
    return {k: getattr(o, k).cpu().numpy() for k in o.__slots__}


THREADS = max(8, os.cpu_count() or 8)


def _compare(kind, n, seed, mask, refs_np=None, threads=THREADS, overrides=(), tag=None):
    """One batch through the CUDA tick (C ABI) and through the oracle (common.compare_outputs); every number
    measured is written to gpurun_out/parity_r02.json (tests/parity_log.py)."""
    s = setup(kind, overrides=tuple(overrides))
    ctrl = _controller(kind, n, overrides)
    q, v = synth.random_states(s["q0"], n, seed)
    out = _run(ctrl, q, v, mask, refs_np)
    res, ref = compare_outputs(kind, q, v, mask, refs_np, out, threads=threads, overrides=tuple(overrides))
    name = tag or os.environ.get("PYTEST_CURRENT_TEST", f"{kind}-{n}-{seed}").split("::")[-1].split(" ")[0]
    parity_log.record(name, res)
    print(f"\n[{kind} n={n} seed={seed}] " + " ".join(f"{k}={v:.3g}" for k, v in res.items() if isinstance(v, (int, float))))
    return res, out, ref


def _assert_parity(res, kind="v1"):
    assert_parity(res, kind)


def test_kinematics_match_oracle():
    s = setup("v1")
    ctrl = _controller("v1", 256)
    q, v = synth.random_states(s["q0"], 256, 3)
    com, lf, rf = ctrl.engine.kinematics(torch.as_tensor(q, device=ctrl.device), torch.as_tensor(v, device=ctrl.device))
    torch.cuda.synchronize()
    com, lf, rf = com.cpu().numpy(), lf.cpu().numpy(), rf.cpu().numpy()
    for i in range(0, 256, 8):
        r = s["oracle"].tick(q[i], v[i], 3, s["refs"])
        assert np.abs(com[i] - r["com"]).max() < 1e-12
        assert np.abs(lf[i] - r["foot"][0]).max() < 1e-13 and np.abs(rf[i] - r["foot"][1]).max() < 1e-13


@pytest.mark.parametrize("kind", ["v1", "v0"])
def test_dynamics_terms_match_oracle(kind, monkeypatch):
    """what RobotWrapper::computeAllTerms and the solver's H, g build produce
    inside computeProblemData / solve (ref:main.py:119,121) — the joint-space inertia M, the non-linear effects h, the
    LOCAL sole Jacobians, the dv block of the Hessian and of the gradient — read back from the dynamics kernel's
    hand-off images (tsidb_debug_terms) and compared with the oracle's dump for every env of a small batch."""
    s = setup(kind)
    n = 24
    monkeypatch.setenv("TSIDB_SMALL_N", "1024")  # read by tsidb_create: the single-launch path keeps slot == env
    monkeypatch.setenv("TSIDB_SMALL_LOCAL_N", "0")  # ... and leaves its hand-off images in global memory
    ctrl = _controller(kind, n)
    q, v = synth.random_states(s["q0"], n, 17)
    step = (0.3, 0.2, 0.2, 0.5) if kind == "v1" else (0.1, 0.1275, 0.05, 0.7)
    mask, refs = synth.walking_batch(s["refs"], n, 17, *step, float(s["refs"]["com"][2]))
    mask[5] = 0
    _run(ctrl, q, v, mask, refs)  # n <= TSIDB_SMALL_N: no class sort, slot == env
    nv = ctrl.engine.nv
    worst = {}
    for e in range(n):
        m = int(mask[e])
        t = ctrl.engine.debug_terms(e, (m & 1) + (m >> 1))
        d = s["oracle"].tick(q[e], v[e], m, {k: a[e] for k, a in refs.items()}, dump=True)["dump"]
        tril = np.tril_indices(nv)
        pairs = {"M": (t["M"], d["M"]), "nle": (t["nle"], d["nle"]), "H": (t["H"][tril], d["H"][:nv, :nv][tril]),
                 "g": (t["g"], d["g"][:nv])}
        for f, bit in ((0, 1), (1, 2)):
            pairs[f"JF{f}"] = (t["JF"][f], d["JF"][f])
        for k, (a, b) in pairs.items():
            worst[k] = max(worst.get(k, 0.0), float(np.max(np.abs(a - b)) / (1e-2 + np.max(np.abs(b)))))
    print("\ndynamics terms vs oracle (max |a-b| / (1e-2 + max|b|)):", {k: f"{x:.1e}" for k, x in worst.items()})
    parity_log.record(f"dynamics_terms_{kind}", worst)
    assert all(x < 1e-10 for x in worst.values()), worst


def test_dynamics_kernel_against_the_references_mujoco_export(monkeypatch):
    """Row a2 against reference-held data, no oracle in between: the joint-space inertia the dynamics kernel hands to the
    solver (its base block = total mass and rotational inertia about the torso origin) and the CoM of the tick's aux outputs,
    at the 12 configurations of tests/golden/mjcf_v1.json (made from ref:robot/v1/mujoco/robot.xml)."""
    from common import assert_whole_body_matches_mjcf, mjcf_case_q, mjcf_golden

    s, g = setup("v1"), mjcf_golden()
    m, n = s["model"], len(g["cases"])
    monkeypatch.setenv("TSIDB_SMALL_N", "1024")  # slot == env for tsidb_debug_terms
    monkeypatch.setenv("TSIDB_SMALL_LOCAL_N", "0")  # hand-off images in global memory
    ctrl = _controller("v1", n)
    q = np.stack([mjcf_case_q(m, c) for c in g["cases"]])
    v = np.zeros((n, m.nv))
    refs = {k: np.broadcast_to(a, (n,) + np.shape(a)).copy() for k, a in s["refs"].items()}
    out = _run(ctrl, q, v, np.full(n, 3, dtype=np.uint8), refs)
    for e, case in enumerate(g["cases"]):
        assert_whole_body_matches_mjcf(m, case, ctrl.engine.debug_terms(e, 2)["M"], out["com"][e, :3])


def test_constructor_references_match_reference_semantics():
    s = setup("v1")
    ctrl = _controller("v1", 4)
    assert np.abs(ctrl.q - s["q0"]).max() < 1e-15  # z-shift of ref:ctrl/WalkController.py:74
    assert ctrl.q0 is ctrl.q  # aliased like the reference (:23)
    for k in ("com", "foot_lf", "foot_rf", "contact_lf", "contact_rf", "posture"):
        assert np.abs(ctrl.default_refs[k] - s["refs"][k]).max() < 1e-13, k
    assert (ctrl.formulation.nVar, ctrl.formulation.nEq, ctrl.formulation.nIn) == (50, 18, 80)


def test_standing_balance_config2_batch4096():
    """BASELINE.json configs[1]: robot/v1 double-support standing balance, 4096 perturbed states."""
    n = 4096
    res, out, ref = _compare("v1", n, 1, np.full(n, 3, np.uint8))
    assert (out["status"] == 0).all()
    _assert_parity(res)


@pytest.mark.parametrize("seed", [2, 3])
def test_walking_phases_v1(seed):
    n = 1024
    s = setup("v1")
    com_h = s["refs"]["com"][2]
    mask, refs = synth.walking_batch(s["refs"], n, seed, 0.3, 0.2, 0.2, 0.5, com_h)
    res, out, ref = _compare("v1", n, seed, mask, refs)
    _assert_parity(res)
    assert set(np.unique(mask)) == {1, 2, 3}


def test_flight_phase_no_contacts():
    n = 128
    res, out, ref = _compare("v1", n, 4, np.zeros(n, np.uint8))
    _assert_parity(res)
    assert np.all(out["f"] == 0.0)


@pytest.mark.parametrize("maskval", [3, 1, 2])
def test_legacy_op3_v0(maskval):
    """BASELINE.json configs[3]: legacy model (op3_conf + Biped): AM task, RF-first order, fMin = 0."""
    n = 1024
    res, out, ref = _compare("v0", n, 5, np.full(n, maskval, np.uint8))
    _assert_parity(res, "v0")


def test_soa_layout_and_host_call_agree_with_device_call():
    import ctypes as C

    from tsid_control_b200._capi import TsidbRefs, check

    s = setup("v1")
    n = 200
    ctrl = _controller("v1", n)
    e = ctrl.engine
    q, v = synth.random_states(s["q0"], n, 6)
    mask = np.array([3, 1, 2, 3, 0] * (n // 5), np.uint8)
    base = _run(ctrl, q, v, mask)
    # SoA through the raw C ABI
    dev = e.device
    qs = torch.as_tensor(np.ascontiguousarray(q.T), device=dev)
    vs = torch.as_tensor(np.ascontiguousarray(v.T), device=dev)
    md = torch.as_tensor(mask, device=dev)
    tau = torch.empty((e.na, n), dtype=torch.float64, device=dev)
    ddq = torch.empty((e.nv, n), dtype=torch.float64, device=dev)
    f = torch.empty((24, n), dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    it = torch.empty(n, dtype=torch.int32, device=dev)
    r = TsidbRefs()
    check(e.lib.tsidb_compute(e.h, n, 1, qs.data_ptr(), vs.data_ptr(), md.data_ptr(), C.byref(r), tau.data_ptr(), ddq.data_ptr(),
                              f.data_ptr(), st.data_ptr(), it.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream), "soa")
    torch.cuda.synchronize()
    assert np.array_equal(tau.cpu().numpy().T, base["tau"]) and np.array_equal(ddq.cpu().numpy().T, base["ddq"])
    assert np.array_equal(f.cpu().numpy().T, base["f"]) and np.array_equal(st.cpu().numpy(), base["status"])
    # host buffers through tsidb_compute_host
    h = e.compute_host(q, v, mask)
    assert np.array_equal(h["tau"], base["tau"]) and np.array_equal(h["ddq"], base["ddq"]) and np.array_equal(h["f"], base["f"])
    assert np.array_equal(h["status"], base["status"]) and np.array_equal(h["iters"], base["iters"])
    assert np.array_equal(h["active_set"].astype(np.int64), base["active_set"])


@pytest.mark.parametrize("n,pinned", [(4100, False), (4100, True), (16500, True)])
def test_chunked_host_call_agrees_with_device_call(n, pinned):
    """tsidb_compute_host cuts the batch into 2 (n >= 2048) or 4 (n >= 16384) chunks on rotating streams; pinned
    caller buffers are DMA'd directly, pageable ones go through the staging copy.  Both must reproduce the
    single-launch device call bit for bit, including a ragged last chunk and per-env references."""
    s = setup("v1")
    ctrl = _controller("v1", n)
    e = ctrl.engine
    q, v = synth.random_states(s["q0"], n, 12)
    mask, refs = synth.walking_batch(s["refs"], n, 12, 0.3, 0.2, 0.2, 0.5, float(s["refs"]["com"][2]))
    base = _run(ctrl, q, v, mask, refs)
    if pinned:
        hq, hv, hm = e.pin(q), e.pin(v), e.pin(mask)
        hr = {k: e.pin(a) for k, a in refs.items()}
        out = e.host_buffers(n, pinned=True)
        h = e.compute_host(hq, hv, hm, hr, out=out)
    else:
        h = e.compute_host(q, v, mask, refs)
    for k in ("tau", "ddq", "f", "status", "iters"):
        assert np.array_equal(h[k], base[k]), k
    assert np.array_equal(h["active_set"].astype(np.int64), base["active_set"])
    # the same call with the references and contact phases resident on the device: only q and v are uploaded
    dev = e.device
    d = e.compute_host_devrefs(q, v, torch.as_tensor(mask, device=dev),
                               {k: torch.as_tensor(np.ascontiguousarray(a), device=dev) for k, a in refs.items()})
    for k in ("tau", "ddq", "f", "status", "iters"):
        assert np.array_equal(d[k], base[k]), k
    assert np.array_equal(d["active_set"].astype(np.int64), base["active_set"])
    # torque-only call (ddq and f stay on the device): tau, status, iters as before, the other buffers untouched
    out2 = e.host_buffers(n, pinned=True)
    out2["ddq"][:] = -7.0
    out2["f"][:] = -7.0
    t = e.compute_host(q, v, mask, refs, out=out2, tau_only=True)
    for k in ("tau", "status", "iters"):
        assert np.array_equal(t[k], base[k]), k
    assert (t["ddq"] == -7.0).all() and (t["f"] == -7.0).all()


@pytest.mark.parametrize("kind,n", [("v1", 1), ("v1", 97), ("v1", 1024), ("v0", 64), ("v0", 333)])
def test_single_launch_small_batch_tick_equals_the_batched_pipeline(kind, n, monkeypatch):
    """Ticks of at most TSIDB_SMALL_N envs (default 1024) run as one launch (tsidb_tick_small_kernel: one warp per env
    through dynamics, elimination and active set).  It calls the stage functions of the batched kernels, so every
    output — tau, ddq, f, status, iters, working sets, multipliers, kinematics — must equal the batched pipeline's bit
    for bit, on all three contact classes.  Up to 2 envs per SM (1, 97, 64 here) the kernel keeps its hand-off images in
    shared memory (TSIDB_SMALL_LOCAL_N), above that (333, 1024) they travel through global memory: both variants are covered."""
    s = setup(kind)
    q, v = synth.random_states(s["q0"], n, 41)
    step = (0.3, 0.2, 0.2, 0.5) if kind == "v1" else (0.1, 0.1275, 0.05, 0.7)
    mask, refs = synth.walking_batch(s["refs"], n, 41, *step, float(s["refs"]["com"][2]))
    if n > 4:
        mask[3] = 0  # one env in flight
    outs = []
    for small in ("1024", "0"):
        monkeypatch.setenv("TSIDB_SMALL_N", small)  # read by tsidb_create
        ctrl = _controller(kind, n)
        launches0 = ctrl.engine.launch_count()
        outs.append(_run(ctrl, q, v, mask, refs))
        launches = ctrl.engine.launch_count() - launches0
        assert launches == (1 if small == "1024" else 9), launches
    a, b = outs
    assert (a["status"] == 0).sum() >= 0.9 * n
    for k in a:
        if a[k] is not None:
            assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("use_graph", [True, False])
def test_device_rollout_matches_host_driven_loop(use_graph):
    """tsidb_rollout (tick -> integrate -> gait step, no host round trip, optionally one CUDA graph replayed) against
    the same loop driven from the host with the numpy gait restatement (tests/gait_ref.py)."""
    from gait_ref import GaitRef

    s = setup("v1")
    n, steps = 96, 40
    ctrl = _controller("v1", n)
    e = ctrl.engine
    dev = e.device
    conf = s["conf"]
    rng = np.random.Generator(np.random.PCG64(31))
    q, v = synth.random_states(s["q0"], n, 15)
    v *= 0.2
    phase0 = rng.uniform(0, 1, n)
    vcmd = np.c_[rng.uniform(-0.3, 0.3, n), rng.uniform(-0.1, 0.1, n)]
    h0 = float(ctrl.default_refs["com"][2])
    gait = dict(dt=conf.dt, step_duration=conf.step_duration, step_length=conf.step_length, step_height=conf.step_height, com_height=h0)
    # device
    e.gait_reset(n, phase0=torch.as_tensor(phase0, device=dev), vcmd=torch.as_tensor(vcmd, device=dev), **gait)
    qd, vd = torch.as_tensor(q, device=dev).clone(), torch.as_tensor(v, device=dev).clone()
    out = e.rollout(qd, vd, steps, use_graph=use_graph)
    torch.cuda.synchronize()
    gs = {k: t.cpu().numpy().copy() for k, t in e.gait_state().items()}
    # host-driven
    g = GaitRef(n, defaults=ctrl.default_refs, phase0=phase0, vcmd=vcmd, **gait)
    qh, vh = torch.as_tensor(q, device=dev).clone(), torch.as_tensor(v, device=dev).clone()
    post = torch.as_tensor(np.tile(ctrl.default_refs["posture"], (n, 1)), device=dev)
    for _ in range(steps):
        refs = {k: torch.as_tensor(np.ascontiguousarray(a), device=dev) for k, a in g.refs().items()}
        refs["posture"] = post
        o = e.compute(qh, vh, torch.as_tensor(g.mask, device=dev), refs, aux=True)
        e.integrate(qh, vh, o.ddq, conf.dt)
        g.step(o.foot_lf.cpu().numpy(), o.foot_rf.cpu().numpy(), o.status.cpu().numpy())
    torch.cuda.synchronize()
    assert np.array_equal(gs["mask"], g.mask) and np.array_equal(gs["fails"], g.fails)
    assert len(set(int(m) for m in g.mask)) == 3  # all three contact classes are present
    ok = g.fails == 0
    assert ok.mean() > 0.9
    assert np.abs(qd.cpu().numpy()[ok] - qh.cpu().numpy()[ok]).max() < 1e-8
    assert np.abs(vd.cpu().numpy()[ok] - vh.cpu().numpy()[ok]).max() < 1e-6
    for k in ("com", "foot_lf", "foot_rf", "contact_lf", "contact_rf"):
        assert np.abs(gs[k][ok] - g.refs()[k][ok]).max() < 1e-8, k
    assert np.abs(out.tau.cpu().numpy()[ok] - o.tau.cpu().numpy()[ok]).max() < 1e-5


def test_integrate_matches_oracle():
    s = setup("v1")
    n = 64
    ctrl = _controller("v1", n)
    q, v = synth.random_states(s["q0"], n, 7)
    rng = np.random.default_rng(7)
    dv = rng.uniform(-20, 20, v.shape)
    qd, vd = torch.as_tensor(q, device=ctrl.device).clone(), torch.as_tensor(v, device=ctrl.device).clone()
    ctrl.integrate_dv(qd, vd, torch.as_tensor(dv, device=ctrl.device), 0.002)
    torch.cuda.synchronize()
    for i in range(n):
        qo, vo = s["oracle"].integrate(q[i], v[i], dv[i], 0.002)
        assert np.abs(qd[i].cpu().numpy() - qo).max() < 1e-14 and np.abs(vd[i].cpu().numpy() - vo).max() < 1e-14
    # single-robot numpy form mutates v in place like the reference (ref:ctrl/WalkController.py:293)
    v1 = v[0].copy()
    q1, v1b = ctrl.integrate_dv(q[0], v1, dv[0], 0.002)
    assert v1b is v1 and np.abs(v1 - (v[0] + 0.002 * dv[0])).max() < 1e-15


def test_single_robot_reference_tick_sequence():
    """ref:main.py:119-128 verbatim against the mirror objects, closed loop for a few ticks."""
    s = setup("v1")
    ctrl = _controller("v1", 1)
    conf = ctrl.conf
    q, v = ctrl.q.copy(), ctrl.v.copy()
    q[7:] += 0.05
    qo, vo = q.copy(), v.copy()
    t = 0.0
    for _ in range(5):
        HQPData = ctrl.formulation.computeProblemData(t, q, v)
        sol = ctrl.solver.solve(HQPData)
        assert sol.status == 0
        tau = ctrl.formulation.getActuatorForces(sol)
        dv = ctrl.formulation.getAccelerations(sol)
        r = s["oracle"].tick(qo, vo, 3, s["refs"])
        assert _err(tau, r["tau"]) < 5 * TOL and _err(dv, r["dv"]) < 5 * TOL
        assert sol.x.shape == (50,) and sol.iterations == r["iters"]
        cop = ctrl.get_cop(sol)
        assert cop is not None and abs(cop[2]) == 0.0
        q, v = ctrl.integrate_dv(q, v, dv, conf.dt)
        qo, vo = s["oracle"].integrate(qo, vo, r["dv"], conf.dt)
        t += conf.dt
    assert np.abs(q - qo).max() < 1e-9


def test_contact_switching_legacy_semantics():
    s = setup("v1")
    ctrl = _controller("v1", 1)
    q, v = ctrl.q.copy(), ctrl.v.copy()
    sol = ctrl.solver.solve(ctrl.formulation.computeProblemData(0.0, q, v))
    assert ctrl.formulation.nVar == 50
    ctrl.remove_contact(left_foot=True, right_foot=False)
    assert (ctrl.formulation.nVar, ctrl.formulation.nEq, ctrl.formulation.nIn) == (38, 12, 63)
    sol = ctrl.solver.solve(ctrl.formulation.computeProblemData(0.0, q, v))
    r = s["oracle"].tick(q, v, 2, s["refs"])
    assert sol.status == r["status"] == 0
    assert _err(ctrl.formulation.getActuatorForces(sol), r["tau"]) < 5 * TOL
    assert np.all(ctrl.formulation.getContactForce("contact_lfoot", sol) == 0.0)
    ctrl.add_contact(left_foot=True, right_foot=False)
    # [UPSTREAM] the re-added contact goes last in x: x = [dv; f_RF; f_LF]
    assert ctrl._contact_order == [1, 0] and ctrl.formulation.nVar == 50
    sol = ctrl.solver.solve(ctrl.formulation.computeProblemData(0.0, q, v))
    assert np.array_equal(sol.x[26:38], ctrl.formulation.getContactForce("contact_rfoot", sol))


def test_per_env_contact_switch_after_a_plain_compute():
    """The natural loop `ctrl.compute(q, v); ctrl.update_tasks(sLF, sRF, cl, cr)` with per-env contact tensors
    (legacy semantics, ref:legacy/biped.py:168-212): compute() leaves the sole placements of ITS tick in ctrl.last,
    lift-off copies them into the foot-task reference, touch-down into the contact reference; a tick without aux
    outputs leaves them None and the switch refuses to run on stale data."""
    s = setup("v1")
    n = 64
    ctrl = _controller("v1", n)
    dev = ctrl.device
    q, v = synth.random_states(s["q0"], n, 51)
    qd, vd = torch.as_tensor(q, device=dev), torch.as_tensor(v, device=dev)
    ctrl.contact_mask = torch.as_tensor(np.array([3, 3, 1, 2] * (n // 4), np.uint8), device=dev)
    ctrl._tick(qd, vd)
    assert ctrl.last.foot_lf is None and ctrl.last.wrench is None and ctrl.last.active_set is not None
    cl = torch.as_tensor(np.array([False, True, True, True] * (n // 4)), device=dev)   # env 0: lift LF, env 3: land LF
    cr = torch.as_tensor(np.array([True, True, True, True] * (n // 4)), device=dev)    # env 2: land RF
    with pytest.raises(RuntimeError, match="foot placements"):
        ctrl.set_contact_phase(cl, cr)
    ctrl.compute(qd, vd, 0.0)
    sample = ctrl.traj_LF.computeNext()
    ctrl.update_tasks(sample, ctrl.traj_RF.computeNext(), cl, cr)
    mask = ctrl.contact_mask.cpu().numpy()
    assert np.array_equal(mask, np.array([2, 3, 3, 3] * (n // 4), np.uint8))
    refs = {k: t.cpu().numpy() for k, t in ctrl.refs.items()}
    for i in range(n):
        r = s["oracle"].tick(q[i], v[i], 3, s["refs"])
        lf12, rf12 = r["foot"][0], r["foot"][1]
        if i % 4 == 0:
            assert np.abs(refs["foot_lf"][i, :12] - lf12).max() < 1e-13 and np.all(refs["foot_lf"][i, 12:] == 0)
        if i % 4 == 3:
            assert np.abs(refs["contact_lf"][i] - lf12).max() < 1e-13
        if i % 4 == 2:
            assert np.abs(refs["contact_rf"][i] - rf12).max() < 1e-13
    out = _run(ctrl, q, v, mask, refs)
    ref = s["oracle"].batch(q, v, mask, refs, n_threads=4)
    assert np.array_equal(out["status"], ref["status"]) and (ref["status"] == 0).all()
    assert _err(out["tau"], ref["tau"]) < TOL and _err(out["ddq"], ref["dv"]) < TOL


def test_infeasible_envs_do_not_abort_the_batch():
    """per-env status, never abort the batch.  An env whose state is NaN must not disturb its
    neighbours; an infeasible problem reports a non-zero status and zero outputs."""
    s = setup("v1")
    n = 32
    ctrl = _controller("v1", n)
    q, v = synth.random_states(s["q0"], n, 8)
    v[5] *= 400.0  # absurd velocities: joint-velocity bounds and torque limits collide
    base = _run(ctrl, q, v, np.full(n, 3, np.uint8))
    ref = s["oracle"].batch(q, v, np.full(n, 3, np.uint8), s["refs"], n_threads=4)
    assert np.array_equal(base["status"], ref["status"])
    good = np.arange(n) != 5
    assert (base["status"][good] == 0).all()
    assert _err(base["tau"][good], ref["tau"][good]) < 5 * TOL


def test_full_size_batch65536_config3_every_env_against_the_oracle():
    """BASELINE.json configs[2] at its quoted size and on the bench's own inputs (seed 0): ALL 65536 envs against the
    fp64 oracle and its 80-bit build (tolerances and working-set rules of _assert_parity), plus the properties that
    need no oracle: feasibility of every solution, run-to-run determinism, permutation (sharding) invariance."""
    s = setup("v1")
    n = 65536
    com_h = s["refs"]["com"][2]
    mask, refs = synth.walking_batch(s["refs"], n, 0, 0.3, 0.2, 0.2, 0.5, com_h)
    res, a, ref = _compare("v1", n, 0, mask, refs)
    _assert_parity(res)
    assert res["n_optimal"] == n
    q, v = synth.random_states(s["q0"], n, 0)
    ctrl = _LIVE[-1]
    ok = a["status"] == 0
    f = a["f"][ok].reshape(-1, 8, 3)
    assert (np.abs(f[:, :, 0]) <= 0.5 * f[:, :, 2] + 1e-5).all() and (np.abs(f[:, :, 1]) <= 0.5 * f[:, :, 2] + 1e-5).all()
    fz = f[:, :, 2].reshape(-1, 2, 4).sum(-1)
    on = np.stack([(mask[ok] & 1) != 0, (mask[ok] & 2) != 0], axis=1)
    assert (fz[on] >= 10.0 - 1e-5).all() and (fz[~on] == 0).all()
    assert (np.abs(a["tau"][ok]) <= 50.0 + 1e-5).all()
    b = _run(ctrl, q, v, mask, refs)
    for k in ("tau", "ddq", "f", "status", "iters", "active_set"):
        assert np.array_equal(a[k], b[k]), k
    perm = np.random.default_rng(0).permutation(n)
    c = _run(ctrl, q[perm], v[perm], mask[perm], {k: r[perm] for k, r in refs.items()})
    for k in ("tau", "ddq", "f", "status", "iters"):
        assert np.array_equal(a[k][perm], c[k]), k


def test_legacy_walking_batch16384_config4_every_env_against_the_oracle():
    """BASELINE.json configs[3] at its quoted size: legacy OP3 model (robot/v0 + op3_conf), 16384 envs with per-env
    contact phases (single/double support) and walking references (step 0.1 x 0.1275 x 0.05 m, 0.7 s,
    ref:legacy/op3_conf.py:9-12): ALL envs against the oracle and its 80-bit build, and feasibility of every solution."""
    s = setup("v0")
    n = 16384
    mask, refs = synth.walking_batch(s["refs"], n, 4, 0.1, 0.1275, 0.05, 0.7, float(s["refs"]["com"][2]))
    res, a, ref = _compare("v0", n, 4, mask, refs)
    _assert_parity(res, "v0")
    ok = a["status"] == 0
    assert ok.mean() > 0.99
    f = a["f"][ok].reshape(-1, 8, 3)
    assert (np.abs(f[:, :, 0]) <= 0.5 * f[:, :, 2] + 1e-5).all() and (np.abs(f[:, :, 1]) <= 0.5 * f[:, :, 2] + 1e-5).all()
    fz = f[:, :, 2].reshape(-1, 2, 4).sum(-1)
    on = np.stack([(mask[ok] & 1) != 0, (mask[ok] & 2) != 0], axis=1)
    assert (fz[on] >= -1e-6).all() and (fz[~on] == 0).all()  # fMin = 0 in the legacy conf


@pytest.mark.parametrize("name,overrides,min_torque,min_jb", [
    ("torque", (("tau_max_scaling", 0.06),), 0.5, 0.0),
    ("jointvel", (("v_max_scaling", 0.09),), 0.0, 0.2),
    ("both", (("tau_max_scaling", 0.08), ("v_max_scaling", 0.1)), 0.3, 0.1),
])
def test_active_actuation_and_joint_velocity_bounds(name, overrides, min_torque, min_jb):
    """SURVEY a7/a8 (ref:ctrl/WalkController.py:168-184): with the reference's limits (50 N m, 100 rad/s) no actuation
    or joint-velocity row is ever active on the synthetic states, so the dense-row path of the active-set kernel
    (M_a | -J_a^T rows entering the working set) would go untested.  Limits tight enough that they bind — 0.6 N m:
    torque rows active in > 50 % of the envs; 0.9 rad/s: joint-velocity rows in > 20 % — on walking inputs with all
    three contact classes; same parity rules."""
    s = setup("v1", overrides=overrides)
    n = 2048
    mask, refs = synth.walking_batch(s["refs"], n, 33, 0.3, 0.2, 0.2, 0.5, float(s["refs"]["com"][2]))
    res, out, ref = _compare("v1", n, 33, mask, refs, overrides=overrides)
    _assert_parity(res)
    assert res["n_optimal"] >= 0.9 * n
    act = res["envs_with_active_force_lf_rf_torque_jointvel_rows"]
    assert act[2] >= min_torque * res["n_optimal"] and act[3] >= min_jb * res["n_optimal"], act
    ok = out["status"] == 0
    tmax = dict(overrides).get("tau_max_scaling", 5.0) * 10.0
    assert (np.abs(out["tau"][ok]) <= tmax * (1 + 1e-9) + 1e-9).all()
    vmax = dict(overrides).get("v_max_scaling", 10.0) * 10.0
    q, v = synth.random_states(s["q0"], n, 33)
    # [UPSTREAM TaskJointBounds, constructed with dt: m_dt = 2 dt] (v_min - v) / (2 dt) <= dv <= (v_max - v) / (2 dt)
    dvj, vj = out["ddq"][ok, 6:], v[ok, 6:]
    assert (dvj <= (vmax - vj) / 0.004 + 1e-7).all() and (dvj >= (-vmax - vj) / 0.004 - 1e-7).all()


def test_mixed_models_two_handles_config5():
    """BASELINE.json configs[4]: robot/v0 and robot/v1 envs side by side (contiguous by model, one handle per model
    on the same device, each with its own constant-memory slot and workspaces): interleaved ticks give what each
    model gives alone."""
    n = 3000
    c1, c0 = _controller("v1", n), _controller("v0", n)
    s1, s0 = setup("v1"), setup("v0")
    q1, v1 = synth.random_states(s1["q0"], n, 9)
    q0, v0 = synth.random_states(s0["q0"], n, 9)
    m1, r1 = synth.walking_batch(s1["refs"], n, 9, 0.3, 0.2, 0.2, 0.5, float(s1["refs"]["com"][2]))
    m0, r0 = synth.walking_batch(s0["refs"], n, 9, 0.1, 0.1275, 0.05, 0.7, float(s0["refs"]["com"][2]))
    solo1, solo0 = _run(c1, q1, v1, m1, r1), _run(c0, q0, v0, m0, r0)
    # interleaved, no synchronisation in between
    dev = c1.device
    c1.contact_mask, c0.contact_mask = torch.as_tensor(m1, device=dev), torch.as_tensor(m0, device=dev)
    c1.refs = {k: torch.as_tensor(np.ascontiguousarray(a), device=dev) for k, a in r1.items()}
    c0.refs = {k: torch.as_tensor(np.ascontiguousarray(a), device=dev) for k, a in r0.items()}
    t = [torch.as_tensor(x, device=dev) for x in (q1, v1, q0, v0)]
    for _ in range(3):
        o1 = c1._tick(t[0], t[1])
        o0 = c0._tick(t[2], t[3])
    torch.cuda.synchronize()
    for k in ("tau", "ddq", "f", "status", "iters"):
        assert np.array_equal(getattr(o1, k).cpu().numpy(), solo1[k]), k
        assert np.array_equal(getattr(o0, k).cpu().numpy(), solo0[k]), k
    assert solo1["tau"].shape[1] == 20 and solo0["tau"].shape[1] == 18
    # and both agree with the oracle (same parity rules as the single-model configs)
    for kind, qq, vv, mm, rr, solo in (("v1", q1, v1, m1, r1, solo1), ("v0", q0, v0, m0, r0, solo0)):
        res, _ = compare_outputs(kind, qq, vv, mm, rr, solo, threads=THREADS)
        parity_log.record(f"test_mixed_models_two_handles_config5[{kind}]", res)
        assert_parity(res, kind)


def test_device_diagnostics_match_the_reference_formulas():
    """tsidb_diagnostics against the formulas of ref:ctrl/WalkController.py:255-289 (CoP, evaluated here in numpy per env
    exactly as the reference writes them) and ref:legacy/biped.py:224-234 (capture point, support polygon)."""
    s = setup("v1")
    n = 500
    ctrl = _controller("v1", n)
    q, v = synth.random_states(s["q0"], n, 17)
    mask = np.array([3, 3, 1, 2, 3] * (n // 5), np.uint8)
    out = _run(ctrl, q, v, mask)
    w = float(np.sqrt(9.80665 / s["refs"]["com"][2]))
    d = ctrl.engine.diagnostics(ctrl.last, ctrl.contact_mask, w)
    torch.cuda.synchronize()
    cop, cp, sup = (d[k].cpu().numpy() for k in ("cop", "capture_point", "support"))
    for i in range(n):
        num, den = np.zeros(2), 0.0
        for f, key in ((0, "foot_lf"), (1, "foot_rf")):
            if not (mask[i] >> f) & 1:
                continue
            wr = out["wrench"][i, 6 * f:6 * f + 6]
            loc = np.array([wr[4] / wr[2], wr[3] / wr[2], 0.0]) if wr[2] > 1e-3 else np.zeros(3)
            p, R = out[key][i, :3], out[key][i, 3:].reshape(3, 3).T
            num += (R @ loc + p)[:2] * wr[2]
            den += wr[2]
        exp = np.r_[num / den, 0.0] if den != 0 else np.zeros(3)
        assert np.abs(cop[i] - exp).max() < 1e-12
        c = out["com"][i]
        assert np.abs(cp[i] - np.r_[c[:2] + c[3:5] / w, 0.0]).max() < 1e-13
        assert np.array_equal(sup[i], np.r_[out["foot_lf"][i, :2], out["foot_rf"][i, :2]])
    # the host-side batched CoP (torch) agrees too
    assert np.abs(ctrl.cop_batch().cpu().numpy() - cop).max() < 1e-12


@pytest.mark.parametrize("n", [1, 7, 33, 1185])
def test_ragged_batch_sizes(n):
    """Batch sizes that do not fill a warp round, a CTA or the persistent grid (148 x 8 + 1): every env still gets
    its own answer, equal to the oracle's."""
    s = setup("v1")
    ctrl = _controller("v1", n)
    q, v = synth.random_states(s["q0"], n, 40 + n)
    mask = np.array([3, 1, 2, 0, 3, 2, 1][:7] * (n // 7 + 1), np.uint8)[:n]
    out = _run(ctrl, q, v, mask)
    idx = np.unique(np.r_[np.arange(min(n, 16)), np.arange(max(0, n - 16), n)])
    ref = s["oracle"].batch(q[idx], v[idx], mask[idx], s["refs"], n_threads=4)
    assert np.array_equal(out["status"][idx], ref["status"])
    ok = ref["status"] == 0
    assert _err(out["tau"][idx][ok], ref["tau"][ok]) < 5e-8 and _err(out["ddq"][idx][ok], ref["dv"][ok]) < 5e-8
    assert np.array_equal(out["iters"][idx][ok] > 0, np.ones(ok.sum(), bool))


def test_api_misuse_is_reported_not_executed():
    """Error behaviour of the boundary: negative return + message, no launch (ref: the tsid binding raises on bad
    sizes; per-env solver failures are NOT errors, they are status values — test_infeasible_envs_...)."""
    import ctypes as C

    from tsid_control_b200._capi import TsidbError, TsidbRefs

    s = setup("v1")
    n = 8
    ctrl = _controller("v1", n)
    e = ctrl.engine
    q, v = synth.random_states(s["q0"], 2 * n, 3)
    dev = e.device
    with pytest.raises(TsidbError, match="max_envs"):
        e.compute(torch.as_tensor(q, device=dev), torch.as_tensor(v, device=dev))          # batch larger than the handle
    with pytest.raises((TypeError, ValueError)):
        e.compute(torch.as_tensor(q[:n, :-1].copy(), device=dev), torch.as_tensor(v[:n], device=dev))  # wrong nq
    with pytest.raises((TypeError, ValueError)):
        e.compute(torch.as_tensor(q[:n], device=dev).float(), torch.as_tensor(v[:n], device=dev))      # wrong dtype
    r = TsidbRefs()
    rc = e.lib.tsidb_compute(e.h, n, 7, 0, 0, None, C.byref(r), 0, 0, 0, 0, 0, None, None, None)     # null pointers, bad layout
    assert rc < 0 and e.lib.tsidb_last_error()
    assert e.lib.tsidb_compute_host(e.h, 0, 0, 0, None, None, 0, 0, 0, 0, 0, None) < 0
    with pytest.raises(TsidbError, match="gait_reset"):
        e._gait_n = n
        e.rollout(torch.as_tensor(q[:n], device=dev), torch.as_tensor(v[:n], device=dev), 2)
    # the handle still works afterwards
    out = _run(ctrl, q[:n], v[:n], np.full(n, 3, np.uint8))
    assert (out["status"] == 0).all()


def test_class_chains_on_streams_equal_the_single_stream_tick(monkeypatch):
    """The three contact-class chains E -> G -> A run on forked streams by default; TSIDB_CLASS_STREAMS=0 (read when
    the handle is created) keeps every kernel on the caller's stream.  Both orders of execution give bit-identical
    results on a mixed batch, also when ticks follow each other without a synchronisation in between."""
    s = setup("v1")
    n = 3000
    q, v = synth.random_states(s["q0"], n, 77)
    mask, refs = synth.walking_batch(s["refs"], n, 77, 0.3, 0.2, 0.2, 0.5, float(s["refs"]["com"][2]))
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("TSIDB_CLASS_STREAMS", flag)
        ctrl = _controller("v1", n)
        dev = ctrl.device
        ctrl.contact_mask = torch.as_tensor(mask, device=dev)
        ctrl.refs = {k: torch.as_tensor(np.ascontiguousarray(a), device=dev) for k, a in refs.items()}
        qd, vd = torch.as_tensor(q, device=dev), torch.as_tensor(v, device=dev)
        for _ in range(3):  # back to back: the next tick's class sort must wait for the side streams of this one
            ctrl._tick(qd, vd, aux=True)
        torch.cuda.synchronize()
        o = ctrl.last
        outs.append({k: getattr(o, k).cpu().numpy() for k in o.__slots__})
    for k in ("tau", "ddq", "f", "status", "iters", "active_set"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    assert (outs[0]["status"] == 0).mean() > 0.99


def test_longest_first_order_of_the_class_sort_changes_no_result(monkeypatch):
    """tsidb_set_sched_hint: from the second tick of a handle on, the class sort orders the envs of a contact class by their
    iteration counts in the previous tick.  The slot an env lands in must not matter: every output of a second tick with
    the hint equals the tick of a handle without it bit for bit (and the hint can be switched per handle at run time)."""
    kind, n = "v1", 6000
    s = setup(kind)
    q, v = synth.random_states(s["q0"], n, 77)
    mask, refs = synth.walking_batch(s["refs"], n, 77, 0.3, 0.2, 0.2, 0.5, float(s["refs"]["com"][2]))
    mask[11] = 0
    monkeypatch.setenv("TSIDB_SMALL_N", "0")
    monkeypatch.setenv("TSIDB_SCHED_HINT", "0")
    plain = _run(_controller(kind, n), q, v, mask, refs)
    monkeypatch.setenv("TSIDB_SCHED_HINT", "1")
    ctrl = _controller(kind, n)
    first = _run(ctrl, q, v, mask, refs)    # no previous tick: one bucket per class
    second = _run(ctrl, q, v, mask, refs)   # ordered by the first tick's iteration counts
    ctrl.engine.set_sched_hint(False)
    third = _run(ctrl, q, v, mask, refs)
    assert int(second["iters"].max()) >= 12  # several buckets are in use
    for out in (first, second, third):
        for k in plain:
            if plain[k] is not None:
                assert np.array_equal(out[k], plain[k], equal_nan=True), k


@pytest.mark.parametrize("tag,rr", [("r50", 0.5), ("r10", 0.1)])
def test_device_foot_trajectory_kernel_against_the_reference_golden(tag, rr):
    """tsidb_foot_trajectory (CUDA) against the samples of the reference's own scipy CubicSplines
    (tests/golden/planners.npz, generated by importing ref:ctrl/Foot_Trajectory.py): 3-knot and 4-knot z, linear x/y/yaw."""
    from test_planners import _device_foot_trajectory

    ctrl = _controller("v1", 1)
    e, dev = ctrl.engine, ctrl.device

    def run(t0, t1, start, target, h, r, t):
        out = e.foot_trajectory(t0, t1, torch.as_tensor(start, device=dev), torch.as_tensor(target, device=dev), h, r,
                                torch.as_tensor(t, device=dev))
        torch.cuda.synchronize()
        return out.cpu().numpy().reshape(len(t), 16)

    _device_foot_trajectory(run, tag, rr)


def test_device_footstep_plan_kernel_against_the_reference_golden():
    """tsidb_footstep_plan (CUDA) against the footsteps the reference's FootstepPlanner.plan produces for its own demo path
    (ref:ctrl/Footstep_Planner.py:127-150) and for a second, curved path — positions, yaw and sides."""
    from test_planners import _device_footstep_plan

    ctrl = _controller("v1", 1)
    e, dev = ctrl.engine, ctrl.device

    def run(path, n_pts, init, L, W, max_steps):
        steps, ns = e.footstep_plan(torch.as_tensor(path, device=dev), torch.as_tensor(init, device=dev), L, W,
                                    n_pts=torch.as_tensor(n_pts, device=dev), max_steps=max_steps)
        torch.cuda.synchronize()
        return steps.cpu().numpy(), ns.cpu().numpy()

    _device_footstep_plan(run)


def test_device_rollout_along_a_device_made_footstep_plan():
    """f1 end to end on the device: tsidb_footstep_plan (FootstepPlanner.plan per env) -> tsidb_gait_set_plan -> tsidb_rollout
    (tick -> integrate -> phase machine whose swing references are FootTrajectory with yaw and a 4-knot z spline), one CUDA
    graph replayed — against the same loop driven from the host with the numpy restatement (tests/gait_ref.py, built on the
    host planner classes that tests/golden/planners.npz pins to the reference's own Python)."""
    from gait_ref import GaitRef
    from tsid_control_b200.ctrl.Footstep_Planner import Footstep, FootstepPlanner

    s = setup("v1")
    n, steps_n = 48, 60
    ctrl = _controller("v1", n)
    e, dev, conf = ctrl.engine, ctrl.device, s["conf"]
    rng = np.random.Generator(np.random.PCG64(77))
    q, v = synth.random_states(s["q0"], n, 19)
    v *= 0.2
    phase0 = rng.uniform(0, 1, n)
    lf, rf = ctrl.default_refs["foot_lf"], ctrl.default_refs["foot_rf"]
    P = 40
    path = np.zeros((n, P, 2))
    for i in range(n):
        w, x, y, th = rng.uniform(-0.5, 0.5), 0.0, 0.0, 0.0
        for k in range(P):
            x += 0.04 * np.cos(th); y += 0.04 * np.sin(th); th += w * 0.1
            path[i, k] = (x, y)
    init = np.tile(np.array([[lf[0], lf[1], 0, 0], [rf[0], rf[1], 0, 1]]), (n, 1, 1))
    h0 = float(ctrl.default_refs["com"][2])
    gait = dict(dt=conf.dt, step_duration=conf.step_duration, step_length=conf.step_length, step_height=0.05, com_height=h0)
    e.gait_reset(n, phase0=torch.as_tensor(phase0, device=dev), **gait)
    steps_d, ns_d = e.footstep_plan(torch.as_tensor(path, device=dev), torch.as_tensor(init, device=dev), 0.1, 2 * abs(lf[1]))
    e.gait_set_plan(steps_d, ns_d, rise_ratio=0.3)
    torch.cuda.synchronize()
    steps_h, ns_h = steps_d.cpu().numpy(), ns_d.cpu().numpy()
    # the device plan is the host planner's plan (itself pinned to the reference)
    fs = FootstepPlanner(step_width=2 * abs(lf[1]), step_length=0.1).plan(
        [p for p in path[0]], [Footstep(init[0, 0, :2], np.array([0, 0, 0.0]), 0), Footstep(init[0, 1, :2], np.array([0, 0, 0.0]), 1)])
    assert ns_h[0] == len(fs) and np.abs(steps_h[0, :len(fs), :2] - np.array([f.position for f in fs])).max() < 1e-14
    qd, vd = torch.as_tensor(q, device=dev).clone(), torch.as_tensor(v, device=dev).clone()
    e.rollout(qd, vd, steps_n, use_graph=True)
    torch.cuda.synchronize()
    gs = {k: t.cpu().numpy().copy() for k, t in e.gait_state().items()}
    # host-driven
    g = GaitRef(n, defaults=ctrl.default_refs, phase0=phase0, steps=steps_h, n_steps=ns_h, rise_ratio=0.3, **gait)
    qh, vh = torch.as_tensor(q, device=dev).clone(), torch.as_tensor(v, device=dev).clone()
    post = torch.as_tensor(np.tile(ctrl.default_refs["posture"], (n, 1)), device=dev)
    for _ in range(steps_n):
        refs = {k: torch.as_tensor(np.ascontiguousarray(a), device=dev) for k, a in g.refs().items()}
        refs["posture"] = post
        o = e.compute(qh, vh, torch.as_tensor(g.mask, device=dev), refs, aux=True)
        e.integrate(qh, vh, o.ddq, conf.dt)
        g.step(o.foot_lf.cpu().numpy(), o.foot_rf.cpu().numpy(), o.status.cpu().numpy())
    torch.cuda.synchronize()
    assert np.array_equal(gs["mask"], g.mask) and np.array_equal(gs["fails"], g.fails)
    ok = g.fails == 0
    assert ok.mean() > 0.9
    assert np.abs(qd.cpu().numpy()[ok] - qh.cpu().numpy()[ok]).max() < 1e-8
    for k in ("com", "foot_lf", "foot_rf", "contact_lf", "contact_rf"):
        assert np.abs(gs[k][ok] - g.refs()[k][ok]).max() < 1e-8, k
    assert np.abs(gs["foot_lf"][:, 17]).max() + np.abs(gs["foot_rf"][:, 17]).max() > 0  # yaw-rate references are in play
