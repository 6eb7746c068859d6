"""CPU checks of the CUDA kernel LOGIC: tsidb_kernels.cuh is compiled with g++ against tests/emu/emu_cuda.h,
which runs the 32 lanes of a warp as lock-step fibers (debug/test infrastructure, not a fallback: it is not in
libtsidb.so).  Both per-env kernel bodies run back to back: prepare_env (dynamics, assembly, equality
elimination) and activeset_env (Goldfarb-Idnani iterations, decode)."""
import numpy as np
import pytest

from common import bits_to_rows, canonical_active, setup
from emu_py import Emu, active_bits
from tsid_control_b200 import synth


def _err(a, b):
    return float((np.abs(a - b) / (1e-2 + np.abs(b))).max())


@pytest.mark.parametrize("kind,maskval,n", [("v1", 3, 24), ("v1", 1, 16), ("v1", 2, 16), ("v1", 0, 8), ("v0", 3, 16), ("v0", 1, 12)])
def test_emulated_kernel_matches_oracle(kind, maskval, n):
    s = setup(kind)
    orc = s["oracle"]
    emu = Emu(s["cm"], s["cc"], s["refs"])
    q, v = synth.random_states(s["q0"], n, 21)
    mask = np.full(n, maskval, np.uint8)
    out = emu.tick(q, v, mask)
    canon = 0
    for i in range(n):
        r = orc.tick(q[i], v[i], maskval, s["refs"])
        assert out["status"][i] == r["status"] == 0
        assert _err(out["tau"][i], r["tau"]) < 5e-8 and _err(out["ddq"][i], r["dv"]) < 5e-8
        assert _err(out["f"][i], r["f"]) < 1e-5
        assert np.abs(out["com"][i] - r["com"]).max() < 1e-12
        assert np.abs(out["foot_lf"][i] - r["foot"][0]).max() < 1e-13
        rows = orc.ci_rows(maskval)
        ra = set(rows[k] for k in r["active"])
        rb = set(bits_to_rows(orc.na, orc.nv, active_bits(out["active"][:, i])))
        canon += canonical_active(ra) == canonical_active(rb)
    assert canon >= n - 1


def test_emulated_kernel_per_env_refs_soa_layout_and_mixed_masks():
    s = setup("v1")
    orc = s["oracle"]
    emu = Emu(s["cm"], s["cc"], s["refs"])
    n = 12
    q, v = synth.random_states(s["q0"], n, 22)
    mask, refs = synth.walking_batch(s["refs"], n, 22, 0.3, 0.2, 0.2, 0.5, float(s["refs"]["com"][2]))
    a = emu.tick(q, v, mask, refs, layout=0)
    b = emu.tick(q, v, mask, refs, layout=1)
    for k in ("tau", "ddq", "f", "status", "iters"):
        assert np.array_equal(a[k], b[k]), k
    for i in range(n):
        r = orc.tick(q[i], v[i], int(mask[i]), {k: x[i] for k, x in refs.items()})
        assert a["status"][i] == r["status"]
        assert _err(a["tau"][i], r["tau"]) < 5e-8 and _err(a["ddq"][i], r["dv"]) < 5e-8


def test_emulated_kernel_reports_infeasible_env():
    s = setup("v1")
    emu = Emu(s["cm"], s["cc"], s["refs"])
    q, v = synth.random_states(s["q0"], 3, 8)
    v[1] *= 400.0
    out = emu.tick(q, v, np.full(3, 3, np.uint8))
    ref = s["oracle"].batch(q, v, np.full(3, 3, np.uint8), s["refs"])
    assert np.array_equal(out["status"], ref["status"]) and out["status"][1] != 0
    assert np.all(out["tau"][1] == 0.0) and np.all(out["ddq"][1] == 0.0)


def test_emulated_kinematics_only():
    s = setup("v0")
    emu = Emu(s["cm"], s["cc"], s["refs"])
    q, v = synth.random_states(s["q0"], 4, 23)
    out = emu.tick(q, v, np.full(4, 3, np.uint8), kin_only=True)
    for i in range(4):
        r = s["oracle"].tick(q[i], v[i], 3, s["refs"])
        assert np.abs(out["com"][i] - r["com"]).max() < 1e-12
        assert np.abs(out["foot_rf"][i] - r["foot"][1]).max() < 1e-13


def test_emulated_kernel_with_active_torque_bounds():
    """Torque limits tight enough that actuation rows enter the working set: exercises the dense-row path and the
    slack lower bound that lets the solver skip actuation rows it can prove satisfied."""
    import copy
    import ctypes as C

    from oracle_py import Oracle

    s = setup("v1")
    cc = copy.deepcopy(s["cc"]) if False else type(s["cc"]).from_buffer_copy(bytes(s["cc"]))
    for i in range(20):
        cc.tau_max[i] = 0.6
        cc.tau_min[i] = -0.6
    orc = Oracle(s["cm"], cc, "liboracle.so")
    emu = Emu(s["cm"], cc, s["refs"])
    n = 20
    q, v = synth.random_states(s["q0"], n, 33)
    v *= 0.3
    mask = np.array([3, 1, 2, 3] * 5, np.uint8)
    out = emu.tick(q, v, mask)
    hit = 0
    for i in range(n):
        r = orc.tick(q[i], v[i], int(mask[i]), s["refs"])
        assert out["status"][i] == r["status"]
        if r["status"] != 0:
            continue
        assert _err(out["tau"][i], r["tau"]) < 5e-7 and _err(out["ddq"][i], r["dv"]) < 5e-7
        rows = orc.ci_rows(int(mask[i]))
        ra = set(rows[k] for k in r["active"])
        rb = set(bits_to_rows(orc.na, orc.nv, active_bits(out["active"][:, i])))
        assert canonical_active(ra) == canonical_active(rb)
        hit += any(blk == 2 for blk, _, _ in ra)
        assert (np.abs(out["tau"][i]) <= 0.6 + 1e-7).all()
    assert hit >= 5  # torque rows are really active in this batch


@pytest.mark.parametrize("overrides", [(("v_max_scaling", 0.09),), (("tau_max_scaling", 0.08), ("v_max_scaling", 0.1))])
def test_emulated_kernel_with_active_joint_velocity_bounds(overrides):
    """Joint-velocity limits of 0.9 / 1.0 rad/s (the synthetic joint rates are U(-1, 1)): TaskJointBounds rows
    (ref:ctrl/WalkController.py:178-184) end in the working set; the GPU tests' parity rules on the emulated kernels."""
    from common import assert_parity, compare_outputs

    s = setup("v1", overrides=overrides)
    emu = Emu(s["cm"], s["cc"], s["refs"])
    n = 48
    q, v = synth.random_states(s["q0"], n, 33)
    mask, refs = synth.walking_batch(s["refs"], n, 33, 0.3, 0.2, 0.2, 0.5, float(s["refs"]["com"][2]))
    out = emu.tick(q, v, mask, refs)
    out["active_set"], out["lam"], out["lam_row"] = out["active"], out["lambda"], out["lambda_row"]
    res, ref = compare_outputs("v1", q, v, mask, refs, out, overrides=overrides)
    assert res["lambda_envs_compared"] >= 30
    assert_parity(res, "v1")
    assert res["envs_with_active_force_lf_rf_torque_jointvel_rows"][3] >= 8
