"""The oracle against its frozen vectors (tests/golden/tick_*.npz, made by tests/golden/make_tick_golden.py).
These pin the oracle, not the reference binaries (parity unpinned)."""
import os

import numpy as np
import pytest

from common import setup

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("kind,masks", [("v1", (3, 1, 2, 0)), ("v0", (3, 1))])
def test_oracle_reproduces_golden(kind, masks):
    g = np.load(os.path.join(HERE, "golden", f"tick_{kind}.npz"))
    s = setup(kind)
    assert np.abs(g["q0"] - s["q0"]).max() < 1e-15
    refs = {k: g["ref_" + k] for k in ("com", "foot_lf", "foot_rf", "contact_lf", "contact_rf", "posture")}
    n = g["q"].shape[0]
    for m in masks:
        res = s["oracle"].batch(g["q"], g["v"], np.full(n, m, np.uint8), refs)
        assert np.array_equal(res["status"], g[f"m{m}_status"]) and np.array_equal(res["iters"], g[f"m{m}_iters"])
        for k in ("tau", "dv", "com", "foot"):
            assert np.abs(res[k] - g[f"m{m}_{k}"]).max() < 1e-9, (m, k)
        assert np.abs(res["f"] - g[f"m{m}_f"]).max() < 1e-6
