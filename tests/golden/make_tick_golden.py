#!/usr/bin/env python
"""Freeze oracle outputs as golden vectors for the tick (tests/golden/tick_*.npz).

PARITY UNPINNED: the reference's own tick cannot run in this image (pinocchio/tsid/eiquadprog are not
installable) and it ships no test vectors, so these files pin the ORACLE (oracle/), not the
reference binaries.  They exist to catch regressions of the oracle itself; if a real tsid install ever shows up
under baseline/_ref, regenerate them from it.

    python tests/golden/make_tick_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

from common import setup  # noqa: E402
from tsid_control_b200 import synth  # noqa: E402


def main():
    for kind, masks in (("v1", (3, 1, 2, 0)), ("v0", (3, 1))):
        s = setup(kind)
        n = 6
        q, v = synth.random_states(s["q0"], n, 99)
        out = {"q": q, "v": v, "q0": s["q0"]}
        for k, a in s["refs"].items():
            out["ref_" + k] = a
        for m in masks:
            res = s["oracle"].batch(q, v, np.full(n, m, np.uint8), s["refs"])
            for k in ("tau", "dv", "f", "status", "iters", "com", "foot"):
                out[f"m{m}_{k}"] = res[k]
        np.savez_compressed(os.path.join(HERE, f"tick_{kind}.npz"), **out)
        print("wrote", f"tick_{kind}.npz")


if __name__ == "__main__":
    main()
