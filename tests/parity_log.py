"""Durable record of the GPU parity measurements: every comparison of the CUDA tick with the oracle appends its
numbers (errors vs the fp64 oracle and vs its 80-bit build, working-set and iteration agreement) to
gpurun_out/parity_r02.json — the one directory that travels back from the GPU box.  The copy under profiles/ is
the committed evidence (tools/collect_parity.py copies it)."""
from __future__ import annotations

import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "gpurun_out", "parity_r02.json")
_LOG: dict = {}


def record(name: str, res: dict) -> None:
    try:
        import torch

        gpu = torch.cuda.get_device_name(0) if torch.cuda.is_available() else None
    except Exception:
        gpu = None
    _LOG[name] = {k: (v if not hasattr(v, "item") else v.item()) for k, v in res.items()}
    doc = {"written": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "gpu": gpu,
           "tolerance": "err = |a-b| / (1e-2 + |b|), bar 1e-8 (north_star: 1e-8 relative / 1e-10 absolute)",
           "cases": _LOG}
    try:
        os.makedirs(os.path.dirname(PATH), exist_ok=True)
        prev = {}
        if os.path.exists(PATH):
            try:
                prev = json.load(open(PATH)).get("cases", {})
            except Exception:
                prev = {}
        prev.update(_LOG)
        doc["cases"] = prev
        with open(PATH, "w") as fh:
            json.dump(doc, fh, indent=1, sort_keys=True)
    except OSError:
        pass
