"""Helpers shared by the GPU parity tests: error measures and a JSON report of what was measured."""
import json
import os
from pathlib import Path

import numpy as np

REPORT = Path(os.environ.get("B200ID_PARITY_REPORT", Path(__file__).resolve().parent.parent / "gpurun_out" / "parity_report.json"))
_entries = {}


def rel_max(a, b):
    """max |a-b| / max |b| (norm-wise relative error)."""
    a, b = np.asarray(a), np.asarray(b)
    den = float(np.max(np.abs(b))) if b.size else 1.0
    return float(np.max(np.abs(a - b))) / (den if den > 0 else 1.0)


def rel_each(a, b):
    """max_k |a_k-b_k| / |b_k| (element-wise relative error)."""
    a, b = np.asarray(a), np.asarray(b)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(a - b) / np.abs(b)
    r = r[np.isfinite(r)]
    return float(np.max(r)) if r.size else 0.0


def record(name, value):
    _entries[name] = value
    try:
        REPORT.parent.mkdir(parents=True, exist_ok=True)
        old = json.loads(REPORT.read_text()) if REPORT.exists() else {}
        old.update(_entries)
        REPORT.write_text(json.dumps(old, indent=1, sort_keys=True))
    except OSError:
        pass
    return value
