"""Shared helpers for the parity tests."""
from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

# Parity tolerances (BASELINE.json north_star): relative L2 <= 1e-4 on the image, L1 <= 1e-4 on the PSF.
TOL_IMAGE_REL_L2 = 1e-4
TOL_PSF_L1 = 1e-4


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def psf_l1(a, b) -> float:
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).sum())


def golden_names():
    return sorted(p.stem for p in GOLDEN.glob("*.npz") if not p.stem.startswith("normalize"))


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    d = {k: z[k] for k in z.files}
    d["window"] = tuple(int(v) for v in d["window"])
    for k in ("tau", "step_factor", "lambd"):
        d[k] = float(d[k])
    for k in ("iterations", "ref_iterations"):
        d[k] = int(d[k])
    for k in ("blind", "correlation"):
        d[k] = bool(d[k])
    return d
