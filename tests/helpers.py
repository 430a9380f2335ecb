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
    """Fixtures of the UNMODIFIED reference solver (tests/golden/make_golden.py)."""
    return sorted(p.stem for p in GOLDEN.glob("*.npz") if not p.stem.startswith(("normalize", "tv_", "tvmm_")))


def tv_golden_names():
    """TV() stencil fixtures of the patched reference (tests/golden/make_golden_tv.py)."""
    return sorted(p.stem for p in GOLDEN.glob("tv_*.npz"))


def tvmm_golden_names():
    """Solver fixtures of the patched reference with the TV term alive (tests/golden/make_golden_tv.py)."""
    return sorted(p.stem for p in GOLDEN.glob("tvmm_*.npz"))


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


def smooth_case(M, N, K, blind, seed, box=5):
    """Box-smoothed scene + a little noise: the non-blind whiteness statistic turns around after ~20
    outer iterations, so the reference's stop rule (lib/deconvolution.pyx:643-654) fires.
    Returns (image, u0, psf0, window) as the golden-vector generator and the row-band worker use them."""
    from image_cases_studies_b200 import synthetic
    from image_cases_studies_b200.lib import utils
    rng = np.random.default_rng(seed)
    p = K // 2
    s = 0.1 + 0.8 * rng.random((M + 2 * p, N + 2 * p, 3), dtype=np.float32)
    # separable box filter in plain numpy (no scipy.ndimage dependency)
    ker = np.ones(box) / box
    for ax in (0, 1):
        s = np.apply_along_axis(lambda v: np.convolve(np.pad(v, box // 2, mode="reflect"), ker, mode="valid"), ax, s)
    s = s.astype(np.float32)
    kt = utils.stack3(utils.gaussian_kernel(K, K / 4))
    image = synthetic._valid_conv_fft(s, kt)
    image += (0.002 * rng.standard_normal(image.shape)).astype(np.float32)
    u0 = np.ascontiguousarray(np.pad(image, ((p, p), (p, p), (0, 0)), mode="edge"))
    psf0 = utils.stack3(utils.uniform_kernel(K)) if blind else kt
    return image, u0, psf0, synthetic.default_window(M, N, p)
