"""Drop-in for the reference's ``lib/deconvolution.pyx`` solver module, running on a B200.

    from image_cases_studies_b200.lib import deconvolution as dc      # was: from lib import deconvolution as dc

Same public names (``richardson_lucy_MM``, ``normalize_kernel``, ``DTYPE``; lib/deconvolution.pyx:31,:73,:341),
same positional/keyword arguments, same in-place contract: ``u`` is updated in place including its pad
ring, ``psf`` is refined in place, and the return value is the VIEW ``u[pad:pad+M, pad:pad+N]`` of the
caller's array (pyx:675).  All arithmetic runs in the CUDA kernels behind include/rltv_b200.h; there is
no CPU path.  ``p, norm, order, priority, refocus, C`` are accepted and ignored, as in the reference's
live arithmetic.
"""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from ..solver import Solver

DTYPE = np.float32

VERBOSE = False          # the reference prints progress (pyx:593,:648,:659,:665-669); off by default here
last_stats: dict | None = None

_cache: dict = {}
_CACHE_MAX = 2


def _check_buffer(a, name):
    if not isinstance(a, np.ndarray):
        raise TypeError(f"Argument '{name}' has incorrect type (expected numpy.ndarray, got {type(a).__name__})")
    if a.dtype != np.float32:
        raise ValueError(f"Buffer dtype mismatch, expected 'DTYPE_t' but got '{a.dtype.name}'")
    if a.ndim != 3:
        raise ValueError(f"Buffer has wrong number of dimensions (expected 3, got {a.ndim})")


def _solver(M, N, MK) -> Solver:
    key = (M, N, MK)
    s = _cache.pop(key, None)
    if s is None:
        while len(_cache) >= _CACHE_MAX:
            _cache.pop(next(iter(_cache))).close()
        s = Solver(M, N, MK)
    _cache[key] = s
    return s


def clear_cache():
    """Release the cached device contexts (each holds ~72 bytes per padded pixel of HBM)."""
    while _cache:
        _cache.popitem()[1].close()


def richardson_lucy_MM(image, u, psf, top, bottom, left, right, tau, M, N, C, MK, iterations, step_factor, lambd,
                       blind=True, correlation=False, p=1., norm=1, order=2, priority=0, refocus=0, mode="mm"):
    """Richardson-Lucy blind / non-blind deconvolution by Majorization-Minimization (lib/deconvolution.pyx:341-675).

    ``mode`` is an extension (not an argument of the reference): "mm" -- the default -- is the reference's shipped
    arithmetic, in which the TV regulariser is computed but never used; "mm_tv" runs the TV branches of
    pyx:516-517 / :542-549 as they would if the commented ``TV(ut, ...)`` calls of pyx:464-465 were alive and wrote
    ``TV_ut_L1`` / ``TV_ut_L2`` (pinned against a patched build of the reference, oracle/build_ref_tv.py).  As in that
    code path the blurry ``image`` is then denoised IN PLACE."""
    global last_stats
    _check_buffer(image, "image")
    _check_buffer(u, "u")
    _check_buffer(psf, "psf")
    M, N, MK = int(M), int(N), int(MK)
    if image.shape != (M, N, 3):
        raise ValueError(f"image has shape {image.shape}, expected (M, N, 3) = {(M, N, 3)}")
    if psf.shape != (MK, MK, 3):
        raise ValueError(f"psf has shape {psf.shape}, expected (MK, MK, 3) = {(MK, MK, 3)}")
    if MK < 3 or MK % 2 == 0:
        raise ValueError("MK must be odd and >= 3")
    if u.shape != (M + MK - 1, N + MK - 1, 3):
        raise ValueError(f"u has shape {u.shape}, expected (M+MK-1, N+MK-1, 3) = {(M + MK - 1, N + MK - 1, 3)}")
    pad = (u.shape[0] - M) // 2                                         # pyx:376
    s = _solver(M, N, MK)
    s.upload(image, u, psf)
    params = Solver.make_params((top, bottom, left, right), tau, iterations, step_factor, lambd, blind, correlation, mode)
    last_stats = s.solve(params)
    s.download(u=u, psf_caller=psf if blind else None)
    if mode == "mm_tv":
        s.download_image(image)                                         # pyx:549 writes the caller's image
    if VERBOSE:
        it = last_stats["iterations"]
        if last_stats["stopped"]:
            print("white autocorellation condition met")
            print("Convergence after %i iterations." % it)
        else:
            print("Did not converge after %i iterations. Don't use the result." % it)
    return u[pad:pad + M, pad:pad + N, ...]                             # pyx:675


def normalize_kernel(kern, MK):
    """Clip negative taps to zero and normalise every channel to sum 1, in place (lib/deconvolution.pyx:73-75)."""
    _check_buffer(kern, "kern")
    MK = int(MK)
    if kern.shape != (MK, MK, 3):
        raise ValueError(f"kern has shape {kern.shape}, expected {(MK, MK, 3)}")
    work = kern if kern.flags.c_contiguous else np.ascontiguousarray(kern)
    from ..solver import default_device
    nat.check(nat.lib.rltv_normalize_kernel(nat.ptr(work), MK, default_device()))
    if work is not kern:
        kern[...] = work
    return None
