"""PSF builders -- drop-in for the kernel constructors of the reference's ``lib/utils.py:134-170``.

Same names, arguments and return values (float64 ``(size, size)`` arrays that sum to 1; the driver
stacks them to 3 channels, ``deconvolve.py:178-179``).  ``scipy.signal.gaussian`` /
``scipy.signal.exponential`` (used by the reference, removed from scipy >= 1.13) are restated in
closed form so nothing here depends on scipy.
"""
from __future__ import annotations

import numpy as np


def _gaussian_window(M: int, std: float) -> np.ndarray:
    n = np.arange(M, dtype=np.float64) - (M - 1) / 2.0
    return np.exp(-0.5 * (n / std) ** 2)


def _exponential_window(M: int, tau: float) -> np.ndarray:
    n = np.arange(M, dtype=np.float64) - (M - 1) / 2.0
    return np.exp(-np.abs(n) / tau)


def disc_blur(x):
    """lib/utils.py:134-136."""
    return [1 / (np.pi * i ** 2) for i in range(1, int(x / 2) + 1)]


def lens_blur(size):
    """lib/utils.py:139-143 (note: the reference's window has int(size/2) taps)."""
    window = disc_blur(size)
    kern = np.outer(window, window)
    return kern / kern.sum()


def uniform_kernel(size):
    """lib/utils.py:146-149."""
    kern = np.ones((size, size))
    kern /= np.sum(kern)
    return kern


def gaussian_kernel(radius, std):
    """lib/utils.py:152-156."""
    window = _gaussian_window(radius, std)
    kern = np.outer(window, window)
    return kern / kern.sum()


def kaiser_kernel(radius, beta):
    """lib/utils.py:159-163."""
    window = np.kaiser(radius, beta)
    kern = np.outer(window, window)
    return kern / kern.sum()


def poisson_kernel(radius, tau):
    """lib/utils.py:166-170."""
    window = _exponential_window(radius, tau)
    kern = np.outer(window, window)
    return kern / kern.sum()


def motion_kernel(size, angle_deg=30.0):
    """A line ("motion blur") PSF.  NOT in the reference (``blur="motion"`` only sets the channel
    correlation flag, ``deconvolve.py:154-157``); defined here for the 31x31 synthetic workload."""
    kern = np.zeros((size, size))
    c = (size - 1) / 2.0
    t = np.linspace(-c, c, 8 * size)
    a = np.deg2rad(angle_deg)
    ys = np.clip(np.rint(c + t * np.sin(a)).astype(int), 0, size - 1)
    xs = np.clip(np.rint(c + t * np.cos(a)).astype(int), 0, size - 1)
    np.add.at(kern, (ys, xs), 1.0)
    return kern / kern.sum()


def stack3(kern) -> np.ndarray:
    """``np.dstack((psf, psf, psf))`` as float32 C-contiguous (deconvolve.py:179, then ``.astype(np.float32)`` :249)."""
    k = np.asarray(kern)
    return np.ascontiguousarray(np.dstack((k, k, k)), dtype=np.float32)
