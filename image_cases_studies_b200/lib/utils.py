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


# ---- timing / output helpers of lib/utils.py used by the driver -----------------------------------------------
def timeit(method):
    """lib/utils.py:30-42: print the wall time of a call."""
    import functools
    import time

    @functools.wraps(method)
    def timed(*args, **kw):
        ts = time.time()
        result = method(*args, **kw)
        print("%r %2.2f sec" % (method.__name__, time.time() - ts))
        return result
    return timed


def write_tiff_rgb16(path, pic) -> None:
    """Minimal baseline TIFF writer: uncompressed little-endian 16-bit RGB, one strip.

    Stands in for the vendored ``tifffile.imsave(..., dtype=np.uint16, photometric='rgb')`` of lib/utils.py:303-312
    (the 9 000-line tifffile module is file I/O, out of scope)."""
    import struct
    a = np.ascontiguousarray(np.asarray(pic).astype(np.uint16))
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected an (H, W, 3) array")
    h, w = a.shape[:2]
    data = a.astype("<u2").tobytes()
    n_tags = 10
    ifd_off = 8
    bps_off = ifd_off + 2 + n_tags * 12 + 4
    data_off = bps_off + 6
    tags = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 3, bps_off), (259, 3, 1, 1), (262, 3, 1, 2), (273, 4, 1, data_off),
            (277, 3, 1, 3), (278, 4, 1, h), (279, 4, 1, len(data)), (284, 3, 1, 1)]
    with open(path, "wb") as f:
        f.write(struct.pack("<2sHI", b"II", 42, ifd_off))
        f.write(struct.pack("<H", n_tags))
        for tag, typ, cnt, val in tags:
            f.write(struct.pack("<HHI", tag, typ, cnt))
            f.write(struct.pack("<HH", val, 0) if (typ == 3 and cnt == 1) else struct.pack("<I", val))
        f.write(struct.pack("<I", 0))
        f.write(struct.pack("<HHH", 16, 16, 16))
        f.write(data)


def save(pic, name, dest_path):
    """lib/utils.py:303-312: 16-bit RGB TIFF named ``<name>.tif`` under ``dest_path``."""
    from os.path import join
    write_tiff_rgb16(join(dest_path, name + ".tif"), pic)
