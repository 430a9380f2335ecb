"""Seeded synthetic workloads of the shapes BASELINE.json names.

``make_case`` returns host (numpy, float32, HWC) arrays exactly as a caller of
``richardson_lucy_MM`` would hold them: the blurry ``image`` (M,N,3), the edge-padded initial
estimate ``u0`` (M+2p, N+2p, 3) (``deconvolve.py:303``), the initial PSF and the whiteness window
``(p+1, 255-p-1, p+1, 255-p-1)`` (``deconvolve.py:308``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .lib import utils


@dataclass
class Case:
    name: str
    image: np.ndarray
    u0: np.ndarray
    psf0: np.ndarray
    psf_true: np.ndarray
    window: tuple
    blind: bool
    iterations: int
    tau: float
    step_factor: float = 1e-3      # quality="normal", deconvolve.py:106-107
    lambd: float = 10000.0         # confidence=10 -> lambd = confidence*1000, deconvolve.py:66,:200

    @property
    def MK(self):
        return self.psf0.shape[0]

    @property
    def shape(self):
        return self.image.shape[:2]


def _valid_conv_fft(s: np.ndarray, k: np.ndarray) -> np.ndarray:
    """Per-channel 'valid' true convolution in float64 FFT, returned float32 (input generation only)."""
    H, W = s.shape[:2]
    K = k.shape[0]
    out = np.empty((H - K + 1, W - K + 1, 3), dtype=np.float32)
    for c in range(3):
        fa = np.fft.rfft2(s[..., c].astype(np.float64), (H, W))
        fb = np.fft.rfft2(k[..., c].astype(np.float64), (H, W))
        full = np.fft.irfft2(fa * fb, (H, W))       # circular == linear on the valid region
        out[..., c] = full[K - 1:, K - 1:]
    return out


def make_inputs(M: int, N: int, psf_true: np.ndarray, seed: int = 0):
    K = psf_true.shape[0]
    p = K // 2
    rng = np.random.default_rng(seed)
    s = (0.1 + 0.8 * rng.random((M + 2 * p, N + 2 * p, 3), dtype=np.float32)).astype(np.float32)
    image = _valid_conv_fft(s, psf_true)
    u0 = np.ascontiguousarray(np.pad(image, ((p, p), (p, p), (0, 0)), mode="edge"), dtype=np.float32)
    return image, u0


def default_window(M: int, N: int, p: int, mask: int = 255):
    m = min(mask, M, N)
    return (p + 1, m - p - 1, p + 1, m - p - 1)


WORKLOADS = {
    # name: (M, N, K, true-PSF builder, blind, outer iterations)
    "c1_nonblind_512_g5": (512, 512, 5, lambda: utils.gaussian_kernel(5, 1.0), False, 10),
    "c2_blind_2mp_k9": (1080, 1920, 9, lambda: utils.gaussian_kernel(9, 2.0), True, 50),
    "c3_blind_24mp_k15": (4000, 6000, 15, lambda: utils.gaussian_kernel(15, 3.0), True, 100),
    "c4_blind_61mp_k31": (6336, 9504, 31, lambda: utils.motion_kernel(31, 30.0), True, 100),
    "c5_nonblind_4k_kaiser7": (2160, 3840, 7, lambda: utils.kaiser_kernel(7, 4.0), False, 10),
}


def make_case(name: str, seed: int = 0, scale: float = 1.0, iterations: int | None = None, shape=None) -> Case:
    """Build a named workload; ``scale`` < 1 shrinks the frame (parity tests run reduced sizes), ``shape=(M, N)``
    overrides the frame size (e.g. to time one row band of a sharded frame on a single GPU)."""
    M, N, K, builder, blind, iters = WORKLOADS[name]
    M = max(int(M * scale), 3 * K + 8)
    N = max(int(N * scale), 3 * K + 8)
    if shape is not None:
        M, N = int(shape[0]), int(shape[1])
    k_true = utils.stack3(builder())
    image, u0 = make_inputs(M, N, k_true, seed)
    psf0 = utils.stack3(utils.uniform_kernel(K)) if blind else k_true.copy()
    return Case(name=name, image=image, u0=u0, psf0=psf0, psf_true=k_true,
                window=default_window(M, N, K // 2), blind=blind,
                iterations=iters if iterations is None else iterations,
                tau=0.0 if blind else 1.0)
