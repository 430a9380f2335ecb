"""numpy statement of the collaborative L-infinity,1,1 TV (sub)gradient used by the B200 solver's ``mode="pam_ctv"``
(TEST INFRASTRUCTURE ONLY).

**PARITY UNPINNED.**  The reference snapshot contains no PAM solver and no collaborative norm: both appear in its
README only (README.md:42-44, :113-117), and the legacy ``divTV`` of ``lib/utils.py:319-351`` evaluates
to zero as written.  There is nothing to execute, so this file is a DEFINITION, not a restatement; the CUDA kernel
(``k_ctv_grad``, csrc/rltv_tvmode.cuh) is tested against it and against properties of the norm, never against reference
output.  Definition (Duran, Moeller, Sbert, Cremers, IPOL 2016, "collaborative TV", l^{inf,1,1}: l-infinity over the
colour channels, l1 over the two derivatives, l1 over the pixels):

    TV(u) = sum_pixels sum_{d in {x, y}} max_c |D_d u_c|        D_x u[y, x] = u[y, x+1] - u[y, x]  (0 in the last column)
                                                                D_y u[y, x] = u[y+1, x] - u[y, x]  (0 in the last row)
    p_{d,c} = D_d u_c / max(eps, |D_d u_c|)  if c is the (first) channel attaining the maximum, else 0   (eps = 1e-3: the
                                                                floor of lib/utils.py:335)
    T = -div p,   (div p)_c[y, x] = p_{x,c}[y, x] - p_{x,c}[y, x-1] + p_{y,c}[y, x] - p_{y,c}[y-1, x]   (p = 0 outside)

The PAM u-step of the mode is  u -= dt_c (lambda g + T),  dt_c = step max(u_c) / (max|lambda g_c + T_c| + 1e-15), followed by
the reference's DoF blend and PSF step unchanged (lib/deconvolution.pyx:499-502, :552, :555-589).
"""
from __future__ import annotations

import numpy as np


def ctv_gradient(u: np.ndarray, eps: float = 1e-3) -> np.ndarray:
    u = np.asarray(u, dtype=np.float64)
    H, W, C = u.shape
    dx = np.zeros_like(u)
    dy = np.zeros_like(u)
    dx[:, :-1] = u[:, 1:] - u[:, :-1]
    dy[:-1, :] = u[1:, :] - u[:-1, :]

    def field(d):
        a = np.abs(d)
        cstar = np.argmax(a, axis=2)                       # first channel attaining the maximum
        sel = np.zeros_like(d)
        np.put_along_axis(sel, cstar[..., None], 1.0, axis=2)
        return sel * d / np.maximum(eps, a)

    px, py = field(dx), field(dy)
    div = px.copy()
    div[:, 1:] -= px[:, :-1]
    div += py
    div[1:, :] -= py[:-1, :]
    return -div


def tv_inf11(u: np.ndarray) -> float:
    u = np.asarray(u, dtype=np.float64)
    return float(np.abs(u[:, 1:] - u[:, :-1]).max(axis=2).sum() + np.abs(u[1:, :] - u[:-1, :]).max(axis=2).sum())
