"""numpy restatement of the reference's TV / divergence stencil (TEST INFRASTRUCTURE ONLY).

``TV(u, out, M, N, epsilon, order, norm, div)`` -- lib/deconvolution.pyx:137-239.  Interior pixels only
(``:159``, ``:239``: "borders are ignored"); ``out`` and ``div`` keep their previous content on the border ring
(zeros in the reference, which allocates them with ``np.zeros`` at ``:384-389``).

The solver calls it twice per inner step (``:495-496``) and discards the results; it is restated --
and built as a CUDA kernel, ``rltv_stage_tv`` -- because SURVEY.md section 8(a) row a4 lists it on the path and the TV-alive
mode (8(f3), ``mode="mm_tv"``) needs it.

PINNED by reference execution: ``TV`` is a ``cdef inline`` function with no Python entry point, so
``oracle/build_ref_tv.py`` appends a ``cpdef`` wrapper to a COPY of the reference source and compiles that; the
fixtures ``tests/golden/tv_o{1,2}n{1,2}.npz`` are outputs of the reference's own ``TV`` through that wrapper
(``tests/golden/make_golden_tv.py``), and ``tests/test_oracle.py`` checks this restatement against them (agreement
1e-7) and against the patched build live when it is present.
"""
from __future__ import annotations

import numpy as np


def _norm(x, y, eps, norm):
    if norm == 1:
        return np.abs(x) + np.abs(y) + eps                       # norm_L1, pyx:133-134
    return np.sqrt(x * x + y * y + eps * eps)                    # norm_L2, pyx:129-130


def tv(u: np.ndarray, epsilon: float, order: int, norm: int):
    """Returns (out, div), both shaped like ``u`` (H, W, 3) with a zero border ring."""
    u = np.asarray(u, dtype=np.float64)
    out = np.zeros_like(u)
    div = np.zeros_like(u)
    d = np.sqrt(2.0)                                             # pyx:146
    adjust = 4.0 * (1 + 1 / d) if norm == 1 else 2.0 * (1 + d)   # pyx:149-152
    c = u[1:-1, 1:-1]
    n, s = u[:-2, 1:-1], u[2:, 1:-1]                             # i-1, i+1
    w, e = u[1:-1, :-2], u[1:-1, 2:]                             # j-1, j+1
    nw, se = u[:-2, :-2], u[2:, 2:]                              # (i-1,j-1), (i+1,j+1)
    ne, sw = u[:-2, 2:], u[2:, :-2]                              # (i-1,j+1), (i+1,j-1)
    if order == 2:                                               # pyx:161-172 / :178-189
        udx = -2 * c + n + s
        udy = -2 * c + w + e
        udxdy = (-2 * c + nw + se) / d
        udydx = (-2 * c + ne + sw) / d
        div[1:-1, 1:-1] = (-udx - udy - udxdy - udydx) / adjust
        out[1:-1, 1:-1] = (_norm(udx, udy, epsilon, norm) + _norm(udxdy, udydx, epsilon, norm)) / adjust
    elif order == 1:                                             # pyx:196-213 / :219-237
        udx_b, udy_b = c - n, c - w
        udx_f, udy_f = -c + s, -c + e
        udxdy_b, udydx_b = (c - nw) / d, (c - ne) / d
        udydx_f, udxdy_f = (-c + sw) / d, (-c + se) / d
        div[1:-1, 1:-1] = (udx_b + udy_b - udx_f - udy_f + udxdy_b + udydx_b - udxdy_f - udydx_f) / adjust
        out[1:-1, 1:-1] = (_norm(udx_b, udy_b, epsilon, norm) + _norm(udx_f, udy_f, epsilon, norm)
                           + _norm(udxdy_b, udydx_b, epsilon, norm) + _norm(udxdy_f, udydx_f, epsilon, norm)) / adjust
    else:
        raise ValueError("order must be 1 or 2")
    return out, div
