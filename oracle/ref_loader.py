"""Loader for the compiled reference solver in oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this.  The product package never does.

``load()`` returns the reference's own ``lib.deconvolution`` extension module (built by
``oracle/build_ref.py`` from ``/root/reference/lib/deconvolution.pyx``), or ``None`` when the
``.so`` is not present.  ``run()`` calls its ``richardson_lucy_MM`` exactly as
``deconvolve.py:277-286`` / ``:304-313`` do, with the solver's prints silenced.
"""
from __future__ import annotations

import contextlib
import importlib.machinery
import importlib.util
import io
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_mod = None


def so_path() -> Path | None:
    cands = sorted((HERE / "_ref" / "lib").glob("deconvolution*.so"))
    return cands[0] if cands else None


def load():
    global _mod
    if _mod is not None:
        return _mod
    so = so_path()
    if so is None:
        return None
    try:
        import matplotlib.pyplot  # noqa: F401  (lib/deconvolution.pyx:11 imports it, never uses it)
    except Exception:
        stubs = str(HERE / "_ref" / "stubs")
        if stubs not in sys.path:
            sys.path.append(stubs)
    name = "oracle_ref.deconvolution"
    loader = importlib.machinery.ExtensionFileLoader(name, str(so))
    spec = importlib.util.spec_from_file_location(name, str(so), loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _mod = mod
    return mod


def run(image, u, psf, window, tau, iterations, step_factor, lambd, blind, correlation=False, quiet=True):
    """Run the compiled reference on COPIES of the inputs.

    Returns ``(out, u, psf, log)``: ``out`` is the returned view ``u[pad:pad+M, pad:pad+N]``
    (lib/deconvolution.pyx:675), ``u``/``psf`` the arrays the reference mutated in place, ``log``
    its stdout (the executed outer-iteration count is parsed from it by ``executed_iterations``).
    """
    mod = load()
    if mod is None:
        raise RuntimeError("oracle/_ref is not built (run python oracle/build_ref.py where /root/reference exists)")
    image = np.array(image, dtype=np.float32, order="C", copy=True)
    u = np.array(u, dtype=np.float32, order="C", copy=True)
    psf = np.array(psf, dtype=np.float32, order="C", copy=True)
    M, N = image.shape[:2]
    MK = psf.shape[0]
    top, bottom, left, right = window
    buf = io.StringIO()
    ctx = contextlib.redirect_stdout(buf) if quiet else contextlib.nullcontext()
    with ctx:
        out = mod.richardson_lucy_MM(image, u, psf, top, bottom, left, right, tau, M, N, 3, MK,
                                     iterations, step_factor, lambd, blind=int(blind),
                                     correlation=int(correlation))
    return out, u, psf, buf.getvalue()


def executed_iterations(log: str) -> int:
    """Outer iterations the reference executed, from its final print (lib/deconvolution.pyx:665-667)."""
    import re

    m = re.search(r"Convergence after (\d+) iterations", log) or re.search(r"Did not converge after (\d+) iterations", log)
    return int(m.group(1)) if m else -1
