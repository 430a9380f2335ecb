"""Build a PATCHED copy of the reference solver into oracle/_ref_tv/ (TEST INFRASTRUCTURE ONLY).

Two things in ``/root/reference/lib/deconvolution.pyx`` cannot be reached from the unmodified build (oracle/_ref):

1. ``TV`` (``:137``) is ``cdef inline``: no Python entry point.  Patch: a ``cpdef`` wrapper ``tv_stencil`` is APPENDED
   to the copy; it calls the reference's own ``TV`` unchanged.  This pins the TV / divergence stencil (SURVEY.md
   section 8 row a4) by reference execution.
2. The TV regulariser is dead in the shipped arithmetic: the two ``TV(ut, ...)`` calls at ``:464-465`` are
   commented out, and as written they would store into ``TV_u_L1`` / ``TV_u_L2`` (overwritten a few lines later), so
   ``TV_ut_L1`` would stay zero even when un-commented.  Patch: the two lines are un-commented AND their output
   arrays renamed to ``TV_ut_L1`` / ``TV_ut_L2`` -- what the variable names, the test at ``:516`` and the formula at
   ``:517`` evidently intend.  The result is labelled "patched reference" everywhere; it pins the TV-alive MM mode
   (SURVEY.md 8(f3) mode (i), ``mode="mm_tv"`` of the B200 solver).  Nothing else is touched.

The patched text exists only under oracle/_ref_tv/ (git-ignored, travels to the GPU box like oracle/_ref); the
repository holds this recipe, not reference source.

Usage:  python oracle/build_ref_tv.py [--force]
"""
from __future__ import annotations

import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from oracle.build_ref import CFLAGS, GCC, LDFLAGS, REF_SRC  # noqa: E402

OUT = HERE / "_ref_tv"
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")
SO_PATH = OUT / "lib" / f"deconvolution_tv{EXT_SUFFIX}"

OLD_1 = "        #TV(ut, TV_u_L1, u_M, u_N, epsilon, 2, 1, div_ut)\n"
NEW_1 = "        TV(ut, TV_ut_L1, u_M, u_N, epsilon, 2, 1, div_ut)\n"
OLD_2 = "        #TV(ut, TV_u_L2, u_M, u_N, epsilon, 2, 2, div_ut)\n"
NEW_2 = "        TV(ut, TV_ut_L2, u_M, u_N, epsilon, 2, 2, div_ut)\n"
WRAPPER = '''

# ---- appended by oracle/build_ref_tv.py (not part of the reference): Python entry point of the reference's TV() ----
cpdef tv_stencil(np.ndarray[DTYPE_t, ndim=3] u, np.ndarray[DTYPE_t, ndim=3] out, int M, int N, float epsilon,
                 int order, int norm, np.ndarray[DTYPE_t, ndim=3] div):
    cdef float[:, :, :] u_v = u
    cdef float[:, :, :] out_v = out
    cdef float[:, :, :] div_v = div
    with nogil:
        TV(u_v, out_v, M, N, epsilon, order, norm, div_v)
'''


def available() -> bool:
    return SO_PATH.exists()


def build(force: bool = False) -> Path | None:
    if SO_PATH.exists() and not force:
        return SO_PATH
    if not REF_SRC.exists():
        return None
    import numpy
    src = REF_SRC.read_text()
    if src.count(OLD_1) != 1 or src.count(OLD_2) != 1:
        raise RuntimeError("the reference no longer contains the two commented TV(ut, ...) calls this recipe patches")
    src = src.replace(OLD_1, NEW_1).replace(OLD_2, NEW_2) + WRAPPER
    (OUT / "lib").mkdir(parents=True, exist_ok=True)
    (OUT / "build").mkdir(parents=True, exist_ok=True)
    (OUT / "lib" / "__init__.py").write_text("")
    pyx = OUT / "build" / "deconvolution_tv.pyx"
    pyx.write_text(src)
    c_file = OUT / "build" / "deconvolution_tv.c"
    subprocess.check_call([sys.executable, "-m", "cython", "-2", str(pyx), "-o", str(c_file)],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
    subprocess.check_call([GCC, *CFLAGS, *inc, str(c_file), *LDFLAGS, "-o", str(SO_PATH)])
    return SO_PATH


def load():
    """The patched module (``richardson_lucy_MM`` with the TV term alive, ``tv_stencil``), or None if not built."""
    if not SO_PATH.exists():
        return None
    import importlib.machinery
    import importlib.util
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        stubs = str(HERE / "_ref" / "stubs")
        if not (HERE / "_ref" / "stubs" / "matplotlib").exists():
            (HERE / "_ref" / "stubs" / "matplotlib").mkdir(parents=True, exist_ok=True)
            (HERE / "_ref" / "stubs" / "matplotlib" / "__init__.py").write_text("")
            (HERE / "_ref" / "stubs" / "matplotlib" / "pyplot.py").write_text("")
        if stubs not in sys.path:
            sys.path.append(stubs)
    name = "oracle_ref_tv.deconvolution_tv"
    if name in sys.modules:
        return sys.modules[name]
    loader = importlib.machinery.ExtensionFileLoader(name, str(SO_PATH))
    spec = importlib.util.spec_from_file_location(name, str(SO_PATH), loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("oracle/_ref_tv:", p if p else "reference tree absent, nothing built")
