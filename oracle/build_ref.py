"""Build the reference's own CPU solver into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference's hot path is one Cython file, ``/root/reference/lib/deconvolution.pyx``.
This recipe translates it with the local ``cython`` and compiles the generated C with
``/usr/bin/gcc`` -- straight from where the source lies; nothing from the reference is
copied into the repository.  All outputs go to ``oracle/_ref/`` (git-ignored, NOT
gpurun-ignored, so the built ``.so`` travels to the GPU box, where ``/root/reference``
does not exist).

Flags follow the reference's ``setup.py:27-28`` (``-O3 -fopenmp -finline-functions
-ffast-math -msse4``) with one deliberate change: ``-march=x86-64-v3`` instead of
``-march=native``, because the ``.so`` is built in this container and executed on a
different host (the GPU box).  x86-64-v3 = AVX2+FMA, present on every B200 host CPU.

The module imports ``matplotlib.pyplot`` (``lib/deconvolution.pyx:11``) without using
it; an empty stub package is written to ``oracle/_ref/stubs`` so the import succeeds.

Usage:  python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/lib/deconvolution.pyx")
OUT = HERE / "_ref"
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")
SO_PATH = OUT / "lib" / f"deconvolution{EXT_SUFFIX}"

CFLAGS = ["-O3", "-fopenmp", "-march=x86-64-v3", "-finline-functions", "-ffast-math", "-msse4",
          "-fPIC", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
LDFLAGS = ["-shared", "-fopenmp"]
# /opt/gcc/bin/gcc (first on PATH in some shells) has no libgomp.spec; use the distro gcc.
GCC = "/usr/bin/gcc"


def available() -> bool:
    return SO_PATH.exists()


def build(force: bool = False) -> Path | None:
    """Returns the path of the built module, or None when the reference tree is absent."""
    if SO_PATH.exists() and not force:
        return SO_PATH
    if not REF_SRC.exists():
        return None
    import numpy

    (OUT / "lib").mkdir(parents=True, exist_ok=True)
    (OUT / "build").mkdir(parents=True, exist_ok=True)
    stubs = OUT / "stubs" / "matplotlib"
    stubs.mkdir(parents=True, exist_ok=True)
    (stubs / "__init__.py").write_text("")
    (stubs / "pyplot.py").write_text("")
    (OUT / "lib" / "__init__.py").write_text("")

    c_file = OUT / "build" / "deconvolution.c"
    # language_level=2 is what Cython 0.28 (the reference's era) defaulted to; the only
    # int/int divisions are C divisions under cdivision=True either way.
    subprocess.check_call([sys.executable, "-m", "cython", "-2", str(REF_SRC), "-o", str(c_file)],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
    subprocess.check_call([GCC, *CFLAGS, *inc, str(c_file), *LDFLAGS, "-o", str(SO_PATH)])
    return SO_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("oracle/_ref:", p if p else "reference tree absent, nothing built")
