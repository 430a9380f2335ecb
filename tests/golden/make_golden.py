"""Generate tests/golden/*.npz by running the REFERENCE'S OWN compiled solver (oracle/_ref).

Run in the build container (where /root/reference exists and `python oracle/build_ref.py` has
produced oracle/_ref):      python tests/golden/make_golden.py [--force]

Each fixture holds the exact float32 inputs handed to ``richardson_lucy_MM``
(lib/deconvolution.pyx:341) and what the unmodified reference returned / mutated: the output view,
the full padded ``u``, the caller's ``psf`` array and the executed outer-iteration count.  The
reference has no tests or golden vectors of its own; these fixtures are the pin.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from image_cases_studies_b200 import synthetic  # noqa: E402
from image_cases_studies_b200.lib import utils  # noqa: E402
from oracle import ref_loader  # noqa: E402

sys.path.insert(0, str(ROOT / "tests"))
from helpers import smooth_case  # noqa: E402

OUT = Path(__file__).resolve().parent


def white_case(M, N, K, true_psf, blind, seed):
    kt = utils.stack3(true_psf)
    image, u0 = synthetic.make_inputs(M, N, kt, seed)
    psf0 = utils.stack3(utils.uniform_kernel(K)) if blind else kt.copy()
    return image, u0, psf0, synthetic.default_window(M, N, K // 2)


CASES = {
    # name: (builder, kwargs for the solver)
    "nonblind_64x64_k5": (lambda: white_case(64, 64, 5, utils.gaussian_kernel(5, 1.0), False, 11),
                          dict(tau=1.0, iterations=3, step_factor=1e-3, lambd=1e4, blind=False)),
    "blind_72x80_k9": (lambda: white_case(72, 80, 9, utils.gaussian_kernel(9, 2.0), True, 12),
                       dict(tau=0.0, iterations=4, step_factor=1e-3, lambd=1e4, blind=True)),
    "blind_corr_64x72_k7": (lambda: white_case(64, 72, 7, utils.kaiser_kernel(7, 4.0), True, 13),
                            dict(tau=0.0, iterations=3, step_factor=1e-3, lambd=1e4, blind=True, correlation=True)),
    "blind_48x56_k3": (lambda: white_case(48, 56, 3, utils.gaussian_kernel(3, 0.8), True, 14),
                       dict(tau=0.0, iterations=5, step_factor=5e-3, lambd=5e3, blind=True)),
    "blind_80x64_k15": (lambda: white_case(80, 64, 15, utils.gaussian_kernel(15, 3.0), True, 15),
                        dict(tau=0.0, iterations=2, step_factor=1e-3, lambd=1e4, blind=True)),
    # K >= 11: the row-FFT stencils (and, up to 17, the fused residual + PSF-gradient kernel) are the default path
    "blind_96x120_k11": (lambda: white_case(96, 120, 11, utils.gaussian_kernel(11, 2.5), True, 21),
                         dict(tau=0.0, iterations=3, step_factor=1e-3, lambd=1e4, blind=True)),
    "blind_corr_88x104_k13": (lambda: white_case(88, 104, 13, utils.kaiser_kernel(13, 5.0), True, 22),
                              dict(tau=0.0, iterations=3, step_factor=1e-3, lambd=1e4, blind=True, correlation=True)),
    "blind_100x90_k17": (lambda: white_case(100, 90, 17, utils.gaussian_kernel(17, 3.5), True, 23),
                         dict(tau=0.0, iterations=2, step_factor=1e-3, lambd=1e4, blind=True)),
    "nonblind_90x110_k15": (lambda: white_case(90, 110, 15, utils.gaussian_kernel(15, 3.0), False, 24),
                            dict(tau=1.0, iterations=3, step_factor=1e-3, lambd=1e4, blind=False)),
    "blind_96x96_k25": (lambda: white_case(96, 96, 25, utils.gaussian_kernel(25, 5.0), True, 25),
                        dict(tau=0.0, iterations=2, step_factor=1e-3, lambd=1e4, blind=True)),
    "blind_100x100_k31": (lambda: white_case(100, 100, 31, utils.gaussian_kernel(31, 6.0), True, 26),
                          dict(tau=0.0, iterations=2, step_factor=1e-3, lambd=1e4, blind=True)),
    # K > 31 (the reference's own job list goes up to 45, deconvolve.py:409): row-FFT stencils with 80 / 96-column segments
    "blind_120x110_k45": (lambda: white_case(120, 110, 45, utils.gaussian_kernel(45, 8.0), True, 27),
                          dict(tau=0.0, iterations=2, step_factor=1e-3, lambd=1e4, blind=True)),
    "nonblind_104x128_k33": (lambda: white_case(104, 128, 33, utils.gaussian_kernel(33, 6.0), False, 28),
                             dict(tau=1.0, iterations=2, step_factor=1e-3, lambd=1e4, blind=False)),
    "nonblind_stop_80x96_k7": (lambda: smooth_case(80, 96, 7, False, 0),
                               dict(tau=0.0, iterations=30, step_factor=1e-3, lambd=1e4, blind=False)),
    "nonblind_stop2_80x96_k7": (lambda: smooth_case(80, 96, 7, False, 1),
                                dict(tau=0.0, iterations=30, step_factor=1e-3, lambd=1e4, blind=False)),
}


def main():
    mod = ref_loader.load()
    if mod is None:
        sys.exit("oracle/_ref missing: run python oracle/build_ref.py first")
    force = "--force" in sys.argv
    for name, (builder, kw) in CASES.items():
        if (OUT / f"{name}.npz").exists() and not force:      # committed fixtures are only rewritten with --force
            print(f"{name}: kept")
            continue
        image, u0, psf0, window = builder()
        out, u, psf, log = ref_loader.run(image, u0, psf0, window, kw["tau"], kw["iterations"], kw["step_factor"],
                                          kw["lambd"], kw["blind"], kw.get("correlation", False))
        its = ref_loader.executed_iterations(log)
        np.savez_compressed(OUT / f"{name}.npz", image=image, u0=u0, psf0=psf0, window=np.array(window),
                            tau=kw["tau"], iterations=kw["iterations"], step_factor=kw["step_factor"],
                            lambd=kw["lambd"], blind=kw["blind"], correlation=kw.get("correlation", False),
                            ref_out=np.ascontiguousarray(out), ref_u=u, ref_psf=psf, ref_iterations=its)
        print(f"{name}: executed {its}/{kw['iterations']} outer iterations, out {out.shape}")
    if (OUT / "normalize_kernel_k7.npz").exists() and not force:
        return
    # normalize_kernel (lib/deconvolution.pyx:73-75) on a kernel with negative taps
    rng = np.random.default_rng(21)
    kern = (rng.random((7, 7, 3), dtype=np.float32) - 0.3).astype(np.float32)
    ref = kern.copy()
    mod.normalize_kernel(ref, 7)
    np.savez_compressed(OUT / "normalize_kernel_k7.npz", kern=kern, ref=ref)
    print("normalize_kernel_k7: done")


if __name__ == "__main__":
    main()
