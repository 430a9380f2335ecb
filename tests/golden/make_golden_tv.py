"""Generate tests/golden/tv_*.npz and tvmm_*.npz by running the PATCHED reference (oracle/_ref_tv, built by
oracle/build_ref_tv.py from /root/reference/lib/deconvolution.pyx: a cpdef wrapper around the reference's own TV(), and
the two TV(ut, ...) calls of :464-465 alive).  Run in the build container:  python tests/golden/make_golden_tv.py [--force]

tv_o{order}n{norm}.npz   TV(u, out, M, N, epsilon, order, norm, div) of lib/deconvolution.pyx:137-239, all four pairs
tvmm_*.npz               richardson_lucy_MM of the patched reference ("TV alive"): inputs, outputs, the denoised image
"""
from __future__ import annotations

import contextlib
import io
import re
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from image_cases_studies_b200 import synthetic  # noqa: E402
from image_cases_studies_b200.lib import utils  # noqa: E402
from oracle import build_ref_tv  # noqa: E402

OUT = Path(__file__).resolve().parent

TVMM = {
    # name: (M, N, K, blind, outer iterations, seed)
    "tvmm_nonblind_64x72_k5": (64, 72, 5, False, 3, 31),
    "tvmm_blind_72x80_k9": (72, 80, 9, True, 3, 32),
    "tvmm_blind_90x110_k15": (90, 110, 15, True, 2, 33),
    "tvmm_nonblind_100x96_k13": (100, 96, 13, False, 2, 34),
}


def main():
    mod = build_ref_tv.load()
    if mod is None:
        sys.exit("oracle/_ref_tv missing: run python oracle/build_ref_tv.py first")
    force = "--force" in sys.argv
    rng = np.random.default_rng(41)
    u = rng.random((37, 53, 3), dtype=np.float32)
    for order in (1, 2):
        for norm in (1, 2):
            f = OUT / f"tv_o{order}n{norm}.npz"
            if f.exists() and not force:
                continue
            eps = 1e-2 if norm == 1 else 1e-6
            out, div = np.zeros_like(u), np.zeros_like(u)
            mod.tv_stencil(u, out, u.shape[0], u.shape[1], eps, order, norm, div)
            np.savez_compressed(f, u=u, epsilon=eps, order=order, norm=norm, ref_out=out, ref_div=div)
            print(f.name, "done")
    for name, (M, N, K, blind, its, seed) in TVMM.items():
        f = OUT / f"{name}.npz"
        if f.exists() and not force:
            continue
        kt = utils.stack3(utils.gaussian_kernel(K, K / 4))
        image, u0 = synthetic.make_inputs(M, N, kt, seed=seed)
        psf0 = utils.stack3(utils.uniform_kernel(K)) if blind else kt.copy()
        window = synthetic.default_window(M, N, K // 2)
        tau = 0.0 if blind else 1.0
        img, uu, psf = image.copy(), u0.copy(), psf0.copy()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            out = mod.richardson_lucy_MM(img, uu, psf, *window, tau, M, N, 3, K, its, 1e-3, 1e4, blind=int(blind), correlation=0)
        m = re.search(r"(?:Convergence after|Did not converge after) (\d+) iterations", buf.getvalue())
        np.savez_compressed(f, image=image, u0=u0, psf0=psf0, window=np.array(window), tau=tau, iterations=its, step_factor=1e-3,
                            lambd=1e4, blind=blind, correlation=False, ref_out=np.ascontiguousarray(out), ref_u=uu, ref_psf=psf,
                            ref_image=img, ref_iterations=int(m.group(1)))
        print(f.name, "executed", m.group(1), "outer iterations")


if __name__ == "__main__":
    main()
