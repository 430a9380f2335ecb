#!/usr/bin/env python
"""Small solves through the chain / row-FFT kernels for compute-sanitizer (memcheck, racecheck, synccheck).
usage: compute-sanitizer --tool memcheck python tools/sanitize_chain.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from image_cases_studies_b200 import synthetic
from image_cases_studies_b200.lib import deconvolution as dc

for name, scale, iters in (("c3_blind_24mp_k15", 0.05, 1), ("c2_blind_2mp_k9", 0.15, 1), ("c1_nonblind_512_g5", 0.3, 1)):
    c = synthetic.make_case(name, seed=3, scale=scale, iterations=iters)
    M, N = c.shape
    u, psf = c.u0.copy(), c.psf0.copy()
    out = dc.richardson_lucy_MM(c.image, u, psf, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor, c.lambd, blind=c.blind)
    print(name, M, N, c.MK, "finite:", bool(np.isfinite(out).all()), "its", dc.last_stats["iterations"], flush=True)
