"""B200-native Richardson-Lucy / MM deconvolution: drop-in for the solver hot path of
aurelienpierre/Image-Cases-Studies (``lib/deconvolution.pyx``).

    from image_cases_studies_b200.lib import deconvolution as dc    # was: from lib import deconvolution as dc
    out = dc.richardson_lucy_MM(image, u, psf, top, bottom, left, right, tau, M, N, C, MK,
                                iterations, step_factor, lambd, blind=True)

Everything numerical runs in hand-written sm_100a CUDA kernels behind the C-ABI declared in
``include/rltv_b200.h`` (``csrc/``); there is no CPU fallback.
"""
__version__ = "0.1.0"
