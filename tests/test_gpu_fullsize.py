"""Parity at the sizes BASELINE.json names (run on the B200 box with -m gpu).

The compiled reference (oracle/_ref: the reference's own lib/deconvolution.pyx, built by oracle/build_ref.py) is run
LIVE on the box's host cores on the same seeded inputs as the CUDA path, at FULL frame size:

    C2  1080 x 1920, 9 x 9 blind     10 of its 50 outer iterations   ("run C1, C2 in full")
    C3  4000 x 6000, 15 x 15 blind    2 outer iterations  (10 inner steps; ~40 s of CPU per outer iteration)
    C4  6336 x 9504, 31 x 31 blind    1 outer iteration
    C1  512 x 512, 5 x 5 non-blind    all 10 outer iterations

Tolerances (BASELINE.json north_star): image rel-L2 <= 1e-4, PSF L1 <= 1e-4, equal executed outer-iteration counts;
the update itself, rel-L2 of (out - image), is bounded as well because in blind mode the output stays within ~3e-5 of
the input (DoF ~ 1), which would make the image tolerance vacuous on its own.
"""
import time

import numpy as np
import pytest

from helpers import TOL_IMAGE_REL_L2, TOL_PSF_L1, psf_l1, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dc():
    import __graft_entry__ as g
    g.build()
    from image_cases_studies_b200.lib import deconvolution
    yield deconvolution
    deconvolution.clear_cache()


@pytest.mark.timeout(1500)
@pytest.mark.parametrize("name,iters", [("c1_nonblind_512_g5", 10), ("c2_blind_2mp_k9", 10), ("c3_blind_24mp_k15", 2),
                                        ("c4_blind_61mp_k31", 1)])
def test_full_frame_against_live_reference(dc, name, iters):
    from oracle import ref_loader
    if ref_loader.so_path() is None:
        pytest.skip("oracle/_ref did not travel")
    from image_cases_studies_b200 import synthetic
    c = synthetic.make_case(name, seed=21, scale=1.0, iterations=iters)
    M, N = c.shape
    assert (M, N) == synthetic.WORKLOADS[name][:2]                       # no scale factor below 1
    t0 = time.time()
    out_r, u_r, psf_r, log = ref_loader.run(c.image, c.u0, c.psf0, c.window, c.tau, c.iterations, c.step_factor,
                                            c.lambd, c.blind)
    t_ref = time.time() - t0
    u, psf = c.u0.copy(), c.psf0.copy()
    t0 = time.time()
    out = dc.richardson_lucy_MM(c.image, u, psf, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor, c.lambd,
                                blind=c.blind)
    t_gpu = time.time() - t0
    st = dict(dc.last_stats)
    dc.clear_cache()
    r_img, r_u = rel_l2(out, out_r), rel_l2(u, u_r)
    upd_ref = out_r.astype(np.float64) - c.image
    upd = out.astype(np.float64) - c.image
    r_upd = float(np.linalg.norm(upd - upd_ref) / max(np.linalg.norm(upd_ref), 1e-30))
    l1 = psf_l1(psf, psf_r)
    print(f"\n{name} {M}x{N} K={c.MK}: its {st['iterations']} (reference {ref_loader.executed_iterations(log)}) "
          f"image rel-L2 {r_img:.2e}  u rel-L2 {r_u:.2e}  update rel-L2 {r_upd:.2e}  psf L1 {l1:.2e}  "
          f"reference {t_ref:.1f} s, drop-in call {t_gpu:.2f} s")
    assert st["iterations"] == ref_loader.executed_iterations(log)
    assert r_img <= TOL_IMAGE_REL_L2 and r_u <= TOL_IMAGE_REL_L2
    assert l1 <= TOL_PSF_L1
    assert r_upd <= 2e-2
    assert np.linalg.norm(upd_ref) > 0


def _valid_corr_fft(u, e):
    """float64: convolve(rot180(u), e, "valid") (lib/deconvolution.pyx:567-571) for u (Hu, Wu), e (M, N) -> (K, K), by a
    circular FFT of size (Hu, Wu): the valid indices [M-1, Hu-1] x [N-1, Wu-1] never wrap."""
    Hu, Wu = u.shape
    M, N = e.shape
    fa = np.fft.rfft2(u[::-1, ::-1].astype(np.float64), (Hu, Wu))
    fb = np.fft.rfft2(e.astype(np.float64), (Hu, Wu))
    full = np.fft.irfft2(fa * fb, (Hu, Wu))
    return full[M - 1:, N - 1:]


@pytest.mark.timeout(900)
def test_psf_gradient_kernel_at_24mp():
    """k_gradk_fft<15, FUSED> (residual + PSF gradient in one kernel) on the whole 4000 x 6000 frame against a float64
    FFT correlation, with a solver-like residual (small, zero-mean).  Every per-CTA fp32 frequency-domain accumulator
    sums ~4 400 row products here (about 100 in the small-frame tests), so this is the size that bounds its error."""
    from image_cases_studies_b200 import synthetic
    from image_cases_studies_b200.solver import Solver
    from oracle import rl_mm_oracle as orc
    c = synthetic.make_case("c3_blind_24mp_k15", seed=5, scale=1.0, iterations=1)
    M, N = c.shape
    K = c.MK
    s = Solver(M, N, K)
    s.upload(c.image, c.u0, c.psf0)
    gk = s.stage_gradk()                      # computes e2 = conv(u, psf) - image itself (pyx:557-565), then pyx:567-571
    err = s.debug_residual()
    s.close()
    worst_e, worst_g = 0.0, 0.0
    for ch in range(3):
        e_ref = orc.conv2(c.u0[..., ch], c.psf0[..., ch], "valid") - c.image[..., ch]
        worst_e = max(worst_e, rel_l2(err[..., ch], e_ref))
        gk_ref = _valid_corr_fft(c.u0[..., ch], e_ref)
        worst_g = max(worst_g, rel_l2(gk[..., ch], gk_ref))
    print(f"\n24 MP K=15: residual rel-L2 {worst_e:.2e}, PSF gradient rel-L2 {worst_g:.2e}")
    assert worst_e < 1e-4
    # dtpsf * |gk| <= step/K * max(psf): a relative error r of gk moves the PSF by at most r * 1e-3/15 * 225 taps in L1,
    # i.e. r = 1e-3 costs 1.5e-5 of the 1e-4 PSF budget per step.  Measured: see the printed value.
    assert worst_g < 1e-3


def test_one_shot_c_entry_point():
    """rltv_richardson_lucy_mm (include/rltv_b200.h): create + upload + solve + download + destroy in one C call, the
    entry point INTEGRATION.md advertises for a cgo/ctypes binding; must agree with the persistent-context path."""
    import ctypes as C
    import __graft_entry__ as g
    g.build()
    from image_cases_studies_b200 import _native as nat, synthetic
    from image_cases_studies_b200.lib import deconvolution as dc
    from image_cases_studies_b200.solver import Solver
    c = synthetic.make_case("c3_blind_24mp_k15", seed=13, scale=0.05, iterations=3)
    M, N = c.shape
    u1, p1 = c.u0.copy(), c.psf0.copy()
    dc.richardson_lucy_MM(c.image, u1, p1, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor, c.lambd, blind=True)
    u2, p2, pr = c.u0.copy(), c.psf0.copy(), np.empty_like(c.psf0)
    params = Solver.make_params(c.window, c.tau, c.iterations, c.step_factor, c.lambd, True)
    st = nat.Stats()
    nat.check(nat.lib.rltv_richardson_lucy_mm(nat.ptr(c.image), c.image.strides[0], nat.ptr(u2), u2.strides[0], nat.ptr(p2),
                                              M, N, c.MK, C.byref(params), C.byref(st), nat.ptr(pr), 0))
    assert st.as_dict()["iterations"] == dc.last_stats["iterations"] == 3
    assert np.array_equal(u1, u2) and np.array_equal(p1, p2) and np.array_equal(p2, pr)
