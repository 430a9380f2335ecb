"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI / the
drop-in module, against (a) the committed golden vectors of the compiled reference, (b) the numpy oracle on
seeded inputs, (c) oracle/_ref run live when it travelled, (d) size-independent properties at full size.

Tolerances (BASELINE.json north_star): image rel-L2 <= 1e-4, PSF L1 <= 1e-4, equal outer-iteration counts.
"""
import numpy as np
import pytest

from helpers import TOL_IMAGE_REL_L2, TOL_PSF_L1, golden_names, load_golden, psf_l1, rel_l2  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dc():
    import __graft_entry__ as g
    g.build()
    from image_cases_studies_b200.lib import deconvolution
    return deconvolution


def _run_dc(dc, g):
    M, N = g["image"].shape[:2]
    MK = g["psf0"].shape[0]
    u = g["u0"].copy()
    psf = g["psf0"].copy()
    out = dc.richardson_lucy_MM(g["image"].copy(), u, psf, *g["window"], g["tau"], M, N, 3, MK, g["iterations"],
                                g["step_factor"], g["lambd"], blind=g["blind"], correlation=g["correlation"])
    return out, u, psf, dict(dc.last_stats)


@pytest.mark.parametrize("name", golden_names())
def test_golden_vectors(dc, name):
    g = load_golden(name)
    out, u, psf, st = _run_dc(dc, g)
    assert np.shares_memory(out, u)                      # the result is a view of the caller's u (pyx:675)
    assert st["iterations"] == g["ref_iterations"]
    assert rel_l2(out, g["ref_out"]) <= TOL_IMAGE_REL_L2
    assert rel_l2(u, g["ref_u"]) <= TOL_IMAGE_REL_L2
    assert psf_l1(psf, g["ref_psf"]) <= TOL_PSF_L1
    # in blind mode the output moves away from the input by only ~3e-5 (DoF ~ 1), so also check the update itself
    upd_ref = g["ref_out"].astype(np.float64) - g["image"]
    upd = out.astype(np.float64) - g["image"]
    assert np.linalg.norm(upd - upd_ref) <= 0.02 * np.linalg.norm(upd_ref) + 1e-6 * np.linalg.norm(g["image"])


@pytest.mark.parametrize("K", [3, 5, 7, 9, 11, 13, 15, 17, 19, 25, 31])
def test_stages_against_oracle(K):
    """Forward residual, adjoint and PSF gradient, each kernel alone against the float64 definition."""
    from image_cases_studies_b200.solver import Solver
    from oracle import rl_mm_oracle as orc
    rng = np.random.default_rng(K)
    M, N = 70 + 3 * K, 150 + K            # not multiples of the tile sizes
    u = rng.random((M + K - 1, N + K - 1, 3), dtype=np.float32)
    image = rng.random((M, N, 3), dtype=np.float32)
    psf = rng.random((K, K, 3), dtype=np.float32)
    psf /= psf.sum(axis=(0, 1), keepdims=True)
    s = Solver(M, N, K)
    s.upload(image, u, psf)
    err = s.stage_residual()
    g = s.stage_adjoint()
    gk = s.stage_gradk()
    s.close()
    for c in range(3):
        e_ref = orc.conv2(u[..., c], psf[..., c], "valid") - image[..., c]
        assert rel_l2(err[..., c], e_ref) < 2e-6
        g_ref = orc.conv2(e_ref, orc.rot180(psf[..., c]), "full")
        assert rel_l2(g[..., c], g_ref) < 3e-6
        gk_ref = orc.conv2(orc.rot180(u[..., c]), e_ref, "valid")
        # all-positive random data = a huge DC bin: the fp32 frequency-domain accumulation is good to ~1e-5 here
        # (a real residual is small and zero-mean; the solver-level PSF tolerance is checked separately)
        assert rel_l2(gk[..., c], gk_ref) < 5e-5


@pytest.mark.parametrize("K", [3, 9, 15, 31, 45])
@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (17, 5), (47, 113), (48, 112), (49, 111), (97, 225)])
def test_stages_tiny_and_tile_edge_shapes(K, shape):
    """Frames smaller than one tile, exactly one tile, one pixel past a tile (direct tiles: 16 x 128, row-FFT tiles:
    48 x 112 / 64 x 96): every kernel alone against the float64 definition, K on both sides of the direct/FFT switch."""
    from image_cases_studies_b200.solver import Solver
    from oracle import rl_mm_oracle as orc
    M, N = shape
    rng = np.random.default_rng(1000 * K + 10 * M + N)
    u = rng.random((M + K - 1, N + K - 1, 3), dtype=np.float32)
    image = rng.random((M, N, 3), dtype=np.float32)
    psf = rng.random((K, K, 3), dtype=np.float32)
    psf /= psf.sum(axis=(0, 1), keepdims=True)
    s = Solver(M, N, K)
    s.upload(image, u, psf)
    err = s.stage_residual()
    g = s.stage_adjoint()
    gk = s.stage_gradk()
    s.close()
    assert err.shape == (M, N, 3) and g.shape == u.shape and gk.shape == (K, K, 3)
    for c in range(3):
        e_ref = orc.conv2(u[..., c], psf[..., c], "valid") - image[..., c]
        assert rel_l2(err[..., c], e_ref) < 5e-6
        g_ref = orc.conv2(e_ref, orc.rot180(psf[..., c]), "full")
        assert rel_l2(g[..., c], g_ref) < 5e-6
        gk_ref = orc.conv2(orc.rot180(u[..., c]), e_ref, "valid")
        assert rel_l2(gk[..., c], gk_ref) < 5e-5


@pytest.mark.parametrize("win", [(4, 60, 4, 60), (3, 120, 10, 97), (0, 33, 5, 200)])
def test_whiteness_statistic(win):
    from image_cases_studies_b200.solver import Solver
    from oracle import rl_mm_oracle as orc
    rng = np.random.default_rng(5)
    M, N, K = 140, 230, 5
    u = rng.random((M + K - 1, N + K - 1, 3), dtype=np.float32)
    image = rng.random((M, N, 3), dtype=np.float32)
    psf = np.full((K, K, 3), 1.0 / (K * K), np.float32)
    s = Solver(M, N, K)
    s.upload(image, u, psf)
    err = s.stage_residual()
    m_gpu = s.stage_whiteness(win)
    s.close()
    top, bottom, left, right = win
    m_ref = orc.whiteness(err[top:bottom, left:right].astype(np.float64), orc.whiteness_weights(bottom - top, right - left))
    assert abs(m_gpu - m_ref) <= 2e-6 * abs(m_ref)


@pytest.mark.parametrize("name,scale,iters", [("c1_nonblind_512_g5", 1.0, 10), ("c2_blind_2mp_k9", 0.25, 6),
                                              ("c3_blind_24mp_k15", 0.06, 3), ("c5_nonblind_4k_kaiser7", 0.1, 3)])
def test_workloads_against_oracle(dc, name, scale, iters):
    from image_cases_studies_b200 import synthetic
    from oracle import rl_mm_oracle as orc
    c = synthetic.make_case(name, seed=7, scale=scale, iterations=iters)
    M, N = c.shape
    g = dict(image=c.image, u0=c.u0, psf0=c.psf0, window=c.window, tau=c.tau, iterations=c.iterations,
             step_factor=c.step_factor, lambd=c.lambd, blind=c.blind, correlation=False)
    out, u, psf, st = _run_dc(dc, g)
    ref = orc.richardson_lucy_MM(c.image, c.u0, c.psf0, *c.window, c.tau, M, N, 3, c.MK, c.iterations,
                                 c.step_factor, c.lambd, blind=c.blind)
    assert st["iterations"] == ref.iterations
    assert rel_l2(out, ref.out) <= TOL_IMAGE_REL_L2
    assert psf_l1(psf, ref.psf) <= TOL_PSF_L1
    assert np.allclose(st["M_r_history"], ref.M_r, rtol=2e-5)
    for a, b in zip(st["dt"], ref.dt[-1]):
        assert abs(a - b) <= 1e-4 * abs(b)


def test_live_reference_when_present(dc):
    from oracle import ref_loader
    if ref_loader.so_path() is None:
        pytest.skip("oracle/_ref did not travel")
    from image_cases_studies_b200 import synthetic
    for name, scale, iters in (("c2_blind_2mp_k9", 0.2, 5), ("c1_nonblind_512_g5", 0.75, 5)):
        c = synthetic.make_case(name, seed=9, scale=scale, iterations=iters)
        M, N = c.shape
        out_r, u_r, psf_r, log = ref_loader.run(c.image, c.u0, c.psf0, c.window, c.tau, c.iterations, c.step_factor,
                                                c.lambd, c.blind)
        g = dict(image=c.image, u0=c.u0, psf0=c.psf0, window=c.window, tau=c.tau, iterations=c.iterations,
                 step_factor=c.step_factor, lambd=c.lambd, blind=c.blind, correlation=False)
        out, u, psf, st = _run_dc(dc, g)
        assert st["iterations"] == ref_loader.executed_iterations(log)
        assert rel_l2(out, out_r) <= TOL_IMAGE_REL_L2 and psf_l1(psf, psf_r) <= TOL_PSF_L1


def test_strided_views_and_inplace_contract(dc):
    """The reference's driver passes row/column-sliced views of larger frames (deconvolve.py:277-286)."""
    from image_cases_studies_b200 import synthetic
    c = synthetic.make_case("c2_blind_2mp_k9", seed=4, scale=0.08, iterations=2)
    M, N = c.shape
    p = c.MK // 2
    big_img = np.zeros((M + 10, N + 14, 3), np.float32)
    big_u = np.zeros((M + 2 * p + 6, N + 2 * p + 8, 3), np.float32)
    img_v = big_img[3:3 + M, 5:5 + N]
    u_v = big_u[2:2 + M + 2 * p, 4:4 + N + 2 * p]
    img_v[...] = c.image
    u_v[...] = c.u0
    psf_a, psf_b = c.psf0.copy(), c.psf0.copy()
    out_v = dc.richardson_lucy_MM(img_v, u_v, psf_a, *c.window, c.tau, M, N, 3, c.MK, 2, c.step_factor, c.lambd, blind=True)
    u_c = c.u0.copy()
    out_c = dc.richardson_lucy_MM(c.image, u_c, psf_b, *c.window, c.tau, M, N, 3, c.MK, 2, c.step_factor, c.lambd, blind=True)
    assert np.shares_memory(out_v, big_u)
    assert np.array_equal(out_v, out_c) and np.array_equal(psf_a, psf_b)      # bit-deterministic
    assert np.array_equal(big_u[2:2 + M + 2 * p, 4:4 + N + 2 * p], u_c)
    assert not big_u[:2].any() and not big_u[:, :4].any()                       # nothing outside the view was touched


def test_normalize_kernel_golden(dc):
    from helpers import GOLDEN
    z = np.load(GOLDEN / "normalize_kernel_k7.npz")
    k = z["kern"].copy()
    assert dc.normalize_kernel(k, 7) is None
    assert np.abs(k - z["ref"]).max() < 1e-6


def test_nonblind_leaves_psf_untouched_and_zero_iterations(dc):
    from image_cases_studies_b200 import synthetic
    c = synthetic.make_case("c1_nonblind_512_g5", seed=2, scale=0.2, iterations=0)
    M, N = c.shape
    u, psf = c.u0.copy(), c.psf0.copy()
    out = dc.richardson_lucy_MM(c.image, u, psf, *c.window, c.tau, M, N, 3, c.MK, 0, c.step_factor, c.lambd, blind=False)
    assert np.array_equal(u, c.u0) and np.array_equal(psf, c.psf0) and dc.last_stats["iterations"] == 0
    assert out.shape == (M, N, 3)


def test_full_size_properties(dc):
    """At BASELINE's 24 MP / 15x15 size the oracle is too slow; check size-independent properties instead:
    PSF stays on the simplex, u ring/interior finite, determinism, and adjointness <Ku, e> == <u, K^T e>."""
    from image_cases_studies_b200.solver import Solver
    rng = np.random.default_rng(0)
    M, N, K = 4000, 6000, 15
    u = (0.1 + 0.8 * rng.random((M + K - 1, N + K - 1, 3), dtype=np.float32))
    image = (0.1 + 0.8 * rng.random((M, N, 3), dtype=np.float32))
    from image_cases_studies_b200.lib import utils
    psf = utils.stack3(utils.gaussian_kernel(K, 3.0))
    s = Solver(M, N, K)
    s.upload(image, u, psf)
    err = s.stage_residual()                       # err = K u - image
    g = s.stage_adjoint()                          # g = K^T err
    lhs = float(np.sum((err.astype(np.float64) + image) * err))        # <K u, err>
    rhs = float(np.sum(u.astype(np.float64) * g))                      # <u, K^T err>
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)
    params = Solver.make_params((8, 247, 8, 247), 0.0, 1, 1e-3, 1e4, True)
    st = s.solve(params)
    u1, p1 = np.empty_like(u), np.empty_like(psf)
    s.download(u1, p1)
    assert st["iterations"] == 1 and np.isfinite(u1).all()
    assert np.allclose(p1.sum(axis=(0, 1)), 1.0, atol=1e-5) and p1.min() >= 0
    s.upload(image, u, psf)
    s.solve(params)
    u2, p2 = np.empty_like(u), np.empty_like(psf)
    s.download(u2, p2)
    s.close()
    assert np.array_equal(u1, u2) and np.array_equal(p1, p2)


@pytest.mark.parametrize("order,norm,eps", [(2, 1, 1e-2), (2, 2, 1e-2), (1, 1, 1e-6), (1, 2, 1e-6)])
def test_tv_stencil_against_oracle(order, norm, eps):
    """TV(u, out, ..., div) of lib/deconvolution.pyx:137-239 (the solver calls it with order 2, norm 1 and 2)."""
    from image_cases_studies_b200.solver import Solver
    from oracle import tv_oracle
    rng = np.random.default_rng(order * 10 + norm)
    M, N, K = 97, 203, 5
    u = rng.random((M + K - 1, N + K - 1, 3), dtype=np.float32)
    s = Solver(M, N, K)
    s.upload(np.zeros((M, N, 3), np.float32), u, np.full((K, K, 3), 1 / 25, np.float32))
    out, div = s.stage_tv(order, norm, eps)
    s.close()
    out_ref, div_ref = tv_oracle.tv(u, eps, order, norm)
    assert rel_l2(out, out_ref) < 2e-6 and rel_l2(div, div_ref) < 2e-6
    assert not out[0].any() and not out[:, 0].any() and not div[-1].any() and not div[:, -1].any()


def test_pyramid_driver_on_gpu_matches_oracle_driven_flow(tmp_path):
    """deblur_module (deconvolve.py:65-368) end to end: the same driver run with the B200 solver and with the oracle."""
    from image_cases_studies_b200 import deconvolve as drv
    from image_cases_studies_b200.lib import utils
    from oracle import rl_mm_oracle as orc
    from test_driver import OracleSolver
    rng = np.random.default_rng(8)
    sharp = rng.random((90, 110, 3)) * 200 + 20
    k = utils.gaussian_kernel(5, 1.2)
    blurred = np.stack([orc.conv2(np.pad(sharp[..., c], 2, mode="edge"), k, "valid") for c in range(3)], axis=2)
    kw = dict(mask=[45, 55], mask_size=41, iterations=3, display=False, save=False)
    out_gpu = drv.deblur_module(blurred, "g", str(tmp_path), 5, **kw)
    psf_gpu = drv.deblur_module.last_psf.copy()
    out_ref = drv.deblur_module(blurred, "r", str(tmp_path), 5, solver=OracleSolver, **kw)
    psf_ref = drv.deblur_module.last_psf.copy()
    assert rel_l2(out_gpu, out_ref) <= TOL_IMAGE_REL_L2
    assert psf_l1(psf_gpu, psf_ref) <= TOL_PSF_L1


@pytest.mark.parametrize("inverse", [0, 1])
def test_fft128_engine(inverse):
    """csrc/rltv_fft.cuh (16 x 8 register passes over shared memory) against numpy's FFT."""
    import __graft_entry__ as g
    g.build()
    from image_cases_studies_b200 import _native as nat
    rng = np.random.default_rng(3)
    nrows = 37
    x = (rng.standard_normal((nrows, 128)) + 1j * rng.standard_normal((nrows, 128))).astype(np.complex64)
    out = np.empty_like(x)
    nat.check(nat.lib.rltv_debug_fft128(nat.ptr(x.view(np.float32)), nat.ptr(out.view(np.float32)), nrows, inverse, 0))
    ref = np.fft.ifft(x.astype(np.complex128), axis=1) * 128 if inverse else np.fft.fft(x.astype(np.complex128), axis=1)
    assert np.abs(out - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("K", [9, 11, 13, 15, 17])
def test_fused_residual_in_psf_gradient_kernel(K, monkeypatch):
    """k_gradk_fft<K, FUSED>: the PSF-gradient kernel that computes the residual of pyx:557-565 itself (no forward-blur
    launch before it), and the residual it leaves behind for the whiteness statistic, against (a) the two-kernel
    sequence it replaces (RLTV_FUSE=0: same arithmetic, so nearly bit-equal) and (b) the float64 definition."""
    from image_cases_studies_b200.solver import Solver
    from oracle import rl_mm_oracle as orc
    monkeypatch.setenv("RLTV_CONV", "fft")
    rng = np.random.default_rng(100 + K)
    M, N = 171 + K, 233 + 2 * K           # ragged against the 80 x 112 tiles
    u = (0.1 + 0.8 * rng.random((M + K - 1, N + K - 1, 3))).astype(np.float32)
    psf = rng.random((K, K, 3), dtype=np.float32)
    psf /= psf.sum(axis=(0, 1), keepdims=True)
    # a blurred scene plus a small perturbation: the residual is small and zero-mean like in the solver
    image = np.stack([orc.conv2(u[..., c], psf[..., c], "valid") for c in range(3)], axis=2)
    image = (image * (1.0 + 0.05 * rng.standard_normal(image.shape))).astype(np.float32)
    res = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("RLTV_FUSE", fuse)
        s = Solver(M, N, K)
        s.upload(image, u, psf)
        if fuse == "0":
            s.stage_residual()
        gk = s.stage_gradk()
        res[fuse] = (gk, s.debug_residual())
        s.close()
    assert rel_l2(res["1"][1], res["0"][1]) < 1e-5, "residual left by the fused kernel vs k_conv_fft"
    assert rel_l2(res["1"][0], res["0"][0]) < 1e-4, "PSF gradient, fused vs two kernels"
    for c in range(3):
        e_ref = orc.conv2(u[..., c], psf[..., c], "valid") - image[..., c]
        gk_ref = orc.conv2(orc.rot180(u[..., c]), e_ref, "valid")
        for fuse in ("1", "0"):
            assert rel_l2(res[fuse][1][..., c], e_ref) < 1e-4, f"residual vs float64, RLTV_FUSE={fuse}"
            # zero-mean residual against a DC-heavy estimate: the fp32 row spectra bound this at a few 1e-4
            assert rel_l2(res[fuse][0][..., c], gk_ref) < 1e-3, f"PSF gradient vs float64, RLTV_FUSE={fuse}"


@pytest.mark.parametrize("K", [9, 11, 13, 15, 17, 19, 25, 31, 33, 37, 41, 45, 47])
def test_row_fft_stencils_against_oracle(K, monkeypatch):
    """k_conv_fft (row-FFT hybrid forward blur / adjoint, csrc/rltv_stencil_fft.cuh) against the float64 definition."""
    from image_cases_studies_b200.solver import Solver
    from oracle import rl_mm_oracle as orc
    monkeypatch.setenv("RLTV_CONV", "fft")
    rng = np.random.default_rng(K)
    M, N = 150 + 3 * K, 260 + K           # several 64 x 112 tiles, ragged edges
    u = rng.random((M + K - 1, N + K - 1, 3), dtype=np.float32)
    image = rng.random((M, N, 3), dtype=np.float32)
    psf = rng.random((K, K, 3), dtype=np.float32)
    psf /= psf.sum(axis=(0, 1), keepdims=True)
    s = Solver(M, N, K)
    s.upload(image, u, psf)
    err = s.stage_residual()
    g = s.stage_adjoint()
    gk = s.stage_gradk()
    s.close()
    for c in range(3):
        e_ref = orc.conv2(u[..., c], psf[..., c], "valid") - image[..., c]
        assert rel_l2(err[..., c], e_ref) < 5e-6
        g_ref = orc.conv2(e_ref, orc.rot180(psf[..., c]), "full")
        assert rel_l2(g[..., c], g_ref) < 5e-6
        gk_ref = orc.conv2(orc.rot180(u[..., c]), e_ref, "valid")
        # all-positive random data = a huge DC bin: the fp32 frequency-domain accumulation of the row-FFT path
        # (default for 11 <= K <= 17) is good to ~2e-5 here; a real residual is small and zero-mean
        assert rel_l2(gk[..., c], gk_ref) < 5e-5


def test_row_fft_solver_against_oracle(dc, monkeypatch):
    from image_cases_studies_b200 import synthetic
    from oracle import rl_mm_oracle as orc
    monkeypatch.setenv("RLTV_CONV", "fft")
    dc.clear_cache()
    try:
        c = synthetic.make_case("c3_blind_24mp_k15", seed=7, scale=0.06, iterations=3)
        M, N = c.shape
        g = dict(image=c.image, u0=c.u0, psf0=c.psf0, window=c.window, tau=c.tau, iterations=c.iterations,
                 step_factor=c.step_factor, lambd=c.lambd, blind=c.blind, correlation=False)
        out, u, psf, st = _run_dc(dc, g)
        ref = orc.richardson_lucy_MM(c.image, c.u0, c.psf0, *c.window, c.tau, M, N, 3, c.MK, c.iterations,
                                     c.step_factor, c.lambd, blind=c.blind)
        assert st["iterations"] == ref.iterations
        assert rel_l2(out, ref.out) <= TOL_IMAGE_REL_L2
        assert psf_l1(psf, ref.psf) <= TOL_PSF_L1
        assert np.allclose(st["M_r_history"], ref.M_r, rtol=2e-5)
    finally:
        dc.clear_cache()


@pytest.mark.parametrize("K", [5, 7, 9, 11, 13, 15, 17])
@pytest.mark.parametrize("shape", [(5, 9), (61, 97), (100, 100), (203, 251), (331, 420)])
def test_chain_kernel_against_oracle(K, shape, monkeypatch):
    """k_chain_fft (csrc/rltv_chain_fft.cuh): forward blur, residual and adjoint in one pass with the residual kept in
    the row-frequency domain, against the float64 definition  g = conv(conv(u, psf, 'valid') - image, rot180 psf, 'full')
    (pyx:477-491) and its own step statistics (pyx:524 with lambda = 1, ut = u).  Shapes: smaller than one segment,
    around one segment (V = 128 - 2(K-1) columns), several segments and several 16-row steps per piece."""
    from image_cases_studies_b200.solver import Solver
    from oracle import rl_mm_oracle as orc
    monkeypatch.setenv("RLTV_CHAIN_MINPIX", "0")      # K <= 9 take the chain kernel on megapixel frames only by default
    M, N = shape
    rng = np.random.default_rng(7000 + 100 * K + M)
    u = (0.1 + 0.8 * rng.random((M + K - 1, N + K - 1, 3))).astype(np.float32)
    psf = rng.random((K, K, 3), dtype=np.float32)
    psf /= psf.sum(axis=(0, 1), keepdims=True)
    image = np.stack([orc.conv2(u[..., c], psf[..., c], "valid") for c in range(3)], axis=2)
    image = (image * (1.0 + 0.05 * rng.standard_normal(image.shape))).astype(np.float32)
    s = Solver(M, N, K)
    s.upload(image, u, psf)
    g, mx = s.stage_chain()
    s.close()
    for c in range(3):
        e_ref = orc.conv2(u[..., c], psf[..., c], "valid") - image[..., c]
        g_ref = orc.conv2(e_ref, orc.rot180(psf[..., c]), "full")
        assert rel_l2(g[..., c], g_ref) < 2e-5, (c, rel_l2(g[..., c], g_ref))
        assert abs(mx[c] - u[..., c].max()) == 0
        assert abs(mx[3 + c] - np.abs(g_ref).max()) <= 1e-4 * np.abs(g_ref).max()


@pytest.mark.parametrize("name", __import__("helpers").tv_golden_names())
def test_tv_stencil_against_reference_executed_fixtures(name):
    """rltv_stage_tv (k_tv<ORDER, NORM>) against the reference's OWN TV() (pyx:137-239) run through the cpdef wrapper of
    oracle/build_ref_tv.py -- fixtures tests/golden/tv_o*n*.npz, all four (order, norm) pairs."""
    from helpers import GOLDEN
    from image_cases_studies_b200.solver import Solver
    z = np.load(GOLDEN / f"{name}.npz")
    u = z["u"]
    K = 3
    M, N = u.shape[0] - K + 1, u.shape[1] - K + 1
    s = Solver(M, N, K)
    s.upload(np.zeros((M, N, 3), np.float32), u, np.full((K, K, 3), 1 / 9, np.float32))
    out, div = s.stage_tv(int(z["order"]), int(z["norm"]), float(z["epsilon"]))
    s.close()
    assert np.abs(out - z["ref_out"]).max() <= 2e-6 * np.abs(z["ref_out"]).max()
    assert np.abs(div - z["ref_div"]).max() <= 2e-6 * np.abs(z["ref_div"]).max()


@pytest.mark.parametrize("name", __import__("helpers").tvmm_golden_names())
def test_tv_alive_mode_against_patched_reference(dc, name):
    """mode="mm_tv" (k_tv_maps / k_tv_grad / k_update_tv) against the PATCHED reference with the TV(ut) calls of pyx:464-465
    alive (oracle/build_ref_tv.py; fixtures tests/golden/tvmm_*.npz): estimate, PSF, the in-place denoised image."""
    g = load_golden(name)
    M, N = g["image"].shape[:2]
    MK = g["psf0"].shape[0]
    image, u, psf = g["image"].copy(), g["u0"].copy(), g["psf0"].copy()
    out = dc.richardson_lucy_MM(image, u, psf, *g["window"], g["tau"], M, N, 3, MK, g["iterations"], g["step_factor"],
                                g["lambd"], blind=g["blind"], mode="mm_tv")
    assert dc.last_stats["iterations"] == g["ref_iterations"]
    assert rel_l2(out, g["ref_out"]) <= TOL_IMAGE_REL_L2 and rel_l2(u, g["ref_u"]) <= TOL_IMAGE_REL_L2
    assert rel_l2(image, g["ref_image"]) <= TOL_IMAGE_REL_L2
    # (non-blind: epsilon = 1e-6 makes the image step ~1e-8 relative, below float32 resolution in the reference too)
    assert np.array_equal(image, g["image"]) == np.array_equal(g["ref_image"], g["image"])
    assert psf_l1(psf, g["ref_psf"]) <= TOL_PSF_L1
    # ... and it is a different result from the shipped arithmetic on the same inputs
    u0, p0 = g["u0"].copy(), g["psf0"].copy()
    dc.richardson_lucy_MM(g["image"].copy(), u0, p0, *g["window"], g["tau"], M, N, 3, MK, g["iterations"], g["step_factor"],
                          g["lambd"], blind=g["blind"])
    assert rel_l2(u0, g["ref_u"]) > 5e-4


def test_tv_alive_mode_against_oracle_larger_frame(dc):
    from image_cases_studies_b200 import synthetic
    from oracle import rl_mm_oracle as orc
    c = synthetic.make_case("c3_blind_24mp_k15", seed=17, scale=0.07, iterations=3)
    M, N = c.shape
    image, u, psf = c.image.copy(), c.u0.copy(), c.psf0.copy()
    out = dc.richardson_lucy_MM(image, u, psf, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor, c.lambd,
                                blind=True, mode="mm_tv")
    ref = orc.richardson_lucy_MM(c.image, c.u0, c.psf0, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor, c.lambd,
                                 blind=True, tv_alive=True)
    assert dc.last_stats["iterations"] == ref.iterations
    assert rel_l2(out, ref.out) <= TOL_IMAGE_REL_L2 and rel_l2(image, ref.image) <= TOL_IMAGE_REL_L2
    assert psf_l1(psf, ref.psf) <= TOL_PSF_L1


@pytest.mark.parametrize("name,scale,iters", [("c3_blind_24mp_k15", 0.05, 3), ("c1_nonblind_512_g5", 0.3, 3)])
def test_pam_collaborative_tv_mode_against_its_definition(dc, name, scale, iters):
    """mode="pam_ctv" (k_ctv_grad + k_update_tv<PAM>): UNPINNED by the reference -- no PAM solver or collaborative norm
    exists in the snapshot (README.md:42-44, :113-117 only).  Checked against the numpy statement of its definition
    (oracle/ctv_oracle.py, oracle/rl_mm_oracle.py pam_ctv=True), and against the default mode (it must differ)."""
    from image_cases_studies_b200 import synthetic
    from oracle import rl_mm_oracle as orc
    c = synthetic.make_case(name, seed=19, scale=scale, iterations=iters)
    M, N = c.shape
    image, u, psf = c.image.copy(), c.u0.copy(), c.psf0.copy()
    out = dc.richardson_lucy_MM(image, u, psf, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor, c.lambd,
                                blind=c.blind, mode="pam_ctv")
    ref = orc.richardson_lucy_MM(c.image, c.u0, c.psf0, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor, c.lambd,
                                 blind=c.blind, pam_ctv=True)
    assert dc.last_stats["iterations"] == ref.iterations
    assert rel_l2(out, ref.out) <= TOL_IMAGE_REL_L2 and psf_l1(psf, ref.psf) <= TOL_PSF_L1
    assert np.array_equal(image, c.image)                      # this mode leaves the blurry image alone
    mm = orc.richardson_lucy_MM(c.image, c.u0, c.psf0, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor, c.lambd,
                                blind=c.blind)
    assert rel_l2(mm.u, ref.u) > 1e-6


@pytest.mark.parametrize("name,scale,iters", [("c5_nonblind_4k_kaiser7", 0.12, 4), ("c1_nonblind_512_g5", 0.6, 4),
                                               ("c2_blind_2mp_k9", 0.2, 3)])
def test_small_psf_chain_path_against_oracle(dc, name, scale, iters, monkeypatch):
    """K = 5, 7, 9 through the chain kernel (the default on megapixel frames; forced here on small ones): non-blind K = 7
    and K = 5 (gradient by k_chain_fft, whiteness residual by the direct forward stencil) and blind K = 9 (chain kernel +
    fused PSF gradient), whole solves against the float64 oracle (pyx:460-656)."""
    from image_cases_studies_b200 import synthetic
    from oracle import rl_mm_oracle as orc
    monkeypatch.setenv("RLTV_CHAIN_MINPIX", "0")
    dc.clear_cache()
    try:
        c = synthetic.make_case(name, seed=5, scale=scale, iterations=iters)
        M, N = c.shape
        u, psf = c.u0.copy(), c.psf0.copy()
        out = dc.richardson_lucy_MM(c.image, u, psf, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor, c.lambd,
                                    blind=c.blind)
        ref = orc.richardson_lucy_MM(c.image, c.u0, c.psf0, *c.window, c.tau, M, N, 3, c.MK, c.iterations, c.step_factor,
                                     c.lambd, blind=c.blind)
        assert dc.last_stats["iterations"] == ref.iterations
        assert rel_l2(out, ref.out) <= 1e-4
        assert psf_l1(psf, ref.psf) <= 1e-4
        assert np.allclose(dc.last_stats["M_r_history"], ref.M_r, rtol=2e-5)
    finally:
        dc.clear_cache()
