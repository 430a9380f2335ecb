"""CPU tests of the pyramid driver (image_cases_studies_b200/deconvolve.py) glue: the solver is injected, so the
whole deblur_module flow (views, in-place hand-offs, pyramid sizes, gamma, TIFF) runs here with the numpy oracle."""
import numpy as np
import pytest

from image_cases_studies_b200 import deconvolve as drv
from image_cases_studies_b200.lib import utils
from oracle import rl_mm_oracle as orc


class OracleSolver:
    """Adapts the float64 oracle to the in-place contract of lib/deconvolution.pyx (pyx:531, :581, :675)."""
    calls = []

    @staticmethod
    def normalize_kernel(kern, MK):
        k = kern.astype(np.float64)
        orc.normalize_kernel(k)
        kern[...] = k

    @classmethod
    def richardson_lucy_MM(cls, image, u, psf, top, bottom, left, right, tau, M, N, C, MK, iterations, step_factor,
                           lambd, blind=True, correlation=False, **kw):
        r = orc.richardson_lucy_MM(image, u, psf, top, bottom, left, right, tau, M, N, C, MK, iterations, step_factor,
                                   lambd, blind=blind, correlation=correlation)
        u[...] = r.u
        if blind:
            psf[...] = r.psf_caller
        cls.calls.append((blind, M, N, MK, r.iterations))
        pad = (u.shape[0] - M) // 2
        return u[pad:pad + M, pad:pad + N]


def test_build_pyramid_matches_reference_examples():
    images, kernels = drv.build_pyramid(15, 10)
    assert kernels == [15, 11, 7, 5, 3]                       # SURVEY.md section 3 (probe of deconvolve.py:40-60)
    assert np.allclose(images, [2 ** (-k / 2) for k in range(5)])
    assert drv.build_pyramid(3, 1) == ([1.], [3])
    assert drv.build_pyramid(7, 1)[1] == [7, 5, 3]


def test_pad_image_and_resize():
    img = np.arange(2 * 3 * 3, dtype=np.float32).reshape(2, 3, 3)
    p = drv.pad_image(img, (1, 1))
    assert p.shape == (4, 5, 3) and p.dtype == np.float32 and p.flags.c_contiguous
    assert np.array_equal(p[0, 0], img[0, 0]) and np.array_equal(p[-1, -1], img[-1, -1])
    r = drv.resize(np.ones((10, 12, 3)), (7, 9, 3))
    assert r.shape == (7, 9, 3) and np.allclose(r, 1.0)


def test_tiff_writer_roundtrip(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    a = rng.integers(0, 65535, (13, 17, 3)).astype(np.uint16)
    utils.save(a.astype(np.float64), "x", str(tmp_path))
    b = cv2.imread(str(tmp_path / "x.tif"), cv2.IMREAD_UNCHANGED)
    assert b is not None and b.dtype == np.uint16 and b.shape == (13, 17, 3)
    assert np.array_equal(b[..., ::-1], a)                    # OpenCV returns BGR


def test_deblur_module_flow_with_injected_oracle(tmp_path):
    rng = np.random.default_rng(5)
    sharp = rng.random((72, 84, 3)) * 200 + 20
    k = utils.gaussian_kernel(3, 0.8)
    blurred = np.stack([orc.conv2(np.pad(sharp[..., c], 1, mode="edge"), k, "valid") for c in range(3)], axis=2)
    OracleSolver.calls.clear()
    out = drv.deblur_module(blurred, "t", str(tmp_path), 3, mask=[36, 42], mask_size=31, iterations=2, display=False,
                            solver=OracleSolver)
    assert out.shape == (72, 84, 3) and np.isfinite(out).all() and out.min() >= 0 and out.max() <= 65535
    assert (tmp_path / "t.tif").exists()
    # one pyramid level for a 3-px blur: a blind call on the 31-px mask crop, then a non-blind call on the full frame
    assert [c[0] for c in OracleSolver.calls] == [True, False]
    assert OracleSolver.calls[0][1:4] == (33, 33, 3)
    assert OracleSolver.calls[1][1:4] == (72 + 2 + 1 + 2, 84 + 2 + 1 + 2, 3)   # +1 px pad, made odd, +1 px pad
    psf = drv.deblur_module.last_psf
    assert psf.shape == (3, 3, 3) and np.allclose(psf.sum(axis=(0, 1)), 1, atol=1e-5)
    # gamma round trip: the output stays close to the (linear) input scaled to 16 bit
    ref16 = (blurred / 255.0) * 65535
    assert np.abs(out - ref16).mean() / ref16.mean() < 0.05
    with pytest.raises(ValueError):
        drv.deblur_module(blurred, "t", str(tmp_path), 4, solver=OracleSolver, save=False)
    with pytest.raises(ValueError):
        drv.deblur_module(blurred, "t", str(tmp_path), 3, mask=[2, 2], solver=OracleSolver, save=False)
