"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/rltv_b200.h declares,
fails loudly (no CPU fallback) without a device, and the Python drop-in validates arguments like the reference."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def nat():
    import __graft_entry__ as g
    g.build()
    from image_cases_studies_b200 import _native
    return _native


def test_header_symbols_all_exported_and_bound(nat):
    header = (ROOT / "include" / "rltv_b200.h").read_text()
    declared = set(re.findall(r"\b(rltv_[a-z_0-9]+)\s*\(", header))
    declared -= {"rltv_status"}
    bound = {name for name, _, _ in nat.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(nat.lib, name)
    assert nat.lib.rltv_abi_version() == 2


def test_struct_layouts_match_header(nat):
    import ctypes as C
    assert C.sizeof(nat.Params) == 44
    assert C.sizeof(nat.Stats) == 4 * (2 + 2 + 3 + 1 + 1 + 1 + 1) + 4 * nat.RLTV_MAX_HISTORY


def test_no_cpu_fallback(nat):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert nat.lib.rltv_device_count() == 0
    from image_cases_studies_b200.solver import Solver
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Solver(32, 32, 5)
    from image_cases_studies_b200.lib import deconvolution as dc
    k = np.ones((3, 3, 3), np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dc.normalize_kernel(k, 3)


def test_dropin_argument_errors_match_reference(nat):
    from image_cases_studies_b200.lib import deconvolution as dc
    img = np.zeros((16, 16, 3), np.float32)
    u = np.zeros((20, 20, 3), np.float32)
    psf = np.ones((5, 5, 3), np.float32) / 25
    args = (1, 15, 1, 15, 0.0, 16, 16, 3, 5, 1, 1e-3, 1e4)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):       # same text Cython raises
        dc.richardson_lucy_MM(img.astype(np.float64), u, psf, *args)
    with pytest.raises(ValueError, match="wrong number of dimensions"):
        dc.richardson_lucy_MM(img[..., 0], u, psf, *args)
    with pytest.raises(ValueError):
        dc.richardson_lucy_MM(img, u[:-1], psf, *args)
    assert dc.DTYPE is np.float32


def test_psf_builders_match_reference_formulas():
    from image_cases_studies_b200.lib import utils
    for k in (utils.uniform_kernel(7), utils.gaussian_kernel(9, 2.0), utils.kaiser_kernel(7, 4.0),
              utils.poisson_kernel(5, 1.5), utils.motion_kernel(31), utils.lens_blur(9)):
        assert abs(k.sum() - 1) < 1e-12 and k.min() >= 0
    scipy_windows = pytest.importorskip("scipy.signal.windows")
    g = scipy_windows.gaussian(9, std=2.0)
    assert np.allclose(utils.gaussian_kernel(9, 2.0), np.outer(g, g) / np.outer(g, g).sum())
    e = scipy_windows.exponential(5, tau=1.5)
    assert np.allclose(utils.poisson_kernel(5, 1.5), np.outer(e, e) / np.outer(e, e).sum())


def test_synthetic_case_shapes():
    from image_cases_studies_b200 import synthetic
    c = synthetic.make_case("c2_blind_2mp_k9", seed=0, scale=0.05)
    M, N = c.shape
    assert c.u0.shape == (M + 8, N + 8, 3) and c.psf0.shape == (9, 9, 3) and c.image.dtype == np.float32
    assert c.image.min() > 0
