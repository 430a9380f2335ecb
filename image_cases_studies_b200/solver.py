"""Device-resident solver handle over the persistent-context half of the C ABI.

``lib/deconvolution.py`` (the drop-in) and ``bench.py`` are both built on this: upload once, run outer
iterations on the GPU, download.  All arithmetic happens in the CUDA kernels under csrc/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as nat

FAMILIES = ("conv_fwd", "conv_adj", "update", "gradk", "psf", "stats", "copy", "halo")


def default_device() -> int:
    if "RLTV_DEVICE" in os.environ:
        return int(os.environ["RLTV_DEVICE"])
    return int(os.environ.get("LOCAL_RANK", "0"))


def _rows_ok(a: np.ndarray) -> bool:
    """Packed RGB float32 pixels with a non-negative row stride: what the C ABI takes without a copy."""
    return (a.dtype == np.float32 and a.ndim == 3 and a.shape[2] == 3 and a.strides[2] == 4
            and a.strides[1] == 12 and a.strides[0] >= 12 * a.shape[1])


class Solver:
    def __init__(self, M: int, N: int, MK: int, device: int | None = None, stream: int | None = None):
        self.M, self.N, self.MK = int(M), int(N), int(MK)
        self.device = default_device() if device is None else int(device)
        self._ctx = C.c_void_p()
        nat.check(nat.lib.rltv_create(C.byref(self._ctx), self.device, self.M, self.N, self.MK,
                                      C.c_void_p(stream) if stream else None))
        self.stats = None

    # -- lifetime -----------------------------------------------------------------------------------
    def close(self):
        if self._ctx:
            nat.lib.rltv_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def u_shape(self):
        return (self.M + self.MK - 1, self.N + self.MK - 1, 3)

    # -- data movement ------------------------------------------------------------------------------
    def upload(self, image=None, u=None, psf=None):
        keep = []
        def prep(a, shape):
            if a is None:
                return None, 0
            if a.shape != shape:
                raise ValueError(f"array of shape {a.shape}, expected {shape}")
            if not _rows_ok(a):
                a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
            return nat.ptr(a), a.strides[0]
        ip, irs = prep(image, (self.M, self.N, 3))
        up, urs = prep(u, self.u_shape)
        pp = None
        if psf is not None:
            if psf.shape != (self.MK, self.MK, 3):
                raise ValueError(f"psf of shape {psf.shape}, expected {(self.MK, self.MK, 3)}")
            psf = np.ascontiguousarray(psf, dtype=np.float32)
            keep.append(psf)
            pp = nat.ptr(psf)
        nat.check(nat.lib.rltv_upload(self._ctx, ip, irs, up, urs, pp))

    def download(self, u=None, psf_caller=None, psf_refined=None):
        """Writes into the given arrays (u may be a row-strided view) and returns them."""
        tmp_u = None
        up, urs = None, 0
        if u is not None:
            if u.shape != self.u_shape:
                raise ValueError("bad u shape")
            if _rows_ok(u) and u.flags.writeable:
                up, urs = nat.ptr(u), u.strides[0]
            else:
                tmp_u = np.empty(self.u_shape, dtype=np.float32)
                up, urs = nat.ptr(tmp_u), tmp_u.strides[0]
        def pk(a):
            if a is None:
                return None, None
            t = a if (a.flags.c_contiguous and a.dtype == np.float32) else np.empty((self.MK, self.MK, 3), np.float32)
            return t, nat.ptr(t)
        tc, pc = pk(psf_caller)
        tr, pr = pk(psf_refined)
        nat.check(nat.lib.rltv_download(self._ctx, up, urs, pc, pr))
        if tmp_u is not None:
            u[...] = tmp_u
        if psf_caller is not None and tc is not psf_caller:
            psf_caller[...] = tc
        if psf_refined is not None and tr is not psf_refined:
            psf_refined[...] = tr
        return u, psf_caller, psf_refined

    # -- running ------------------------------------------------------------------------------------
    @staticmethod
    def make_params(window, tau, iterations, step_factor, lambd, blind, correlation=False, mode="mm") -> nat.Params:
        top, bottom, left, right = (int(v) for v in window)
        modes = {"mm": nat.MODE_MM, "mm_tv": nat.MODE_MM_TV, "pam_ctv": nat.MODE_PAM_CTV}
        if mode not in modes:
            raise ValueError("mode must be 'mm' (the reference's shipped arithmetic), 'mm_tv' (its TV term alive) or "
                             "'pam_ctv' (PAM step with the collaborative TV norm; unpinned extension)")
        return nat.Params(top, bottom, left, right, float(tau), int(iterations), float(step_factor), float(lambd),
                          int(bool(blind)), int(bool(correlation)), modes[mode])

    def download_image(self, image: np.ndarray) -> np.ndarray:
        """The blurry image as it is on the device (mode="mm_tv" denoises it in place, pyx:547-549)."""
        if image.shape != (self.M, self.N, 3):
            raise ValueError("bad image shape")
        t = image if (_rows_ok(image) and image.flags.writeable) else np.empty((self.M, self.N, 3), np.float32)
        nat.check(nat.lib.rltv_download_image(self._ctx, nat.ptr(t), t.strides[0]))
        if t is not image:
            image[...] = t
        return image

    def solve(self, params: nat.Params) -> dict:
        st = nat.Stats()
        nat.check(nat.lib.rltv_solve(self._ctx, C.byref(params), C.byref(st)))
        self.stats = st.as_dict()
        return self.stats

    def begin(self, params: nat.Params):
        nat.check(nat.lib.rltv_begin(self._ctx, C.byref(params)))

    def enqueue_outer(self, n: int = 1):
        nat.check(nat.lib.rltv_enqueue_outer(self._ctx, int(n)))

    def finish(self) -> dict:
        st = nat.Stats()
        nat.check(nat.lib.rltv_finish(self._ctx, C.byref(st)))
        self.stats = st.as_dict()
        return self.stats

    @property
    def stream(self) -> int:
        return int(nat.lib.rltv_stream(self._ctx) or 0)

    def ignore_stop(self, on: bool = True):
        """Benchmark stepping: evaluate the stop rule but keep iterating (see include/rltv_b200.h)."""
        nat.check(nat.lib.rltv_set_ignore_stop(self._ctx, int(on)))

    def profile_enable(self, on: bool = True):
        nat.check(nat.lib.rltv_profile_enable(self._ctx, int(on)))

    def profile(self) -> dict:
        out = {}
        for fam in FAMILIES:
            ms, n = C.c_float(), C.c_int32()
            nat.check(nat.lib.rltv_profile_get(self._ctx, fam.encode(), C.byref(ms), C.byref(n)))
            out[fam] = (float(ms.value), int(n.value))
        return out

    # -- stage-level entry points (parity tests) ----------------------------------------------------
    def stage_residual(self) -> np.ndarray:
        out = np.empty((self.M, self.N, 3), np.float32)
        nat.check(nat.lib.rltv_stage_residual(self._ctx, nat.ptr(out)))
        return out

    def stage_adjoint(self) -> np.ndarray:
        out = np.empty(self.u_shape, np.float32)
        nat.check(nat.lib.rltv_stage_adjoint(self._ctx, nat.ptr(out)))
        return out

    def stage_chain(self):
        """g = adjoint(forward residual) from the spectral chain kernel, and the step statistics it reduces
        (max u_c, max |g_c|): returns (g, max6)."""
        out = np.empty(self.u_shape, np.float32)
        mx = np.empty(6, np.float32)
        nat.check(nat.lib.rltv_stage_chain(self._ctx, nat.ptr(out), nat.ptr(mx)))
        return out, mx

    def stage_gradk(self) -> np.ndarray:
        out = np.empty((self.MK, self.MK, 3), np.float32)
        nat.check(nat.lib.rltv_stage_gradk(self._ctx, nat.ptr(out)))
        return out

    def debug_residual(self) -> np.ndarray:
        """The residual buffer currently on the device (whatever kernel wrote it last)."""
        out = np.empty((self.M, self.N, 3), np.float32)
        nat.check(nat.lib.rltv_debug_download_err(self._ctx, nat.ptr(out)))
        return out

    def stage_tv(self, order: int, norm: int, epsilon: float):
        """TV norm map and divergence of the device-resident estimate (lib/deconvolution.pyx:137-239)."""
        out = np.empty(self.u_shape, np.float32)
        div = np.empty(self.u_shape, np.float32)
        ms = C.c_float()
        nat.check(nat.lib.rltv_stage_tv(self._ctx, int(order), int(norm), float(epsilon), nat.ptr(out), nat.ptr(div), C.byref(ms)))
        self.last_tv_ms = float(ms.value)
        return out, div

    def stage_whiteness(self, window) -> float:
        v = C.c_float()
        top, bottom, left, right = (int(x) for x in window)
        nat.check(nat.lib.rltv_stage_whiteness(self._ctx, top, bottom, left, right, C.byref(v)))
        return float(v.value)
