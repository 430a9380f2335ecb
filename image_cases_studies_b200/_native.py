"""ctypes binding of the C ABI in include/rltv_b200.h (csrc/ -> librltv_b200.so).

There is no fallback: if the shared library is missing, importing this module raises ImportError;
if no CUDA device is present every compute call raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
# RLTV_LIB: an alternative build of the same library (kernel experiments, e.g. another warp-role split); never a fallback
LIB_PATH = Path(os.environ["RLTV_LIB"]) if os.environ.get("RLTV_LIB") else PKG / "librltv_b200.so"

RLTV_MAX_HISTORY = 4096
RLTV_MAX_MK = 47
INNER_ITER = 5
MODE_MM, MODE_MM_TV, MODE_PAM_CTV = 0, 1, 2


class Params(C.Structure):
    _fields_ = [("top", C.c_int32), ("bottom", C.c_int32), ("left", C.c_int32), ("right", C.c_int32),
                ("tau", C.c_float), ("iterations", C.c_int32), ("step_factor", C.c_float), ("lambd", C.c_float),
                ("blind", C.c_int32), ("correlation", C.c_int32), ("mode", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("iterations_executed", C.c_int32), ("stopped", C.c_int32), ("M_r", C.c_float),
                ("M_r_prev", C.c_float), ("dt", C.c_float * 3), ("dtpsf", C.c_float), ("solve_ms", C.c_float),
                ("kernel_launches", C.c_int32), ("n_history", C.c_int32),
                ("M_r_history", C.c_float * RLTV_MAX_HISTORY)]

    def as_dict(self):
        return dict(iterations=int(self.iterations_executed), stopped=bool(self.stopped), M_r=float(self.M_r),
                    M_r_prev=float(self.M_r_prev), dt=tuple(float(x) for x in self.dt), dtpsf=float(self.dtpsf),
                    solve_ms=float(self.solve_ms), kernel_launches=int(self.kernel_launches),
                    M_r_history=[float(self.M_r_history[i]) for i in range(int(self.n_history))])


class Band(C.Structure):
    _fields_ = [("row_lo", C.c_int32), ("row_hi", C.c_int32), ("own_lo", C.c_int32), ("own_hi", C.c_int32)]


PH_OUTER_BEGIN, PH_GRAD, PH_UPDATE, PH_PSF_GRAD, PH_PSF_STEP, PH_OUTER_END = range(6)

# every symbol include/rltv_b200.h declares: (name, restype, argtypes)
_FP = C.POINTER(C.c_float)
SYMBOLS = [
    ("rltv_abi_version", C.c_int, []),
    ("rltv_last_error", C.c_char_p, []),
    ("rltv_device_count", C.c_int, []),
    ("rltv_richardson_lucy_mm", C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int32,
                                          C.c_int32, C.c_int32, C.POINTER(Params), C.POINTER(Stats), C.c_void_p, C.c_int32]),
    ("rltv_normalize_kernel", C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    ("rltv_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    ("rltv_destroy", C.c_int, [C.c_void_p]),
    ("rltv_upload", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("rltv_download", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    ("rltv_solve", C.c_int, [C.c_void_p, C.POINTER(Params), C.POINTER(Stats)]),
    ("rltv_begin", C.c_int, [C.c_void_p, C.POINTER(Params)]),
    ("rltv_enqueue_outer", C.c_int, [C.c_void_p, C.c_int32]),
    ("rltv_finish", C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    ("rltv_stream", C.c_void_p, [C.c_void_p]),
    ("rltv_set_ignore_stop", C.c_int, [C.c_void_p, C.c_int32]),
    ("rltv_profile_enable", C.c_int, [C.c_void_p, C.c_int32]),
    ("rltv_profile_get", C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    ("rltv_create_band", C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Band), C.c_void_p]),
    ("rltv_upload_band", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("rltv_download_rows", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    ("rltv_ipc_export", C.c_int, [C.c_void_p, C.c_void_p]),
    ("rltv_set_rank", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    ("rltv_ipc_attach", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]),
    ("rltv_set_whiteness_owner", C.c_int, [C.c_void_p, C.c_int32]),
    ("rltv_download_image", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    ("rltv_gather_alloc", C.c_int, [C.c_void_p, C.c_void_p]),
    ("rltv_gather_attach", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    ("rltv_gather_push", C.c_int, [C.c_void_p, C.c_uint32]),
    ("rltv_gather_download", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    ("rltv_enqueue_phase", C.c_int, [C.c_void_p, C.c_int32]),
    ("rltv_device_ptr", C.c_void_p, [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t)]),
    ("rltv_poll_record", C.c_int, [C.c_void_p, C.c_int32]),
    ("rltv_poll_wait", C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    ("rltv_stage_residual", C.c_int, [C.c_void_p, C.c_void_p]),
    ("rltv_stage_adjoint", C.c_int, [C.c_void_p, C.c_void_p]),
    ("rltv_debug_chain_roles", C.c_int, [C.c_int32]),
    ("rltv_stage_chain", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("rltv_stage_gradk", C.c_int, [C.c_void_p, C.c_void_p]),
    ("rltv_debug_download_err", C.c_int, [C.c_void_p, C.c_void_p]),
    ("rltv_stage_tv", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    ("rltv_debug_phase_cycles", C.c_int, [C.POINTER(C.c_uint64)]),
    ("rltv_debug_fft128", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    ("rltv_stage_whiteness", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
]

if not LIB_PATH.exists():
    raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). There is no CPU fallback.")
lib = C.CDLL(str(LIB_PATH))
for _name, _res, _args in SYMBOLS:
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.rltv_last_error() or b"").decode()


def check(rc: int):
    if rc == 0:
        return
    msg = last_error()
    if rc == -1:
        raise ValueError(msg)
    raise RuntimeError(f"rltv error {rc}: {msg}")


def ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)
