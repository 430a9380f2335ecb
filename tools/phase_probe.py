"""Per-phase cycle breakdown of the row-FFT stencil kernel on the headline workload (debug tool).

Needs a probe build:  RLTV_NVCC_EXTRA=-DRLTV_PHASE_PROBE python -c "import __graft_entry__ as g; g.build(force=True)"
(the production build leaves the counters out: they cost the kernel register spills)."""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g; g.build()
from image_cases_studies_b200 import synthetic, _native as nat
from image_cases_studies_b200.solver import Solver
case = synthetic.make_case("c3_blind_24mp_k15", seed=0)
M, N = case.shape
s = Solver(M, N, case.MK)
s.upload(case.image, case.u0, case.psf0)
s.ignore_stop(True)
p = Solver.make_params(case.window, case.tau, 10**6, case.step_factor, case.lambd, case.blind)
buf = (C.c_uint64 * 8)()
s.begin(p); s.enqueue_outer(2); s.finish()
nat.check(nat.lib.rltv_debug_phase_cycles(buf))
s.begin(p); s.enqueue_outer(2); st = s.finish()
nat.check(nat.lib.rltv_debug_phase_cycles(buf))
v = np.array(list(buf)[:6], dtype=np.float64)
names = ["wait_tma", "fwd_fft", "mac", "inv_fft", "epilogue", "flush"]
print({n: round(x / v.sum(), 3) for n, x in zip(names, v)}, "total cycles/CTA-launch", v.sum() / 148 / 30)
