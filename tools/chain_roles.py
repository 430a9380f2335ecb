#!/usr/bin/env python
"""Time the chain kernel with subsets of its warp roles disabled (results are garbage; timing only).
usage: python tools/chain_roles.py [M N K]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from image_cases_studies_b200 import _native as nat, synthetic
from image_cases_studies_b200.solver import Solver

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (4000, 6000, 15)
rng = np.random.default_rng(0)
u = (0.1 + 0.8 * rng.random((M + K - 1, N + K - 1, 3), dtype=np.float32))
img = (0.1 + 0.8 * rng.random((M, N, 3), dtype=np.float32))
psf = np.full((K, K, 3), 1.0 / (K * K), np.float32)
s = Solver(M, N, K)
s.upload(img, u, psf)
# bits: 1 forward FFT, 2 E, 4 inverse FFT, 8 TMA loads, 16 G, 32 epilogue
names = {0: "all roles", 63: "loop + barriers only", 61: "E only", 47: "G only", 45: "E + G only", 54: "fwdFFT + TMA only",
         59: "IFFT only", 31: "EPI only", 18: "no E, no G", 32: "no EPI", 4 | 32: "no IFFT, no EPI", 1 | 8: "no TMA + fwdFFT",
         2: "no E", 16: "no G"}
for mask, name in names.items():
    nat.check(nat.lib.rltv_debug_chain_roles(mask))
    ts = []
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            nat.check(nat.lib.rltv_stage_chain(s._ctx, None, None))
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) / 5)
    print(f"mask {mask:2d} {name:32s} {min(ts) * 1e3:.3f} ms per stage call (includes ~0.02 ms of launch + memset + readback)")
nat.check(nat.lib.rltv_debug_chain_roles(0))
