#!/usr/bin/env python
"""Headline benchmark: MPix*iter/s of blind RL/MM deconvolution, 24 MP RGB, 15x15 PSF (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A *step* is one OUTER iteration of the solver (lib/deconvolution.pyx:460-656) on the whole frame: 5 inner
steps (forward blur, adjoint, update, PSF step) plus the whiteness statistic.  MPix*iter/s counts INNER
steps, as BASELINE.md defines it:  value = M*N*5*steps / seconds / 1e6.

* `value`    : frame resident in HBM, K steps timed with CUDA events on the solver's stream.
* `e2e`      : the drop-in call `dc.richardson_lucy_MM(image, u, psf, ...)` on pinned HOST arrays, one call =
               the workload's full outer-iteration count; H2D of image/u/psf and D2H of u/psf inside the timing.
* `roofline` : the dominant kernel family, algorithmic bytes per launch / its mean launch duration (CUDA
               events around every launch in a separate profiling pass), against MEASURED_PEAKS.json.
* `cpu_baseline` : the reference's own compiled Cython solver (oracle/_ref) on a bounded crop of the workload.
* `--impl reference` : times that same reference solver as the main metric (one step = one outer iteration
               on the bounded crop).

N > 1 (torchrun, one rank per GPU): the SAME frame is split into row bands, one per GPU (strong scaling):
halo rows, step scalars, PSF-gradient sums and the stop flag all move by in-kernel peer stores over NVLink
(image_cases_studies_b200/distributed.py; `--comm nccl` = host all-reduce baseline); device time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "MPix*iter/s, blind RL/MM deconvolution (TV-PAM lineage), 24MP RGB, 15x15 PSF"
UNIT = "MPix*iter/s"
DEFAULT_WORKLOAD = "c3_blind_24mp_k15"
INNER = 5
# algorithmic bytes per pixel per inner step (fp32 RGB = 12 B per full-frame array touch)
BYTES_PER_PX_STEP = {True: 132.0, False: 108.0}
# algorithmic bytes per pixel per LAUNCH of each kernel family (minimal operands of that stage)
FAMILY_BYTES_PER_PX = {"conv_fwd": 24.0,   # read u, image (residual consumed on chip in the ideal fused form)
                       "conv_adj": 36.0,   # read u, ut (for the step statistics), write g
                       "update": 60.0,     # read g, u, ut, image; write u
                       "gradk": 12.0}      # read u (residual on chip)


def csrc_hash() -> str:
    """sha256 (first 16 hex digits) over the CUDA sources: an ncu capture is only quoted for the kernels it measured."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / "image_cases_studies_b200" / "csrc").glob("*.cu*")):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def ncu_traffic(workload, family):
    """DRAM bytes per launch of `family` from the committed ncu --set full capture of this workload, or None when there
    is no capture for this workload or the capture was taken with different kernel sources (stamp mismatch)."""
    p = ROOT / "profiles" / "ncu_traffic_r02.json"
    if not p.exists():
        return None, None
    d = json.loads(p.read_text())
    if d.get("workload") != workload:
        return None, None
    if d.get("csrc_sha16") != csrc_hash():
        return None, f"profiles/{p.name} is stale (kernel sources changed since the capture): not quoted"
    return d["per_launch_dram_bytes"].get(family), f"profiles/{p.name} ({d.get('source')}, csrc {d.get('csrc_sha16')})"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp32_peak():
    p = ROOT / "profiles" / "microbench_r01.json"
    if p.exists():
        d = json.loads(p.read_text())
        return max(float(d.get("ffma_tflops_ilp16", 0)), float(d.get("ffma2_tflops_ilp16", 0))), "measured (profiles/microbench_r01.json)"
    return 148 * 128 * 2 * 1.965e9 / 1e12, "nominal 148 SM x 128 FMA x 1.965 GHz"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def pinned(a: np.ndarray) -> np.ndarray:
    import torch
    t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
    out = t.numpy()
    out[...] = a
    return out


# --------------------------------------------------------------------------------------------------------
def cpu_reference_sample(workload: str, steps: int, warmup: int, force_port: bool = False):
    """Times the reference's CPU implementation on a bounded crop: one step = one outer iteration.

    kind "reference": the reference's own compiled solver (oracle/_ref, all host cores through its OpenMP loops);
    kind "port": the numpy restatement (oracle/rl_mm_oracle.py, float32) when oracle/_ref is not present."""
    from image_cases_studies_b200 import synthetic
    from oracle import ref_loader
    have_ref = (not force_port) and ref_loader.load() is not None
    _, _, K, _, blind, _ = synthetic.WORKLOADS[workload]
    M = 600 if have_ref else 300
    full_M = synthetic.WORKLOADS[workload][0]
    c = synthetic.make_case(workload, seed=0, scale=min(1.0, M / full_M), iterations=1)
    M, N = c.shape
    if have_ref:
        def one():
            ref_loader.run(c.image, c.u0, c.psf0, c.window, c.tau, 1, c.step_factor, c.lambd, c.blind)
    else:
        from oracle import rl_mm_oracle

        def one():
            rl_mm_oracle.richardson_lucy_MM(c.image, c.u0, c.psf0, *c.window, c.tau, M, N, 3, c.MK, 1, c.step_factor,
                                            c.lambd, blind=c.blind, dtype=np.float32)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    how = ("the reference's compiled solver; throughput per pixel is size-independent: it convolves by FFT" if have_ref
           else "numpy port of the reference algorithm, single thread")
    return {"value": M * N * INNER * len(times) / total / 1e6, "unit": UNIT, "cores": os.cpu_count() if have_ref else 1,
            "kind": "reference" if have_ref else "port",
            "sample": f"{M}x{N} crop of {workload}, MK={K}, {'blind' if blind else 'non-blind'}, {len(times)} outer iteration(s) "
                      f"of 5 inner steps each ({how})",
            "seconds": total, "ms_per_step": 1e3 * total / len(times)}


def workload_config(workload, M, N, K, blind):
    """The `config` object of the JSON line: the workload, identical for the GPU arm and the reference arm."""
    return {"workload": workload, "frame": [M, N, 3], "psf": K, "blind": blind,
            "step": "one outer iteration = 5 inner steps + whiteness statistic",
            "l2": "working set (5 planar frame copies, 1.4 GB at 24 MP) far exceeds the 126 MB L2; no flush needed"}


def cpu_reference_full_frame(workload: str):
    """The reference's own compiled solver on the FULL frame of the workload (time >= 2 outer
    iterations and extrapolate per step, the per-step cost is constant).  One untimed call on a small crop first (loads
    the extension, spins up its OpenMP team and scipy's FFT plans)."""
    from image_cases_studies_b200 import synthetic
    from oracle import ref_loader
    if ref_loader.load() is None:
        return None
    full_M, full_N, K, _, blind, _ = synthetic.WORKLOADS[workload]
    small = synthetic.make_case(workload, seed=0, scale=min(1.0, 256.0 / full_M), iterations=1)
    ref_loader.run(small.image, small.u0, small.psf0, small.window, small.tau, 1, small.step_factor, small.lambd, small.blind)
    n_outer = 2 if full_M * full_N <= 30e6 else 1
    c = synthetic.make_case(workload, seed=0, scale=1.0, iterations=n_outer)
    M, N = c.shape
    t0 = time.perf_counter()
    _, _, _, log = ref_loader.run(c.image, c.u0, c.psf0, c.window, c.tau, n_outer, c.step_factor, c.lambd, c.blind)
    total = time.perf_counter() - t0
    done = max(ref_loader.executed_iterations(log), 1)
    return {"value": M * N * INNER * done / total / 1e6, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
            "sample": f"the FULL {M}x{N} frame of {workload}, MK={K}, {'blind' if blind else 'non-blind'}: {done} outer iteration(s) of 5 "
                      f"inner steps timed in one call of the reference's compiled solver ({total:.1f} s); per-step cost is constant, "
                      "so the per-step figure stands for every step of the run (BASELINE.md section 3)",
            "seconds": total, "ms_per_step": 1e3 * total / done, "shape": (M, N, K, blind), "steps_timed": done}


def run_reference_impl(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from image_cases_studies_b200 import synthetic
    r = None if args.reference_crop else cpu_reference_full_frame(args.workload)
    if r is None:                                           # oracle/_ref absent: the numpy port on a crop
        r = cpu_reference_sample(args.workload, min(args.steps, 3), min(args.warmup, 1))
        M, N, K, _, blind, _ = synthetic.WORKLOADS[args.workload]
        r["steps_timed"] = min(args.steps, 3)
    else:
        M, N, K, blind = r["shape"]
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args.workload, M, N, K, blind),
            "steps_timed": r["steps_timed"],
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------------------------------------
def run_frame_batch(args, rank, dev, world):
    """BASELINE config 5: a batch of independent frames, sharded across ranks (embarrassingly parallel, no collective).
    Every frame goes host -> device -> solve (workload's outer iterations) -> host; two contexts on two streams per GPU
    so the copies of one frame overlap the kernels of the other.  One 'step' = one frame."""
    import torch
    import torch.distributed as dist
    from image_cases_studies_b200 import synthetic
    from image_cases_studies_b200.solver import Solver
    per_rank = args.frames // world
    base = synthetic.make_case(args.workload, seed=rank, scale=args.scale)
    M, N = base.shape
    K = base.MK
    params = Solver.make_params(base.window, base.tau, base.iterations, base.step_factor, base.lambd, base.blind)
    # synthetic frames differ by a per-frame offset of the same scene (generation of 256 distinct 4K frames on the host
    # would dominate the run); contents do not change the work per frame
    img_h, u_h, psf_h = pinned(base.image), pinned(base.u0), pinned(base.psf0)
    out_h = [pinned(base.u0), pinned(base.u0)]
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    solvers = [Solver(M, N, K, device=dev, stream=s.cuda_stream) for s in streams]
    def one(i):
        s = solvers[i & 1]
        s.upload(img_h, u_h, psf_h)
        st = s.solve(params)
        s.download(u=out_h[i & 1])
        return st
    for i in range(2):
        one(i)                                             # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    import concurrent.futures as cf
    t0 = time.perf_counter()
    iters = 0
    launches = 0
    with cf.ThreadPoolExecutor(2) as ex:                   # two host threads, one per context/stream
        for st in ex.map(one, range(per_rank)):
            iters += st["iterations"]
            launches += st["kernel_launches"]
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    t = torch.tensor([secs, float(iters), float(launches)], device=f"cuda:{dev}", dtype=torch.float64)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        secs, iters, launches = float(tm[0]), float(t[1]), float(t[2])
    value = M * N * INNER * iters / secs / 1e6
    if rank == 0:
        line = {"metric": "MPix*iter/s, non-blind RL/MM deconvolution, batch of 4K RGB frames, 7x7 PSF", "value": value, "unit": UNIT,
                "n_gpus": world, "steps": per_rank * world, "warmup": 2, "ms_per_step": 1e3 * secs / max(per_rank, 1),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": args.workload, "frames": per_rank * world, "frame": [M, N, 3], "psf": K,
                           "blind": base.blind, "outer_iterations_per_frame": base.iterations,
                           "parallelism": f"{per_rank} frames per GPU x {world} GPUs, no collective; 2 streams per GPU",
                           "step": "one frame: H2D, solve, D2H"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": img_h.nbytes + u_h.nbytes + psf_h.nbytes,
                        "d2h_bytes_per_step": u_h.nbytes},
                "roofline": None, "cpu_baseline": None, "gpu_launches": int(launches)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------
def emit(line):
    """The ONE JSON line of the contract, on the real stdout (see main: fd 1 is parked on stderr while running)."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


_REAL_STDOUT = None


def main():
    # libraries (NCCL's version banner, torchrun notices) write to fd 1: keep stdout for the JSON line alone
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the frame (debugging only; invalid as a result)")
    ap.add_argument("--frame", default=None, help="MxN override of the frame size (debugging only; invalid as a result)")
    ap.add_argument("--e2e-calls", type=int, default=2)
    ap.add_argument("--e2e-iterations", type=int, default=None)
    ap.add_argument("--comm", default="fused", choices=["fused", "nccl"],
                    help="N>1: in-kernel peer-memory exchanges (default) or the host/NCCL all-reduce baseline")
    ap.add_argument("--frames", type=int, default=0,
                    help="batch mode (BASELINE config 5): this many independent frames of the workload, sharded across the "
                         "ranks (no collective), each solved end to end from pinned host memory")
    ap.add_argument("--reference-crop", action="store_true",
                    help="--impl reference: time the 600x900 crop (the quick cpu_baseline sample) instead of the full frame")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_impl(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank, local_rank, world = dist_env()
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ge.build()
    from image_cases_studies_b200 import synthetic
    from image_cases_studies_b200.lib import deconvolution as dc
    from image_cases_studies_b200.solver import Solver

    dev = local_rank
    torch.cuda.set_device(dev)
    if args.frames > 0:
        return run_frame_batch(args, rank, dev, world)
    shape = tuple(int(v) for v in args.frame.lower().split("x")) if args.frame else None
    case = synthetic.make_case(args.workload, seed=0, scale=args.scale, shape=shape)   # every rank builds the same frame
    M, N = case.shape
    K = case.MK
    params = Solver.make_params(case.window, case.tau, 10 ** 6, case.step_factor, case.lambd, case.blind)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        stream = torch.cuda.Stream(device=dev)
        solver = Solver(M, N, K, device=dev, stream=stream.cuda_stream)
        solver.upload(case.image, case.u0, case.psf0)
    else:
        from image_cases_studies_b200.distributed import BandSolver
        from image_cases_studies_b200 import distributed as rl_dist
        solver = BandSolver(M, N, K, case.window, device=dev, comm=args.comm)
        stream = solver.tstream
        solver.upload(case.image, case.u0, case.psf0)

    # ---- device-resident steps (value) -------------------------------------------------------------
    # The blind stop rule fires after 3 outer iterations on some synthetic inputs (it is data dependent); the
    # timed steady-state steps keep evaluating it but do not act on it.  The e2e call below obeys it.
    solver.ignore_stop(True)
    with torch.cuda.stream(stream):
        solver.begin(params)
        with ClockSampler(dev) as clocks:          # sampled under the same load: warm-up, timed steps, untimed tail
            solver.enqueue_outer(args.warmup)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            solver.enqueue_outer(args.steps)
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            tail = max(0, min(200, int(600.0 / max(ms / args.steps, 1e-3)) - args.steps - args.warmup))
            solver.enqueue_outer(tail)             # keeps the GPU busy ~0.6 s in total so nvidia-smi sees the load
            barrier()
        st = solver.finish()
    launches_total = st["kernel_launches"]
    launches_timed = round(launches_total * args.steps / (args.steps + args.warmup + tail))
    executed = st["iterations"]
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = M * N * INNER * args.steps / (ms * 1e-3) / 1e6
    # a stop inside the timed window would turn later steps into no-ops: flag it
    valid_steps = (executed == args.steps + args.warmup + tail)

    # ---- per-family profile pass (roofline of the dominant kernel; rank 0's band when sharded) ---------
    with torch.cuda.stream(stream):
        solver.upload(case.image, case.u0, case.psf0)
        solver.profile_enable(True)
        solver.begin(params)
        solver.enqueue_outer(2)
        solver.finish()
        prof = solver.profile()
        solver.profile_enable(False)
    rows_frac = 1.0 if world == 1 else (solver.band[3] - solver.band[2]) / float(M + K - 1)
    hbm_peak, peak_src = peaks()
    fam_ms = {f: (t / n if n else 0.0) for f, (t, n) in prof.items()}
    fam_tot = {f: t for f, (t, n) in prof.items()}
    tot = sum(fam_tot.values()) or 1.0
    dom = max(("conv_fwd", "conv_adj", "update", "gradk"), key=lambda f: fam_tot.get(f, 0.0))
    # default kernel family of the library (rltv_api.cu): row-FFT / chain kernels for K >= 11, and for K = 9 on megapixel frames
    conv_mode = os.environ.get("RLTV_CONV", "fft" if (K >= 11 or (K == 9 and M * N >= (1 << 20))) else "direct")
    fused_residual = case.blind and conv_mode == "fft" and K <= 17 and os.environ.get("RLTV_FUSE", "1") != "0"
    fam_bytes = dict(FAMILY_BYTES_PER_PX)
    # chain kernel (9 <= K <= 17 on the row-FFT path): forward blur, residual and adjoint are ONE launch (k_chain_fft), timed under the conv_adj family:
    # reads u, ut, image (as packed spectra), writes g
    chain = conv_mode == "fft" and 9 <= K <= 17 and os.environ.get("RLTV_CHAIN", "1") != "0" and (world == 1 or args.comm == "fused")
    if chain:
        fam_bytes["conv_adj"] = 48.0
    if fused_residual:
        fam_bytes["gradk"] = 24.0        # the PSF-gradient kernel computes the residual itself: reads u and the image
    alg_bytes = fam_bytes[dom] * M * N * rows_frac
    # the row-FFT PSF gradient is two launches (accumulate + finish) timed under one family: report them per gradient
    gk_launches = 2 if conv_mode == "fft" else 1
    if chain:
        flops_note = "forward + adjoint in one launch"
    if prof["gradk"][1]:
        fam_ms["gradk"] = fam_tot["gradk"] / max(prof["gradk"][1] // gk_launches, 1)
    dom_ms = fam_ms[dom]
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms else 0.0
    fpk, fpk_src = fp32_peak()
    flops_launch = 2.0 * 3 * K * K * M * N * rows_frac * (2.0 if (chain and dom == "conv_adj") else 1.0)
    step_bytes = BYTES_PER_PX_STEP[case.blind] * M * N * INNER
    ms_step = ms / args.steps
    # measured DRAM traffic of the same kernel: only meaningful for the whole frame on one GPU at the profiled size
    traffic, traffic_src = (ncu_traffic(case.name, dom) if (world == 1 and not args.frame and args.scale == 1.0) else (None, None))
    kname = {"conv_adj": "chain" if chain else "conv_adj"}.get(dom, dom)
    roofline = {"bound": "hbm", "kernel": f"{kname}<{K}>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dom_ms,
                "share_of_step": fam_tot[dom] / tot,
                # FLOPs of the DIRECT K x K stencil (2*3*K*K per pixel).  For 11 <= K <= 17 the kernels evaluate it through
                # row FFTs (4K FMAs per row-bin + two 128-point FFTs), so the direct-equivalent rate may exceed the FMA peak:
                # that ratio is the FLOP reduction, not a utilisation.
                "fp32": {"direct_flop_per_launch": flops_launch,
                         "direct_equivalent_tflops": flops_launch / (dom_ms * 1e-3) / 1e12 if dom_ms else 0.0,
                         "peak_tflops": fpk, "peak_source": fpk_src,
                         "direct_equivalent_over_peak": (flops_launch / (dom_ms * 1e-3) / 1e12) / fpk if dom_ms else 0.0,
                         "algorithm": "row-FFT hybrid" if conv_mode == "fft" else "direct stencil"},
                "step": {"algorithmic_bytes": step_bytes, "achieved_gbs": step_bytes / (ms_step * 1e-3) / 1e9,
                         "frac_of_hbm_all_gpus": step_bytes / (ms_step * 1e-3) / 1e9 / (hbm_peak * world)},
                "family_ms_per_launch": fam_ms, "family_share": {f: t / tot for f, t in fam_tot.items()}}

    # ---- end to end through the drop-in API, host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        iters = args.e2e_iterations or case.iterations
        img_h, u_h, psf_h = pinned(case.image), pinned(case.u0), pinned(case.psf0)
        dc.clear_cache()
        solver.close()
        h2d = img_h.nbytes + u_h.nbytes + psf_h.nbytes
        d2h = u_h.nbytes + psf_h.nbytes
        done_iters, secs, launches_call = 0, 0.0, 0
        for i in range(1 + args.e2e_calls):       # first call is warm-up
            u_h[...] = case.u0
            psf_h[...] = case.psf0
            barrier()
            t0 = time.perf_counter()
            if world == 1:
                dc.richardson_lucy_MM(img_h, u_h, psf_h, *case.window, case.tau, M, N, 3, K, iters, case.step_factor,
                                      case.lambd, blind=case.blind)
                stats = dc.last_stats
            else:
                rl_dist.richardson_lucy_MM(img_h, u_h, psf_h, *case.window, case.tau, M, N, 3, K, iters,
                                           case.step_factor, case.lambd, blind=case.blind, comm=args.comm, gather="root")
                stats = rl_dist.richardson_lucy_MM.last_stats
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i > 0:
                done_iters += stats["iterations"]
                secs += dt
                launches_call = stats["kernel_launches"]
        if world > 1:
            t = torch.tensor([secs], device=f"cuda:{dev}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        e2e = {"value": M * N * INNER * done_iters / secs / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d if world == 1 else h2d // world, "d2h_bytes_per_step": d2h if world == 1 else d2h // world,
               "step": f"one richardson_lucy_MM call of {iters} outer iterations "
               f"({done_iters // max(args.e2e_calls, 1)} executed) on pinned host arrays"
               + ("" if world == 1 else "; per-rank band upload/download, result bands gathered on rank 0 over NCCL"),
               "calls": args.e2e_calls, "seconds_per_call": secs / max(args.e2e_calls, 1), "gpu_launches_per_call": launches_call}
        dc.clear_cache()
        if world > 1:
            rl_dist.clear_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_sample(args.workload, 2, 1)
        if r:
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": workload_config(args.workload, M, N, K, case.blind),
                "run": {"parallelism": "single GPU" if world == 1 else
                           (f"one frame in {world} row bands, one per GPU; halo rows, step scalars, PSF-gradient sums and the stop "
                            "flag all move by in-kernel peer stores over NVLink (no NCCL in the loop)" if args.comm == "fused" else
                            f"one frame in {world} row bands, one per GPU: halo rows by NVLink peer stores, NCCL all-reduce of "
                            "6 step scalars + 3*MK^2 PSF-gradient sums per inner step (baseline)"),
                           "all_steps_live": valid_steps,
                           "stop_rule": "evaluated every step, not acted upon in the timed region (obeyed in e2e)"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_timed,
                "clocks": clocks.summary()}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
