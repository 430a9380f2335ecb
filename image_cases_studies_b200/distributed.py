"""Row-band sharding of ONE frame across the GPUs of one NVSwitch box.

One process per GPU (torchrun); every rank calls the same functions with the same host arrays (SPMD).
The padded estimate ``u`` is cut into contiguous row bands; each band also holds a halo of 2*(MK/2) rows of
its neighbours.  Per inner step (lib/deconvolution.pyx:473-591):

    GRAD      forward blur + adjoint on the band; the adjoint kernel's last CTA publishes the band's
              max(u_c), max|G_c| (pyx:524) into every band's memory, then waits for everybody's and takes the max
    UPDATE    gradient step + blend, then the band PUSHES its edge rows into the neighbours' halos; the push
              kernel ends when the neighbours' rows have landed here too
    PSF_GRAD  residual, PSF-gradient partial sums; the finishing CTA publishes the band's 3*MK*MK sums
              (pyx:571) to every band
    PSF_STEP  waits for all bands' sums, adds them in rank order, identical tiny update on every rank
              (the replicated PSF stays bit-identical without a broadcast)

and once per outer iteration the band that holds the whiteness window evaluates the stop rule (pyx:623-654) and
publishes the decision.  All of this is peer stores over NVLink into CUDA-IPC-mapped memory ordered by
step-numbered flags (csrc/rltv_band.cuh): with ``comm="fused"`` (default) an outer iteration is a pure kernel
sequence -- no NCCL, no host round trip.  ``comm="nccl"`` keeps the baseline in which the host all-reduces the
three tiny buffers between phases through torch.distributed; it exists to measure what the fusion buys.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat
from .solver import Solver, _rows_ok, default_device

INNER_ITER = nat.INNER_ITER


# ------------------------------------------------------------------------------------------------------
# host-side planning (pure Python: unit-tested on CPU, also under gloo with world_size 2)
# ------------------------------------------------------------------------------------------------------
def plan_bands(M: int, MK: int, world: int, window=None):
    """Cut the Hu = M+MK-1 rows of u into `world` bands.

    Returns ``(bands, owner)``: bands[r] = (row_lo, row_hi, own_lo, own_hi) in frame rows, ``owner`` the rank
    whose band evaluates the whiteness statistic for ``window = (top, bottom, left, right)`` (image rows).
    Every band owns >= 2P rows (halos come from the immediate neighbours only); the window must lie inside the
    rows its owner OWNS (the fused PSF-gradient kernel leaves the residual of the owned rows only); when an even cut
    would split the window, the cuts are chosen to minimise the largest band under that constraint.
    """
    P = MK // 2
    Hu = M + MK - 1
    if world < 1:
        raise ValueError("world size must be >= 1")
    cuts = [round(i * Hu / world) for i in range(world + 1)]
    owner = 0
    if window is not None and world > 1:
        wt, wb = window[0] + P, window[1] + P            # window rows in u coordinates
        owner = max(i for i in range(world) if cuts[i] <= wt)
        if wb > cuts[owner + 1]:
            # The window straddles an even cut.  Smallest possible largest band: k bands above a <= wt, the owner
            # [a, b) with b >= wb, world-1-k bands below -- search the band height T and the owner index k.  (Round 2's
            # first version only moved the cut below the window: at 8 bands of a 4000-row frame the owner then had 629
            # rows against 470 below it, and every kernel of every step waited for that band.)
            best = None
            for T in range(-(-Hu // world), Hu + 1):
                for k in range(world):
                    a = 0 if k == 0 else min(wt, k * T)
                    b = Hu if k == world - 1 else max(wb, Hu - (world - 1 - k) * T)
                    if a <= wt and b >= wb and b - a <= T and a >= 2 * P * k and Hu - b >= 2 * P * (world - 1 - k):
                        best = (k, a, b)
                        break
                if best:
                    break
            if best is None:
                raise ValueError("whiteness window straddles a band boundary")
            k, a, b = best
            owner = k
            cuts = ([round(i * a / k) for i in range(k)] if k else []) + [a] + \
                   [b + round(j * (Hu - b) / (world - 1 - k)) for j in range(world - 1 - k)] + [Hu]
        if wt < cuts[owner] or wb > cuts[owner + 1]:
            raise ValueError("whiteness window straddles a band boundary")
    bands = []
    for r in range(world):
        own_lo, own_hi = cuts[r], cuts[r + 1]
        if world > 1 and own_hi - own_lo < max(2 * P, 1):
            raise ValueError(f"band {r} would own {own_hi - own_lo} rows < 2*(MK//2) = {2 * P}: use fewer GPUs")
        bands.append((max(own_lo - 2 * P, 0), min(own_hi + 2 * P, Hu), own_lo, own_hi))
    return bands, owner


def image_rows_of_band(band, M: int, MK: int):
    """Image rows [i0, i1) whose u-rows (image row + P) fall inside the rows a band holds."""
    P = MK // 2
    row_lo, row_hi = band[0], band[1]
    return max(row_lo - P, 0), min(row_hi - P, M)


class _DevMem:
    """Zero-copy torch view of a raw device pointer (through __cuda_array_interface__)."""
    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


# ------------------------------------------------------------------------------------------------------
class BandSolver:
    """One rank's band of a frame.  Needs an initialised torch.distributed NCCL process group."""

    def __init__(self, M, N, MK, window, device=None, group=None, comm="fused"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.M, self.N, self.MK = int(M), int(N), int(MK)
        self.device = default_device() if device is None else int(device)
        torch.cuda.set_device(self.device)
        self.bands, self.owner = plan_bands(self.M, self.MK, self.world, window)
        self.band = self.bands[self.rank]
        self.tstream = torch.cuda.Stream(device=self.device)
        self._ctx = C.c_void_p()
        b = nat.Band(*self.band)
        nat.check(nat.lib.rltv_create_band(C.byref(self._ctx), self.device, self.M, self.N, self.MK, C.byref(b),
                                           C.c_void_p(self.tstream.cuda_stream)))
        if comm not in ("fused", "nccl"):
            raise ValueError("comm must be 'fused' or 'nccl'")
        self.fused = comm == "fused"
        nat.check(nat.lib.rltv_set_whiteness_owner(self._ctx, int(self.rank == self.owner)))
        nat.check(nat.lib.rltv_set_rank(self._ctx, self.rank, self.world, int(self.fused)))
        # exchange CUDA IPC handles of the u allocations (they carry the halo flags and all-gather slots);
        # fused: map every band, NCCL baseline: the two neighbours are enough
        h = (C.c_char * 64)()
        nat.check(nat.lib.rltv_ipc_export(self._ctx, h))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(h.raw), group=group)
        for peer in range(self.world):
            if peer == self.rank or (not self.fused and abs(peer - self.rank) != 1):
                continue
            hb = (C.c_char * 64).from_buffer_copy(handles[peer])
            nat.check(nat.lib.rltv_ipc_attach(self._ctx, peer, hb, self.bands[peer][0], self.bands[peer][1]))
        dist.barrier(group=group)
        dev = f"cuda:{self.device}"
        def view(name, typestr, itemsize):
            n = C.c_size_t()
            p = nat.lib.rltv_device_ptr(self._ctx, name.encode(), C.byref(n))
            return torch.as_tensor(_DevMem(p, n.value // itemsize, typestr), device=dev)
        self.t_step_max = view("step_max", "<i4", 4)
        self.t_gk_sum = view("gk_sum", "<f8", 8)
        self.t_stop = view("stop", "<i4", 4)
        self.stats = None

    def close(self):
        if self._ctx:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)      # nobody unmaps a band a neighbour may still push into
            nat.lib.rltv_destroy(self._ctx)
            self._ctx = C.c_void_p()

    # -- data movement: every rank passes the FULL host arrays and moves only its band -----------------
    def upload(self, image, u, psf):
        row_lo, row_hi = self.band[0], self.band[1]
        i0, i1 = image_rows_of_band(self.band, self.M, self.MK)
        img = image[i0:i1]
        ub = u[row_lo:row_hi]
        if not _rows_ok(img):
            img = np.ascontiguousarray(img, dtype=np.float32)
        if not _rows_ok(ub):
            ub = np.ascontiguousarray(ub, dtype=np.float32)
        psf = np.ascontiguousarray(psf, dtype=np.float32)
        nat.check(nat.lib.rltv_upload_band(self._ctx, nat.ptr(img), img.strides[0], i0, i1 - i0, nat.ptr(ub),
                                           ub.strides[0], nat.ptr(psf)))
        # No barrier here with comm="fused": the first peer store into a neighbour's halo happens in the first update
        # kernel, which follows the first gradient kernel, whose tail waits for the step statistics of EVERY band --
        # i.e. for every band's first kernel, which its own upload precedes in stream order.
        if not self.fused:
            self.dist.barrier(group=self.group)

    def download(self, u, psf_caller=None, psf_refined=None, gather="all"):
        """Writes this band's owned rows into ``u``.  gather="all": every rank ends up with the full frame (the SPMD
        drop-in contract); "root": only rank 0 does (the other ranks keep their own band); False/None: no exchange."""
        torch, dist = self.torch, self.dist
        if gather is True:
            gather = "all"
        own_lo, own_hi = self.band[2], self.band[3]
        if gather and self.world > 1 and (gather == "all" or self.rank == 0):
            own_hi = own_lo                          # this rank receives the whole frame through the device gather below
        rows = u[own_lo:own_hi]
        tmp = rows if (_rows_ok(rows) and rows.flags.writeable) else np.empty(rows.shape, np.float32)
        pc = psf_caller if psf_caller is None or psf_caller.flags.c_contiguous else np.empty_like(psf_caller, order="C")
        pr = psf_refined if psf_refined is None or psf_refined.flags.c_contiguous else np.empty_like(psf_refined, order="C")
        nat.check(nat.lib.rltv_download_rows(self._ctx, nat.ptr(tmp), tmp.strides[0], own_lo, own_hi - own_lo,
                                             nat.ptr(pc) if pc is not None else None,
                                             nat.ptr(pr) if pr is not None else None))
        if tmp is not rows:
            rows[...] = tmp
        if psf_caller is not None and pc is not psf_caller:
            psf_caller[...] = pc
        if psf_refined is not None and pr is not psf_refined:
            psf_refined[...] = pr
        if gather and self.world > 1:
            self._gather_on_device(u, gather)
        return u

    def _gather_on_device(self, u, gather):
        """Result bands travel GPU to GPU: every band converts its owned rows planar -> HWC straight into the full-frame
        staging buffer of the destination rank(s) (peer stores over NVLink into CUDA-IPC-mapped memory), then each
        destination copies the whole frame to its host array in ONE transfer.  (Round 1 bounced every band through the
        host twice and through NCCL send/recv: ~145 ms of a 305 ms call at 8 GPUs.)"""
        dist = self.dist
        dests = list(range(self.world)) if gather == "all" else [0]
        key = tuple(dests)
        if getattr(self, "_gather_ready", None) != key:
            h = (C.c_char * 64)()
            mine = self.rank in dests
            if mine:
                nat.check(nat.lib.rltv_gather_alloc(self._ctx, h))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(h.raw) if mine else None, group=self.group)
            for r in dests:
                if r != self.rank:
                    hb = (C.c_char * 64).from_buffer_copy(handles[r])
                    nat.check(nat.lib.rltv_gather_attach(self._ctx, r, hb))
            self._gather_ready = key
        mask = 0
        for r in dests:
            mask |= 1 << r
        nat.check(nat.lib.rltv_gather_push(self._ctx, mask))       # synchronises this rank's stream
        dist.barrier(group=self.group)                             # every band has landed in every destination
        if self.rank in dests:
            if _rows_ok(u) and u.flags.writeable:
                nat.check(nat.lib.rltv_gather_download(self._ctx, nat.ptr(u), u.strides[0]))
            else:
                tmp = np.empty(u.shape, np.float32)
                nat.check(nat.lib.rltv_gather_download(self._ctx, nat.ptr(tmp), tmp.strides[0]))
                u[...] = tmp
        return u

    # -- stepping ------------------------------------------------------------------------------------------
    def begin(self, params: nat.Params):
        self.params = params
        with self.torch.cuda.stream(self.tstream):
            nat.check(nat.lib.rltv_begin(self._ctx, C.byref(params)))

    def _phase(self, ph):
        nat.check(nat.lib.rltv_enqueue_phase(self._ctx, ph))

    def enqueue_outer(self, n: int = 1, first_it: int | None = None):
        """Enqueue n outer iterations on this rank's stream; no host synchronisation."""
        if self.fused:
            with self.torch.cuda.stream(self.tstream):
                for k in range(n):
                    nat.check(nat.lib.rltv_enqueue_outer(self._ctx, 1))
                    if first_it is not None:
                        nat.check(nat.lib.rltv_poll_record(self._ctx, first_it + k))
            return
        dist, group = self.dist, self.group
        MAX, SUM = dist.ReduceOp.MAX, dist.ReduceOp.SUM
        blind = bool(self.params.blind)
        with self.torch.cuda.stream(self.tstream):
            for k in range(n):
                self._phase(nat.PH_OUTER_BEGIN)
                for _ in range(INNER_ITER):
                    self._phase(nat.PH_GRAD)
                    dist.all_reduce(self.t_step_max, op=MAX, group=group)
                    self._phase(nat.PH_UPDATE)
                    if blind:
                        self._phase(nat.PH_PSF_GRAD)
                        dist.all_reduce(self.t_gk_sum, op=SUM, group=group)
                        self._phase(nat.PH_PSF_STEP)
                self._phase(nat.PH_OUTER_END)
                dist.all_reduce(self.t_stop, op=MAX, group=group)
                if first_it is not None:
                    nat.check(nat.lib.rltv_poll_record(self._ctx, first_it + k))

    def finish(self) -> dict:
        st = nat.Stats()
        with self.torch.cuda.stream(self.tstream):
            nat.check(nat.lib.rltv_finish(self._ctx, C.byref(st)))
        stats = st.as_dict()
        # the whiteness history lives on the owning band
        box = [stats if self.rank == self.owner else None]
        src = self.dist.get_global_rank(self.group, self.owner) if self.group is not None else self.owner
        self.dist.broadcast_object_list(box, src=src, group=self.group)
        own = box[0]
        stats.update(M_r=own["M_r"], M_r_prev=own["M_r_prev"], M_r_history=own["M_r_history"])
        self.stats = stats
        return stats

    def solve(self, params: nat.Params) -> dict:
        """The reference's outer loop (pyx:460): every rank takes the same stop decision from the same flag."""
        self.begin(params)
        for it in range(int(params.iterations)):
            if it >= 2:
                stop = C.c_int32()
                nat.check(nat.lib.rltv_poll_wait(self._ctx, it - 2, C.byref(stop)))
                if stop.value:
                    break
            self.enqueue_outer(1, first_it=it)
        return self.finish()

    def ignore_stop(self, on=True):
        nat.check(nat.lib.rltv_set_ignore_stop(self._ctx, int(on)))

    def profile_enable(self, on=True):
        nat.check(nat.lib.rltv_profile_enable(self._ctx, int(on)))

    def profile(self) -> dict:
        out = {}
        for fam in ("conv_fwd", "conv_adj", "update", "gradk", "psf", "stats", "copy", "halo"):
            ms, n = C.c_float(), C.c_int32()
            nat.check(nat.lib.rltv_profile_get(self._ctx, fam.encode(), C.byref(ms), C.byref(n)))
            out[fam] = (float(ms.value), int(n.value))
        return out


_cache: dict = {}


def clear_cache():
    """Release the cached band contexts (collective: every rank must call it)."""
    while _cache:
        _cache.popitem()[1].close()


def richardson_lucy_MM(image, u, psf, top, bottom, left, right, tau, M, N, C_, MK, iterations, step_factor, lambd,
                       blind=True, correlation=False, group=None, comm="fused", gather="all", **ignored):
    """SPMD drop-in: every rank of the process group calls this with the same arrays; on return every rank's ``u``
    and ``psf`` hold the full result, exactly as the single-GPU ``lib.deconvolution.richardson_lucy_MM`` leaves them
    (``gather="root"``: only rank 0's ``u`` is complete).  Band contexts (device buffers, IPC mappings) are cached per
    geometry, like the single-GPU module caches its contexts; ``clear_cache()`` releases them."""
    M, N, MK = int(M), int(N), int(MK)
    key = (M, N, MK, int(top), int(bottom), int(left), int(right), comm, id(group))
    s = _cache.get(key)
    if s is None:
        clear_cache()
        s = _cache[key] = BandSolver(M, N, MK, (top, bottom, left, right), group=group, comm=comm)
    try:
        s.upload(image, u, psf)
        params = Solver.make_params((top, bottom, left, right), tau, iterations, step_factor, lambd, blind, correlation)
        stats = s.solve(params)
        s.download(u, psf_caller=psf if blind else None, gather=gather)
    except Exception:
        _cache.pop(key, None)
        raise
    pad = (u.shape[0] - M) // 2
    richardson_lucy_MM.last_stats = stats
    return u[pad:pad + M, pad:pad + N, ...]
