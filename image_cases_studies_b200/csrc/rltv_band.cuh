// Row-band sharding across the GPUs of one NVSwitch box.  Everything that crosses GPUs is done by the
// kernels themselves with peer stores over NVLink into IPC-mapped memory of the other bands, ordered by
// step-numbered flags -- no host involvement and no NCCL inside an outer iteration:
//
//   halo exchange   the update kernel itself (k_update, rltv_elementwise.cuh) stores the first / last 2P owned rows it
//                   rewrites (lib/deconvolution.pyx:527-531, :552) into the bottom / top halo of the neighbours as
//                   well, and ends only when the neighbours' rows have landed here;
//   step scalars    the last CTA of the adjoint kernel publishes the band's max(u_c), max|G_c| (pyx:524) into slot
//                   [rank] of EVERY band's Comm block, then waits for all slots and takes the max;
//   PSF gradient    the last CTA of k_gradk / k_gradk_fft_finish sums the per-CTA partials (double, fixed order) and publishes the
//                   band's 3*K*K sums (pyx:571); k_psf_update waits for all bands and adds them in rank order, so
//                   every band computes the bit-identical PSF without a broadcast;
//   stop flag       the band holding the whiteness window publishes the stop decision (pyx:643-654).
//
// All-gather slots are double-buffered by the parity of the step number: a band can be at most one exchange
// ahead of any other (it needs everybody's contribution to pass), so parity is enough to rule out overwrites.
#pragma once
#include "rltv_common.cuh"

namespace rltv {

constexpr int MAXR = 8;             // GPUs of one box
constexpr int MAXKK = 47 * 47;   // RLTV_MAX_MK squared

// Lives at the tail of every band's u allocation (one IPC handle maps the band and its Comm block).
struct Comm {
  int halo_flag[2];                 // raised by the band above / below: their push number `seq` has landed
  int pad0[14];
  int max_flag[2][MAXR];            // [parity][source rank] = seq of the published step scalars
  int max_val[2][MAXR][8];          // 6 used: ord(max u_c), ord(max |G_c|)
  int gk_flag[2][MAXR];
  int stop_flag[2];                 // [parity] = seq of the published stop decision
  int stop_val[2];
  float mr_val[2];
  double gk_val[2][MAXR][3 * MAXKK];
};

struct CommPeers {
  Comm* peer[MAXR];                 // peer[r] = band r's Comm block (peer[rank] = own, local memory)
  int nranks, rank;
};

struct HaloSide {
  float* peer_u;        // neighbour's u planes (IPC-mapped), nullptr if there is no neighbour on this side
  int* peer_flag;       // flag word inside the neighbour's Comm block that THIS band raises
  size_t peer_plane;    // floats per plane in the neighbour's band
  int src_row;          // first local row to send
  int dst_row;          // first row in the neighbour's local coordinates
  int nrows;
};

__device__ __forceinline__ void spin_until(const int* flag, int seq) {
  const long long t0 = clock64();
  while (*reinterpret_cast<const volatile int*>(flag) < seq) {
    if (clock64() - t0 > 20000000000LL) __trap();   // a lost peer faults the launch (~10 s) instead of hanging
  }
}

// Called by ONE thread of the last CTA of the adjoint kernel: publish this band's step scalars to every band.
__device__ __forceinline__ void publish_step_max(State* st, int slot, const CommPeers& cp, int seq) {
  const int par = seq & 1;
  int v[6];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = atomicMax(&st->smax[slot][i], ORD_LOWEST);      // atomic read of the final value
    v[3 + i] = atomicMax(&st->smax[slot][3 + i], ORD_LOWEST);
  }
  for (int r = 0; r < cp.nranks; ++r) {
    volatile int* dst = cp.peer[r]->max_val[par][cp.rank];
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[i] = v[i];
  }
  __threadfence_system();
  for (int r = 0; r < cp.nranks; ++r) *reinterpret_cast<volatile int*>(&cp.peer[r]->max_flag[par][cp.rank]) = seq;
  __threadfence_system();
}

// Warp 0 of the last CTA of the adjoint kernel, right after publish_step_max: wait for every band's scalars of step
// `seq` and reduce them with max into the local State (pyx:524).  No band's publication depends on this wait, so
// there is no cycle; the update kernel that follows in stream order sees the global maxima.
__device__ __forceinline__ void gather_step_max(State* st, int slot, Comm* mine, int nranks, int seq, int lane) {
  const int par = seq & 1;
  if (lane < nranks) spin_until(&mine->max_flag[par][lane], seq);
  __syncwarp();
  __threadfence_system();
  if (lane < 6) {
    int m = ORD_LOWEST;
    for (int r = 0; r < nranks; ++r) m = max(m, *reinterpret_cast<volatile int*>(&mine->max_val[par][r][lane]));
    st->smax[slot][lane] = m;
  }
}

// Tail of the adjoint kernels with row bands: the last CTA to finish publishes this band's step scalars to every band
// (peer stores) and gathers everybody's.  `slot` is one free word of shared memory; call with all threads.
__device__ __forceinline__ void band_step_max_tail_slot(State* st, int sslot, const CommPeers& cp, int seq, unsigned* done_counter, int* slot) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
    if (last) {
      *done_counter = 0u;
      __threadfence();
      publish_step_max(st, sslot, cp, seq);
    }
    *slot = last;
  }
  __syncthreads();
  if (*slot && threadIdx.x < 32) gather_step_max(st, sslot, cp.peer[cp.rank], cp.nranks, seq, threadIdx.x);
}
__device__ __forceinline__ void band_step_max_tail(State* st, const CommPeers& cp, int seq, unsigned* done_counter, int* slot) {
  band_step_max_tail_slot(st, 0, cp, seq, done_counter, slot);
}

// Non-owning bands: wait for the stop decision of outer iteration `seq` from the band that holds the window.
__global__ void k_stop_gather(State* __restrict__ st, Comm* __restrict__ mine, int seq) {
  if (st->stop) return;
  if (threadIdx.x != 0) return;
  const int par = seq & 1;
  spin_until(&mine->stop_flag[par], seq);
  __threadfence_system();
  const int stop = *reinterpret_cast<volatile int*>(&mine->stop_val[par]);
  const float mr = *reinterpret_cast<volatile float*>(&mine->mr_val[par]);
  const int it = st->it;
  if (it > 0) st->M_r_prev = st->M_r;
  st->M_r = mr;
  if (it < 4096) st->hist[it] = mr;
  st->n_hist = min(it + 1, 4096);
  st->it = it + 1;
  if (stop) st->stop = 1;
}

// Owning band, one thread, after k_outer_finalize: publish (stop, M_r) of outer iteration `seq` to every band.
__global__ void k_stop_publish(const State* __restrict__ st, CommPeers cp, int seq, int it_expected) {
  if (threadIdx.x != 0) return;
  // Runs even when st->stop is set: the iteration that SET the flag must still tell the others.  Iterations the
  // host enqueued after the stop never advanced st->it, publish nothing, and nobody waits for them.
  if (st->it != it_expected) return;
  const int par = seq & 1;
  for (int r = 0; r < cp.nranks; ++r) {
    if (r == cp.rank) continue;
    *reinterpret_cast<volatile int*>(&cp.peer[r]->stop_val[par]) = st->stop;
    *reinterpret_cast<volatile float*>(&cp.peer[r]->mr_val[par]) = st->M_r;
  }
  __threadfence_system();
  for (int r = 0; r < cp.nranks; ++r)
    if (r != cp.rank) *reinterpret_cast<volatile int*>(&cp.peer[r]->stop_flag[par]) = seq;
  __threadfence_system();
}

}  // namespace rltv
