// Row-band sharding across the GPUs of one NVSwitch box: halo exchange of the estimate `u` by direct peer
// stores over NVLink (the neighbour's band is mapped through CUDA IPC), ordered by step-numbered flags that
// live in the same peer allocation.  No host involvement, no NCCL on this path.
//
// After the update kernel rewrote the owned rows of u (lib/deconvolution.pyx:527-531, :552), each band pushes
// its first / last 2P owned rows into the bottom / top halo of the previous / next band; the next kernel that
// reads u (forward blur, pyx:477 / :557) is preceded by k_halo_wait on the flags the neighbours raise.
#pragma once
#include "rltv_common.cuh"

namespace rltv {

struct HaloSide {
  float* peer_u;        // neighbour's u planes (IPC-mapped), nullptr if there is no neighbour on this side
  int* peer_flag;       // flag word inside the neighbour's allocation that THIS band raises
  size_t peer_plane;    // floats per plane in the neighbour's band
  int src_row;          // first local row to send
  int dst_row;          // first row in the neighbour's local coordinates
  int nrows;
};

// grid-stride float4 copy of nrows x pitch x 3 planes to each neighbour, then (last CTA) raise the flags.
__global__ void __launch_bounds__(256)
k_halo_push(Geom g, const State* __restrict__ st, const float* __restrict__ u, HaloSide top, HaloSide bot,
            unsigned* __restrict__ done_counter, int seq) {
  if (st->stop) return;
  const int row4 = g.pitch / 4;
  const HaloSide sides[2] = {top, bot};
#pragma unroll
  for (int sd = 0; sd < 2; ++sd) {
    const HaloSide& h = sides[sd];
    if (!h.peer_u) continue;
    const int per_plane = h.nrows * row4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * per_plane; i += gridDim.x * blockDim.x) {
      const int c = i / per_plane, r = i - c * per_plane;
      const float4 v = reinterpret_cast<const float4*>(u + size_t(c) * g.plane + size_t(h.src_row) * g.pitch)[r];
      reinterpret_cast<float4*>(h.peer_u + size_t(c) * h.peer_plane + size_t(h.dst_row) * g.pitch)[r] = v;
    }
  }
  __threadfence_system();          // peer stores of this thread are visible system-wide before the flag
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *done_counter = 0u;
    __threadfence_system();
    if (top.peer_u) *reinterpret_cast<volatile int*>(top.peer_flag) = seq;
    if (bot.peer_u) *reinterpret_cast<volatile int*>(bot.peer_flag) = seq;
    __threadfence_system();
  }
}

// One warp: lanes 0/1 spin until the neighbours' pushes number `seq` have landed in this band's halo rows.
// Bounded: a lost peer faults the launch after ~10 s instead of hanging the GPU.
__global__ void k_halo_wait(const State* __restrict__ st, const int* flag_from_top, const int* flag_from_bot, int seq) {
  if (st->stop) return;
  const int* f = threadIdx.x == 0 ? flag_from_top : (threadIdx.x == 1 ? flag_from_bot : nullptr);
  if (f) {
    const long long t0 = clock64();
    while (*reinterpret_cast<const volatile int*>(f) < seq) {
      if (clock64() - t0 > 20000000000LL) __trap();
    }
  }
  __threadfence_system();
}

}  // namespace rltv
