// Elementwise / reduction kernels of one inner step: the fused gradient-step + depth-of-field blend,
// the PSF step + clip-and-normalise projection, and layout conversion at the host boundary.
#pragma once
#include "rltv_band.cuh"
#include "rltv_common.cuh"

namespace rltv {

// Start of a solve: both statistics slots empty (State was zeroed just before).
__global__ void k_state_init(State* __restrict__ st) {
  if (threadIdx.x < 6) {
    const int s = threadIdx.x / 3, c = threadIdx.x % 3;
    st->smax[s][c] = ORD_LOWEST;
    st->smax[s][3 + c] = 0;
    st->tvmax[s][c] = 0;
    st->tvmax[s][3 + c] = 0;
    st->tvmax[s][6 + c] = ORD_LOWEST;
  }
}

// ------------------------------------------------------------------------------------------------
// K3: u -= dt_c * (lambda*g + (u-ut)/2) on the whole padded domain (pyx:519-531), then on the interior
//     DoF = ((g - I)/(g + I))^2 [/lambda if non-blind] ; u = (1-DoF)*u + DoF*I          (pyx:499-502, :552)
// One float4 per thread per plane row; all five operands share the same (aligned) geometry.
// ------------------------------------------------------------------------------------------------
// `first`: first inner step of an outer iteration.  ut == u there (pyx:462 `ut[:] = u.copy()`), so instead of a separate
// 24 B/px copy pass the update reads u only and WRITES the majoriser: ut <- u_old.  Same bytes as a normal step.
// Row bands: the update kernel also does the halo exchange.  Threads that rewrite one of the first / last 2P owned rows
// store the new values straight into the neighbour's halo rows as well (peer stores over NVLink); the CTA that completes
// a side raises the neighbour's step-numbered flag; and one designated CTA waits for the NEIGHBOURS' flags, so that
// when this kernel ends the band's own halo rows are up to date for whatever reads u next.  (Round 1 did this in a
// separate copy kernel, k_halo_push: 22 us per inner step that did not shrink with the band.)
struct HaloPush {
  HaloSide top, bot;
  unsigned* counters;          // [0] top side, [1] bottom side
  const int* flag_from_top;    // raised by the band above when ITS rows have landed here (nullptr: no neighbour)
  const int* flag_from_bot;
  int seq;
  int enabled;
};

template <bool FIRST>
__global__ void __launch_bounds__(256)
k_update(Geom g, State* __restrict__ st, float* __restrict__ u, const float* __restrict__ ut, float* __restrict__ ut_out,
         const float* __restrict__ gbuf, const float* __restrict__ img, float step, float lambd, int blind,
         int slot, int reset_slot, HaloPush hp) {
  if (st->stop) return;
  const int c = blockIdx.z;
  const int Y = g.own0 + blockIdx.y;          // only owned rows are updated; halo rows arrive from the neighbours
  const int X = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  // dt = step_factor * (amax(u_c) + 0) / (amax|gradu_c| + 1e-15), float32 like the reference (pyx:524)
  const float dt = step * ord2f(st->smax[slot][c]) / (ord2f(st->smax[slot][3 + c]) + 1e-15f);
  if (X == 0 && blockIdx.y == 0) {
    st->dt[c] = dt;
    if (reset_slot >= 0) {               // the statistics slot of the NEXT inner step (nobody touches it now)
      st->smax[reset_slot][c] = ORD_LOWEST;
      st->smax[reset_slot][3 + c] = 0;
    }
  }
  if (X >= g.Wu && !hp.enabled) return;
  const bool active = X < g.Wu;
  const size_t off = size_t(c) * g.plane + size_t(Y) * g.pitch + (active ? X : 0);
  const float4 uv = *reinterpret_cast<const float4*>(u + off);
  const float4 tv = FIRST ? uv : *reinterpret_cast<const float4*>(ut + off);
  const float4 gv = *reinterpret_cast<const float4*>(gbuf + off);
  const float4 iv = *reinterpret_cast<const float4*>(img + off);
  const float uu[4] = {uv.x, uv.y, uv.z, uv.w}, tt[4] = {tv.x, tv.y, tv.z, tv.w};
  const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, ii[4] = {iv.x, iv.y, iv.z, iv.w};
  const int gy = g.row0 + Y;
  const bool rowin = (gy >= g.P) && (gy < g.P + g.M);
  float o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float G = fmaf(lambd, gg[i], 0.5f * (uu[i] - tt[i]));
    float un = fmaf(-dt, G, uu[i]);
    if (rowin && (X + i) >= g.P && (X + i) < g.P + g.N) {
      const float d = (gg[i] - ii[i]) / (gg[i] + ii[i]);
      float dof = d * d;
      if (!blind) dof = dof / lambd;
      un = (1.f - dof) * un + dof * ii[i];
    }
    o[i] = ((X + i) < g.Wu) ? un : uu[i];
  }
  if (active) {
    *reinterpret_cast<float4*>(u + off) = make_float4(o[0], o[1], o[2], o[3]);
    if (FIRST) *reinterpret_cast<float4*>(ut_out + off) = uv;
  }
  if (hp.enabled) {
    const bool te = hp.top.peer_u && Y >= hp.top.src_row && Y < hp.top.src_row + hp.top.nrows;     // CTA-uniform
    const bool be = hp.bot.peer_u && Y >= hp.bot.src_row && Y < hp.bot.src_row + hp.bot.nrows;
    if (active && te)
      *reinterpret_cast<float4*>(hp.top.peer_u + size_t(c) * hp.top.peer_plane + size_t(hp.top.dst_row + Y - hp.top.src_row) * g.pitch + X) =
          make_float4(o[0], o[1], o[2], o[3]);
    if (active && be)
      *reinterpret_cast<float4*>(hp.bot.peer_u + size_t(c) * hp.bot.peer_plane + size_t(hp.bot.dst_row + Y - hp.bot.src_row) * g.pitch + X) =
          make_float4(o[0], o[1], o[2], o[3]);
    if (te || be) {                  // CTA-uniform: only the CTAs that pushed rows pay for the system-scope fence
      __threadfence_system();        // this thread's peer stores are visible system-wide before any flag
      __syncthreads();
      if (threadIdx.x == 0) {
        if (te && atomicAdd(hp.counters + 0, 1u) == unsigned(gridDim.x * hp.top.nrows * 3 - 1)) {
          hp.counters[0] = 0u;
          __threadfence_system();
          *reinterpret_cast<volatile int*>(hp.top.peer_flag) = hp.seq;
        }
        if (be && atomicAdd(hp.counters + 1, 1u) == unsigned(gridDim.x * hp.bot.nrows * 3 - 1)) {
          hp.counters[1] = 0u;
          __threadfence_system();
          *reinterpret_cast<volatile int*>(hp.bot.peer_flag) = hp.seq;
        }
      }
    }
    // ONE designated CTA (the last one in launch order) does not leave before the neighbours' rows have landed here, so
    // the kernel as a whole does not complete before that.  It depends on the NEIGHBOURS' update kernels only.
    if (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1 && threadIdx.x < 2) {
      const int* f = threadIdx.x == 0 ? hp.flag_from_top : hp.flag_from_bot;
      if (f) spin_until(f, hp.seq);
      __threadfence_system();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5: PSF step (pyx:574-589).  One block.  psf / psf_caller layout: planar [3][K*K].
//   gk[q] = gk'[K-1-q] (gk' = gk_sum, the double-precision sum over tiles -- and over row bands); dtpsf = step/K * amax(psf) / (amax|gk| + 1e-15);
//   psf -= dtpsf * gk ; (correlation: every channel = channel mean) ; clip < 0 ; divide by channel sum.
// `psf_caller` reproduces what the CALLER's array holds in the reference: it tracks psf, except that with
// `correlation` it freezes after the first un-normalised step (pyx:581 writes in place, pyx:585 rebinds).
// ------------------------------------------------------------------------------------------------
// Body of the PSF step for one whole CTA (any multiple of 32 threads up to 512); `sm` = 6*K*K floats of shared memory.
__device__ __forceinline__ void psf_update_body(State* __restrict__ st, const double* __restrict__ gk_sum, int K, float step,
                                                int correlation, float* __restrict__ psf, float* __restrict__ psf_caller,
                                                Comm* __restrict__ mine, int nranks, int seq, float* __restrict__ sm) {
  const int par = seq & 1;
  if (nranks > 1) {
    // row bands: wait for every band's published PSF-gradient sums of this step (peer stores into `mine`)
    if (threadIdx.x < nranks) spin_until(&mine->gk_flag[par][threadIdx.x], seq);
    __syncthreads();
    __threadfence_system();
  }
  const int KK2 = K * K;
  float* gk = sm;              // [3][KK2]
  float* pk = sm + 3 * KK2;    // [3][KK2]
  __shared__ float red[16];
  __shared__ float s_dtp;
  __shared__ double s_sum[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mg = 0.f, mp = -INFINITY;
  for (int i = tid; i < 3 * KK2; i += blockDim.x) {
    const int c = i / KK2, q = i - c * KK2;
    const int qy = q / K, qx = q - qy * K;
    const int o = (K - 1 - qy) * K + (K - 1 - qx);
    double gsum;
    if (nranks > 1) {
      gsum = 0.0;                       // rank order: every band adds the same numbers in the same order
      for (int r = 0; r < nranks; ++r) gsum += *reinterpret_cast<volatile double*>(&mine->gk_val[par][r][c * KK2 + o]);
    } else {
      gsum = gk_sum[c * KK2 + o];
    }
    const float v = float(gsum);
    gk[i] = v;
    const float p = psf[i];
    pk[i] = p;
    mg = fmaxf(mg, fabsf(v));
    mp = fmaxf(mp, p);
  }
  mg = warp_max(mg);
  if (lane == 0) red[warp] = mg;
  __syncthreads();
  if (tid == 0) { float m = 0.f; for (int w = 0; w < int(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]); s_dtp = m; }
  __syncthreads();
  const float amax_gk = s_dtp;
  __syncthreads();
  mp = warp_max(mp);
  if (lane == 0) red[warp] = mp;
  __syncthreads();
  if (tid == 0) {
    float m = -INFINITY; for (int w = 0; w < int(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    const float dtp = (step / float(K)) * m / (amax_gk + 1e-15f);      // pyx:574
    s_dtp = dtp;
    st->dtpsf = dtp;
  }
  __syncthreads();
  const float dtp = s_dtp;
  const bool first = (st->blind_steps == 0);
  for (int i = tid; i < 3 * KK2; i += blockDim.x) {
    const float v = fmaf(-dtp, gk[i], pk[i]);                          // pyx:581
    pk[i] = v;
    if (correlation && first) psf_caller[i] = v;                       // the caller's array stops here (pyx:585)
  }
  __syncthreads();
  if (correlation) {
    for (int q = tid; q < KK2; q += blockDim.x) {
      const float m = (pk[q] + pk[KK2 + q] + pk[2 * KK2 + q]) / 3.f;   // np.mean(psf, axis=2), pyx:585
      pk[q] = m; pk[KK2 + q] = m; pk[2 * KK2 + q] = m;
    }
    __syncthreads();
  }
  // clip + per-channel sum (pyx:58-64); warp c sums channel c in a fixed order
  if (warp < 3) {
    double s = 0.0;
    for (int q = lane; q < KK2; q += 32) {
      float v = pk[warp * KK2 + q];
      if (v < 0.f) v = 0.f;
      pk[warp * KK2 + q] = v;
      s += double(v);
    }
    s = warp_sum(s);
    if (lane == 0) s_sum[warp] = s;
  }
  __syncthreads();
  for (int i = tid; i < 3 * KK2; i += blockDim.x) {
    const float v = pk[i] / float(s_sum[i / KK2]);                     // pyx:67-70
    psf[i] = v;
    if (!correlation) psf_caller[i] = v;
  }
  if (tid == 0) st->blind_steps += 1;
}

__global__ void __launch_bounds__(256)
k_psf_update(State* __restrict__ st, const double* __restrict__ gk_sum, int K, float step, int correlation,
             float* __restrict__ psf, float* __restrict__ psf_caller, Comm* __restrict__ mine, int nranks, int seq) {
  if (st->stop) return;
  extern __shared__ float sm[];
  psf_update_body(st, gk_sum, K, step, correlation, psf, psf_caller, mine, nranks, seq, sm);
}

// normalize_kernel (pyx:73-75): same clip + normalise on a planar [3][K*K] buffer, one block.
__global__ void __launch_bounds__(128)
k_normalize_kernel(float* __restrict__ kern, int KK2) {
  __shared__ double s_sum[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 3) {
    double s = 0.0;
    for (int q = lane; q < KK2; q += 32) {
      float v = kern[warp * KK2 + q];
      if (v < 0.f) v = 0.f;
      kern[warp * KK2 + q] = v;
      s += double(v);
    }
    s = warp_sum(s);
    if (lane == 0) s_sum[warp] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * KK2; i += blockDim.x) kern[i] = kern[i] / float(s_sum[i / KK2]);
}

// ------------------------------------------------------------------------------------------------
// Host-boundary layout conversion: packed HWC rows <-> planar u-geometry planes.
// src: rows x cols pixels, `src_pitch` floats per row (packed RGB); dst planes at offset (oy, ox).
// ------------------------------------------------------------------------------------------------
__global__ void k_hwc_to_planar(const float* __restrict__ src, size_t src_pitch, int rows, int cols,
                                float* __restrict__ dst, Geom g, int oy, int ox) {
  const int y = blockIdx.y;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < cols * 3; e += gridDim.x * blockDim.x) {
    const int x = e / 3, c = e - 3 * x;
    dst[size_t(c) * g.plane + size_t(y + oy) * g.pitch + (x + ox)] = src[size_t(y) * src_pitch + e];
  }
}

__global__ void k_planar_to_hwc(const float* __restrict__ src, Geom g, int oy, int ox, int rows, int cols,
                                float* __restrict__ dst, size_t dst_pitch) {
  const int y = blockIdx.y;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < cols * 3; e += gridDim.x * blockDim.x) {
    const int x = e / 3, c = e - 3 * x;
    dst[size_t(y) * dst_pitch + e] = src[size_t(c) * g.plane + size_t(y + oy) * g.pitch + (x + ox)];
  }
}

// [K][K][3] packed <-> [3][K*K] planar (PSF)
__global__ void k_psf_pack(float* __restrict__ hwc, float* __restrict__ planar, int KK2, int to_planar) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * KK2; i += gridDim.x * blockDim.x) {
    const int q = i / 3, c = i - 3 * q;
    if (to_planar) planar[c * KK2 + q] = hwc[i];
    else hwc[i] = planar[c * KK2 + q];
  }
}

}  // namespace rltv
