// Direct K x K stencil kernels: forward blur + residual, adjoint + step statistics, PSF-gradient
// correlation.  FP32 FMA-pipe kernels: register-blocked (each thread owns X=4 adjacent outputs on R rows
// and slides a rolling window of R input rows through registers), inputs staged in shared memory with a
// zero-filled halo, PSF taps broadcast from shared memory as float4.
//
// Reference arithmetic being replaced (all three run as FFTs inside scipy.signal.convolve there):
//   forward  : synth = convolve(u, psf, "valid"); error = synth - image        lib/deconvolution.pyx:477-488
//   adjoint  : gradu = convolve(error, rot180(psf), "full")                    lib/deconvolution.pyx:490-491
//   PSF grad : gradk = convolve(rot180(u), error, "valid")                     lib/deconvolution.pyx:567-571
// In u-geometry (see rltv_common.cuh) the first two are the centred stencil
//   out[Y][X] = sum_{a,b} w[a][b] * in[Y-P+a][X-P+b]      (zero outside the plane)
// with w = rot180(psf) for the forward blur (true convolution) and w = psf for the adjoint.
#pragma once
#include "rltv_common.cuh"

namespace rltv {

// ------------------------------------------------------------------------------------------------
// Tile configuration of the two convolution kernels
// ------------------------------------------------------------------------------------------------
template <int K>
struct ConvCfg {
  static constexpr int X = 4;                       // outputs per thread along x (one float4)
  static constexpr int R = (K <= 17) ? 4 : 2;       // output rows per thread == depth of the rolling window
  static constexpr int WARPS = 8;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int TW = 32 * X;                 // 128 output columns per block
  static constexpr int TH = WARPS * R;              // 32 (or 16) output rows per block
  static constexpr int P = K / 2;
  static constexpr int NV = (K + X - 1 + 3) / 4;    // float4 loads per window row (window = X + K - 1 floats)
  static constexpr int SP = TW - X + 4 * NV;        // shared-memory row pitch (floats), multiple of 4
  static constexpr int SROWS = TH + K - 1;
  static constexpr int KP = (K + 3) & ~3;           // padded tap-row length
  static constexpr int SMEM_FLOATS = SROWS * SP + K * KP;
  static constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * sizeof(float);
};

// tile(r, c) = plane[Y0 - P + r][X0 - P + c], zero outside [0,Hu) x [0,Wu)
template <int ROWS, int SP, int THREADS>
__device__ __forceinline__ void load_tile_zero(float* __restrict__ tile, const float* __restrict__ plane,
                                               int Hu, int Wu, int pitch, int ytop, int xleft) {
  for (int idx = threadIdx.x; idx < ROWS * SP; idx += THREADS) {
    const int r = idx / SP;
    const int c = idx - r * SP;
    const int y = ytop + r, x = xleft + c;
    float v = 0.f;
    if (y >= 0 && y < Hu && x >= 0 && x < Wu) v = __ldg(plane + size_t(y) * pitch + x);
    tile[idx] = v;
  }
}

// One ky step of the register-blocked stencil: pulls input row (ky + R - 1) of the thread's column window
// into the rolling register window, loads tap row ky, and issues R * X * K FMAs.
template <int K, int KK>
__device__ __forceinline__ void stencil_step(const float* __restrict__ base, const float* __restrict__ wS, int ky,
                                             float (&win)[ConvCfg<K>::R][4 * ConvCfg<K>::NV],
                                             float (&acc)[ConvCfg<K>::R][4]) {
  using C = ConvCfg<K>;
  {
    const float4* p = reinterpret_cast<const float4*>(base + (ky + C::R - 1) * C::SP);
#pragma unroll
    for (int v = 0; v < C::NV; ++v) {
      const float4 t = p[v];
      win[(KK + C::R - 1) % C::R][4 * v + 0] = t.x;
      win[(KK + C::R - 1) % C::R][4 * v + 1] = t.y;
      win[(KK + C::R - 1) % C::R][4 * v + 2] = t.z;
      win[(KK + C::R - 1) % C::R][4 * v + 3] = t.w;
    }
  }
  float w[C::KP];
  {
    const float4* p = reinterpret_cast<const float4*>(wS + ky * C::KP);
#pragma unroll
    for (int v = 0; v < C::KP / 4; ++v) {
      const float4 t = p[v];
      w[4 * v + 0] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
    }
  }
#pragma unroll
  for (int j = 0; j < C::R; ++j)
#pragma unroll
    for (int kx = 0; kx < K; ++kx)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(w[kx], win[(KK + j) % C::R][i + kx], acc[j][i]);
}

template <int K, int KK, int N>
struct StencilUnroll {
  __device__ __forceinline__ static void run(const float* base, const float* wS, int kyb,
                                             float (&win)[ConvCfg<K>::R][4 * ConvCfg<K>::NV],
                                             float (&acc)[ConvCfg<K>::R][4]) {
    stencil_step<K, KK>(base, wS, kyb + KK, win, acc);
    StencilUnroll<K, KK + 1, N>::run(base, wS, kyb, win, acc);
  }
};
template <int K, int N>
struct StencilUnroll<K, N, N> {
  __device__ __forceinline__ static void run(const float*, const float*, int,
                                             float (&)[ConvCfg<K>::R][4 * ConvCfg<K>::NV], float (&)[ConvCfg<K>::R][4]) {}
};

// acc[j][i] = sum_{ky,kx} w[ky][kx] * tile[warp*R + j + ky][4*lane + i + kx]
template <int K>
__device__ __forceinline__ void stencil_core(const float* __restrict__ tile, const float* __restrict__ wS,
                                             float (&acc)[ConvCfg<K>::R][4]) {
  using C = ConvCfg<K>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* base = tile + (warp * C::R) * C::SP + lane * 4;
  float win[C::R][4 * C::NV];
#pragma unroll
  for (int j = 0; j < C::R; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  // prologue: input rows 0 .. R-2 into window slots 0 .. R-2
#pragma unroll
  for (int j = 0; j < C::R - 1; ++j) {
    const float4* p = reinterpret_cast<const float4*>(base + j * C::SP);
#pragma unroll
    for (int v = 0; v < C::NV; ++v) {
      const float4 t = p[v];
      win[j][4 * v + 0] = t.x; win[j][4 * v + 1] = t.y; win[j][4 * v + 2] = t.z; win[j][4 * v + 3] = t.w;
    }
  }
  constexpr int KMAIN = (K / C::R) * C::R;
#pragma unroll 1
  for (int kyb = 0; kyb < KMAIN; kyb += C::R) StencilUnroll<K, 0, C::R>::run(base, wS, kyb, win, acc);
  StencilUnroll<K, 0, K - KMAIN>::run(base, wS, KMAIN, win, acc);
}

// ------------------------------------------------------------------------------------------------
// K1 / K4a: err = valid_conv(u, psf) - image       (pyx:477-488 and :557-565)
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(ConvCfg<K>::THREADS, 2)
k_conv_fwd(Geom g, State* __restrict__ st, const float* __restrict__ u, const float* __restrict__ img,
           const float* __restrict__ psf, float* __restrict__ err) {
  using C = ConvCfg<K>;
  if (st->stop) return;
  extern __shared__ float4 smem4[];
  float* tile = reinterpret_cast<float*>(smem4);
  float* wS = tile + C::SROWS * C::SP;
  const int c = blockIdx.z;
  const int X0 = blockIdx.x * C::TW, Y0 = blockIdx.y * C::TH;
  // First kernel of an inner step: reset the step-size reductions the adjoint kernel accumulates into
  // (stream order guarantees the previous update kernel has consumed them).
  if (blockIdx.x == 0 && blockIdx.y == 0 && c == 0 && threadIdx.x < 3) {
    st->max_u[threadIdx.x] = 0u;
    st->max_G[threadIdx.x] = 0u;
  }
  // true convolution == correlation with the 180-degree rotated PSF (pyx:242-252 does this on the CPU)
  for (int i = threadIdx.x; i < K * C::KP; i += C::THREADS) {
    const int ky = i / C::KP, kx = i - ky * C::KP;
    wS[i] = (kx < K) ? __ldg(psf + size_t(c) * K * K + (K - 1 - ky) * K + (K - 1 - kx)) : 0.f;
  }
  const float* up = u + size_t(c) * g.plane;
  load_tile_zero<C::SROWS, C::SP, C::THREADS>(tile, up, g.Hu, g.Wu, g.pitch, Y0 - C::P, X0 - C::P);
  __syncthreads();
  float acc[C::R][4];
  stencil_core<K>(tile, wS, acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int X = X0 + 4 * lane;
  if (X >= g.pitch) return;
  const float* ip = img + size_t(c) * g.plane;
  float* ep = err + size_t(c) * g.plane;
#pragma unroll
  for (int j = 0; j < C::R; ++j) {
    const int Y = Y0 + warp * C::R + j;
    if (Y >= g.Hu) break;
    const bool rowin = (Y >= C::P) && (Y < C::P + g.M);
    const float4 iv = *reinterpret_cast<const float4*>(ip + size_t(Y) * g.pitch + X);
    float4 o;
    o.x = (rowin && X + 0 >= C::P && X + 0 < C::P + g.N) ? acc[j][0] - iv.x : 0.f;
    o.y = (rowin && X + 1 >= C::P && X + 1 < C::P + g.N) ? acc[j][1] - iv.y : 0.f;
    o.z = (rowin && X + 2 >= C::P && X + 2 < C::P + g.N) ? acc[j][2] - iv.z : 0.f;
    o.w = (rowin && X + 3 >= C::P && X + 3 < C::P + g.N) ? acc[j][3] - iv.w : 0.f;
    *reinterpret_cast<float4*>(ep + size_t(Y) * g.pitch + X) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// K2: g = full_conv(err, rot180(psf)); per-channel max(u), max|lambda*g + (u-ut)/2|   (pyx:490-491, :519, :524)
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(ConvCfg<K>::THREADS, 2)
k_conv_adj(Geom g, State* __restrict__ st, const float* __restrict__ err, const float* __restrict__ psf,
           const float* __restrict__ u, const float* __restrict__ ut, float lambd, float* __restrict__ gout) {
  using C = ConvCfg<K>;
  if (st->stop) return;
  extern __shared__ float4 smem4[];
  float* tile = reinterpret_cast<float*>(smem4);
  float* wS = tile + C::SROWS * C::SP;
  __shared__ float red_u[C::WARPS], red_G[C::WARPS];
  const int c = blockIdx.z;
  const int X0 = blockIdx.x * C::TW, Y0 = blockIdx.y * C::TH;
  for (int i = threadIdx.x; i < K * C::KP; i += C::THREADS) {
    const int ky = i / C::KP, kx = i - ky * C::KP;
    wS[i] = (kx < K) ? __ldg(psf + size_t(c) * K * K + ky * K + kx) : 0.f;
  }
  load_tile_zero<C::SROWS, C::SP, C::THREADS>(tile, err + size_t(c) * g.plane, g.Hu, g.Wu, g.pitch, Y0 - C::P, X0 - C::P);
  __syncthreads();
  float acc[C::R][4];
  stencil_core<K>(tile, wS, acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int X = X0 + 4 * lane;
  float mu = -INFINITY, mG = 0.f;
  if (X < g.pitch) {
    const float* up = u + size_t(c) * g.plane;
    const float* utp = ut + size_t(c) * g.plane;
    float* gp = gout + size_t(c) * g.plane;
#pragma unroll
    for (int j = 0; j < C::R; ++j) {
      const int Y = Y0 + warp * C::R + j;
      if (Y >= g.Hu) break;
      const size_t off = size_t(Y) * g.pitch + X;
      const float4 uv = *reinterpret_cast<const float4*>(up + off);
      const float4 tv = *reinterpret_cast<const float4*>(utp + off);
      const float uu[4] = {uv.x, uv.y, uv.z, uv.w};
      const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool in = (X + i) < g.Wu;
        o[i] = in ? acc[j][i] : 0.f;
        if (in) {
          const float G = fmaf(lambd, acc[j][i], 0.5f * (uu[i] - tt[i]));   // pyx:519
          mu = fmaxf(mu, uu[i]);
          mG = fmaxf(mG, fabsf(G));
        }
      }
      *reinterpret_cast<float4*>(gp + off) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  mu = warp_max(mu);
  mG = warp_max(mG);
  if (lane == 0) { red_u[warp] = mu; red_G[warp] = mG; }
  __syncthreads();
  if (warp == 0) {
    mu = lane < C::WARPS ? red_u[lane] : -INFINITY;
    mG = lane < C::WARPS ? red_G[lane] : 0.f;
    mu = warp_max(mu);
    mG = warp_max(mG);
    if (lane == 0) {
      atomicMax(&st->max_u[c], f2ord(mu));
      atomicMax(&st->max_G[c], f2ord(mG));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4b: PSF gradient.  gk'[dy][dx] = sum_{Y,X} err[Y][X] * u[Y-P+dy][X-P+dx];  gradk[q] = gk'[K-1-q].
// Each lane owns a 4-pixel segment of the tile row, each warp a group of DYG displacement rows dy; the
// DYG*K accumulators per lane are reduced across lanes through shared memory, and each block writes one
// deterministic partial per (dy,dx) that k_gradk_reduce / k_psf_update sum in a fixed order.
// ------------------------------------------------------------------------------------------------
template <int K>
struct GradkCfg {
  static constexpr int P = K / 2;
  static constexpr int DYG = (K <= 17) ? 3 : 1;                  // displacement rows per thread
  static constexpr int NG = (K + DYG - 1) / DYG;                 // groups needed to cover all dy
  static constexpr int NGB = (K <= 17) ? NG : 8;                 // groups (warps per row-part) in one block
  static constexpr int DYB = NGB * DYG;                          // dy handled by one block
  static constexpr int NCHUNK = (K + DYB - 1) / DYB;             // blocks along dy
  static constexpr int RP = (K > 17) ? 1 : (NG >= 5 ? 1 : (NG >= 3 ? 2 : (NG == 2 ? 3 : 4)));
  static constexpr int RPP = (DYG == 3) ? 30 : 32;               // tile rows per row-part (multiple of DYG)
  static constexpr int TH = RP * RPP;
  static constexpr int TW = 128;
  static constexpr int NV = (K + 3 + 3) / 4;
  static constexpr int SP = TW - 4 + 4 * NV;
  static constexpr int SROWS = TH + DYB - 1;
  static constexpr int WARPS = NGB * RP;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int RSTRIDE = 36;                             // scratch row stride (floats): conflict-free float4 reads
  static constexpr int TILE_FLOATS = SROWS * SP + TH * TW;
  static constexpr int SCRATCH_FLOATS = WARPS * DYG * K * RSTRIDE;
  static constexpr int SMEM_FLOATS = TILE_FLOATS > SCRATCH_FLOATS ? TILE_FLOATS : SCRATCH_FLOATS;
  static constexpr size_t SMEM_BYTES = size_t(SMEM_FLOATS) * sizeof(float);
};

template <int K, int KK>
__device__ __forceinline__ void gradk_step(const float* __restrict__ ubase, const float* __restrict__ ebase, int y,
                                           float (&win)[GradkCfg<K>::DYG][4 * GradkCfg<K>::NV],
                                           float (&acc)[GradkCfg<K>::DYG][K]) {
  using C = GradkCfg<K>;
  {
    const float4* p = reinterpret_cast<const float4*>(ubase + (y + C::DYG - 1) * C::SP);
#pragma unroll
    for (int v = 0; v < C::NV; ++v) {
      const float4 t = p[v];
      win[(KK + C::DYG - 1) % C::DYG][4 * v + 0] = t.x;
      win[(KK + C::DYG - 1) % C::DYG][4 * v + 1] = t.y;
      win[(KK + C::DYG - 1) % C::DYG][4 * v + 2] = t.z;
      win[(KK + C::DYG - 1) % C::DYG][4 * v + 3] = t.w;
    }
  }
  const float4 e4 = *reinterpret_cast<const float4*>(ebase + y * C::TW);
  const float e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
  for (int d = 0; d < C::DYG; ++d)
#pragma unroll
    for (int dx = 0; dx < K; ++dx)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[d][dx] = fmaf(e[i], win[(KK + d) % C::DYG][i + dx], acc[d][dx]);
}

template <int K, int KK, int N>
struct GradkUnroll {
  __device__ __forceinline__ static void run(const float* ubase, const float* ebase, int yb,
                                             float (&win)[GradkCfg<K>::DYG][4 * GradkCfg<K>::NV],
                                             float (&acc)[GradkCfg<K>::DYG][K]) {
    gradk_step<K, KK>(ubase, ebase, yb + KK, win, acc);
    GradkUnroll<K, KK + 1, N>::run(ubase, ebase, yb, win, acc);
  }
};
template <int K, int N>
struct GradkUnroll<K, N, N> {
  __device__ __forceinline__ static void run(const float*, const float*, int,
                                             float (&)[GradkCfg<K>::DYG][4 * GradkCfg<K>::NV],
                                             float (&)[GradkCfg<K>::DYG][K]) {}
};

// partial layout: [c][tile][dy][dx], tile = blockIdx.y * gridDim.x + blockIdx.x; grid.z = 3 * NCHUNK
template <int K>
__global__ void __launch_bounds__(GradkCfg<K>::THREADS, (GradkCfg<K>::THREADS <= 256) ? 2 : 1)
k_gradk(Geom g, const State* __restrict__ st, const float* __restrict__ err, const float* __restrict__ u,
        float* __restrict__ partial) {
  using C = GradkCfg<K>;
  if (st->stop) return;
  extern __shared__ float4 smem4[];
  float* utile = reinterpret_cast<float*>(smem4);
  float* etile = utile + C::SROWS * C::SP;
  const int c = blockIdx.z / C::NCHUNK;
  const int dyb = (blockIdx.z - c * C::NCHUNK) * C::DYB;
  const int X0 = blockIdx.x * C::TW, Y0 = blockIdx.y * C::TH;
  load_tile_zero<C::SROWS, C::SP, C::THREADS>(utile, u + size_t(c) * g.plane, g.Hu, g.Wu, g.pitch,
                                              Y0 - C::P + dyb, X0 - C::P);
  load_tile_zero<C::TH, C::TW, C::THREADS>(etile, err + size_t(c) * g.plane, g.Hu, g.Wu, g.pitch, Y0, X0);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp % C::NGB, part = warp / C::NGB;
  float acc[C::DYG][K];
#pragma unroll
  for (int d = 0; d < C::DYG; ++d)
#pragma unroll
    for (int dx = 0; dx < K; ++dx) acc[d][dx] = 0.f;
  {
    const float* ubase = utile + (part * C::RPP + grp * C::DYG) * C::SP + 4 * lane;
    const float* ebase = etile + (part * C::RPP) * C::TW + 4 * lane;
    float win[C::DYG][4 * C::NV];
#pragma unroll
    for (int j = 0; j < C::DYG - 1; ++j) {
      const float4* p = reinterpret_cast<const float4*>(ubase + j * C::SP);
#pragma unroll
      for (int v = 0; v < C::NV; ++v) {
        const float4 t = p[v];
        win[j][4 * v + 0] = t.x; win[j][4 * v + 1] = t.y; win[j][4 * v + 2] = t.z; win[j][4 * v + 3] = t.w;
      }
    }
#pragma unroll 1
    for (int yb = 0; yb < C::RPP; yb += C::DYG) GradkUnroll<K, 0, C::DYG>::run(ubase, ebase, yb, win, acc);
  }
  __syncthreads();   // all warps are done with the tiles: reuse the memory as reduction scratch
  float* scratch = reinterpret_cast<float*>(smem4);
#pragma unroll
  for (int d = 0; d < C::DYG; ++d)
#pragma unroll
    for (int dx = 0; dx < K; ++dx) scratch[(warp * C::DYG * K + d * K + dx) * C::RSTRIDE + lane] = acc[d][dx];
  __syncthreads();
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  const int ntiles = gridDim.x * gridDim.y;
  for (int o = threadIdx.x; o < C::DYB * K; o += C::THREADS) {
    const int ld = o / K, dx = o - ld * K;
    const int dy = dyb + ld;
    if (dy >= K) continue;
    const int grp2 = ld / C::DYG, d = ld - grp2 * C::DYG;
    float s = 0.f;
    for (int p = 0; p < C::RP; ++p) {
      const float4* row = reinterpret_cast<const float4*>(scratch + ((p * C::NGB + grp2) * C::DYG * K + d * K + dx) * C::RSTRIDE);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const float4 t = row[v];
        s += (t.x + t.y) + (t.z + t.w);
      }
    }
    partial[((size_t(c) * ntiles + tile) * K + dy) * K + dx] = s;
  }
}

// Stage 1 of the cross-tile reduction: partial2[c][chunk][o] = sum over the chunk's tiles (double).
template <int NCH>
__global__ void k_gradk_reduce(const State* __restrict__ st, const float* __restrict__ partial, int ntiles, int KK2,
                               double* __restrict__ partial2) {
  if (st->stop) return;
  const int c = blockIdx.y, chunk = blockIdx.x;
  const int per = (ntiles + NCH - 1) / NCH;
  const int t0 = chunk * per, t1 = min(ntiles, t0 + per);
  for (int o = threadIdx.x; o < KK2; o += blockDim.x) {
    double s = 0.0;
    for (int t = t0; t < t1; ++t) s += double(partial[(size_t(c) * ntiles + t) * KK2 + o]);
    partial2[(size_t(c) * NCH + chunk) * KK2 + o] = s;
  }
}

}  // namespace rltv
