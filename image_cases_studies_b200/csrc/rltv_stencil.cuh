// Direct K x K stencil kernels: forward blur + residual, adjoint + step statistics, PSF-gradient
// correlation.  FP32 FMA-pipe kernels, sm_100a:
//   * persistent CTAs walk a static tile schedule; the input tile (+ K-1 halo, zero-filled at the frame
//     border by the TMA unit) is double-buffered in shared memory, the next tile's cp.async.bulk.tensor
//     load overlapping the current tile's FMAs; epilogue operands (image / u, ut) are TMA-staged too;
//   * register blocking: each thread owns X=4 adjacent outputs on R rows and slides a rolling window of R
//     input rows through registers (one LDS.128 row per ky, R*4*K FFMA per ky) -- ~96 % of the inner-loop
//     instructions are FFMA; PSF taps are broadcast from shared memory as float4.
//
// Reference arithmetic being replaced (all three run as FFTs inside scipy.signal.convolve there):
//   forward  : synth = convolve(u, psf, "valid"); error = synth - image        lib/deconvolution.pyx:477-488
//   adjoint  : gradu = convolve(error, rot180(psf), "full")                    lib/deconvolution.pyx:490-491
//   PSF grad : gradk = convolve(rot180(u), error, "valid")                     lib/deconvolution.pyx:567-571
// In u-geometry (see rltv_common.cuh) the first two are the centred stencil
//   out[Y][X] = sum_{a,b} w[a][b] * in[Y-P+a][X-P+b]      (zero outside the plane)
// with w = rot180(psf) for the forward blur (true convolution) and w = psf for the adjoint.
#pragma once
#include "rltv_band.cuh"
#include "rltv_common.cuh"
#include "rltv_tma.cuh"

namespace rltv {

constexpr int align128(int bytes) { return (bytes + 127) & ~127; }

// ------------------------------------------------------------------------------------------------
// Tile configuration of the two convolution kernels
// ------------------------------------------------------------------------------------------------
template <int K>
struct ConvCfg {
  static constexpr int X = 4;                       // outputs per thread along x (one float4)
  static constexpr int R = (K <= 17) ? 4 : 2;       // output rows per thread == depth of the rolling window
  // 4 warps x 4 rows: 128x16-output tiles, 4 CTAs per SM.  Small tiles keep the static schedule balanced when a
  // row band is only ~500 rows tall (8 GPUs), and four independent load/compute pipelines per SM overlap better.
  static constexpr int WARPS = (K <= 17) ? 4 : 8;
  static constexpr int CTAS_PER_SM = (K <= 17) ? 4 : 2;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int TW = 32 * X;                 // 128 output columns per tile
  static constexpr int TH = WARPS * R;              // 16 output rows per tile
  static constexpr int P = K / 2;
  // The TMA unit requires the box start to be 16-byte aligned along x (measured: tools/tma_test.cu), so the
  // tile starts P4 = roundup(P, 4) columns left of the outputs and every window index is shifted by DELTA.
  static constexpr int P4 = (P + 3) & ~3;
  static constexpr int DELTA = P4 - P;
  static constexpr int NV = (DELTA + K + X - 1 + 3) / 4;   // float4 loads per window row
  static constexpr int SP = TW - X + 4 * NV;        // shared-memory row pitch (floats) == TMA box width
  static constexpr int SROWS = TH + K - 1;          // TMA box height
  static constexpr int KP = (K + 3) & ~3;           // padded tap-row length
  static constexpr int IN_BYTES = SROWS * SP * 4;
  static constexpr int IN_STRIDE = align128(IN_BYTES);
  static constexpr int EPI_BYTES = TH * TW * 4;     // one epilogue operand tile
  static constexpr int W_BYTES = align128(K * KP * 4);
  // [in0][in1][epi0][epi1 (adjoint only)][weights][3 mbarriers]
  static constexpr int smem_bytes(bool adj) { return 2 * IN_STRIDE + (adj ? 2 : 1) * EPI_BYTES + W_BYTES + 64 + 128; }
};

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
      for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(w[kx], win[(KK + j) % C::R][i + kx + C::DELTA], acc[j][i]);
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
// K1 / K4a (ADJ = false): err = valid_conv(u, psf) - image                          pyx:477-488, :557-565
//     tm_in = u (box SP x SROWS), tm_e0 = image (box TW x TH)
// K2 (ADJ = true): g = full_conv(err, rot180 psf); max(u_c), max|lambda*g + (u-ut)/2|   pyx:490-491, :519, :524
//     tm_in = err, tm_e0 = u, tm_e1 = ut
// Output rows [ybeg, yend) of the local band (whole frame: all rows; row band: the rows this band needs).
// Tiles are numbered x-fastest inside a channel; CTA b processes tiles b, b+grid, b+2*grid, ...
// ------------------------------------------------------------------------------------------------
template <int K, bool ADJ>
__global__ void __launch_bounds__(ConvCfg<K>::THREADS, ConvCfg<K>::CTAS_PER_SM)
k_conv(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_e0,
       const __grid_constant__ CUtensorMap tm_e1, Geom g, State* __restrict__ st, const float* __restrict__ psf,
       float lambd, float* __restrict__ out, int ntx, int nty, int ybeg, int yend, CommPeers cp, int seq,
       unsigned* __restrict__ done_counter) {
  using C = ConvCfg<K>;
  if (st->stop) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
  // NB: stage buffers are addressed as smem + s * stride (never through a pointer array): the compiler must
  // see shared-space provenance to emit LDS.128 instead of generic LD.E (measured: 4x shared wavefronts).
  float* e0 = reinterpret_cast<float*>(smem + 2 * C::IN_STRIDE);
  float* e1 = reinterpret_cast<float*>(smem + 2 * C::IN_STRIDE + C::EPI_BYTES);
  float* wS = reinterpret_cast<float*>(smem + 2 * C::IN_STRIDE + (ADJ ? 2 : 1) * C::EPI_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * C::IN_STRIDE + (ADJ ? 2 : 1) * C::EPI_BYTES + C::W_BYTES);
  // bars[0], bars[1]: input stages; bars[2]: epilogue operands
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_per_c = ntx * nty, ntiles = 3 * tiles_per_c;

  if (!ADJ && blockIdx.x == 0 && tid < 3) {
    // first kernel of an inner step: reset the step-size reductions the adjoint kernel accumulates into
    st->smax[0][tid] = ORD_LOWEST;
    st->smax[0][3 + tid] = 0;
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_e0);
    if (ADJ) tma_prefetch_desc(&tm_e1);
  }
  __syncthreads();

  auto issue_in = [&](int t, int s) {
    const int c = t / tiles_per_c, r = t - c * tiles_per_c;
    const int by = r / ntx, bx = r - by * ntx;
    mbar_arrive_expect_tx(&bars[s], C::IN_BYTES);
    tma_load_3d(smem + s * C::IN_STRIDE, &tm_in, bx * C::TW - C::P4, ybeg + by * C::TH - C::P, c, &bars[s]);
  };
  auto issue_epi = [&](int t) {
    const int c = t / tiles_per_c, r = t - c * tiles_per_c;
    const int by = r / ntx, bx = r - by * ntx;
    mbar_arrive_expect_tx(&bars[2], (ADJ ? 2 : 1) * C::EPI_BYTES);
    tma_load_3d(e0, &tm_e0, bx * C::TW, ybeg + by * C::TH, c, &bars[2]);
    if (ADJ) tma_load_3d(e1, &tm_e1, bx * C::TW, ybeg + by * C::TH, c, &bars[2]);
  };

  int t = blockIdx.x;
  if (tid == 0 && t < ntiles) issue_in(t, 0);
  int cur_c = -1;
  float mu = -INFINITY, mG = 0.f;
  for (int k = 0; t < ntiles; ++k, t += gridDim.x) {
    const int s = k & 1;
    const int c = t / tiles_per_c, r = t - c * tiles_per_c;
    const int by = r / ntx, bx = r - by * ntx;
    if (tid == 0) {
      // stage s^1 and the epilogue buffers were released by the __syncthreads that ended iteration k-1
      const int tn = t + gridDim.x;
      if (tn < ntiles) issue_in(tn, s ^ 1);
      issue_epi(t);
    }
    if (c != cur_c) {
      // forward: true convolution == correlation with the 180-degree rotated PSF (pyx:242-252 on the CPU)
      for (int i = tid; i < K * C::KP; i += C::THREADS) {
        const int ky = i / C::KP, kx = i - ky * C::KP;
        const int src = ADJ ? (ky * K + kx) : ((K - 1 - ky) * K + (K - 1 - kx));
        wS[i] = (kx < K) ? __ldg(psf + size_t(c) * K * K + src) : 0.f;
      }
      cur_c = c;
      __syncthreads();
    }
    mbar_wait(&bars[s], (k >> 1) & 1);
    float acc[C::R][4];
    stencil_core<K>(reinterpret_cast<const float*>(smem + s * C::IN_STRIDE), wS, acc);
    mbar_wait(&bars[2], k & 1);

    const int X = bx * C::TW + 4 * lane;
    if (X < g.pitch) {
      float* op = out + size_t(c) * g.plane;
#pragma unroll
      for (int j = 0; j < C::R; ++j) {
        const int Y = ybeg + by * C::TH + warp * C::R + j;
        if (Y >= yend) break;
        const float4 a = *reinterpret_cast<const float4*>(e0 + (warp * C::R + j) * C::TW + 4 * lane);
        const float av[4] = {a.x, a.y, a.z, a.w};
        float o[4];
        if (!ADJ) {
          const int gy = g.row0 + Y;
          const bool rowin = (gy >= C::P) && (gy < C::P + g.M);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = (rowin && (X + i) >= C::P && (X + i) < C::P + g.N) ? acc[j][i] - av[i] : 0.f;
        } else {
          const float4 b = *reinterpret_cast<const float4*>(e1 + (warp * C::R + j) * C::TW + 4 * lane);
          const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool in = (X + i) < g.Wu;
            o[i] = in ? acc[j][i] : 0.f;
            if (in && Y >= g.own0 && Y < g.own1) {   // statistics over owned rows only (row bands)
              const float G = fmaf(lambd, acc[j][i], 0.5f * (av[i] - bv[i]));   // pyx:519
              mu = fmaxf(mu, av[i]);
              mG = fmaxf(mG, fabsf(G));
            }
          }
        }
        *reinterpret_cast<float4*>(op + size_t(Y) * g.pitch + X) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncthreads();   // everyone is done with stage s, e0, e1 (and wS if the channel changes next)
    if (ADJ) {
      // channel boundary (or last tile): flush the per-channel statistics
      const int tn = t + gridDim.x;
      const int cn = tn < ntiles ? tn / tiles_per_c : -1;
      if (cn != c) {
        __shared__ float red_u[C::WARPS], red_G[C::WARPS];
        const float wu = warp_max(mu), wG = warp_max(mG);
        if (lane == 0) { red_u[warp] = wu; red_G[warp] = wG; }
        __syncthreads();
        if (warp == 0) {
          float a = lane < C::WARPS ? red_u[lane] : -INFINITY;
          float b = lane < C::WARPS ? red_G[lane] : 0.f;
          a = warp_max(a);
          b = warp_max(b);
          if (lane == 0) {
            atomicMax(&st->smax[0][c], f2ord(a));
            atomicMax(&st->smax[0][3 + c], f2ord(b));
          }
        }
        __syncthreads();
        mu = -INFINITY;
        mG = 0.f;
      }
    }
  }
  if (ADJ && cp.nranks > 1) band_step_max_tail(st, cp, seq, done_counter, reinterpret_cast<int*>(smem));
}

// ------------------------------------------------------------------------------------------------
// K4b: PSF gradient.  gk'[dy][dx] = sum_{Y,X} err[Y][X] * u[Y-P+dy][X-P+dx];  gradk[q] = gk'[K-1-q].
// Each lane owns a 4-pixel segment of the tile row, each warp a group of DYG displacement rows dy (and one
// of RP row-parts of the tile).  The DYG*K accumulators per lane PERSIST across all the tiles a CTA walks for
// one (channel, dy-chunk) work item and are reduced once (xor-butterfly over lanes, fixed-order sum over
// row-parts); the last CTA to finish sums the per-CTA partials in CTA order (double): bit-reproducible.
// ------------------------------------------------------------------------------------------------
template <int K>
struct GradkCfg {
  static constexpr int P = K / 2;
  static constexpr int DYG = (K <= 17) ? 3 : 1;                  // displacement rows per thread
  static constexpr int NG = (K + DYG - 1) / DYG;                 // groups needed to cover all dy
  static constexpr int NGB = (K <= 17) ? NG : 8;                 // groups (warps per row-part) in one CTA
  static constexpr int DYB = NGB * DYG;                          // dy handled per work item
  static constexpr int NCHUNK = (K + DYB - 1) / DYB;             // work items per channel
  static constexpr int RP = (K > 17) ? 2 : (NG >= 6 ? 2 : (NG >= 4 ? 3 : (NG == 3 ? 4 : (NG == 2 ? 6 : 12))));
  static constexpr int RPP = (K > 17) ? 32 : (NG >= 4 ? 30 : (NG == 3 ? 24 : (NG == 2 ? 15 : 6)));
  static constexpr int TH = RP * RPP;
  static constexpr int TW = 128;
  static constexpr int P4 = (P + 3) & ~3;                        // 16-byte aligned TMA box start (see ConvCfg)
  static constexpr int DELTA = P4 - P;
  static constexpr int NV = (DELTA + K + 3 + 3) / 4;
  static constexpr int SP = TW - 4 + 4 * NV;
  static constexpr int SROWS = TH + DYB - 1;
  static constexpr int WARPS = NGB * RP;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int U_BYTES = SROWS * SP * 4;
  static constexpr int U_STRIDE = align128(U_BYTES);
  static constexpr int E_BYTES = TH * TW * 4;
  static constexpr int STAGE = U_STRIDE + E_BYTES;               // E_BYTES is a multiple of 128
  static constexpr int RED_BYTES = align128(WARPS * DYG * K * 4);
  static constexpr int SMEM_BYTES = 2 * STAGE + RED_BYTES + 64 + 128;
  static_assert(RPP % DYG == 0, "row-part height must be a multiple of the rolling-window depth");
  static_assert(SMEM_BYTES <= 227 * 1024, "tile does not fit shared memory");
  static_assert(WARPS <= 16, "too many warps");
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
      for (int i = 0; i < 4; ++i) acc[d][dx] = fmaf(e[i], win[(KK + d) % C::DYG][i + dx + C::DELTA], acc[d][dx]);
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

// tm_u: u, box SP x SROWS;  tm_e: err, box TW x TH.  partial layout: [c][cta][dy][dx] (gridDim.x CTAs).
template <int K>
__global__ void __launch_bounds__(GradkCfg<K>::THREADS, 1)
k_gradk(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_e, Geom g,
        const State* __restrict__ st, float* __restrict__ partial, int ntx, int nty, double* __restrict__ gk_sum,
        CommPeers cp, int seq, unsigned* __restrict__ done_counter) {
  using C = GradkCfg<K>;
  if (st->stop) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
  // stage s: u tile at smem + s*STAGE, err tile at smem + s*STAGE + U_STRIDE (pointer arithmetic, not a
  // pointer array, so the loads stay LDS)
  float* red = reinterpret_cast<float*>(smem + 2 * C::STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * C::STAGE + C::RED_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = warp % C::NGB, part = warp / C::NGB;
  const int tiles_per_c = ntx * nty;
  // tiles of one work item walked by this CTA: b, b+grid, ...
  const int my_tiles = (tiles_per_c > int(blockIdx.x)) ? (tiles_per_c - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
  constexpr int NITEMS = 3 * C::NCHUNK;
  const int total = NITEMS * my_tiles;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_u);
    tma_prefetch_desc(&tm_e);
  }
  __syncthreads();

  auto issue = [&](int q, int s) {
    const int item = q / my_tiles, tl = blockIdx.x + (q - item * my_tiles) * gridDim.x;
    const int c = item / C::NCHUNK, dyb = (item - c * C::NCHUNK) * C::DYB;
    const int by = tl / ntx, bx = tl - by * ntx;
    mbar_arrive_expect_tx(&bars[s], C::U_BYTES + C::E_BYTES);
    // tm_e covers the OWNED rows only (row 0 of that map = local row own0): rows outside arrive as zeros,
    // so halo rows never contribute to the PSF gradient
    tma_load_3d(smem + s * C::STAGE, &tm_u, bx * C::TW - C::P4, g.own0 + by * C::TH - C::P + dyb, c, &bars[s]);
    tma_load_3d(smem + s * C::STAGE + C::U_STRIDE, &tm_e, bx * C::TW, by * C::TH, c, &bars[s]);
  };

  float acc[C::DYG][K];
#pragma unroll
  for (int d = 0; d < C::DYG; ++d)
#pragma unroll
    for (int dx = 0; dx < K; ++dx) acc[d][dx] = 0.f;

  if (tid == 0 && total > 0) issue(0, 0);
  for (int q = 0; q < total; ++q) {
    const int s = q & 1;
    if (tid == 0 && q + 1 < total) issue(q + 1, s ^ 1);   // stage s^1 was released by the barrier ending iteration q-1
    mbar_wait(&bars[s], (q >> 1) & 1);
    {
      const float* ubase = reinterpret_cast<const float*>(smem + s * C::STAGE) + (part * C::RPP + grp * C::DYG) * C::SP + 4 * lane;
      const float* ebase = reinterpret_cast<const float*>(smem + s * C::STAGE + C::U_STRIDE) + (part * C::RPP) * C::TW + 4 * lane;
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
    const int item = q / my_tiles;
    const bool last_of_item = (q + 1 == total) || ((q + 1) / my_tiles != item);
    if (last_of_item) {
      // reduce over lanes (fixed butterfly), then over row-parts in order, and write this CTA's partial
#pragma unroll
      for (int d = 0; d < C::DYG; ++d)
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {
          float v = acc[d][dx];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) red[(warp * C::DYG + d) * K + dx] = v;
          acc[d][dx] = 0.f;
        }
      __syncthreads();
      const int c = item / C::NCHUNK, dyb = (item - c * C::NCHUNK) * C::DYB;
      for (int o = tid; o < C::DYB * K; o += C::THREADS) {
        const int ld = o / K, dx = o - ld * K;
        const int dy = dyb + ld;
        if (dy >= K) continue;
        const int g2 = ld / C::DYG, d = ld - g2 * C::DYG;
        float sum = 0.f;
        for (int p = 0; p < C::RP; ++p) sum += red[((p * C::NGB + g2) * C::DYG + d) * K + dx];
        partial[((size_t(c) * gridDim.x + blockIdx.x) * K + dy) * K + dx] = sum;
      }
    }
    __syncthreads();   // stage s (and `red`) free for reuse
  }
  if (total == 0) {
    // CTA without tiles: its partial slots must still be defined (zero)
    for (int o = tid; o < 3 * K * K; o += C::THREADS) {
      const int c = o / (K * K), r = o - c * K * K;
      partial[(size_t(c) * gridDim.x + blockIdx.x) * K * K + r] = 0.f;
    }
  }
  // The last CTA to finish sums the per-CTA partials in CTA order (double: deterministic whichever CTA is
  // last) into gk_sum and, with row bands, publishes the band's sums to every band.
  __threadfence();
  __syncthreads();
  __shared__ bool last;
  if (tid == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int par = seq & 1;
  for (int o = tid; o < 3 * K * K; o += C::THREADS) {
    const int c = o / (K * K), r = o - c * K * K;
    // 8 independent accumulators (CTA b goes to accumulator b % 8): 8+ loads in flight per thread, fixed tree
    double a8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float* src = partial + size_t(c) * gridDim.x * K * K + r;
    unsigned b = 0;
    for (; b + 8 <= gridDim.x; b += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldcg(src + size_t(b + j) * K * K);
#pragma unroll
      for (int j = 0; j < 8; ++j) a8[j] += double(v[j]);
    }
    for (; b < gridDim.x; ++b) a8[b & 7] += double(__ldcg(src + size_t(b) * K * K));
    const double sum = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
    gk_sum[o] = sum;
    if (cp.nranks > 1)
      for (int rr = 0; rr < cp.nranks; ++rr) cp.peer[rr]->gk_val[par][cp.rank][o] = sum;
  }
  if (cp.nranks > 1) __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    *done_counter = 0u;
    if (cp.nranks > 1) {
      for (int rr = 0; rr < cp.nranks; ++rr) *reinterpret_cast<volatile int*>(&cp.peer[rr]->gk_flag[par][cp.rank]) = seq;
      __threadfence_system();
    }
  }
}

}  // namespace rltv
