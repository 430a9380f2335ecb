// Row-FFT hybrid stencils for 9 <= K <= 47 (forward blur, adjoint, PSF gradient; fused residual up to K = 17): the FLOP-reducing form of the forward blur and the adjoint
// (lib/deconvolution.pyx:477-491, :557-565), which the reference itself evaluates as FFT convolutions (scipy).
//
// A K x K stencil  out[Y][X] = sum_{ky,kx} w[ky][kx] in[Y-P+ky][X-P+kx]  costs K*K FMAs per output directly.  Here the
// x direction is done in the frequency domain on 128-sample row segments (112 valid outputs each, overlap-save) and
// the y direction stays a direct sum -- but of SPECTRA:   Out^[Y] = sum_ky Wc[ky] (.) In^[Y-P+ky]   -> K complex MACs
// per (row, bin) = 4K FMAs per 112/128 outputs instead of K*K: 34 instead of 225 FMAs per output at K = 15, plus two
// 128-point FFTs per row.  Because the taps are real, two image rows ride through the complex FFT as real and
// imaginary parts (rows Y and Y+24 of a 48-row tile) and come out as the real and imaginary parts of the inverse
// transform: no real-FFT untangling anywhere.
//
// Per 48 x 112 output tile (CTAs of 384 threads, two per SM so that the FFT phases of one overlap the MAC of the
// other; persistent; one TMA stage that is re-armed as soon as the forward FFTs have consumed it):
//   1. forward FFT of the 24+K-1 packed rows straight out of the TMA buffer          (csrc/rltv_fft.cuh)
//   2. vertical MAC over spectra, in place: thread = (bin, chunk of 8 output rows), tap spectra in shared memory
//   3. inverse FFT of the 24 packed output rows
//   4. epilogue as in k_conv (residual / step statistics), 112 columns x 48 rows, operands prefetched to registers
#pragma once
#include "../../include/rltv_b200.h"
#include "rltv_band.cuh"
#include "rltv_common.cuh"
#include "rltv_elementwise.cuh"
#include "rltv_fft.cuh"
#include "rltv_stencil.cuh"
#include "rltv_tma.cuh"

namespace rltv {

template <int K>
struct FftCfg {
  static_assert(K >= 9 && K <= RLTV_MAX_MK, "128-sample segments: 112 outputs + halo up to K = 17, 96 up to 33, fewer above");
  static constexpr int P = K / 2;
  static constexpr int P4 = (P + 3) & ~3;          // 16-byte aligned TMA box start
  // valid output columns per 128-sample segment: what the circular convolution leaves, in multiples of 8
  static constexpr int TWO_MAX = ((FFT_N - P4 - P) / 8) * 8;
  static constexpr int TWO = (K <= 17) ? 112 : (TWO_MAX < 96 ? TWO_MAX : 96);
  static constexpr int CTAS_PER_SM = (K <= 17) ? 2 : 1;
  static constexpr int HB = 24;                    // rows per packed block (real part: rows 0..23, imaginary: 24..47)
  static constexpr int TROWS = 2 * HB;
  static constexpr int IN_ROWS = TROWS + K - 1;    // TMA box height (real rows)
  static constexpr int ZROWS = HB + K - 1;         // packed complex rows
  static constexpr int THREADS = 384;              // 12 warps; two CTAs per SM run their phases out of step
  static constexpr bool TWB = K >= 19;             // batched twiddle fetch (30 registers): only where one CTA per SM leaves room
  static constexpr int NCHUNK = THREADS / FFT_N;   // MAC: 128 bins x 3 row chunks
  static constexpr int CHUNK = HB / NCHUNK;        // output rows per MAC thread
  // epilogue: ER row slots x TWO/2 column pairs; ER = the largest divisor of HB the thread count covers
  static constexpr int ER_MAX = THREADS / (TWO / 2);
  static constexpr int ER = ER_MAX >= 12 ? 12 : (ER_MAX >= 8 ? 8 : 6);
  static constexpr int NTASK = HB / ER;
  // TMA box = INW x IN_ROWS floats.  INW = 136 (not 128): consecutive rows then start 8 banks apart, so the four rows
  // a warp transforms read the dense TMA buffer conflict-free with compile-time offsets (no per-row fetch rotation).
  static constexpr int INW = 136;
  static constexpr int IN_BYTES = IN_ROWS * INW * 4;                          // TMA transaction size
  static constexpr int IN_STRIDE = (IN_BYTES + 127) & ~127;
  static constexpr int ZB_BYTES = ZROWS * FFT_PITCH * 8;
  static constexpr int WS_BYTES = K * FFT_N * 8;                              // tap spectra of the current channel
  static_assert(NCHUNK * CHUNK == HB && ER * NTASK == HB, "MAC chunks / epilogue row slots cover the packed rows");
  static_assert(ZB_BYTES % 128 == 0 && WS_BYTES % 128 == 0, "128-byte aligned sections");
  static constexpr int SMEM_BYTES = IN_STRIDE + ZB_BYTES + WS_BYTES + FFT_TW_BYTES + 64 + 128;
  static_assert(P4 + TWO - 1 + P <= FFT_N - 1, "segment too short for this K");
  static_assert(CTAS_PER_SM * SMEM_BYTES <= 227 * 1024, "shared memory");
};

// Tap spectra: wspec[dir][c][ky][k] = (1/128) sum_{j=-P..P} w[ky][j+P] exp(+2 pi i j k / 128), w = rot180(psf) for
// dir 0 (forward blur = true convolution) and w = psf for dir 1 (adjoint).  grid (K, 3, 2), 128 threads.
__global__ void __launch_bounds__(128)
k_psf_spectrum(const State* __restrict__ st, const float* __restrict__ psf, int K, float2* __restrict__ wspec) {
  if (st->stop) return;
  const int ky = blockIdx.x, c = blockIdx.y, dir = blockIdx.z, k = threadIdx.x, P = K / 2;
  float re = 0.f, im = 0.f;
  for (int kx = 0; kx < K; ++kx) {
    const int src = dir ? (ky * K + kx) : ((K - 1 - ky) * K + (K - 1 - kx));
    const float w = psf[size_t(c) * K * K + src];
    float s, co;
    sincospif(2.0f * float(((kx - P) * k) & (FFT_N - 1)) / float(FFT_N), &s, &co);
    re = fmaf(w, co, re);
    im = fmaf(w, s, im);
  }
  wspec[((size_t(dir) * 3 + c) * K + ky) * FFT_N + k] = make_float2(re * (1.f / FFT_N), im * (1.f / FFT_N));
}

// Per-phase cycle counters of k_conv_fft (thread 0 of every CTA; read by rltv_debug_phase_cycles): 0 wait for the
// TMA tile, 1 forward FFT, 2 MAC (+ operand prefetch), 3 inverse FFT, 4 epilogue, 5 statistics flush.  Only filled
// by a -DRLTV_PHASE_PROBE build; the clock reads are not ordered against the barriers, so the split is approximate.
__device__ unsigned long long g_fft_phase_cycles[8];

template <int K, bool ADJ>
__global__ void __launch_bounds__(FftCfg<K>::THREADS, FftCfg<K>::CTAS_PER_SM)
k_conv_fft(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ e0g, const float* __restrict__ e1g,
           Geom g, State* __restrict__ st, const float2* __restrict__ wspec,
           float lambd, float* __restrict__ out, int ntx, int nty, int ybeg, int yend, CommPeers cp, int seq,
           unsigned* __restrict__ done_counter) {
  using C = FftCfg<K>;
  if (st->stop) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const float* inR = reinterpret_cast<const float*>(smem);                       // real rows, TMA destination
  float2* ZB = reinterpret_cast<float2*>(smem + C::IN_STRIDE);                   // packed spectra, then outputs
  float2* WS = reinterpret_cast<float2*>(smem + C::IN_STRIDE + C::ZB_BYTES);     // tap spectra [K][128]
  float2* tw = reinterpret_cast<float2*>(smem + C::IN_STRIDE + C::ZB_BYTES + C::WS_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(tw) + FFT_TW_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_per_c = ntx * nty, ntiles = 3 * tiles_per_c;

  if (!ADJ && blockIdx.x == 0 && tid < 3) {
    st->smax[0][tid] = ORD_LOWEST;
    st->smax[0][3 + tid] = 0;
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
  }
  fft_fill_twiddles(tw);
  __syncthreads();

  auto issue_in = [&](int t) {
    const int c = t / tiles_per_c, r = t - c * tiles_per_c;
    const int by = r / ntx, bx = r - by * ntx;
    mbar_arrive_expect_tx(bar, C::IN_BYTES);
    tma_load_3d(smem, &tm_in, bx * C::TWO - C::P4, ybeg + by * C::TROWS - C::P, c, bar);
  };
  // MAC role of this thread: frequency bin and chunk of output rows; epilogue role: row slot and column pair
  const int bin = tid & (FFT_N - 1), chunk = tid >> 7;
  const int er = tid / (C::TWO / 2), ep = tid - er * (C::TWO / 2);
  const bool epi_active = er < C::ER;

  int t = blockIdx.x;
  if (tid == 0 && t < ntiles) issue_in(t);
  int cur_c = -1;
  float mu = -INFINITY, mG = 0.f;
  // phase probe (tools/phase_probe.py): compiled in only with -DRLTV_PHASE_PROBE -- its six 64-bit counters cost the
  // production kernel 88 bytes of register spills at the 80-register cap of two 384-thread CTAs per SM
#ifdef RLTV_PHASE_PROBE
  long long ph[6] = {0, 0, 0, 0, 0, 0}, tprev = clock64();
  auto mark = [&](int i) { if (tid == 0) { const long long now = clock64(); ph[i] += now - tprev; tprev = now; } };
#else
  auto mark = [](int) {};
#endif
  for (int k = 0; t < ntiles; ++k, t += gridDim.x) {
    const int c = t / tiles_per_c, r = t - c * tiles_per_c;
    const int by = r / ntx, bx = r - by * ntx;
    if (c != cur_c) {
      // tap spectra of this channel into shared memory (the previous tile's last barrier released WS)
      const float2* wsrc = wspec + (size_t(ADJ ? 1 : 0) * 3 + c) * K * FFT_N;
      for (int i = tid; i < K * FFT_N; i += C::THREADS) WS[i] = __ldg(wsrc + i);
      cur_c = c;
    }
    mbar_wait(bar, k & 1);
    mark(0);

    __syncwarp();
    // 1. forward FFT of the packed rows: Z[r] = in[r] + i in[r + HB], read straight from the TMA buffer
    for (int task = tid; task < C::ZROWS * 8; task += C::THREADS) {
      const unsigned mask = __activemask();
      const int zr = task >> 3, tt = task & 7;
      const float* ra = inR + zr * C::INW + tt;
      fft128_core<false, C::TWB>(ZB + zr * FFT_PITCH, tw, tt, [&](int j) { return make_float2(ra[8 * j], ra[C::HB * C::INW + 8 * j]); }, mask, 0);
    }
    __syncthreads();
    mark(1);
    if (tid == 0) {                         // the real-row stage is consumed: its reload overlaps everything below
      const int tn = t + gridDim.x;
      if (tn < ntiles) issue_in(tn);
    }

    // 2. vertical MAC over spectra, in place: O[y][bin] = sum_ky Wc[ky][bin] * Z[y + ky][bin] -> ZB row y
    {
      float2 z[C::CHUNK + K - 1];
#pragma unroll
      for (int i = 0; i < C::CHUNK + K - 1; ++i) z[i] = ZB[(chunk * C::CHUNK + i) * FFT_PITCH + bin];
      float2 acc[C::CHUNK];
#pragma unroll
      for (int y = 0; y < C::CHUNK; ++y) acc[y] = make_float2(0.f, 0.f);
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const float2 w = WS[ky * FFT_N + bin];
#pragma unroll
        for (int y = 0; y < C::CHUNK; ++y) acc[y] = cfma(z[y + ky], w, acc[y]);
      }
      __syncthreads();                      // every chunk has read its window (chunks overlap by K-1 rows)
#pragma unroll
      for (int y = 0; y < C::CHUNK; ++y) ZB[(chunk * C::CHUNK + y) * FFT_PITCH + bin] = acc[y];
    }
    __syncthreads();

    // Epilogue operands of this thread's outputs go to registers now, so their DRAM latency hides behind the
    // inverse FFT.  Epilogue thread = (row slot er, column pair ep): outputs (er + ER*q [+ HB], 2ep..2ep+1).
    const int X = bx * C::TWO + 2 * ep;
    const int Ybase = ybeg + by * C::TROWS + er;
    const bool xok = epi_active && X < g.pitch;
    const size_t off0 = size_t(c) * g.plane + size_t(Ybase) * g.pitch + X;
    float2 pa[C::NTASK][2], pb[ADJ ? C::NTASK : 1][2];
#pragma unroll
    for (int q = 0; q < C::NTASK; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool ok = xok && (Ybase + C::ER * q + h * C::HB) < yend;
        const size_t goff = ok ? off0 + size_t(C::ER * q + h * C::HB) * g.pitch : 0;
        pa[q][h] = ok ? __ldg(reinterpret_cast<const float2*>(e0g + goff)) : make_float2(0.f, 0.f);
        if (ADJ) pb[q][h] = ok ? __ldg(reinterpret_cast<const float2*>(e1g + goff)) : make_float2(0.f, 0.f);
      }
    mark(2);

    // 3. inverse FFT of the packed output rows, in place
    for (int task = tid; task < C::HB * 8; task += C::THREADS) {
      const unsigned mask = __activemask();
      const int zr = task >> 3, tt = task & 7;
      float2* row = ZB + zr * FFT_PITCH;
      fft128_core<true, C::TWB>(row, tw, tt, [&](int j) { return row[tt + 8 * j]; }, mask, 0);
    }
    __syncthreads();
    mark(3);

    // 4. epilogue: real part -> rows 0..HB-1 of the tile, imaginary part -> rows HB..2HB-1; columns P4 .. P4+111 valid
    {
      // tile-uniform fast path: every output of the tile is inside the image (forward) / owned and inside u (adjoint)
      const int Yt = ybeg + by * C::TROWS, Xt = bx * C::TWO;
      const bool interior = ADJ ? (Xt + C::TWO <= g.Wu && Yt >= g.own0 && Yt + C::TROWS <= g.own1 && Yt + C::TROWS <= yend)
                                : (Xt >= C::P && Xt + C::TWO <= C::P + g.N && g.row0 + Yt >= C::P &&
                                   g.row0 + Yt + C::TROWS <= C::P + g.M && Yt + C::TROWS <= yend);
      float* op = out + off0;
      if (interior) {
        if (epi_active) {
#pragma unroll
          for (int q = 0; q < C::NTASK; ++q) {
            const float4 zz = *reinterpret_cast<const float4*>(ZB + (er + C::ER * q) * FFT_PITCH + C::P4 + 2 * ep);
            const float v[2][2] = {{zz.x, zz.z}, {zz.y, zz.w}};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float2 o;
              if (!ADJ) {
                o = make_float2(v[h][0] - pa[q][h].x, v[h][1] - pa[q][h].y);
              } else {
                o = make_float2(v[h][0], v[h][1]);
                const float G0 = fmaf(lambd, o.x, 0.5f * (pa[q][h].x - pb[q][h].x));   // pyx:519
                const float G1 = fmaf(lambd, o.y, 0.5f * (pa[q][h].y - pb[q][h].y));
                mu = fmaxf(mu, fmaxf(pa[q][h].x, pa[q][h].y));
                mG = fmaxf(mG, fmaxf(fabsf(G0), fabsf(G1)));
              }
              *reinterpret_cast<float2*>(op + size_t(C::ER * q + h * C::HB) * g.pitch) = o;
            }
          }
        }
      } else if (xok) {
        const bool cin[2] = {ADJ ? (X < g.Wu) : (X >= C::P && X < C::P + g.N),
                             ADJ ? (X + 1 < g.Wu) : (X + 1 >= C::P && X + 1 < C::P + g.N)};
#pragma unroll
        for (int q = 0; q < C::NTASK; ++q) {
          const float4 zz = *reinterpret_cast<const float4*>(ZB + (er + C::ER * q) * FFT_PITCH + C::P4 + 2 * ep);   // 2 complex
          const float v[2][2] = {{zz.x, zz.z}, {zz.y, zz.w}};                                                       // [h][column]
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int Y = Ybase + C::ER * q + h * C::HB;
            if (Y >= yend) continue;
            const float av[2] = {pa[q][h].x, pa[q][h].y};
            float o[2];
            if (!ADJ) {
              const int gy = g.row0 + Y;
              const bool rowin = (gy >= C::P) && (gy < C::P + g.M);
#pragma unroll
              for (int j = 0; j < 2; ++j) o[j] = (rowin && cin[j]) ? v[h][j] - av[j] : 0.f;
            } else {
              const float bv[2] = {pb[q][h].x, pb[q][h].y};
              const bool owned = Y >= g.own0 && Y < g.own1;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                o[j] = cin[j] ? v[h][j] : 0.f;
                if (cin[j] && owned) {
                  const float G = fmaf(lambd, v[h][j], 0.5f * (av[j] - bv[j]));   // pyx:519
                  mu = fmaxf(mu, av[j]);
                  mG = fmaxf(mG, fabsf(G));
                }
              }
            }
            *reinterpret_cast<float2*>(op + size_t(C::ER * q + h * C::HB) * g.pitch) = make_float2(o[0], o[1]);
          }
        }
      }
    }
    __syncthreads();   // ZB and WS are free again
    mark(4);
    if (ADJ) {
      const int tn = t + gridDim.x;
      const int cn = tn < ntiles ? tn / tiles_per_c : -1;
      if (cn != c) {
        __shared__ float red_u[16], red_G[16];
        const float wu = warp_max(mu), wG = warp_max(mG);
        if (lane == 0) { red_u[warp] = wu; red_G[warp] = wG; }
        __syncthreads();
        if (warp == 0) {
          float a = lane < C::THREADS / 32 ? red_u[lane] : -INFINITY;
          float b = lane < C::THREADS / 32 ? red_G[lane] : 0.f;
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
#ifdef RLTV_PHASE_PROBE
  if (tid == 0)
    for (int i = 0; i < 6; ++i) atomicAdd(&g_fft_phase_cycles[i], (unsigned long long)ph[i]);
#endif
  if (ADJ && cp.nranks > 1) band_step_max_tail(st, cp, seq, done_counter, reinterpret_cast<int*>(smem));
}

}  // namespace rltv

namespace rltv {

// ------------------------------------------------------------------------------------------------
// PSF gradient in the row-frequency domain (lib/deconvolution.pyx:567-571).
//   gk'[dy][dx] = sum_{Y,X} err[Y][X] u[Y-P+dy][X-P+dx]   is, along x, a cross-correlation:  for one row pair
//   R[s] = sum_n e[n] u[n+s]  <=>  R^ = conj(E^) U^ .  The K displacement rows dy stay a direct loop, the K lags
//   dx = s + P come out of ONE inverse transform at the very end, because the products are ACCUMULATED in the
//   frequency domain over all rows and all tiles:   C[dy][k] += conj(Ze[y][k]) Zu[y+dy][k].
//   Rows are packed in pairs like in k_conv_fft (Z = row y + i row y+HB); the packed product equals A + iB with
//   A = conj(E)U + conj(E')U' (wanted) and B a cross term; both are spectra of real sequences, hence
//   A[k] = (C[k] + conj(C[-k])) / 2 -- untangled once, by k_gradk_fft_finish, before a K-lag inverse DFT in double.
// Per 80 x 112 tile (K <= 17; 64 x 96 above): forward row FFTs + HB*128*K complex MACs; no inverse FFT, no epilogue.
// One real-data stage only: the next tile's TMA is issued as soon as the forward FFTs have consumed the stage and
// overlaps the MAC phase.
// ------------------------------------------------------------------------------------------------
template <int K>
struct GradkFftCfg {
  static_assert(K >= 9 && K <= RLTV_MAX_MK, "see FftCfg");
  static constexpr int P = K / 2;
  static constexpr int P4 = (P + 3) & ~3;
  static constexpr int TWO = FftCfg<K>::TWO;
  static constexpr int HB = (K <= 17) ? 40 : (K <= 31 ? 32 : 24);   // shared memory: (HB + K - 1) spectra rows of u
#ifndef RLTV_GRADK_THREADS
#define RLTV_GRADK_THREADS 512   // (640 threads, five row chunks: 0.58 vs 0.54 ms at 24 MP)
#endif
  static constexpr int THREADS = (K <= 17) ? RLTV_GRADK_THREADS : 256;   // large K: 2K complex accumulators + a K-deep window per thread
  static constexpr int TROWS = 2 * HB;
  static constexpr int U_ROWS = TROWS + K - 1;      // real u rows per tile
  static constexpr int ZU_ROWS = HB + K - 1;
  static constexpr int NCH = THREADS / FFT_N;
  static constexpr int CHUNK = HB / NCH;
  static constexpr int UW = 136;                    // TMA box width of the u rows (floats): rows start 8 banks apart, so the four
                                                    // rows a warp transforms read the dense buffer conflict-free at compile-time offsets
  static constexpr int U_TX = U_ROWS * UW * 4;      // bytes of the TMA transaction
  static constexpr int U_BYTES = (U_TX + 127) & ~127;
  static constexpr int E_BYTES = TROWS * FFT_N * 4;
  static constexpr int ZU_BYTES = ZU_ROWS * FFT_PITCH * 8;
  static constexpr int ZE_BYTES = HB * FFT_PITCH * 8;
  static constexpr int SMEM_BYTES = U_BYTES + E_BYTES + ZU_BYTES + ZE_BYTES + FFT_TW_BYTES + 64 + 128;
  // fused variant (residual computed in the kernel): + the forward tap spectra of the current channel
  static constexpr int WS_BYTES = K * FFT_N * 8;
  static constexpr int SMEM_BYTES_FUSED = SMEM_BYTES + WS_BYTES;
  static constexpr bool CAN_FUSE = (SMEM_BYTES_FUSED <= 227 * 1024) && (CHUNK % 2 == 0);
  static constexpr bool TWB = true;                 // batched twiddle fetch in the FFT engine (30 registers; 0.575 -> 0.545 ms at K = 15)
  static_assert(NCH * CHUNK == HB, "row chunks cover the packed rows");
  static_assert(NCH * K * FFT_N * 8 <= ZU_BYTES + ZE_BYTES, "chunk-reduction scratch fits the spectra buffers");
  static_assert(P4 + TWO - 1 + P <= FFT_N - 1, "segment too short for this K");
  static_assert(U_BYTES % 128 == 0 && E_BYTES % 128 == 0 && ZU_BYTES % 128 == 0 && ZE_BYTES % 128 == 0, "alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

// part: [cta][c][dy][k] float2 per-CTA frequency-domain sums; k_gradk_fft_finish does the rest.
//
// FUSED = true computes the residual of pyx:557-565 in the same kernel instead of reading it: the u rows it needs
// are exactly the rows whose spectra this kernel already holds, so the separate forward-blur launch (its TMA loads,
// its forward FFTs, its residual write and this kernel's read of it) disappears.  tm_e then maps the IMAGE, and per
// tile   Ze <- FFT( mask( IFFT( sum_ky W[ky] Zu[y+ky] ) - image ) )   (vertical MAC, inverse FFT, residual, forward FFT:
// the same steps as k_conv_fft<K,false>, whose results it reproduces), optionally stored to err_out for the whiteness
// statistic (pyx:623-638 reads the residual of the last inner step).
template <int K, bool FUSED>
__global__ void __launch_bounds__(GradkFftCfg<K>::THREADS, 1)
k_gradk_fft(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_e, Geom g,
            const State* __restrict__ st, const float2* __restrict__ wspec, float* __restrict__ err_out,
            float2* __restrict__ part, int ntx, int nty) {
  using C = GradkFftCfg<K>;
  static_assert(!FUSED || C::CAN_FUSE, "fused residual does not fit for this K");
  if (st->stop) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const float* uR = reinterpret_cast<const float*>(smem);
  const float* eR = reinterpret_cast<const float*>(smem + C::U_BYTES);
  float2* ZU = reinterpret_cast<float2*>(smem + C::U_BYTES + C::E_BYTES);
  float2* ZE = reinterpret_cast<float2*>(smem + C::U_BYTES + C::E_BYTES + C::ZU_BYTES);
  float2* tw = reinterpret_cast<float2*>(smem + C::U_BYTES + C::E_BYTES + C::ZU_BYTES + C::ZE_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(tw) + FFT_TW_BYTES);
  float2* WS = reinterpret_cast<float2*>(smem + C::SMEM_BYTES - 128);   // FUSED only: forward tap spectra [K][128]
  const int tid = threadIdx.x;
  const int tiles_per_c = ntx * nty;
  // tiles of all three channels form one list dealt round-robin to the CTAs (no per-channel rounding up)
  const int all_tiles = 3 * tiles_per_c;
  const int total = (all_tiles > int(blockIdx.x)) ? (all_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_u);
    tma_prefetch_desc(&tm_e);
  }
  fft_fill_twiddles(tw);
  __syncthreads();

  auto issue = [&](int q) {
    const int f = blockIdx.x + q * gridDim.x;
    const int c = f / tiles_per_c, tl = f - c * tiles_per_c;
    const int by = tl / ntx, bx = tl - by * ntx;
    mbar_arrive_expect_tx(bar, C::U_TX + C::E_BYTES);
    tma_load_3d(smem, &tm_u, bx * C::TWO - C::P4, g.own0 + by * C::TROWS - C::P, c, bar);
    tma_load_3d(smem + C::U_BYTES, &tm_e, bx * C::TWO - C::P4, by * C::TROWS, c, bar);   // tm_e: owned rows only
  };
  // The single real-data stage can only be re-armed after the forward FFTs, which leaves the load just the MAC phase
  // to land.  An L2 prefetch of the same boxes one phase earlier takes the DRAM latency off it (measured: 0.381 ->
  // 0.347 ms at 24 MP).  The same trick does nothing for k_conv_fft, whose load has three phases to land.
  auto prefetch = [&](int q) {
    const int f = blockIdx.x + q * gridDim.x;
    const int c = f / tiles_per_c, tl = f - c * tiles_per_c;
    const int by = tl / ntx, bx = tl - by * ntx;
    tma_prefetch_3d(&tm_u, bx * C::TWO - C::P4, g.own0 + by * C::TROWS - C::P, c);
    tma_prefetch_3d(&tm_e, bx * C::TWO - C::P4, by * C::TROWS, c);
  };

  const int bin = tid & (FFT_N - 1), chunk = tid >> 7;
  float2 acc[K];
#pragma unroll
  for (int d = 0; d < K; ++d) acc[d] = make_float2(0.f, 0.f);

  unsigned flushed = 0u;                              // channels this CTA wrote a partial for (CTA-uniform)
  int cur_c = -1;
  const int nown = g.own1 - g.own0;
  if (tid == 0 && total > 0) issue(0);
  for (int q = 0; q < total; ++q) {
    if (tid == 0 && q + 1 < total) prefetch(q + 1);
    if constexpr (FUSED) {
      const int c = (blockIdx.x + q * gridDim.x) / tiles_per_c;
      if (c != cur_c) {                               // forward tap spectra of this channel (WS is idle between tiles)
        const float2* wsrc = wspec + size_t(c) * K * FFT_N;
        for (int i = tid; i < K * FFT_N; i += C::THREADS) WS[i] = __ldg(wsrc + i);
        cur_c = c;
      }
    }
    mbar_wait(bar, q & 1);
    __syncwarp();
    if constexpr (!FUSED) {
      // forward FFTs: packed u rows, then packed err rows (columns outside the 112 valid ones are zeroed)
      for (int task = tid; task < (C::ZU_ROWS + C::HB) * 8; task += C::THREADS) {
        const unsigned mask = __activemask();
        const int zr = task >> 3, tt = task & 7;
        if (zr < C::ZU_ROWS) {
          const float* ra = uR + zr * C::UW + tt;
          const float* rb = ra + C::HB * C::UW;
          fft128_core<false, C::TWB>(ZU + zr * FFT_PITCH, tw, tt, [&](int j) { return make_float2(ra[8 * j], rb[8 * j]); }, mask, 0);
        } else {
          const int er = zr - C::ZU_ROWS;
          const float* ra = eR + er * FFT_N;
          const float* rb = eR + (er + C::HB) * FFT_N;
          fft128_row<false, C::TWB>(ZE + er * FFT_PITCH, tw, tt, [&](int n) {
            const bool in = (n >= C::P4) && (n < C::P4 + C::TWO);
            return in ? make_float2(ra[n], rb[n]) : make_float2(0.f, 0.f);
          }, mask, zr & 3);
        }
      }
      __syncthreads();
      if (tid == 0 && q + 1 < total) issue(q + 1);    // the real-data stage is free: overlap the load with the MAC
    } else {
      const int f = blockIdx.x + q * gridDim.x;
      const int c = f / tiles_per_c, tl = f - c * tiles_per_c;
      const int by = tl / ntx, bx = tl - by * ntx;
      // a. forward FFT of the packed u rows
      for (int task = tid; task < C::ZU_ROWS * 8; task += C::THREADS) {
        const unsigned mask = __activemask();
        const int zr = task >> 3, tt = task & 7;
        const float* ra = uR + zr * C::UW + tt;
        const float* rb = ra + C::HB * C::UW;
        fft128_core<false, C::TWB>(ZU + zr * FFT_PITCH, tw, tt, [&](int j) { return make_float2(ra[8 * j], rb[8 * j]); }, mask, 0);
      }
      __syncthreads();
      // b. blur: O[y][bin] = sum_ky Wc[ky][bin] Zu[y + ky][bin] -> ZE row y, two half-chunks to stay within registers
      {
        constexpr int HC = C::CHUNK / 2;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int y0 = chunk * C::CHUNK + half * HC;
          float2 z[HC + K - 1];
#pragma unroll
          for (int i = 0; i < HC + K - 1; ++i) z[i] = ZU[(y0 + i) * FFT_PITCH + bin];
          float2 o[HC];
#pragma unroll
          for (int y = 0; y < HC; ++y) o[y] = make_float2(0.f, 0.f);
#pragma unroll
          for (int ky = 0; ky < K; ++ky) {
            const float2 w = WS[ky * FFT_N + bin];
#pragma unroll
            for (int y = 0; y < HC; ++y) o[y] = cfma(z[y + ky], w, o[y]);
          }
#pragma unroll
          for (int y = 0; y < HC; ++y) ZE[(y0 + y) * FFT_PITCH + bin] = o[y];
        }
      }
      __syncthreads();
      // c. back to the signal domain: real part = blurred row y, imaginary part = blurred row y + HB
      for (int task = tid; task < C::HB * 8; task += C::THREADS) {
        const unsigned mask = __activemask();
        const int zr = task >> 3, tt = task & 7;
        float2* row = ZE + zr * FFT_PITCH;
        fft128_core<true, C::TWB>(row, tw, tt, [&](int j) { return row[tt + 8 * j]; }, mask, 0);
      }
      __syncthreads();
      // (c, d and e are all row-local; merged into one phase without block barriers -- each 8-thread group doing inverse FFT,
      //  residual and forward FFT of its row back to back -- they are SLOWER, 0.528 vs 0.503 ms: only 320 of the 512 threads
      //  have a row, and the reload of the real-data stage can then start only after all three)
      // d. residual = blur - image inside the image and the owned rows, zero elsewhere (and in the 16 invalid columns).
      //    A thread keeps its column (THREADS is a multiple of 128): everything that depends on the column only is
      //    computed once per tile (ncu: written per element, this phase was 16 % of the kernel's instructions, all integer).
      {
        static_assert(C::THREADS % FFT_N == 0, "one column per thread");
        constexpr int RSTEP = C::THREADS / FFT_N;
        const int n = tid & (FFT_N - 1);
        const int X = bx * C::TWO - C::P4 + n;
        const bool nvalid = (n >= C::P4) && (n < C::P4 + C::TWO);
        const bool colin = nvalid && X >= C::P && X < C::P + g.N;
        const bool store = err_out != nullptr && nvalid && X < g.pitch;
        const int rl0 = by * C::TROWS;                             // first tile row inside the owned rows
        const int lim = min(nown, C::P + g.M - g.row0 - g.own0);   // rows rl in [lo, lim) lie inside the image and the band
        const int lo = C::P - g.row0 - g.own0;
        float* eo = err_out + size_t(c) * g.plane + size_t(g.own0 + rl0) * g.pitch + X;
        if (rl0 >= lo && rl0 + C::TROWS <= lim) {
          // all rows of the tile lie inside the image and the band (all but the border tiles): no row tests, straight-line
          // code (with them this phase was still 17 % of the kernel's instructions: 37 per row pair, now 9)
          if (!store) {
#pragma unroll
            for (int r = tid >> 7; r < C::HB; r += RSTEP) {
              const float2 v = ZE[r * FFT_PITCH + n];
              const float ea = colin ? v.x - eR[r * FFT_N + n] : 0.f;
              const float eb = colin ? v.y - eR[(r + C::HB) * FFT_N + n] : 0.f;
              ZE[r * FFT_PITCH + n] = make_float2(ea, eb);
            }
          } else {
#pragma unroll
            for (int r = tid >> 7; r < C::HB; r += RSTEP) {
              const float2 v = ZE[r * FFT_PITCH + n];
              const float ea = colin ? v.x - eR[r * FFT_N + n] : 0.f;
              const float eb = colin ? v.y - eR[(r + C::HB) * FFT_N + n] : 0.f;
              eo[size_t(r) * g.pitch] = ea;
              eo[size_t(r + C::HB) * g.pitch] = eb;
              ZE[r * FFT_PITCH + n] = make_float2(ea, eb);
            }
          }
        } else
#pragma unroll 2
        for (int r = tid >> 7; r < C::HB; r += RSTEP) {
          const float2 v = ZE[r * FFT_PITCH + n];
          const int ra = rl0 + r, rb = ra + C::HB;
          const float ea = (colin && ra >= lo && ra < lim) ? v.x - eR[r * FFT_N + n] : 0.f;
          const float eb = (colin && rb >= lo && rb < lim) ? v.y - eR[(r + C::HB) * FFT_N + n] : 0.f;
          if (store && ra < nown) eo[size_t(r) * g.pitch] = ea;
          if (store && rb < nown) eo[size_t(r + C::HB) * g.pitch] = eb;
          ZE[r * FFT_PITCH + n] = make_float2(ea, eb);
        }
      }
      __syncthreads();
      if (tid == 0 && q + 1 < total) issue(q + 1);    // u and image rows are consumed: reload under the phases below
      // e. spectra of the packed residual rows
      for (int task = tid; task < C::HB * 8; task += C::THREADS) {
        const unsigned mask = __activemask();
        const int zr = task >> 3, tt = task & 7;
        float2* row = ZE + zr * FFT_PITCH;
        fft128_core<false, C::TWB>(row, tw, tt, [&](int j) { return row[tt + 8 * j]; }, mask, 0);
      }
      __syncthreads();
    }
    // MAC: C[dy][bin] += conj(Ze[y][bin]) * Zu[y + dy][bin]; a K-deep register window slides down the packed u rows
    {
      const float2* zu0 = ZU + (chunk * C::CHUNK) * FFT_PITCH + bin;
      const float2* ze0 = ZE + (chunk * C::CHUNK) * FFT_PITCH + bin;
      float2 win[K];
#pragma unroll
      for (int i = 0; i < K - 1; ++i) win[i] = zu0[i * FFT_PITCH];
#pragma unroll
      for (int y = 0; y < C::CHUNK; ++y) {
        win[(y + K - 1) % K] = zu0[(y + K - 1) * FFT_PITCH];
        const float2 e = ze0[y * FFT_PITCH];
#pragma unroll
        for (int d = 0; d < K; ++d) {
          acc[d] = cfma_conj(e, win[(y + d) % K], acc[d]);
        }
      }
    }
    __syncthreads();      // ZU / ZE are free for the next tile
    const int c = (blockIdx.x + q * gridDim.x) / tiles_per_c;
    if (q + 1 == total || int(blockIdx.x + (q + 1) * gridDim.x) / tiles_per_c != c) {
      flushed |= 1u << c;
      // end of a channel: fold the row chunks (fixed order) through the spectra buffers and write this CTA's partial
      float2* red = ZU;   // [chunk][K][128]
#pragma unroll
      for (int d = 0; d < K; ++d) {
        red[(chunk * K + d) * FFT_N + bin] = acc[d];
        acc[d] = make_float2(0.f, 0.f);
      }
      __syncthreads();
      for (int o = tid; o < K * FFT_N; o += C::THREADS) {
        float2 s0 = red[o];
#pragma unroll
        for (int ch = 1; ch < C::NCH; ++ch) {
          const float2 v = red[ch * K * FFT_N + o];
          s0.x += v.x;
          s0.y += v.y;
        }
        part[(size_t(blockIdx.x) * 3 + c) * K * FFT_N + o] = s0;
      }
      __syncthreads();
    }
  }
  for (int c = 0; c < 3; ++c)                         // channels this CTA had no tile of
    if (!(flushed & (1u << c)))
      for (int o = tid; o < K * FFT_N; o += C::THREADS) part[(size_t(blockIdx.x) * 3 + c) * K * FFT_N + o] = make_float2(0.f, 0.f);
}

// Finish of the PSF gradient, one CTA per (channel, dy) row of frequency-domain sums, everything in double and in a
// fixed order (deterministic):  sum the per-CTA partials -> untangle A[k] = (C[k] + conj(C[-k]))/2 (two real rows were
// packed per complex FFT) -> K-lag inverse DFT -> gk_sum.  With row bands the last CTA raises the "published" flags.
// FOLD (single launch for the whole PSF step): the last CTA to finish also runs the PSF update (pyx:574-589, waiting for
// the other bands' sums first when the frame is sharded) and refreshes the tap spectra of the new PSF -- two single-CTA
// launches and their launch gaps less per inner step.
__global__ void __launch_bounds__(512)
k_gradk_fft_finish(State* __restrict__ st, const float2* __restrict__ part, int nparts, int K,
                   double* __restrict__ gk_sum, CommPeers cp, int seq, unsigned* __restrict__ ticket,
                   int fold, float step, int correlation, float* __restrict__ psf, float* __restrict__ psf_caller,
                   float2* __restrict__ wspec) {
  if (st->stop) return;
  extern __shared__ float sm_psf[];
  __shared__ double2 sh[4][FFT_N];
  __shared__ double2 A[FFT_N];
  __shared__ double2 tw64[FFT_N];
  const int tid = threadIdx.x, k = tid & (FFT_N - 1), qtr = tid >> 7, P = K / 2;
  const int row = blockIdx.x;                                  // c * K + dy
  const size_t nelem = size_t(3) * K * FFT_N;
  const float2* src = part + size_t(row) * FFT_N + k;
  const int per = (nparts + 3) / 4, b0 = qtr * per, b1 = min(nparts, b0 + per);
  double ax[4] = {0, 0, 0, 0}, ay[4] = {0, 0, 0, 0};
  int b = b0;
  // 20 loads in flight per thread: the partials sit in L2 (just written), so this loop is pure latency -- with 8 in
  // flight the 37 partials of a quarter cost five L2 round trips (26 us for the kernel), now two
  for (; b + 20 <= b1; b += 20) {
    float2 v[20];
#pragma unroll
    for (int j = 0; j < 20; ++j) v[j] = __ldcg(src + size_t(b + j) * nelem);
#pragma unroll
    for (int j = 0; j < 20; ++j) { ax[j & 3] += double(v[j].x); ay[j & 3] += double(v[j].y); }
  }
  for (; b + 8 <= b1; b += 8) {
    float2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcg(src + size_t(b + j) * nelem);
#pragma unroll
    for (int j = 0; j < 8; ++j) { ax[j & 3] += double(v[j].x); ay[j & 3] += double(v[j].y); }
  }
  for (; b < b1; ++b) {
    const float2 v = __ldcg(src + size_t(b) * nelem);
    ax[0] += double(v.x);
    ay[0] += double(v.y);
  }
  sh[qtr][k] = make_double2((ax[0] + ax[1]) + (ax[2] + ax[3]), (ay[0] + ay[1]) + (ay[2] + ay[3]));
  if (tid < FFT_N) {
    double sn, cs;
    sincospi(2.0 * double(tid) / double(FFT_N), &sn, &cs);
    tw64[tid] = make_double2(cs, sn);
  }
  __syncthreads();
  if (tid < FFT_N)
    sh[0][k] = make_double2((sh[0][k].x + sh[1][k].x) + (sh[2][k].x + sh[3][k].x), (sh[0][k].y + sh[1][k].y) + (sh[2][k].y + sh[3][k].y));
  __syncthreads();
  if (tid < FFT_N) {
    const double2 a = sh[0][k], c = sh[0][(FFT_N - k) & (FFT_N - 1)];
    A[k] = make_double2(0.5 * (a.x + c.x), 0.5 * (a.y - c.y));
  }
  __syncthreads();
  // lag dx - P by 16 threads: 8 bins each, then a fixed xor tree (32 lags per pass of the 512 threads)
  const int par = seq & 1;
  for (int dx0 = 0; dx0 < K; dx0 += 32) {
    const int dx = dx0 + (tid >> 4), j = tid & 15;
    double acc = 0.0;
    if (dx < K) {
      const int s = dx - P;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kk = j * 8 + i;
        const double2 a = A[kk], w = tw64[(kk * s) & (FFT_N - 1)];
        acc += a.x * w.x - a.y * w.y;                            // Re(A[k] e^{+2 pi i k s / N})
      }
    }
#pragma unroll
    for (int m = 1; m < 16; m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (dx < K && j == 0) {
      const double sum = acc * (1.0 / double(FFT_N));
      const int o = row * K + dx;
      gk_sum[o] = sum;
      if (cp.nranks > 1)
        for (int rr = 0; rr < cp.nranks; ++rr) cp.peer[rr]->gk_val[par][cp.rank][o] = sum;
    }
  }
  if (cp.nranks > 1 || fold) {
    __shared__ int s_last;
    // the sums cross GPUs only with row bands; on one GPU a device-scope fence orders them before the ticket (the
    // system-scope fence was 36 % of this kernel's stall samples)
    if (cp.nranks > 1) __threadfence_system(); else __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
      if (last) {
        *ticket = 0u;
        if (cp.nranks > 1) __threadfence_system(); else __threadfence();
        if (cp.nranks > 1) {
          for (int rr = 0; rr < cp.nranks; ++rr) *reinterpret_cast<volatile int*>(&cp.peer[rr]->gk_flag[par][cp.rank]) = seq;
          __threadfence_system();
        }
      }
      s_last = last;
    }
    __syncthreads();
    if (fold && s_last) {
      psf_update_body(st, gk_sum, K, step, correlation, psf, psf_caller, cp.peer[cp.rank], cp.nranks, seq, sm_psf);
      float2* twf = reinterpret_cast<float2*>(sh);     // the partial-sum scratch is free now
      if (tid < FFT_N) twf[tid] = make_float2(float(tw64[tid].x), float(tw64[tid].y));
      __syncthreads();
      // tap spectra of the new PSF (same definition as k_psf_spectrum), twiddles from the table built above
      const int n_out = 2 * 3 * K * FFT_N;
      for (int o = tid; o < n_out; o += blockDim.x) {
        const int kb = o & (FFT_N - 1), r = o >> 7, ky = r % K, c = (r / K) % 3, dir = r / (3 * K);
        float re = 0.f, im = 0.f;                       // float like k_psf_spectrum (B200's FP64 rate would make this the
        for (int kx = 0; kx < K; ++kx) {                // longest part of the PSF step)
          const int src = dir ? (ky * K + kx) : ((K - 1 - ky) * K + (K - 1 - kx));
          const float w = psf[size_t(c) * K * K + src];
          const float2 e = twf[((kx - P) * kb) & (FFT_N - 1)];
          re = fmaf(w, e.x, re);
          im = fmaf(w, e.y, im);
        }
        wspec[o] = make_float2(re * (1.f / FFT_N), im * (1.f / FFT_N));
      }
    }
  }
}

}  // namespace rltv
