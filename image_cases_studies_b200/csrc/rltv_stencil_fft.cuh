// Row-FFT hybrid stencils for 9 <= K <= 17: the FLOP-reducing form of the forward blur and the adjoint
// (lib/deconvolution.pyx:477-491, :557-565), which the reference itself evaluates as FFT convolutions (scipy).
//
// A K x K stencil  out[Y][X] = sum_{ky,kx} w[ky][kx] in[Y-P+ky][X-P+kx]  costs K*K FMAs per output directly.  Here the
// x direction is done in the frequency domain on 128-sample row segments (112 valid outputs each, overlap-save) and
// the y direction stays a direct sum -- but of SPECTRA:   Out^[Y] = sum_ky Wc[ky] (.) In^[Y-P+ky]   -> K complex MACs
// per (row, bin) = 4K FMAs per 112/128 outputs instead of K*K: 34 instead of 225 FMAs per output at K = 15, plus two
// 128-point FFTs per row.  Because the taps are real, two image rows ride through the complex FFT as real and
// imaginary parts (rows Y and Y+48 of a 96-row tile) and come out as the real and imaginary parts of the inverse
// transform: no real-FFT untangling anywhere.
//
// Per 96 x 112 output tile (one CTA of 512 threads per SM, persistent, TMA double-buffered input like k_conv):
//   1. forward FFT of the 48+K-1 packed rows straight out of the TMA buffer          (csrc/rltv_fft.cuh)
//   2. vertical MAC over spectra: thread = (bin, chunk of 12 output rows), K complex taps in registers
//   3. inverse FFT of the 48 packed output rows (into the consumed TMA buffer)
//   4. epilogue as in k_conv (residual / step statistics), 112 columns x 96 rows, operands read from global memory
#pragma once
#include "rltv_band.cuh"
#include "rltv_common.cuh"
#include "rltv_fft.cuh"
#include "rltv_stencil.cuh"
#include "rltv_tma.cuh"

namespace rltv {

template <int K>
struct FftCfg {
  static_assert(K >= 9 && K <= 17, "128-sample segments hold 112 outputs + K-1 halo only up to K = 17");
  static constexpr int P = K / 2;
  static constexpr int P4 = (P + 3) & ~3;          // 16-byte aligned TMA box start
  static constexpr int TWO = 112;                  // valid output columns per 128-sample segment
  static constexpr int HB = 48;                    // rows per packed block (real part: rows 0..47, imaginary: 48..95)
  static constexpr int TROWS = 2 * HB;
  static constexpr int IN_ROWS = TROWS + K - 1;    // TMA box height (real rows)
  static constexpr int ZROWS = HB + K - 1;         // packed complex rows
  static constexpr int CHUNK = 12;                 // output rows per MAC thread (4 chunks x 128 bins = 512 threads)
  static constexpr int THREADS = 512;
  static constexpr int IN_BYTES = IN_ROWS * FFT_N * 4;                        // TMA transaction size
  static constexpr int ZB_BYTES = ZROWS * FFT_PITCH * 8;
  static constexpr int OB_BYTES = HB * FFT_PITCH * 8;
  static constexpr int STAGE_BYTES = IN_BYTES > OB_BYTES ? IN_BYTES : OB_BYTES;   // the output spectra reuse the stage
  static_assert(P4 + TWO - 1 + P <= FFT_N - 1, "segment too short for this K");
  static_assert(OB_BYTES <= STAGE_BYTES, "output spectra alias the consumed input stage");
  static_assert(STAGE_BYTES % 128 == 0 && ZB_BYTES % 128 == 0, "TMA destinations: 128-byte aligned");
  static_assert(4 * CHUNK == HB && HB % 8 == 0, "MAC chunks / epilogue row slots cover the packed rows");
  static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + ZB_BYTES + FFT_N * 8 + 64 + 128;
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

template <int K, bool ADJ>
__global__ void __launch_bounds__(FftCfg<K>::THREADS, 1)
k_conv_fft(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ e0g, const float* __restrict__ e1g,
           Geom g, State* __restrict__ st, const float2* __restrict__ wspec,
           float lambd, float* __restrict__ out, int ntx, int nty, int ybeg, int yend, CommPeers cp, int seq,
           unsigned* __restrict__ done_counter) {
  using C = FftCfg<K>;
  if (st->stop) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float2* ZB = reinterpret_cast<float2*>(smem + 2 * C::STAGE_BYTES);
  float2* tw = reinterpret_cast<float2*>(smem + 2 * C::STAGE_BYTES + C::ZB_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(tw) + FFT_N * 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_per_c = ntx * nty, ntiles = 3 * tiles_per_c;

  if (!ADJ && blockIdx.x == 0 && tid < 3) {
    st->max_u[tid] = ORD_LOWEST;
    st->max_G[tid] = 0;
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
  }
  fft_fill_twiddles(tw);
  __syncthreads();

  auto issue_in = [&](int t, int s) {
    const int c = t / tiles_per_c, r = t - c * tiles_per_c;
    const int by = r / ntx, bx = r - by * ntx;
    mbar_arrive_expect_tx(&bars[s], C::IN_BYTES);
    tma_load_3d(smem + s * C::STAGE_BYTES, &tm_in, bx * C::TWO - C::P4, ybeg + by * C::TROWS - C::P, c, &bars[s]);
  };
  // MAC role of this thread: frequency bin and chunk of output rows
  const int bin = tid & (FFT_N - 1), chunk = tid >> 7;          // 4 chunks x 12 rows = 48 packed output rows
  float2 wreg[K];
  // epilogue role: 8 row slots x 56 column pairs = 448 of the 512 threads
  const int er = tid / (C::TWO / 2), ep = tid - er * (C::TWO / 2);
  const bool epi_active = er < 8;

  int t = blockIdx.x;
  if (tid == 0 && t < ntiles) issue_in(t, 0);
  int cur_c = -1;
  float mu = -INFINITY, mG = 0.f;
  for (int k = 0; t < ntiles; ++k, t += gridDim.x) {
    const int s = k & 1;
    const int c = t / tiles_per_c, r = t - c * tiles_per_c;
    const int by = r / ntx, bx = r - by * ntx;
    if (tid == 0) {
      const int tn = t + gridDim.x;
      if (tn < ntiles) issue_in(tn, s ^ 1);
    }
    if (c != cur_c) {
      const float2* wsrc = wspec + ((size_t(ADJ ? 1 : 0) * 3 + c) * K) * FFT_N + bin;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) wreg[ky] = __ldg(wsrc + ky * FFT_N);
      cur_c = c;
    }
    mbar_wait(&bars[s], (k >> 1) & 1);

    __syncwarp();
    // 1. forward FFT of the packed rows: Z[r] = in[r] + i in[r + HB], read straight from the TMA buffer
    {
      const float* inR = reinterpret_cast<const float*>(smem + s * C::STAGE_BYTES);
      for (int task = tid; task < C::ZROWS * 8; task += C::THREADS) {
        const unsigned mask = __activemask();
        const int zr = task >> 3, tt = task & 7;
        const float* ra = inR + zr * FFT_N;
        const float* rb = inR + (zr + C::HB) * FFT_N;
        fft128_row<false>(ZB + zr * FFT_PITCH, tw, tt, [&](int n) { return make_float2(ra[n], rb[n]); }, mask, zr & 3);
      }
    }
    __syncthreads();

    // 2. vertical MAC over spectra: O[y][bin] = sum_ky Wc[ky][bin] * Z[y + ky][bin]; O overwrites the consumed stage
    float2* OB = reinterpret_cast<float2*>(smem + s * C::STAGE_BYTES);
    {
      float2 z[C::CHUNK + K - 1];
#pragma unroll
      for (int i = 0; i < C::CHUNK + K - 1; ++i) z[i] = ZB[(chunk * C::CHUNK + i) * FFT_PITCH + bin];
#pragma unroll
      for (int y = 0; y < C::CHUNK; ++y) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          acc.x = fmaf(wreg[ky].x, z[y + ky].x, acc.x);
          acc.x = fmaf(-wreg[ky].y, z[y + ky].y, acc.x);
          acc.y = fmaf(wreg[ky].x, z[y + ky].y, acc.y);
          acc.y = fmaf(wreg[ky].y, z[y + ky].x, acc.y);
        }
        OB[(chunk * C::CHUNK + y) * FFT_PITCH + bin] = acc;
      }
    }
    __syncthreads();

    // Epilogue operands of this thread's outputs go to registers now, so their DRAM latency hides behind the
    // inverse FFT.  Epilogue thread = (row slot er of 8, column pair ep of 56): outputs (er + 8q [+ HB], 2ep..2ep+1).
    constexpr int NTASK = C::HB / 8;
    const int X = bx * C::TWO + 2 * ep;
    const int Ybase = ybeg + by * C::TROWS + er;
    const bool xok = epi_active && X < g.pitch;
    const size_t off0 = size_t(c) * g.plane + size_t(Ybase) * g.pitch + X;
    float2 pa[NTASK][2], pb[ADJ ? NTASK : 1][2];
#pragma unroll
    for (int q = 0; q < NTASK; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool ok = xok && (Ybase + 8 * q + h * C::HB) < yend;
        const size_t goff = ok ? off0 + size_t(8 * q + h * C::HB) * g.pitch : 0;
        pa[q][h] = ok ? __ldg(reinterpret_cast<const float2*>(e0g + goff)) : make_float2(0.f, 0.f);
        if (ADJ) pb[q][h] = ok ? __ldg(reinterpret_cast<const float2*>(e1g + goff)) : make_float2(0.f, 0.f);
      }

    // 3. inverse FFT of the packed output rows, in place
    for (int task = tid; task < C::HB * 8; task += C::THREADS) {
      const unsigned mask = __activemask();
      const int zr = task >> 3, tt = task & 7;
      float2* row = OB + zr * FFT_PITCH;
      fft128_row<true>(row, tw, tt, [&](int n) { return row[n]; }, mask);
    }
    __syncthreads();

    // 4. epilogue: real part -> rows 0..HB-1 of the tile, imaginary part -> rows HB..2HB-1; columns P4 .. P4+111 valid
    if (xok) {
      float* op = out + off0;
      const bool cin[2] = {ADJ ? (X < g.Wu) : (X >= C::P && X < C::P + g.N),
                           ADJ ? (X + 1 < g.Wu) : (X + 1 >= C::P && X + 1 < C::P + g.N)};
#pragma unroll
      for (int q = 0; q < NTASK; ++q) {
        const float4 zz = *reinterpret_cast<const float4*>(OB + (er + 8 * q) * FFT_PITCH + C::P4 + 2 * ep);   // 2 complex
        const float v[2][2] = {{zz.x, zz.z}, {zz.y, zz.w}};                                                   // [h][column]
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int Y = Ybase + 8 * q + h * C::HB;
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
          *reinterpret_cast<float2*>(op + size_t(8 * q + h * C::HB) * g.pitch) = make_float2(o[0], o[1]);
        }
      }
    }
    __syncthreads();   // stage s (now holding O) and ZB are free again
    if (ADJ) {
      const int tn = t + gridDim.x;
      const int cn = tn < ntiles ? tn / tiles_per_c : -1;
      if (cn != c) {
        __shared__ float red_u[16], red_G[16];
        const float wu = warp_max(mu), wG = warp_max(mG);
        if (lane == 0) { red_u[warp] = wu; red_G[warp] = wG; }
        __syncthreads();
        if (warp == 0) {
          float a = lane < 16 ? red_u[lane] : -INFINITY;
          float b = lane < 16 ? red_G[lane] : 0.f;
          a = warp_max(a);
          b = warp_max(b);
          if (lane == 0) {
            atomicMax(&st->max_u[c], f2ord(a));
            atomicMax(&st->max_G[c], f2ord(b));
          }
        }
        __syncthreads();
        mu = -INFINITY;
        mG = 0.f;
      }
    }
  }
  if (ADJ && cp.nranks > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      if (atomicAdd(done_counter, 1u) == gridDim.x - 1) {
        *done_counter = 0u;
        __threadfence();
        publish_step_max(st, cp, seq);
      }
    }
  }
}

}  // namespace rltv
