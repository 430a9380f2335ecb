// Spectral chain kernel: forward blur, residual and adjoint of one inner step in ONE pass, 11 <= K <= 17
//   err = conv(u, psf, "valid") - image ;  g = conv(err, rot180 psf, "full")            lib/deconvolution.pyx:477-491
//   + the step statistics max(u_c), max|lambda g + (u - ut)/2|                            pyx:519, :524
//
// Like the row-FFT hybrid kernels (rltv_stencil_fft.cuh) the x direction lives in the frequency domain (128-sample
// row segments, two real rows packed per complex FFT) and the y direction is a direct sum over K tap spectra.  What
// is new: the residual never returns to the signal domain.
//     Err^[y] = 128 * sum_ky W0[ky] U^[y-P+ky] - I^[y]          (I^ = packed spectra of the image rows, computed once
//     G^[y]   =       sum_ky W1[ky] Err^[y-P+ky]                  per solve: the image does not change)
// so a row costs ONE forward FFT (u) and ONE inverse FFT (g) instead of four, at the price of a shorter valid span:
// two chained circular convolutions leave V = 128 - 2(K-1) valid columns per segment (100 at K = 15).
// Rows are processed ROLLING down a column strip in steps of S = 24 packed rows, so no vertical halo is ever
// transformed twice; the spectra of the last 2P rows are kept for the next step.
//
// The residual must be ZERO outside the image (the 'full' convolution of pyx:491 sees zeros there), which a product of
// spectra cannot express.  Only the first and last segment of a row and the first / last P rows of the frame are
// affected; those steps take a slower path that drops to the signal domain once (inverse FFT, mask, forward FFT).
//
// One CTA per SM, 20 warps with fixed roles, ONE block barrier per step (software pipeline over steps t):
//   warps 14-19  forward FFT of the u rows of step t      TMA stage -> U block (t & 1)
//   warps  0-7   MAC of step t-1: thread = (frequency bin, half of the step's 24 packed rows): Err^ then G^, 720 FFMA2
//                per thread; two warps per scheduler so that one's window loads / stores hide behind the other's FMAs;
//                the two halves meet at two 256-thread named barriers (Err^ complete / carried rows free)
//   warps  8-13  inverse FFT of the g rows of step t-2 + epilogue (g store, step statistics), row-local; its u / ut
//                operands are prefetched into L2 one time step ahead
// The FMA pipe is the bound (MAC role: 2880 pipe cycles per step and SM sub-partition, FFT roles ~1000).
#pragma once
#include "rltv_band.cuh"
#include "rltv_common.cuh"
#include "rltv_fft.cuh"
#include "rltv_tma.cuh"

namespace rltv {

// One unit of work: a column strip (channel c, segment s) and two row ranges that ride through the complex FFTs as
// real and imaginary part: g rows [ya, ya + L) and [ya + L, ya + L + Lb), Lb <= L (local rows of the band).
struct ChainPiece {
  int c, s;
  int ya, L, Lb;
  int nsteps;        // ceil((L + 4P) / S)
  int ipk_row0;      // first packed row of this piece in the image-spectra array (nsteps * S rows)
  int xfix;          // 1: the segment touches the left / right border of the image (masking needed every step)
};

template <int K>
struct ChainCfg {
  static_assert(K >= 9 && K <= 17, "two chained circular convolutions on 128-sample segments");
  static constexpr int P = K / 2;
  static constexpr int V = FFT_N - 2 * (K - 1);             // valid g columns per segment
  static constexpr int DX = (4 - ((K - 1) & 3)) & 3;        // FFT sample n sits at TMA box column n + DX (16-byte aligned box start)
  static constexpr int S = 24;                              // packed rows per step
  static constexpr int INW = 136;                           // TMA box width (floats): rows start 8 banks apart
#ifndef RLTV_CHAIN_WMAC
#define RLTV_CHAIN_WMAC 4
#endif
  static constexpr int WMAC = RLTV_CHAIN_WMAC, WIFFT = 6, WFFT = 6;   // warps per role (WMAC = 4: one MAC thread per bin, 8: two)
  static constexpr int THREADS = 32 * (WMAC + WIFFT + WFFT);
  static constexpr int NMAC = 32 * WMAC, NIFFT = 32 * WIFFT, NFFT = 32 * WFFT; // role sizes (threads)
  static constexpr int MH = WMAC / 4;                       // MAC threads per frequency bin
  static constexpr int MROWS = S / MH;                      // rows of a step owned by one MAC thread
  static constexpr int MR = (MH == 1) ? 8 : 6;              // MAC role: rows per register block
  static constexpr int PCACHE = 24;                         // pieces of this CTA kept in shared memory
  static constexpr int T2 = 2 * P;                          // rows carried from one step to the next
  static constexpr int WP = 65;                             // tap spectra are Hermitian: bins 0..64 are stored
  static constexpr int U_BYTES = 2 * S * FFT_PITCH * 8;     // two blocks written by the forward FFT role
  static constexpr int UT_BYTES = T2 * FFT_N * 8;           // last 2P rows of the previous block (private columns)
  static constexpr int GB_BYTES = 2 * S * FFT_PITCH * 8;    // G^ blocks: MAC role writes one, inverse FFT role works on the other
  static constexpr int EC_BYTES = S * FFT_N * 8;            // Err^ of the current step (private columns)
  static constexpr int ET_BYTES = T2 * FFT_N * 8;
  static constexpr int IN_BYTES = 2 * S * INW * 4;          // TMA stage: S real rows of each half
  static constexpr int W_BYTES = ((2 * K * WP * 8) + 127) & ~127;
  static constexpr int TW_BYTES = FFT_TW_BYTES;
  static constexpr int OFF_UT = U_BYTES, OFF_GB = OFF_UT + UT_BYTES, OFF_EC = OFF_GB + GB_BYTES, OFF_ET = OFF_EC + EC_BYTES,
                       OFF_IN = OFF_ET + ET_BYTES, OFF_W = OFF_IN + IN_BYTES, OFF_TW = OFF_W + W_BYTES, OFF_BAR = OFF_TW + TW_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 64 + 128;
  static_assert(V % 4 == 0 && (K - 1 + DX) % 4 == 0, "float4 epilogue / aligned TMA box");
  static_assert(OFF_IN % 128 == 0, "TMA destination alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
  static_assert(T2 <= S, "the carried rows come from one block");
  static_assert(MROWS % MR == 0 && T2 % 2 == 0 && NIFFT == 8 * S && NFFT == 8 * S && (WMAC == 4 || WMAC == 8), "role geometry");
};

__host__ __device__ inline int chain_nseg(int Wu, int V) { return (Wu + V - 1) / V; }

// ------------------------------------------------------------------------------------------------
// Packed image spectra, once per solve:  Ipk[piece row r][k] = FFT_128( img[ya - 3P + r] + i img[ya + L - 3P + r] )
// over the segment's 128 columns (zero outside the planes).  One 8-thread group per packed row.
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256)
k_chain_image_spectra(Geom g, const float* __restrict__ img, const ChainPiece* __restrict__ pieces, int npieces,
                      float2* __restrict__ Ipk) {
  using C = ChainCfg<K>;
  __shared__ float2 buf[32 * FFT_PITCH];
  __shared__ float2 tw[FFT_TW_ENTRIES];
  fft_fill_twiddles(tw);
  __syncthreads();
  const int grp = threadIdx.x >> 3, tt = threadIdx.x & 7;
  for (int p = blockIdx.y; p < npieces; p += gridDim.y) {
    const ChainPiece pc = pieces[p];
    const int nrows = pc.nsteps * C::S;
    const int x0 = C::V * pc.s - (K - 1);
    for (int r0 = blockIdx.x * 32; r0 < nrows; r0 += gridDim.x * 32) {
      const int r = r0 + grp;
      if (r < nrows) {
        const unsigned mask = __activemask();
        const int ra = pc.ya - 3 * C::P + r, rb = ra + pc.L;
        const bool oka = ra >= 0 && ra < g.Hu, okb = rb >= 0 && rb < g.Hu;
        const float* pa = img + size_t(pc.c) * g.plane + size_t(oka ? ra : 0) * g.pitch;
        const float* pb = img + size_t(pc.c) * g.plane + size_t(okb ? rb : 0) * g.pitch;
        float2* row = buf + grp * FFT_PITCH;
        fft128_row<false>(row, tw, tt, [&](int n) {
          const int X = x0 + n;
          const bool in = X >= 0 && X < g.pitch;
          return make_float2((in && oka) ? __ldg(pa + X) : 0.f, (in && okb) ? __ldg(pb + X) : 0.f);
        }, mask, 0);
        __syncwarp(mask);
        float2* dst = Ipk + (size_t(pc.ipk_row0) + r) * FFT_N;
        for (int n = tt; n < FFT_N; n += 8) dst[n] = row[n];
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// role helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float4 lds128(const void* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// MAC over one block of R rows with a ROLLING register window:  acc[y] = sum_ky w[ky] * z[y + ky], where z[0 .. 2P) are
// the last 2P window rows of the previous block (or the rows carried over from the previous step) and z[2P .. 2P + R)
// the block's own rows, loaded here.  The caller moves the window down by R rows afterwards.  The loop over the blocks
// of a step is NOT unrolled: straight-line code of a whole step (~30 KB) made the role instruction-fetch bound (ncu:
// 'no instruction' was its top stall); one block body fits the L0 instruction cache.
template <int K, int R, int CPITCH>
__device__ __forceinline__ void chain_mac_block(float2 (&z)[R + K - 1], const float2* __restrict__ cur,
                                                const float2* __restrict__ w, float sgn, float2 (&acc)[R]) {
  using C = ChainCfg<K>;
#pragma unroll
  for (int m = 0; m < R; ++m) z[C::T2 + m] = cur[m * CPITCH];
#pragma unroll
  for (int y = 0; y < R; ++y) acc[y] = make_float2(0.f, 0.f);
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    float2 wv = w[ky * C::WP];
    wv.y *= sgn;                       // bins above 64 read the spectrum of bin 128 - k: conjugate (real taps)
#pragma unroll
    for (int y = 0; y < R; ++y) acc[y] = cfma(z[y + ky], wv, acc[y]);
  }
}

// Debug: bit mask of roles whose work is skipped (1 forward FFT, 2 MAC, 4 inverse FFT + epilogue, 8 TMA loads); results
// are garbage then -- used only to time the roles in isolation (rltv_debug_chain_roles).
__device__ int g_chain_skip_roles = 0;

struct ChainCursor {   // (piece, step) of one role; advanced once per time step
  int p, j;
};

// ------------------------------------------------------------------------------------------------
// The kernel.  grid = number of SMs; CTA b works through pieces [cta_first[b], cta_first[b + 1]).
// `slot`: which half of State::smax this inner step accumulates its statistics in.
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(ChainCfg<K>::THREADS, 1)
k_chain_fft(const __grid_constant__ CUtensorMap tm_u, Geom g, State* __restrict__ st, const float2* __restrict__ wspec,
            const float2* __restrict__ Ipk, const ChainPiece* __restrict__ pieces_g, const int* __restrict__ cta_first,
            const float* __restrict__ ug, const float* __restrict__ utg, float lambd, float* __restrict__ gout,
            int slot, int ut_is_u, CommPeers cp, int seq, unsigned* __restrict__ done_counter) {
  using C = ChainCfg<K>;
  if (st->stop) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float2* U = reinterpret_cast<float2*>(smem);
  float2* UT = reinterpret_cast<float2*>(smem + C::OFF_UT);
  float2* GB = reinterpret_cast<float2*>(smem + C::OFF_GB);
  float2* EC = reinterpret_cast<float2*>(smem + C::OFF_EC);
  float2* ET = reinterpret_cast<float2*>(smem + C::OFF_ET);
  const float* IN = reinterpret_cast<const float*>(smem + C::OFF_IN);
  float2* WS = reinterpret_cast<float2*>(smem + C::OFF_W);          // [2][K][WP]: dir 0 scaled by 128, dir 1
  float2* tw = reinterpret_cast<float2*>(smem + C::OFF_TW);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p0 = cta_first[blockIdx.x], p1 = cta_first[blockIdx.x + 1];
  if (p0 >= p1) {
    if (cp.nranks > 1) band_step_max_tail_slot(st, slot, cp, seq, done_counter, reinterpret_cast<int*>(smem));
    return;
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_u);
  }
  fft_fill_twiddles(tw);
  __syncthreads();

  // this CTA's pieces in shared memory: with 220 KB of dynamic shared memory there is next to no L1 left, and every
  // role reads the current piece several times per step (ncu: those L2 round trips were the top long-scoreboard stall)
  __shared__ ChainPiece spc[C::PCACHE];
  for (int i = tid; i < C::PCACHE && p0 + i < p1; i += C::THREADS) spc[i] = pieces_g[p0 + i];
  __syncthreads();
  auto piece_nsteps = [&](int p) { return (p - p0 < C::PCACHE) ? spc[p - p0].nsteps : pieces_g[p].nsteps; };
  auto piece = [&](int p) -> ChainPiece { return (p - p0 < C::PCACHE) ? spc[p - p0] : pieces_g[p]; };
  int total_steps = 0;
  for (int p = p0; p < p1; ++p) total_steps += piece_nsteps(p);
  const int imgLo = C::P - g.row0, imgHi = C::P + g.M - g.row0;       // local rows on which the residual exists

  // TMA of the u rows of (piece, step): real rows [ya - 2P + jS, +S) and the same of the second half
  auto issue = [&](const ChainPiece& pc, int j) {
    const int xb = C::V * pc.s - (K - 1) - C::DX;
    const int y = pc.ya - 2 * C::P + j * C::S;
    mbar_arrive_expect_tx(bar, C::IN_BYTES);
    tma_load_3d(smem + C::OFF_IN, &tm_u, xb, y, pc.c, bar);
    tma_load_3d(smem + C::OFF_IN + C::S * C::INW * 4, &tm_u, xb, y + pc.L, pc.c, bar);
  };
  auto advance = [&](ChainCursor& cur) {
    if (++cur.j >= piece_nsteps(cur.p)) { cur.j = 0; ++cur.p; }
  };
  // a step needs the slow (masking) path if its residual rows touch rows outside the image or the segment is a border one
  auto needs_fix = [&](const ChainPiece& pc, int j) {
    const int ea = pc.ya + j * C::S - 3 * C::P, eb = ea + pc.L;
    const bool rowfix = (ea < imgLo) || (ea + C::S > imgHi) || (eb < imgLo) || (eb + C::S > imgHi);
    return pc.xfix || rowfix;
  };

  ChainCursor cf{p0, 0}, cm{p0, 0}, ci{p0, 0};          // forward FFT / MAC / inverse FFT cursors
  const int skip = g_chain_skip_roles;
  if (tid == 0 && !(skip & 8)) issue(piece(p0), 0);
  float mu = -INFINITY, mG = 0.f;                       // inverse-FFT role: statistics of the current piece's channel
  constexpr int NQ = (C::V / 4 + 7) / 8;                // float4 columns per thread of a row's 8-thread group
  float4 da[NQ], db[NQ];                                // inverse-FFT role: (u - ut)/2 under its next g row pair
#pragma unroll
  for (int q = 0; q < NQ; ++q) da[q] = db[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int RB = C::MR;                             // MAC role: rows per block
  int cur_wc = -1;                                      // MAC role: channel whose tap spectra are in shared memory
  // MAC role: thread = (frequency bin k, half h of the step's rows)
  const int mk = tid & (FFT_N - 1), mh = (C::MH == 1) ? 0 : ((tid >> 7) & 1), mr0 = mh * C::MROWS;
  float2 iv[RB];                                        // image spectra of the thread's next block (loaded one block ahead)
#pragma unroll
  for (int y = 0; y < RB; ++y) iv[y] = make_float2(0.f, 0.f);
  if (warp < C::WMAC) {
    const float2* ip = Ipk + (size_t(piece(p0).ipk_row0) + mr0) * FFT_N + mk;
#pragma unroll
    for (int y = 0; y < RB; ++y) iv[y] = __ldg(ip + size_t(y) * FFT_N);
  }

  for (int t = 0; t < total_steps + 2; ++t) {
    if (warp >= C::WMAC + C::WIFFT) {
      // ---------------- forward FFT of step t ----------------
      const int task = tid - 32 * (C::WMAC + C::WIFFT);  // 0..191: row = task >> 3
      const int zr = task >> 3, tt = task & 7;
      if (t < total_steps && !(skip & 1)) {
        {   // the MAC role reads the image spectra of this step during the NEXT time step: bring them into L2 now
          const ChainPiece pf = piece(cf.p);
          prefetch_l2(Ipk + (size_t(pf.ipk_row0) + size_t(cf.j) * C::S) * FFT_N + task * 16);
        }
        if (!(skip & 8)) mbar_wait(bar, t & 1);
        const float* ra = IN + zr * C::INW + C::DX + tt;
        float2* dst = U + ((t & 1) * C::S + zr) * FFT_PITCH;
        fft128_core<false>(dst, tw, tt, [&](int j) { return make_float2(ra[8 * j], ra[C::S * C::INW + 8 * j]); }, 0xffffffffu, 0);
      }
      // slow path of the MAC role's step (t - 1): signal-domain masking of its residual rows
      if (t >= 1 && t <= total_steps) {
        const ChainPiece pc = piece(cm.p);
        if (needs_fix(pc, cm.j)) {
          named_bar_sync(1, C::NMAC + C::NFFT);          // MAC role has written Err^ of this step
          float2* scratch = GB + (((t - 1) & 1) * C::S + zr) * FFT_PITCH;   // G^ block of this step: not written yet
          const float2* erow = EC + zr * FFT_N;
          fft128_core<true>(scratch, tw, tt, [&](int j) { return erow[tt + 8 * j]; }, 0xffffffffu, 0);
          __syncwarp();
          const int ea = pc.ya + cm.j * C::S - 3 * C::P + zr, eb = ea + pc.L;
          const bool rowa = ea >= imgLo && ea < imgHi, rowb = eb >= imgLo && eb < imgHi;
          const int x0 = C::V * pc.s - (K - 1);
          const float sc = 1.f / FFT_N;
          for (int n = tt; n < FFT_N; n += 8) {
            const int X = x0 + n;
            const bool col = X >= C::P && X < C::P + g.N;
            float2 v = scratch[n];
            v.x = (col && rowa) ? v.x * sc : 0.f;
            v.y = (col && rowb) ? v.y * sc : 0.f;
            scratch[n] = v;
          }
          __syncwarp();
          fft128_core<false>(scratch, tw, tt, [&](int j) { return scratch[tt + 8 * j]; }, 0xffffffffu, 0);
          __syncwarp();
          float2* ew = EC + zr * FFT_N;
          for (int n = tt; n < FFT_N; n += 8) ew[n] = scratch[n];
          named_bar_sync(2, C::NMAC + C::NFFT);          // MAC role may read the masked Err^
        }
      }
    } else if (warp < C::WMAC) {
      // ---------------- MAC of step t - 1: thread = (bin mk, rows [mr0, mr0 + S/2) of the step) ----------------
      if (t >= 1 && t <= total_steps && !(skip & 2)) {
        const ChainPiece pc = piece(cm.p);
        const int kk = (mk <= 64) ? mk : FFT_N - mk;
        if (pc.c != cur_wc) {
          // tap spectra of this channel: forward taps scaled by 128 (Err^ must be the unnormalised spectrum, like I^)
          named_bar_sync(3, C::NMAC);
          for (int i = tid; i < 2 * K * C::WP; i += C::NMAC) {
            const int dir = i / (K * C::WP), r = i - dir * K * C::WP, ky = r / C::WP, b = r - ky * C::WP;
            float2 v = __ldg(wspec + ((size_t(dir) * 3 + pc.c) * K + ky) * FFT_N + b);
            if (dir == 0) { v.x *= float(FFT_N); v.y *= float(FFT_N); }
            WS[i] = v;
          }
          named_bar_sync(3, C::NMAC);
          cur_wc = pc.c;
        }
        const int blk = (t - 1) & 1;
        const float2* ucur = U + blk * C::S * FFT_PITCH + mk;
        float2* utail = UT + mk;
        float2* ecur = EC + mk;
        float2* etail = ET + mk;
        float2* gb = GB + blk * C::S * FFT_PITCH + mk;
        const float2* w0 = WS + kk;
        const float2* w1 = WS + K * C::WP + kk;
        const float2* ip = Ipk + (size_t(pc.ipk_row0) + size_t(cm.j) * C::S + mr0) * FFT_N + mk;
        const float sgn = (mk > 64) ? -1.f : 1.f;
        // Err^ = 128 sum W0 U^ - I^ on this thread's rows, blocks of RB rows
        {
          float2 z[RB + K - 1];
#pragma unroll
          for (int m = 0; m < C::T2; ++m) {
            const int x = mr0 + m;                       // window row in [tail | current block]
            z[m] = (x < C::T2) ? utail[x * FFT_N] : ucur[(x - C::T2) * FFT_PITCH];
          }
#pragma unroll 1
          for (int B = 0; B < C::MROWS / RB; ++B) {
            float2 acc[RB];
            chain_mac_block<K, RB, FFT_PITCH>(z, ucur + (mr0 + RB * B) * FFT_PITCH, w0, sgn, acc);
#pragma unroll
            for (int y = 0; y < RB; ++y) ecur[(mr0 + RB * B + y) * FFT_N] = make_float2(acc[y].x - iv[y].x, acc[y].y - iv[y].y);
            if (B + 1 < C::MROWS / RB) {                 // image spectra of the next block: L2 hits, one block of MACs to land
#pragma unroll
              for (int y = 0; y < RB; ++y) iv[y] = __ldg(ip + size_t(RB * (B + 1) + y) * FFT_N);
            }
#pragma unroll
            for (int m = 0; m < C::T2; ++m) z[m] = z[m + RB];
          }
        }
        if (C::MH > 1) named_bar_sync(4, C::NMAC);       // Err^ of the step complete; nobody reads the old U^ tail any more
        // rows the next step needs from this block of U^ (the block itself is overwritten during the next step)
#pragma unroll
        for (int m = 0; m < C::T2 / C::MH; ++m) {
          const int r = mh * (C::T2 / C::MH) + m;
          utail[r * FFT_N] = ucur[(C::S - C::T2 + r) * FFT_PITCH];
        }
        if (needs_fix(pc, cm.j)) {
          named_bar_arrive(1, C::NMAC + C::NFFT);
          named_bar_sync(2, C::NMAC + C::NFFT);
        }
        // G^ = sum W1 Err^
        {
          float2 z[RB + K - 1];
#pragma unroll
          for (int m = 0; m < C::T2; ++m) {
            const int x = mr0 + m;
            z[m] = (x < C::T2) ? etail[x * FFT_N] : ecur[(x - C::T2) * FFT_N];
          }
#pragma unroll 1
          for (int B = 0; B < C::MROWS / RB; ++B) {
            float2 acc[RB];
            chain_mac_block<K, RB, FFT_N>(z, ecur + (mr0 + RB * B) * FFT_N, w1, sgn, acc);
#pragma unroll
            for (int y = 0; y < RB; ++y) gb[(mr0 + RB * B + y) * FFT_PITCH] = acc[y];
#pragma unroll
            for (int m = 0; m < C::T2; ++m) z[m] = z[m + RB];
          }
        }
        if (C::MH > 1) named_bar_sync(4, C::NMAC);       // nobody reads the old Err^ tail any more
#pragma unroll
        for (int m = 0; m < C::T2 / C::MH; ++m) {
          const int r = mh * (C::T2 / C::MH) + m;
          etail[r * FFT_N] = ecur[(C::S - C::T2 + r) * FFT_N];
        }
        // image spectra of this thread's first block of the NEXT step: in flight across the barrier
        if (t < total_steps) {
          ChainCursor nx = cm;
          advance(nx);
          const float2* ipn = Ipk + (size_t(piece(nx.p).ipk_row0) + size_t(nx.j) * C::S + mr0) * FFT_N + mk;
#pragma unroll
          for (int y = 0; y < RB; ++y) iv[y] = __ldg(ipn + size_t(y) * FFT_N);
        }
      }
    } else {
      // ---------------- inverse FFT + epilogue of step t - 2 ----------------
      if (t >= 2 && !(skip & 4)) {
        const ChainPiece pc = piece(ci.p);
        const int task = tid - 32 * C::WMAC, zr = task >> 3, tt = task & 7;
        const int rel = ci.j * C::S - 4 * C::P + zr;                 // g row of the first half, relative to the piece
        const int ya = pc.ya + rel, yb = ya + pc.L;
        const bool va = rel >= 0 && rel < pc.L, vb = rel >= 0 && rel < pc.Lb;
        float2* row = GB + (((t - 2) & 1) * C::S + zr) * FFT_PITCH;
        const int xs = C::V * pc.s;
        const size_t offa = size_t(pc.c) * g.plane + size_t(va ? ya : 0) * g.pitch + xs;
        const size_t offb = size_t(pc.c) * g.plane + size_t(vb ? yb : 0) * g.pitch + xs;
        // (u - ut)/2 under this row pair is already in da / db: loaded at the end of the previous time step (below)
        fft128_core<true>(row, tw, tt, [&](int j) { return row[tt + 8 * j]; }, 0xffffffffu, 0);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int x4 = 4 * (tt + 8 * q);
          const bool ok = x4 < C::V && xs + x4 < g.pitch;
          if (!ok) continue;
          const float4 z01 = lds128(row + (K - 1) + x4);            // (re0, im0, re1, im1)
          const float4 z23 = lds128(row + (K - 1) + x4 + 2);
          const int Xm = xs + x4;
          const bool c0 = Xm < g.Wu, c1 = Xm + 1 < g.Wu, c2 = Xm + 2 < g.Wu, c3 = Xm + 3 < g.Wu;
          if (va) {
            const float4 o = make_float4(c0 ? z01.x : 0.f, c1 ? z01.z : 0.f, c2 ? z23.x : 0.f, c3 ? z23.z : 0.f);
            if (c0) mG = fmaxf(mG, fabsf(fmaf(lambd, o.x, da[q].x)));                         // pyx:519
            if (c1) mG = fmaxf(mG, fabsf(fmaf(lambd, o.y, da[q].y)));
            if (c2) mG = fmaxf(mG, fabsf(fmaf(lambd, o.z, da[q].z)));
            if (c3) mG = fmaxf(mG, fabsf(fmaf(lambd, o.w, da[q].w)));
            *reinterpret_cast<float4*>(gout + offa + x4) = o;
          }
          if (vb) {
            const float4 o = make_float4(c0 ? z01.y : 0.f, c1 ? z01.w : 0.f, c2 ? z23.y : 0.f, c3 ? z23.w : 0.f);
            if (c0) mG = fmaxf(mG, fabsf(fmaf(lambd, o.x, db[q].x)));
            if (c1) mG = fmaxf(mG, fabsf(fmaf(lambd, o.y, db[q].y)));
            if (c2) mG = fmaxf(mG, fabsf(fmaf(lambd, o.z, db[q].z)));
            if (c3) mG = fmaxf(mG, fabsf(fmaf(lambd, o.w, db[q].w)));
            *reinterpret_cast<float4*>(gout + offb + x4) = o;
          }
        }
        // end of a piece: flush the statistics of its channel
        if (ci.j == pc.nsteps - 1) {
          const float wu = warp_max(mu), wG = warp_max(mG);
          if (lane == 0) {
            atomicMax(&st->smax[slot][pc.c], f2ord(wu));
            atomicMax(&st->smax[slot][3 + pc.c], f2ord(wG));
          }
          mu = -INFINITY;
          mG = 0.f;
        }
      }
    }
    // cursors (every thread keeps all three: the roles must agree on them)
    if (t >= 2) advance(ci);
    if (t >= 1 && t <= total_steps) advance(cm);
    if (t < total_steps) advance(cf);
    // inverse-FFT role, end of the time step: (a) load u / ut under its NEXT g row pair (step t - 1) and keep (u - ut)/2 in
    // registers across the barrier -- the rows were prefetched into L2 one time step earlier, so this costs L2 latency;
    // (b) prefetch the rows of the step after that into L2.
    if (warp >= C::WMAC && warp < C::WMAC + C::WIFFT && t >= 1 && t <= total_steps && !(skip & 4)) {
      const ChainPiece pc = piece(ci.p);
      const int task = tid - 32 * C::WMAC, zr = task >> 3, tt = task & 7;
      const int rel = ci.j * C::S - 4 * C::P + zr;
      const int ya = pc.ya + rel, yb = ya + pc.L;
      const bool va = rel >= 0 && rel < pc.L, vb = rel >= 0 && rel < pc.Lb;
      const int xs = C::V * pc.s;
      const size_t offa = size_t(pc.c) * g.plane + size_t(va ? ya : 0) * g.pitch + xs;
      const size_t offb = size_t(pc.c) * g.plane + size_t(vb ? yb : 0) * g.pitch + xs;
      if (t < total_steps) {                               // (b) first: the prefetches do not wait for anything
        ChainCursor nx = ci;
        advance(nx);
        const ChainPiece pn = piece(nx.p);
        const int reln = nx.j * C::S - 4 * C::P + zr;
        const bool vna = reln >= 0 && reln < pn.L, vnb = reln >= 0 && reln < pn.Lb;
        const int xn = C::V * pn.s;
        // 2 arrays x 2 rows x (V * 4 bytes = up to 4 lines of 128 B, unaligned: 5) = 20 line addresses over 8 threads
        for (int i = tt; i < 20; i += 8) {
          const int arr = i & 1, part = (i >> 1) & 1, ln = i >> 2;
          const bool v = part ? vnb : vna;
          const int y = pn.ya + reln + (part ? pn.L : 0);
          int x = xn + 32 * ln;
          if (x > xn + C::V - 4) x = xn + C::V - 4;
          if (v && x < g.pitch && !(arr && ut_is_u))
            prefetch_l2((arr ? utg : ug) + size_t(pn.c) * g.plane + size_t(y) * g.pitch + x);
        }
      }
      // all loads of a batch are issued before any is consumed (clamped addresses instead of branches: with a branch
      // per load the compiler serialised eight DRAM round trips per step and this role became the critical path)
      constexpr int QB = (NQ + 1) / 2;
#pragma unroll
      for (int q0 = 0; q0 < NQ; q0 += QB) {
        float4 ua[QB], ta[QB], ub[QB], tb[QB];
#pragma unroll
        for (int i = 0; i < QB; ++i) {
          const int x4 = 4 * (tt + 8 * (q0 + i));
          const bool ok = (q0 + i) < NQ && x4 < C::V && xs + x4 < g.pitch;
          const int xc = ok ? x4 : 0;
          ua[i] = __ldg(reinterpret_cast<const float4*>(ug + offa + xc));
          ub[i] = __ldg(reinterpret_cast<const float4*>(ug + offb + xc));
          if (!ut_is_u) {
            ta[i] = __ldg(reinterpret_cast<const float4*>(utg + offa + xc));
            tb[i] = __ldg(reinterpret_cast<const float4*>(utg + offb + xc));
          } else {
            ta[i] = ua[i];
            tb[i] = ub[i];
          }
        }
#pragma unroll
        for (int i = 0; i < QB; ++i) {
          const int q = q0 + i;
          if (q >= NQ) continue;
          const int x4 = 4 * (tt + 8 * q);
          const bool ok = x4 < C::V && xs + x4 < g.pitch;
          const int Xm = xs + x4;
          const bool c0 = ok && Xm < g.Wu, c1 = ok && Xm + 1 < g.Wu, c2 = ok && Xm + 2 < g.Wu, c3 = ok && Xm + 3 < g.Wu;
          da[q] = make_float4(0.5f * (ua[i].x - ta[i].x), 0.5f * (ua[i].y - ta[i].y), 0.5f * (ua[i].z - ta[i].z), 0.5f * (ua[i].w - ta[i].w));
          db[q] = make_float4(0.5f * (ub[i].x - tb[i].x), 0.5f * (ub[i].y - tb[i].y), 0.5f * (ub[i].z - tb[i].z), 0.5f * (ub[i].w - tb[i].w));
          if (va) mu = fmaxf(mu, fmaxf(fmaxf(c0 ? ua[i].x : -INFINITY, c1 ? ua[i].y : -INFINITY), fmaxf(c2 ? ua[i].z : -INFINITY, c3 ? ua[i].w : -INFINITY)));
          if (vb) mu = fmaxf(mu, fmaxf(fmaxf(c0 ? ub[i].x : -INFINITY, c1 ? ub[i].y : -INFINITY), fmaxf(c2 ? ub[i].z : -INFINITY, c3 ? ub[i].w : -INFINITY)));
        }
      }
    }
    __syncthreads();
    // the TMA stage was consumed by the forward FFT of step t: load step t + 1
    if (tid == 0 && t + 1 < total_steps && !(skip & 8)) issue(piece(cf.p), cf.j);
  }
  if (cp.nranks > 1) band_step_max_tail_slot(st, slot, cp, seq, done_counter, reinterpret_cast<int*>(smem));
}

}  // namespace rltv
