// Spectral chain kernel: forward blur, residual and adjoint of one inner step in ONE pass, 5 <= K <= 17
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
// Rows are processed ROLLING down a column strip in steps of S = 16 packed rows, so no vertical halo is ever
// transformed twice; the spectra of the previous step stay in a three-deep ring for the window of the next one.
//
// The residual must be ZERO outside the image (the 'full' convolution of pyx:491 sees zeros there), which a product of
// spectra cannot express.  Only the first and last segment of a row and the first / last P rows of the frame are
// affected; those steps take a slower path that drops to the signal domain once (inverse FFT, mask, forward FFT).
//
// One CTA per SM, 24 warps in FIVE roles -- warp groups (4 warps = one warp per SM sub-partition): E, G, IFFT and FFT one
// each, the epilogue two -- ONE block barrier per step, a five-stage software pipeline over the steps s of 16 packed rows:
//   role     works on step   reads -> writes
//   FFT      t               TMA stage (double-buffered) -> U^ ring block s % 3          [+ the masking path for role G]
//   E        t - 1           U^ block + last 2P rows of the previous ring block, W0 (registers), I^ (L2) -> Err^ ring block
//   G        t - 2           Err^ block + last 2P rows of the previous one, W1 (registers) -> G^ ring block
//   IFFT     t - 3           G^ block, in place (row-local)
//   EPI      t - 4           g rows (signal domain) + u / ut from L2 -> g store, step statistics   (16 threads per row pair)
// Every role has its own register budget (setmaxnreg: E 112, G 96, FFT / IFFT 72, EPI 64 per thread; the launch has 80).
// Round 2's first version had three roles (FFT 6 warps, MAC 4, IFFT + epilogue 6) at 24 rows per step.  ncu's source-level
// stall samples showed the lone MAC warp of each sub-partition busy 83 % of the kernel (2 200 dependent instructions per
// step at ~4 cycles each: it WAS the step time) and the IFFT + epilogue warps 74 % (latency chains: shared-memory round
// trips of the FFT passes, then the L2 round trips of the epilogue operands).  Splitting both along the pipeline -- not
// across rows, which needs barriers between the halves -- halves the serial chain of every role; three-deep rings replace
// the tail copies; the taps moved from shared memory into registers.  0.60 -> 0.50 ms at 24 MP, K = 15; the FMA pipe
// (48 % busy) and instruction issue (51 %) are now contended by six warps per sub-partition rather than idle behind one.
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
  static_assert(K >= 5 && K <= 17, "two chained circular convolutions on 128-sample segments");
  static constexpr int P = K / 2;
  static constexpr int V = FFT_N - 2 * (K - 1);             // valid g columns per segment
  static constexpr int DX = (4 - ((K - 1) & 3)) & 3;        // FFT sample n sits at TMA box column n + DX (16-byte aligned box start)
  static constexpr int S = 16;                              // packed rows per step (8 threads per row: 128 threads per FFT role)
  static constexpr int INW = 136;                           // TMA box width (floats): rows start 8 banks apart
  // roles in this order of thread ids: E, G, IFFT, FFT one warp group each, EPI two
  static constexpr int ROLE = 128;
  static constexpr int T_E = 0, T_G = ROLE, T_IFFT = 2 * ROLE, T_FFT = 3 * ROLE, T_EPI = 4 * ROLE;
  static constexpr int THREADS = 6 * ROLE;
  // registers per thread of every role (setmaxnreg; the kernel is launched with 80): 112 + 96 + 72 + 72 + 2 * 64 = 480 = 6 * 80
  static constexpr int REG_E = 112, REG_G = 96, REG_FFT = 72, REG_EPI = 64;
  static constexpr int DEPTH = 4;                           // the last role runs DEPTH time steps behind the first
  static constexpr int RB = 8;                              // MAC roles: rows per register block
  static constexpr int NB = S / RB;
  static constexpr int PCACHE = 24;                         // pieces of this CTA kept in shared memory
  static constexpr int T2 = 2 * P;                          // window rows a step needs from the previous ring block
  static constexpr int RING = 3;                            // blocks per ring: written / read with its predecessor / free
  static constexpr int U_BYTES = RING * S * FFT_PITCH * 8;  // U^ ring (forward FFT role writes, E reads)
  static constexpr int EC_BYTES = RING * S * FFT_N * 8;     // Err^ ring (E writes, G reads: thread = bin on both sides)
  static constexpr int GB_BYTES = RING * S * FFT_PITCH * 8; // G^ ring (G writes, IFFT transforms in place, EPI reads)
  static constexpr int IN_STAGE = 2 * S * INW * 4;          // one TMA stage: S real rows of each half
  static constexpr int IN_BYTES = 2 * IN_STAGE;
  static constexpr int TW_BYTES = FFT_TW_BYTES;
  static constexpr int OFF_GB = U_BYTES, OFF_EC = OFF_GB + GB_BYTES, OFF_IN = OFF_EC + EC_BYTES, OFF_TW = OFF_IN + IN_BYTES,
                       OFF_BAR = OFF_TW + TW_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 64 + 128;
  static_assert(V % 4 == 0 && (K - 1 + DX) % 4 == 0, "float4 epilogue / aligned TMA box");
  static_assert(OFF_IN % 128 == 0 && IN_STAGE % 128 == 0, "TMA destination alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
  static_assert(T2 <= S, "the window rows come from one ring block");
  static_assert(S % RB == 0 && ROLE == 8 * S && ROLE == FFT_N, "role geometry: 8 threads per row, one MAC thread per bin");
  static_assert(REG_E + REG_G + 2 * REG_FFT + 2 * REG_EPI == 6 * 80, "setmaxnreg budget = registers of the launch");
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
// Register budget of the executing warp group from here on (sm_90+: setmaxnreg; ptxas allocates the code that follows with it)
template <int N>
__device__ __forceinline__ void set_maxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void set_maxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// MAC over one block of R rows with a ROLLING register window:  acc[y] = sum_ky w[ky] * z[y + ky], where z[0 .. 2P) are
// the last 2P window rows of the previous block (or of the previous ring block) and z[2P .. 2P + R) the block's own
// rows, loaded here.  The caller moves the window down by R rows afterwards.  The loop over the two blocks of a step is
// NOT unrolled: with 24 warps in five different code regions per SM sub-partition the instruction caches are the scarce
// resource (unrolled: every role slowed down, 0.54 -> 0.69 ms; straight-line code of a whole step made round 2's first
// version instruction-fetch bound as well).  The taps sit in registers for a whole channel (conjugated there for the bins
// above 64: real taps), which removes one shared-memory load and one select per tap and block from the loop.
template <int K, int R, int CPITCH>
__device__ __forceinline__ void chain_mac_block(float2 (&z)[R + K - 1], const float2* __restrict__ cur, const float2 (&w)[K],
                                                float2 (&acc)[R]) {
  using C = ChainCfg<K>;
#pragma unroll
  for (int m = 0; m < R; ++m) z[C::T2 + m] = cur[m * CPITCH];
#pragma unroll
  for (int y = 0; y < R; ++y) acc[y] = make_float2(0.f, 0.f);
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
#pragma unroll
    for (int y = 0; y < R; ++y) acc[y] = cfma(z[y + ky], w[ky], acc[y]);
  }
}

// Debug: bit mask of roles whose work is skipped (1 forward FFT, 2 E, 4 inverse FFT, 8 TMA loads, 16 G, 32 epilogue);
// results are garbage then -- used only to time the roles in isolation (rltv_debug_chain_roles).
__device__ int g_chain_skip_roles = 0;

struct ChainCursor {   // (piece, step) of one role; advanced once per time step
  int p, j;
};
struct ChainPieceCursor {   // ... with the piece itself in registers (re-read only when the piece changes)
  int p, j;
  ChainPiece pc;
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
  float2* GB = reinterpret_cast<float2*>(smem + C::OFF_GB);
  float2* EC = reinterpret_cast<float2*>(smem + C::OFF_EC);
  float2* tw = reinterpret_cast<float2*>(smem + C::OFF_TW);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);   // one mbarrier per TMA stage
  const int tid = threadIdx.x, lane = tid & 31;
  const int role = tid / C::ROLE, rt = tid - role * C::ROLE;         // warp-uniform role, thread inside the role
  const int p0 = cta_first[blockIdx.x], p1 = cta_first[blockIdx.x + 1];
  if (p0 >= p1) {
    if (cp.nranks > 1) band_step_max_tail_slot(st, slot, cp, seq, done_counter, reinterpret_cast<int*>(smem));
    return;
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_u);
  }
  fft_fill_twiddles(tw);

  // this CTA's pieces in shared memory: with 220 KB of dynamic shared memory there is next to no L1 left, and every
  // role reads its current piece every step (ncu: those L2 round trips were the top long-scoreboard stall)
  __shared__ ChainPiece spc[C::PCACHE];
  for (int i = tid; i < C::PCACHE && p0 + i < p1; i += C::THREADS) spc[i] = pieces_g[p0 + i];
  __syncthreads();
  auto piece_nsteps = [&](int p) { return (p - p0 < C::PCACHE) ? spc[p - p0].nsteps : pieces_g[p].nsteps; };
  auto piece = [&](int p) -> ChainPiece { return (p - p0 < C::PCACHE) ? spc[p - p0] : pieces_g[p]; };
  int total_steps = 0;
  for (int p = p0; p < p1; ++p) total_steps += piece_nsteps(p);
  const int imgLo = C::P - g.row0, imgHi = C::P + g.M - g.row0;       // local rows on which the residual exists

  // TMA of the u rows of (piece, step): real rows [ya - 2P + jS, +S) and the same of the second half
  auto issue = [&](const ChainPiece& pc, int j, int stage) {
    const int xb = C::V * pc.s - (K - 1) - C::DX;
    const int y = pc.ya - 2 * C::P + j * C::S;
    unsigned char* dst = smem + C::OFF_IN + stage * C::IN_STAGE;
    mbar_arrive_expect_tx(bar + stage, C::IN_STAGE);
    tma_load_3d(dst, &tm_u, xb, y, pc.c, bar + stage);
    tma_load_3d(dst + C::S * C::INW * 4, &tm_u, xb, y + pc.L, pc.c, bar + stage);
  };
  auto advance = [&](ChainCursor& cur) {
    if (++cur.j >= piece_nsteps(cur.p)) { cur.j = 0; ++cur.p; }
  };
  // the MAC and epilogue roles keep their piece in registers: per step that is one compare instead of a table lookup
  // (ncu: ~100 of a MAC role's 680-770 instructions per step were this bookkeeping)
  auto advance_pc = [&](ChainPieceCursor& cur) {
    if (++cur.j >= cur.pc.nsteps) {
      cur.j = 0;
      ++cur.p;
      if (cur.p < p1) cur.pc = piece(cur.p);
    }
  };
  // a step needs the slow (masking) path if its residual rows touch rows outside the image or the segment is a border one
  auto needs_fix = [&](const ChainPiece& pc, int j) {
    const int ea = pc.ya + j * C::S - 3 * C::P, eb = ea + pc.L;
    const bool rowfix = (ea < imgLo) || (ea + C::S > imgHi) || (eb < imgLo) || (eb + C::S > imgHi);
    return pc.xfix || rowfix;
  };

  const int skip = g_chain_skip_roles;
  const bool fix_on = !(skip & (1 | 16));
  const int total = total_steps, nt = total_steps + C::DEPTH;
  constexpr int RB = C::RB;
  // Every role runs its own copy of the time loop (the register budgets differ: setmaxnreg applies to the code that
  // follows it) and meets the others at the end of every time step.
  // (a NAMED barrier: the roles arrive from different places of the code, which barrier 0 = __syncthreads() must not see)
  auto step_barrier = [&]() { named_bar_sync(5, C::THREADS); };

  if (role == 3) {
    // ================= FFT: masking path of role G's step t - 2, then the forward FFT of step t =================
    set_maxnreg_dec<C::REG_FFT>();
    const int zr = rt >> 3, tt = rt & 7;                  // row of the step, thread of the row
    ChainPieceCursor cfix{p0, 0, piece(p0)};              // role G's cursor
    ChainCursor cload{p0, 0};                             // first thread: the TMA load cursor, two steps ahead
    int fp = p0, fj = 0, f_ipk, f_ns;                     // this role's own (piece, step): only the image-spectra row is needed
    { const ChainPiece q = piece(p0); f_ipk = q.ipk_row0; f_ns = q.nsteps; }
    int rbu = 0, rbf = 0;                                 // ring blocks of step t (U^) and of step t - 2 (Err^ / G^)
    if (rt == 0 && !(skip & 8)) {
      issue(piece(p0), 0, 0);
      advance(cload);
      if (total > 1) issue(piece(cload.p), cload.j, 1);
      advance(cload);               // = step 2 (possibly past the end: only dereferenced while t + 2 < total)
    }
    for (int t = 0; t < nt; ++t) {
      if (t >= 2 && t - 2 < total) {
        const ChainPiece& pc = cfix.pc;
        if (fix_on && needs_fix(pc, cfix.j)) {
          float2* scratch = GB + (rbf * C::S + zr) * FFT_PITCH;      // G^ block of that step: not written yet
          float2* erow = EC + (rbf * C::S + zr) * FFT_N;
          fft128_core<true, true>(scratch, tw, tt, [&](int j) { return erow[tt + 8 * j]; }, 0xffffffffu, 0);
          __syncwarp();
          const int ea = pc.ya + cfix.j * C::S - 3 * C::P + zr, eb = ea + pc.L;
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
          fft128_core<false, true>(scratch, tw, tt, [&](int j) { return scratch[tt + 8 * j]; }, 0xffffffffu, 0);
          __syncwarp();
          for (int n = tt; n < FFT_N; n += 8) erow[n] = scratch[n];
          __threadfence_block();
          named_bar_arrive(2, 2 * C::ROLE);                          // role G may read the masked Err^
        }
        advance_pc(cfix);
        rbf = (rbf == C::RING - 1) ? 0 : rbf + 1;
      }
      if (t < total) {
        if (!(skip & 1)) {
          // role E reads the image spectra of this step during the NEXT time step: bring them into L2 now
          prefetch_l2(Ipk + (size_t(f_ipk) + size_t(fj) * C::S) * FFT_N + rt * 16);
          if (!(skip & 8)) mbar_wait(bar + (t & 1), (t >> 1) & 1);
          const float* ra = reinterpret_cast<const float*>(smem + C::OFF_IN + (t & 1) * C::IN_STAGE) + zr * C::INW + C::DX + tt;
          float2* dst = U + (rbu * C::S + zr) * FFT_PITCH;
          fft128_core<false, true>(dst, tw, tt, [&](int j) { return make_float2(ra[8 * j], ra[C::S * C::INW + 8 * j]); }, 0xffffffffu, 0);
        }
        if (++fj >= f_ns) {
          fj = 0;
          if (++fp < p1) { const ChainPiece q = piece(fp); f_ipk = q.ipk_row0; f_ns = q.nsteps; }
        }
        rbu = (rbu == C::RING - 1) ? 0 : rbu + 1;
      }
      step_barrier();
      // the TMA stage t & 1 was consumed by the forward FFT of step t: load step t + 2 into it
      if (rt == 0 && t + 2 < total && !(skip & 8)) {
        issue(piece(cload.p), cload.j, t & 1);
        advance(cload);
      }
    }
  } else if (role == 0) {
    // ================= E, step t - 1:  Err^ = 128 sum W0 U^ - I^  (thread = frequency bin) =================
    set_maxnreg_inc<C::REG_E>();
    const int mk = rt, kk = (mk <= 64) ? mk : FFT_N - mk;
    const float sgn = (mk > 64) ? -1.f : 1.f;
    int cur_wc = -1;                                      // channel whose tap spectra are in the registers
    float2 w[K];
#pragma unroll
    for (int ky = 0; ky < K; ++ky) w[ky] = make_float2(0.f, 0.f);
    float2 iv[RB];                                        // image spectra of the thread's next block (loaded one block ahead)
    // this role's (piece, step) with the three piece fields it needs in registers; ring blocks of the step and its predecessor
    // (measured and dropped: carrying the window in registers from step to step instead of re-reading its first 2P rows
    //  -- no gain, 16 B of spills; the subtraction of I^ as a sixth pipeline stage run by the IFFT role, which idles 63 % of
    //  the time -- the extra pass over the Err^ block costs more shared-memory traffic than it takes off this role,
    //  0.454 -> 0.498 ms; the FFT twiddles of the IFFT role in registers for the whole kernel -- no change)
    int ep = p0, ej = 0, e_c, e_ipk, e_ns, rb = 0, rp = C::RING - 1;
    {
      const ChainPiece q = piece(p0);
      e_c = q.c; e_ipk = q.ipk_row0; e_ns = q.nsteps;
      const float2* ip = Ipk + size_t(e_ipk) * FFT_N + mk;
#pragma unroll
      for (int y = 0; y < RB; ++y) iv[y] = __ldg(ip + size_t(y) * FFT_N);
    }
    for (int t = 0; t < nt; ++t) {
      if (t >= 1 && t - 1 < total) {
        if (!(skip & 2)) {
          if (e_c != cur_wc) {
            // forward tap spectra of this channel, scaled by 128 (Err^ must be the unnormalised spectrum, like I^)
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
              const float2 v = __ldg(wspec + (size_t(e_c) * K + ky) * FFT_N + kk);
              w[ky] = make_float2(v.x * float(FFT_N), v.y * (sgn * float(FFT_N)));
            }
            cur_wc = e_c;
          }
          const float2* ucur = U + rb * C::S * FFT_PITCH + mk;
          const float2* uprev = U + (rp * C::S + C::S - C::T2) * FFT_PITCH + mk;
          float2* ecur = EC + rb * C::S * FFT_N + mk;
          const float2* ip = Ipk + (size_t(e_ipk) + size_t(ej) * C::S) * FFT_N + mk;
          float2 z[RB + K - 1];
#pragma unroll
          for (int m = 0; m < C::T2; ++m) z[m] = uprev[m * FFT_PITCH];
#pragma unroll 1
          for (int B = 0; B < C::NB; ++B) {
            float2 acc[RB];
            chain_mac_block<K, RB, FFT_PITCH>(z, ucur + RB * B * FFT_PITCH, w, acc);
#pragma unroll
            for (int y = 0; y < RB; ++y) ecur[(RB * B + y) * FFT_N] = make_float2(acc[y].x - iv[y].x, acc[y].y - iv[y].y);
            if (B + 1 < C::NB) {                           // image spectra of the next block: L2 hits, one block of MACs to land
#pragma unroll
              for (int y = 0; y < RB; ++y) iv[y] = __ldg(ip + size_t(RB * (B + 1) + y) * FFT_N);
            }
#pragma unroll
            for (int m = 0; m < C::T2; ++m) z[m] = z[m + RB];
          }
        }
        // next step
        if (++ej >= e_ns) {
          ej = 0;
          if (++ep < p1) { const ChainPiece q = piece(ep); e_c = q.c; e_ipk = q.ipk_row0; e_ns = q.nsteps; }
        }
        rp = rb;
        rb = (rb == C::RING - 1) ? 0 : rb + 1;
        // image spectra of this thread's first block of that step, issued before the barrier (measured against issuing
        // them at the start of the step: 0.455 vs 0.460 ms)
        if (t < total && !(skip & 2)) {
          const float2* ipn = Ipk + (size_t(e_ipk) + size_t(ej) * C::S) * FFT_N + mk;
#pragma unroll
          for (int y = 0; y < RB; ++y) iv[y] = __ldg(ipn + size_t(y) * FFT_N);
        }
      }
      step_barrier();
    }
  } else if (role == 1) {
    // ================= G, step t - 2:  G^ = sum W1 Err^  (thread = frequency bin) =================
    set_maxnreg_inc<C::REG_G>();
    const int mk = rt, kk = (mk <= 64) ? mk : FFT_N - mk;
    const float sgn = (mk > 64) ? -1.f : 1.f;
    int cur_wc = -1;
    float2 w[K];
#pragma unroll
    for (int ky = 0; ky < K; ++ky) w[ky] = make_float2(0.f, 0.f);
    ChainPieceCursor cg{p0, 0, piece(p0)};
    int rb = 0, rp = C::RING - 1;                         // ring blocks of the step and its predecessor
    for (int t = 0; t < nt; ++t) {
      if (t >= 2 && t - 2 < total) {
        if (!(skip & 16)) {
          const ChainPiece& pc = cg.pc;
          if (pc.c != cur_wc) {
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
              const float2 v = __ldg(wspec + ((size_t(3) + pc.c) * K + ky) * FFT_N + kk);
              w[ky] = make_float2(v.x, v.y * sgn);
            }
            cur_wc = pc.c;
          }
          if (fix_on && needs_fix(pc, cg.j)) named_bar_sync(2, 2 * C::ROLE);    // the FFT role has masked this step's Err^
          const float2* ecur = EC + rb * C::S * FFT_N + mk;
          const float2* eprev = EC + (rp * C::S + C::S - C::T2) * FFT_N + mk;
          float2* gb = GB + rb * C::S * FFT_PITCH + mk;
          float2 z[RB + K - 1];
#pragma unroll
          for (int m = 0; m < C::T2; ++m) z[m] = eprev[m * FFT_N];
#pragma unroll 1
          for (int B = 0; B < C::NB; ++B) {
            float2 acc[RB];
            chain_mac_block<K, RB, FFT_N>(z, ecur + RB * B * FFT_N, w, acc);
#pragma unroll
            for (int y = 0; y < RB; ++y) gb[(RB * B + y) * FFT_PITCH] = acc[y];
#pragma unroll
            for (int m = 0; m < C::T2; ++m) z[m] = z[m + RB];
          }
        }
        advance_pc(cg);
        rp = rb;
        rb = (rb == C::RING - 1) ? 0 : rb + 1;
      }
      step_barrier();
    }
  } else if (role == 2) {
    // ================= inverse FFT of the g rows of step t - 3, in place =================
    set_maxnreg_dec<C::REG_FFT>();
    const int zr = rt >> 3, tt = rt & 7;
    int rb = 0;
    for (int t = 0; t < nt; ++t) {
      if (t >= 3 && t - 3 < total) {
        if (!(skip & 4)) {
          float2* row = GB + (rb * C::S + zr) * FFT_PITCH;
          fft128_core<true, true>(row, tw, tt, [&](int j) { return row[tt + 8 * j]; }, 0xffffffffu, 0);
        }
        rb = (rb == C::RING - 1) ? 0 : rb + 1;
      }
      step_barrier();
    }
  } else {
    // ================= epilogue of step t - 4: g store, max(u_c), max|lambda g + (u - ut)/2|  (16 threads per row pair) ======
    set_maxnreg_dec<C::REG_EPI>();
    const int re = tid - C::T_EPI, zr = re >> 4, tt = re & 15;
    constexpr int NQ = (C::V / 4 + 15) / 16;              // float4 columns per thread
    float mu = -INFINITY, mG = 0.f;                       // statistics of the current piece's channel
    // one pixel of a g row: the statistic max|lambda g + (u - ut)/2| (pyx:519) and max(u) (pyx:524)
    auto px = [&](float o, float uu, float tv) {
      mG = fmaxf(mG, fabsf(fmaf(lambd, o, 0.5f * (uu - tv))));
      mu = fmaxf(mu, uu);
    };
    ChainPieceCursor ce{p0, 0, piece(p0)};
    int rb = 0;                                           // ring block of the step
    for (int t = 0; t < nt; ++t) {
      if (t >= 4 && t - 4 < total) {
        const int pc_c = ce.pc.c;
        const bool last_of_piece = ce.j == ce.pc.nsteps - 1;
        if (!(skip & 32)) {
          const ChainPiece& pc = ce.pc;
          const int rel = ce.j * C::S - 4 * C::P + zr;                 // g row of the first half, relative to the piece
          const bool va = rel >= 0 && rel < pc.L, vb = rel >= 0 && rel < pc.Lb;
          const int xs = C::V * pc.s;
          const size_t offa = size_t(pc.c) * g.plane + size_t(va ? pc.ya + rel : 0) * g.pitch + xs;
          const size_t offb = size_t(pc.c) * g.plane + size_t(vb ? pc.ya + rel + pc.L : 0) * g.pitch + xs;
          const float* pua = ug + offa;
          const float* pub = ug + offb;
          const float* pta = (ut_is_u ? ug : utg) + offa;
          const float* ptb = (ut_is_u ? ug : utg) + offb;
          // all operand loads first (clamped addresses instead of branches: with a branch per load the compiler
          // serialises the round trips); the rows were brought into L2 one time step ago
          float4 ua[NQ], ta[NQ], ub[NQ], tb[NQ];
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int x4 = 4 * (tt + 16 * q);
            const int xc = (x4 < C::V && xs + x4 < g.pitch) ? x4 : 0;
            ua[q] = __ldg(reinterpret_cast<const float4*>(pua + xc));
            ub[q] = __ldg(reinterpret_cast<const float4*>(pub + xc));
            ta[q] = __ldg(reinterpret_cast<const float4*>(pta + xc));
            tb[q] = __ldg(reinterpret_cast<const float4*>(ptb + xc));
          }
          // the rows of the NEXT step into L2 (a whole time step ahead of their use): one line address per thread --
          // (array, half, line) -- the 16-byte aligned row of V floats touches four lines of 128 B, or a fifth
          advance_pc(ce);                                              // (everything of the current step that depends on its piece is in registers by now)
          if (t - 3 < total) {
            const ChainPiece& pn = ce.pc;
            const int reln = ce.j * C::S - 4 * C::P + zr;
            const int arr = tt & 1, part = (tt >> 1) & 1, ln = tt >> 2;
            const bool v = reln >= 0 && reln < (part ? pn.Lb : pn.L);
            const int xn = C::V * pn.s;
            if (v && !(arr && ut_is_u)) {
              const float* rowp = (arr ? utg : ug) + size_t(pn.c) * g.plane + size_t(pn.ya + reln + (part ? pn.L : 0)) * g.pitch + xn;
              if (xn + 32 * ln < g.pitch) prefetch_l2(rowp + 32 * ln);
              if (ln == 0 && xn + C::V - 4 < g.pitch) prefetch_l2(rowp + C::V - 4);
            }
          }
          const float2* row = GB + (rb * C::S + zr) * FFT_PITCH + (K - 1);
          const bool full = xs + C::V <= g.Wu;                         // the segment lies inside the row: no column checks
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int x4 = 4 * (tt + 16 * q);
            if (x4 >= C::V || xs + x4 >= g.pitch) continue;
            const float4 z01 = lds128(row + x4);                       // (re0, im0, re1, im1)
            const float4 z23 = lds128(row + x4 + 2);
            if (full) {
              if (va) {
                px(z01.x, ua[q].x, ta[q].x); px(z01.z, ua[q].y, ta[q].y); px(z23.x, ua[q].z, ta[q].z); px(z23.z, ua[q].w, ta[q].w);
                *reinterpret_cast<float4*>(gout + offa + x4) = make_float4(z01.x, z01.z, z23.x, z23.z);
              }
              if (vb) {
                px(z01.y, ub[q].x, tb[q].x); px(z01.w, ub[q].y, tb[q].y); px(z23.y, ub[q].z, tb[q].z); px(z23.w, ub[q].w, tb[q].w);
                *reinterpret_cast<float4*>(gout + offb + x4) = make_float4(z01.y, z01.w, z23.y, z23.w);
              }
            } else {
              const int nv = g.Wu - (xs + x4);                         // columns of this float4 inside the row (may be <= 0)
              const bool c0 = nv > 0, c1 = nv > 1, c2 = nv > 2, c3 = nv > 3;
              if (va) {
                const float4 o = make_float4(c0 ? z01.x : 0.f, c1 ? z01.z : 0.f, c2 ? z23.x : 0.f, c3 ? z23.z : 0.f);
                if (c0) px(o.x, ua[q].x, ta[q].x);
                if (c1) px(o.y, ua[q].y, ta[q].y);
                if (c2) px(o.z, ua[q].z, ta[q].z);
                if (c3) px(o.w, ua[q].w, ta[q].w);
                *reinterpret_cast<float4*>(gout + offa + x4) = o;
              }
              if (vb) {
                const float4 o = make_float4(c0 ? z01.y : 0.f, c1 ? z01.w : 0.f, c2 ? z23.y : 0.f, c3 ? z23.w : 0.f);
                if (c0) px(o.x, ub[q].x, tb[q].x);
                if (c1) px(o.y, ub[q].y, tb[q].y);
                if (c2) px(o.z, ub[q].z, tb[q].z);
                if (c3) px(o.w, ub[q].w, tb[q].w);
                *reinterpret_cast<float4*>(gout + offb + x4) = o;
              }
            }
          }
          // end of a piece: flush the statistics of its channel
          if (last_of_piece) {
            const float wu = warp_max(mu), wG = warp_max(mG);
            if (lane == 0) {
              atomicMax(&st->smax[slot][pc_c], f2ord(wu));
              atomicMax(&st->smax[slot][3 + pc_c], f2ord(wG));
            }
            mu = -INFINITY;
            mG = 0.f;
          }
        } else {
          advance_pc(ce);
        }
        rb = (rb == C::RING - 1) ? 0 : rb + 1;
      }
      step_barrier();
    }
  }
  // back to the launch's register count before the common tail (it may call into the row-band exchange)
  if (role <= 1) set_maxnreg_dec<80>(); else set_maxnreg_inc<80>();
  if (cp.nranks > 1) band_step_max_tail_slot(st, slot, cp, seq, done_counter, reinterpret_cast<int*>(smem));
}

}  // namespace rltv
