// Batched 128-point complex FFT over rows held in shared memory (fp32), the engine of the FLOP-reducing
// "row-FFT hybrid" stencils (rltv_stencil_fft.cuh).
//
// Each row is transformed by 8 threads of one warp, 16 points per thread, in two register passes:
//   pass A  thread t loads x[t + 8j] (j = 0..15), does a 16-point DFT over j, multiplies by w128^(t*q) and stores
//           the result at exchange position 9q + t;
//   pass B  thread u loads, for q in {u, u+8}, the 8 values t = 0..7, does an 8-point DFT over t and stores
//           X[q + 16p] in natural order.
// (Cooley-Tukey with n = t + 8j, k = q + 16p.)  One shared-memory round trip per pass: 32 B of traffic per point.
// The exchange pitch of 9 and a row pitch of FFT_PITCH = 152 complex (== 16 banks mod 32) make every LDS.64 /
// STS.64 of a half-warp conflict-free.  Only __syncwarp() is needed inside: a row never leaves its 8 threads.
//
// Arithmetic: every complex value lives in one 64-bit register pair and every operation is a packed sm_100a
// instruction (FADD2 / FMUL2 / FFMA2 through __fadd2_rn / __fmul2_rn / __ffma2_rn).  The hardware operand modifiers
// do the rest for free: a half swap (.LO_HI), a per-half sign (.NP / .PN) and a scalar broadcast (.F32) -- so a complex
// add is ONE instruction, a multiplication by +-i costs nothing (it folds into the consumer), and a complex
// multiplication is TWO:   x * w = FFMA2(x, w.re, FMUL2((-x.im, x.re), w.im)).   Half the issue slots of the scalar
// form for the same FMA-pipe work; the kernels around this engine are issue-bound, not pipe-bound.
#pragma once
#include <cuda_runtime.h>

namespace rltv {

constexpr int FFT_N = 128;
constexpr int FFT_PITCH = 152;   // complex elements per row buffer (>= 9*16 = 144 exchange slots)
// Pass-A twiddle table: tw[q][t'] = exp(-2 pi i t' q / 128), q = 0..15, t' = 0..31.  For a fixed q the 32 lanes of a
// warp read 32 consecutive entries (conflict-free; the plain 128-entry table was read with stride q: up to 8-way
// conflicts), and the index is a compile-time offset from a per-thread base.
constexpr int FFT_TW_ENTRIES = 16 * 32;
constexpr int FFT_TW_BYTES = FFT_TW_ENTRIES * 8;

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a * w, two packed instructions
__device__ __forceinline__ float2 cmulf(float2 a, float2 w) {
  return __ffma2_rn(a, make_float2(w.x, w.x), __fmul2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y)));
}
// acc + a * w (complex), two packed instructions
__device__ __forceinline__ float2 cfma(float2 a, float2 w, float2 acc) {
  return __ffma2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y), __ffma2_rn(a, make_float2(w.x, w.x), acc));
}
// acc + conj(e) * z, two packed instructions
__device__ __forceinline__ float2 cfma_conj(float2 e, float2 z, float2 acc) {
  return __ffma2_rn(make_float2(z.y, -z.x), make_float2(e.y, e.y), __ffma2_rn(z, make_float2(e.x, e.x), acc));
}
// multiply by -i (forward) or +i (inverse): folds into the consumer's operand modifiers
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// 4-point DFT, in place: X[k] = sum_n a[n] w4^(nk), w4 = exp(-+ 2 pi i / 4)
template <bool INV>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mi<INV>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, t3);
  a3 = csub(t1, t3);
}

// w^m for w = exp(-2 pi i / 16) (forward) / its conjugate (inverse), m = 0..9 as compile-time constants
template <bool INV>
__device__ __forceinline__ float2 w16(int m) {
  const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, SQ = 0.70710678118654752f;
  float2 w;
  switch (m) {
    case 0: w = make_float2(1.f, 0.f); break;
    case 1: w = make_float2(C1, -S1); break;
    case 2: w = make_float2(SQ, -SQ); break;
    case 3: w = make_float2(S1, -C1); break;
    case 4: w = make_float2(0.f, -1.f); break;
    case 6: w = make_float2(-SQ, -SQ); break;
    case 9: w = make_float2(-C1, S1); break;
    default: w = make_float2(0.f, 0.f); break;
  }
  if (INV) w.y = -w.y;
  return w;
}

// 16-point DFT in registers (4 x 4): input x[n], output X[k] in natural order, in place.
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&x)[16]) {
  // n = a + 4b, k = c + 4d:  Y[a][c] = DFT4_b(x[a+4b]) ; Y[a][c] *= w16^(a c) ; X[c+4d] = DFT4_a(Y[a][c])
#pragma unroll
  for (int a = 0; a < 4; ++a) dft4<INV>(x[a], x[a + 4], x[a + 8], x[a + 12]);   // result c stored at x[a + 4c]
#pragma unroll
  for (int a = 1; a < 4; ++a)
#pragma unroll
    for (int c = 1; c < 4; ++c) {
      if (a * c == 4) x[a + 4 * c] = mul_mi<INV>(x[a + 4 * c]);                 // w16^4 = -+i: free
      else x[a + 4 * c] = cmulf(x[a + 4 * c], w16<INV>(a * c));
    }
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4<INV>(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);   // result d at x[4c + d]
  // now x[4c + d] = X[c + 4d]: transpose the 4x4 index
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = c + 1; d < 4; ++d) {
      const float2 t = x[4 * c + d];
      x[4 * c + d] = x[4 * d + c];
      x[4 * d + c] = t;
    }
}

// 8-point DFT in registers (2 x 4), natural order in and out.
template <bool INV>
__device__ __forceinline__ void dft8(float2 (&x)[8]) {
  const float SQ = 0.70710678118654752f;
  // n = a + 2b, k = c + 4d:  Y[a][c] = DFT4_b(x[a+2b]) ; Y[1][c] *= w8^c ; X[c] = Y0+Y1 ; X[c+4] = Y0-Y1
  dft4<INV>(x[0], x[2], x[4], x[6]);   // Y[0][c] at x[2c]
  dft4<INV>(x[1], x[3], x[5], x[7]);   // Y[1][c] at x[2c+1]
  // w8^1 = (1 -+ i)/sqrt2, w8^2 = -+i, w8^3 = (-1 -+ i)/sqrt2
  x[3] = cmulf(x[3], make_float2(SQ, INV ? SQ : -SQ));
  x[5] = mul_mi<INV>(x[5]);
  x[7] = cmulf(x[7], make_float2(-SQ, INV ? SQ : -SQ));
  float2 y[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    y[c] = cadd(x[2 * c], x[2 * c + 1]);
    y[c + 4] = csub(x[2 * c], x[2 * c + 1]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = y[k];
}

// tw[q * 32 + t'] = exp(-2 pi i t' q / 128) in shared memory (filled once per CTA)
__device__ __forceinline__ void fft_fill_twiddles(float2* tw) {
  for (int m = threadIdx.x; m < FFT_TW_ENTRIES; m += blockDim.x) {
    const int q = m >> 5, tp = m & 31;
    float s, c;
    sincospif(-2.0f * float((tp * q) & (FFT_N - 1)) / float(FFT_N), &s, &c);
    tw[m] = make_float2(c, s);
  }
}

// Pass A + B of one row for the 8 threads (t = 0..7) that own it.  `load(j)` returns input element t + 8*((j + rot) & 15)
// -- the caller does the addressing, so that a dense source can use compile-time offsets.
// On return the row buffer `row` (FFT_PITCH complex) holds the spectrum / signal in natural order [0,128).
// `rot` (0..3) rotates the order in which the 16 stride-8 samples are fetched (x'[j] = x[(j + rot) mod 16]); by the
// shift theorem that only changes the pass-A twiddle index.  Callers whose four rows per warp would otherwise read
// the same banks of a dense 512-byte-pitch source use different `rot` per row; everybody else passes 0.
template <bool INV, bool BATCH = false, typename Load>
__device__ __forceinline__ void fft128_core(float2* __restrict__ row, const float2* __restrict__ tw, int t, Load load,
                                            unsigned mask, int rot) {
  float2 x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = load(j);
  dft16<INV>(x);                                 // X'[q] = X[q] * w16^(-rot q)   (forward; conjugate for inverse)
  // BATCH: all 15 twiddles are fetched in one batch BEFORE the warp barrier (the compiler cannot move a load across it):
  // fetched one by one after it, each multiplication waits for its own shared-memory round trip (ncu: short-scoreboard
  // stalls on every FMUL2 of this loop, ~17 % of an FFT role's time in the chain kernel).  Costs 30 registers across the
  // barrier, which the phase-structured tile kernels (many warps in the same phase, registers at their cap) do not have.
  const float2* twp = tw + (t + 8 * rot);        // w128^((t + 8 rot) q): w128^(t q) * w16^(+rot q) undoes the rotation
  float2 wq[BATCH ? 16 : 1];
  if constexpr (BATCH) {
#pragma unroll
    for (int q = 1; q < 16; ++q) wq[q] = twp[q * 32];
  }
  __syncwarp(mask);                              // everyone's loads are done before anyone stores (in-place rows)
  row[t] = x[0];
#pragma unroll
  for (int q = 1; q < 16; ++q) {
    float2 w = BATCH ? wq[q] : twp[q * 32];
    if (INV) w.y = -w.y;
    row[9 * q + t] = cmulf(x[q], w);
  }
  __syncwarp(mask);
  // pass B: this thread handles q = t and q = t + 8
  float2 a[8], b[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    a[s] = row[9 * t + s];
    b[s] = row[9 * (t + 8) + s];
  }
  dft8<INV>(a);
  dft8<INV>(b);
  __syncwarp(mask);
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    row[t + 16 * p] = a[p];
    row[t + 8 + 16 * p] = b[p];
  }
}

// Generic form: `load(n)` returns input element n in [0, 128).
template <bool INV, bool BATCH = false, typename Load>
__device__ __forceinline__ void fft128_row(float2* __restrict__ row, const float2* __restrict__ tw, int t, Load load,
                                           unsigned mask = 0xffffffffu, int rot = 0) {
  const int base = t + 8 * rot;
  fft128_core<INV, BATCH>(row, tw, t, [&](int j) { return load((base + 8 * j) & (FFT_N - 1)); }, mask, rot);
}

// Debug / test kernel: transforms `nrows` rows of 128 complex values (global, packed) one CTA per 16 rows.
template <bool INV>
__global__ void __launch_bounds__(128) k_fft128_debug(const float2* __restrict__ in, float2* __restrict__ out, int nrows) {
  __shared__ float2 buf[16 * FFT_PITCH];
  __shared__ float2 tw[FFT_TW_ENTRIES];
  fft_fill_twiddles(tw);
  __syncthreads();
  const int r = threadIdx.x >> 3, t = threadIdx.x & 7;
  const int row = blockIdx.x * 16 + r;
  if (row < nrows) {
    const unsigned mask = __activemask();
    const float2* src = in + size_t(row) * FFT_N;
    fft128_row<INV>(buf + r * FFT_PITCH, tw, t, [&](int n) { return src[n]; }, mask, r & 3);
    __syncwarp(mask);
    for (int n = t; n < FFT_N; n += 8) out[size_t(row) * FFT_N + n] = buf[r * FFT_PITCH + n];
  }
}

}  // namespace rltv
