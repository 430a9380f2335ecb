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
#pragma once
#include <cuda_runtime.h>

namespace rltv {

constexpr int FFT_N = 128;
constexpr int FFT_PITCH = 152;   // complex elements per row buffer (>= 9*16 = 144 exchange slots)

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by -i (forward) or +i (inverse)
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
    for (int c = 1; c < 4; ++c) x[a + 4 * c] = cmulf(x[a + 4 * c], w16<INV>(a * c));
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
  {
    const float2 v = x[3];
    x[3] = INV ? make_float2((v.x - v.y) * SQ, (v.x + v.y) * SQ) : make_float2((v.x + v.y) * SQ, (v.y - v.x) * SQ);
  }
  x[5] = mul_mi<INV>(x[5]);
  {
    const float2 v = x[7];
    x[7] = INV ? make_float2((-v.x - v.y) * SQ, (v.x - v.y) * SQ) : make_float2((v.y - v.x) * SQ, (-v.x - v.y) * SQ);
  }
  float2 y[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    y[c] = cadd(x[2 * c], x[2 * c + 1]);
    y[c + 4] = csub(x[2 * c], x[2 * c + 1]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = y[k];
}

// tw128[m] = exp(-2 pi i m / 128), m = 0..127, in shared memory (filled once per CTA by fft_fill_twiddles)
__device__ __forceinline__ void fft_fill_twiddles(float2* tw128) {
  for (int m = threadIdx.x; m < FFT_N; m += blockDim.x) {
    float s, c;
    sincospif(-2.0f * float(m) / float(FFT_N), &s, &c);
    tw128[m] = make_float2(c, s);
  }
}

// Pass A + B of one row for the 8 threads (t = 0..7) that own it.  `load(n)` returns input element n.
// On return the row buffer `row` (FFT_PITCH complex) holds the spectrum / signal in natural order [0,128).
// `rot` (0..15) rotates the order in which the 16 stride-8 samples are fetched (x'[j] = x[(j + rot) mod 16]); by the
// shift theorem that only changes the pass-A twiddle index.  The four rows a warp handles use different `rot`, so
// their loads from a dense 512-byte-pitch source (the TMA buffer) fall into different banks.
template <bool INV, typename Load>
__device__ __forceinline__ void fft128_row(float2* __restrict__ row, const float2* __restrict__ tw128, int t, Load load,
                                           unsigned mask = 0xffffffffu, int rot = 0) {
  float2 x[16];
  const int base = t + 8 * rot;                  // rot <= 3: only j >= 13 can wrap around
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    int n = base + 8 * j;
    if (j >= 13 && n >= FFT_N) n -= FFT_N;
    x[j] = load(n);
  }
  dft16<INV>(x);                                 // X'[q] = X[q] * w16^(-rot q)   (forward; conjugate for inverse)
  __syncwarp(mask);                              // everyone's loads are done before anyone stores (in-place rows)
  row[t] = x[0];
  int widx = 0;
#pragma unroll
  for (int q = 1; q < 16; ++q) {
    widx = (widx + base) & (FFT_N - 1);          // (t + 8 rot) q mod 128: w128^(t q) * w16^(+rot q) undoes the rotation
    float2 w = tw128[widx];
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

// Debug / test kernel: transforms `nrows` rows of 128 complex values (global, packed) one CTA per 16 rows.
template <bool INV>
__global__ void __launch_bounds__(128) k_fft128_debug(const float2* __restrict__ in, float2* __restrict__ out, int nrows) {
  __shared__ float2 buf[16 * FFT_PITCH];
  __shared__ float2 tw[FFT_N];
  fft_fill_twiddles(tw);
  __syncthreads();
  const int r = threadIdx.x >> 3, t = threadIdx.x & 7;
  const int row = blockIdx.x * 16 + r;
  if (row < nrows) {
    const unsigned mask = __activemask();
    const float2* src = in + size_t(row) * FFT_N;
    fft128_row<INV>(buf + r * FFT_PITCH, tw, t, [&](int n) { return src[n]; }, mask);
    __syncwarp(mask);
    for (int n = t; n < FFT_N; n += 8) out[size_t(row) * FFT_N + n] = buf[r * FFT_PITCH + n];
  }
}

}  // namespace rltv
