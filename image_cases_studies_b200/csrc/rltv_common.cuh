// Shared device-side definitions for the RL/MM deconvolution kernels (sm_100a).
//
// Data layout in HBM ("u-geometry"): every full-frame array -- the estimate u, the majoriser ut, the
// adjoint gradient g, the blurry image and the residual err -- is stored PLANAR (one plane per RGB
// channel, because the PSF differs per channel, lib/deconvolution.pyx:478) with the SAME geometry:
// Hu = M + K - 1 rows of `pitch` floats (pitch % 4 == 0, pitch >= Wu = N + K - 1).  image and err live in
// the interior at offset (P, P), P = K/2, and their ring is kept at zero, which is exactly the
// "zero outside the image" boundary condition of the reference's 'full' convolution (pyx:491).  With
// that, the forward blur, the adjoint and the PSF-gradient correlation are all *centred* K x K stencils
// over one coordinate system, and every elementwise pass is float4-aligned across all its operands.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rltv {

struct Geom {
  int M, N;      // image rows / cols
  int K, P;      // PSF size, pad = K/2
  int Hu, Wu;    // M + K - 1, N + K - 1
  int pitch;     // floats per row of every plane
  size_t plane;  // floats per plane (Hu * pitch)
};

// Device-resident solver state: everything the host would otherwise have to read back between kernels.
struct State {
  int stop;                // stop_flag (pyx:417): once set every kernel returns immediately
  int it;                  // outer iterations executed (pyx:456,:656)
  int blind_steps;         // PSF steps taken (for the `correlation` caller-array quirk, pyx:581-585)
  float M_r, M_r_prev;     // pyx:624,:638
  unsigned max_u[3];       // order-preserving uint encoding of max(u_c)          (pyx:524)
  unsigned max_G[3];       //                                  max|gradu_c|       (pyx:524)
  float dt[3];             // last image step sizes
  float dtpsf;             // last PSF step size (pyx:574)
  double win_sum;          // whiteness window: sum, min, max of the residual (pyx:627-629)
  float win_min, win_max;
  double white_acc;        // sum of autocorr^2 * weights (pyx:634)
  int n_hist;
  float hist[4096];
};

// Monotone float -> uint map so that atomicMax on the encoding is a float max (any sign).
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rltv
