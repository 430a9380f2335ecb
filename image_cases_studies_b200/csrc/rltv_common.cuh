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
  int M, N;      // image rows / cols of the WHOLE frame
  int K, P;      // PSF size, pad = K/2
  int Hu, Wu;    // rows held by this context (whole frame: M + K - 1; row band: owned rows + halos), N + K - 1
  int pitch;     // floats per row of every plane
  size_t plane;  // floats per plane (Hu * pitch)
  // Row-band sharding (one band per GPU): local row Y is row (row0 + Y) of the frame's u-geometry; rows
  // [own0, own1) are owned (updated, counted in reductions), the rest are halo copies of the neighbours' rows.
  int row0, own0, own1;
};

// Device-resident solver state: everything the host would otherwise have to read back between kernels.
struct State {
  int stop;                // stop_flag (pyx:417): once set every kernel returns immediately
  int it;                  // outer iterations executed (pyx:456,:656)
  int blind_steps;         // PSF steps taken (for the `correlation` caller-array quirk, pyx:581-585)
  float M_r, M_r_prev;     // pyx:624,:638
  int smax[2][6];          // step statistics, order-preserving int encoding: [slot][0..2] = max(u_c), [slot][3..5] = max|gradu_c|
                           // (pyx:524; 6 contiguous int32 = one MAX all-reduce across row bands).  Slot 0 is THE slot of
                           // the two-kernel gradient path (reset by its forward kernel); the chain kernel alternates
                           // the slots by inner step and the update kernel resets the one the next step will use.
  int tvmax[2][9];         // TV-alive mode (rltv_tvmode.cuh): [slot][0..2] max|G_c|, [3..5] max|T_c|, [6..8] max(image_c)
  float dt[3];             // last image step sizes
  float dtpsf;             // last PSF step size (pyx:574)
  double win_sum;          // whiteness window: sum, min, max of the residual (pyx:627-629)
  float win_min, win_max;
  double white_acc;        // sum of autocorr^2 * weights (pyx:634)
  int n_hist;
  float hist[4096];
};

// Monotone float -> signed int map: atomicMax (and an int32 MAX all-reduce) on the encoding is a float max.
__device__ __forceinline__ int f2ord(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : (b ^ 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : (o ^ 0x7fffffff)); }
constexpr int ORD_LOWEST = int(0x80000000u);

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
