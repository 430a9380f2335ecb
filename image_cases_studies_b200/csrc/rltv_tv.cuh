// TV / divergence stencil of the reference: TV(u, out, M, N, epsilon, order, norm, div),
// lib/deconvolution.pyx:137-239.  3x3 neighbourhood, interior pixels only (pyx:159, :239), per channel.
// HBM-bound: reads u (12 B/px), writes `out` and `div` (24 B/px); one float4 of outputs per thread, the three
// input rows come through L1 (each u element is touched by 3 rows x 3 columns of neighbouring threads).
// The reference runs it twice per inner step and discards the results, so the pinned solver
// does not launch it; it is exposed through rltv_stage_tv and is the building block of the TV-alive modes.
#pragma once
#include "rltv_common.cuh"

namespace rltv {

template <int NORM>
__device__ __forceinline__ float tv_norm(float x, float y, float eps) {
  if (NORM == 1) return fabsf(x) + fabsf(y) + eps;        // pyx:133-134
  return sqrtf(fmaf(x, x, fmaf(y, y, eps * eps)));         // pyx:129-130
}

template <int ORDER, int NORM>
__global__ void __launch_bounds__(256)
k_tv(Geom g, const float* __restrict__ u, float eps, float* __restrict__ out, float* __restrict__ div) {
  const int c = blockIdx.z;
  const int Y = blockIdx.y;
  const int X = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (X >= g.Wu) return;
  const size_t off = size_t(c) * g.plane + size_t(Y) * g.pitch + X;
  float o[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
  if (Y >= 1 && Y < g.Hu - 1) {
    const float* r0 = u + off - g.pitch;
    const float* r1 = u + off;
    const float* r2 = u + off + g.pitch;
    float a[3][6];   // columns X-1 .. X+4 of rows Y-1, Y, Y+1
    const float* rows[3] = {r0, r1, r2};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float4 m = *reinterpret_cast<const float4*>(rows[r]);
      a[r][0] = X > 0 ? __ldg(rows[r] - 1) : 0.f;
      a[r][1] = m.x; a[r][2] = m.y; a[r][3] = m.z; a[r][4] = m.w;
      a[r][5] = (X + 4 < g.Wu) ? __ldg(rows[r] + 4) : 0.f;
    }
    const float d = 1.41421356237309515f;                                   // powf(2, 0.5), pyx:146
    const float adjust = (NORM == 1) ? 4.f * (1.f + 1.f / d) : 2.f * (1.f + d);   // pyx:149-152
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = X + i;
      if (x < 1 || x >= g.Wu - 1) continue;
      const float cc = a[1][i + 1];
      const float n = a[0][i + 1], s = a[2][i + 1], w = a[1][i], e = a[1][i + 2];
      const float nw = a[0][i], se = a[2][i + 2], ne = a[0][i + 2], sw = a[2][i];
      if (ORDER == 2) {                                                     // pyx:162-172
        const float udx = -2.f * cc + n + s;
        const float udy = -2.f * cc + w + e;
        const float udxdy = (-2.f * cc + nw + se) / d;
        const float udydx = (-2.f * cc + ne + sw) / d;
        dv[i] = (-udx - udy - udxdy - udydx) / adjust;
        o[i] = (tv_norm<NORM>(udx, udy, eps) + tv_norm<NORM>(udxdy, udydx, eps)) / adjust;
      } else {                                                              // pyx:197-213
        const float udx_b = cc - n, udy_b = cc - w;
        const float udx_f = -cc + s, udy_f = -cc + e;
        const float udxdy_b = (cc - nw) / d, udydx_b = (cc - ne) / d;
        const float udydx_f = (-cc + sw) / d, udxdy_f = (-cc + se) / d;
        dv[i] = (udx_b + udy_b - udx_f - udy_f + udxdy_b + udydx_b - udxdy_f - udydx_f) / adjust;
        o[i] = (tv_norm<NORM>(udx_b, udy_b, eps) + tv_norm<NORM>(udx_f, udy_f, eps) +
                tv_norm<NORM>(udxdy_b, udydx_b, eps) + tv_norm<NORM>(udxdy_f, udydx_f, eps)) / adjust;
      }
    }
  }
  *reinterpret_cast<float4*>(out + off) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(div + off) = make_float4(dv[0], dv[1], dv[2], dv[3]);
}

}  // namespace rltv
