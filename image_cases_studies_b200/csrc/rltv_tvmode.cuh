// TV-alive MM mode ("mm_tv", SURVEY.md 8(f3) mode (i)): the reference's regulariser with its two TV(ut, ...) calls alive
// (lib/deconvolution.pyx:464-465 un-commented and writing TV_ut_L1 / TV_ut_L2 -- the "patched reference" built by
// oracle/build_ref_tv.py; the shipped arithmetic never takes these branches).  Per outer iteration
//     TVut1, TVut2 = TV(ut, eps, order 2, norm 1 / 2)                                              pyx:464-465
// and per inner step, after the adjoint g = K^T (K u - image):
//     TVu1, TVu2, div = TV(u, eps, 2, 1), TV(u, eps, 2, 2)         (div of the second call survives)   pyx:495-496
//     T = div/TVu1/TVut1/2 + div/TVu2/TVut2/2   on every pixel off the border ring, 0 on it          pyx:517, :543
//     G = T + lambda g + (u - ut)/4             (ring: lambda g + (u - ut)/2)                        pyx:517, :519
//     u -= dt_c G,  dt_c = step max(u_c) / (max|G_c| + 1e-15)                                        pyx:524-531
//     image -= dti_c T / lambda,  dti_c = step max(image_c) / (max|T_c| + 1e-15)                     pyx:547-549
//     u_int = (1 - DoF) u_int + DoF image   (DoF from g and the image BEFORE its update)            pyx:499-502, :552
// Three kernels: k_tv_maps (once per outer iteration), k_tv_grad (stencil + T + the three reductions, one pass over
// u / ut / g / the two maps), k_update_tv (the update proper).  Whole-frame contexts only: the denoised image rows a
// band's neighbours would need are not exchanged.
#pragma once
#include "rltv_band.cuh"
#include "rltv_common.cuh"
#include "rltv_elementwise.cuh"

namespace rltv {

struct Tv2 {
  float tv1, tv2, div;   // TV(u, .., 2, 1) norm map, TV(u, .., 2, 2) norm map, divergence of the norm-2 call
};

// Order-2 stencil of pyx:159-189 at one pixel from its 3x3 neighbourhood a[row][col].
__device__ __forceinline__ Tv2 tv_order2(const float (&a)[3][3], float eps) {
  const float d = 1.41421356237309515f;                      // powf(2, 0.5), pyx:146
  const float adj1 = 4.f * (1.f + 1.f / d), adj2 = 2.f * (1.f + d);   // pyx:149-152
  const float cc = a[1][1];
  const float udx = -2.f * cc + a[0][1] + a[2][1];
  const float udy = -2.f * cc + a[1][0] + a[1][2];
  const float udxdy = (-2.f * cc + a[0][0] + a[2][2]) / d;
  const float udydx = (-2.f * cc + a[0][2] + a[2][0]) / d;
  Tv2 r;
  r.tv1 = ((fabsf(udx) + fabsf(udy) + eps) + (fabsf(udxdy) + fabsf(udydx) + eps)) / adj1;
  r.tv2 = (sqrtf(fmaf(udx, udx, fmaf(udy, udy, eps * eps))) + sqrtf(fmaf(udxdy, udxdy, fmaf(udydx, udydx, eps * eps)))) / adj2;
  r.div = (-udx - udy - udxdy - udydx) / adj2;
  return r;
}

// rows Y-1, Y, Y+1 and columns X-1 .. X+4 of one plane (zero outside the held rows / the frame width)
__device__ __forceinline__ void tv_load_rows(const Geom& g, const float* __restrict__ plane, int Y, int X, float (&a)[3][6]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int y = Y - 1 + r;
    if (y < 0 || y >= g.Hu) {
#pragma unroll
      for (int i = 0; i < 6; ++i) a[r][i] = 0.f;
      continue;
    }
    const float* row = plane + size_t(y) * g.pitch + X;
    const float4 m = *reinterpret_cast<const float4*>(row);
    a[r][0] = X > 0 ? __ldg(row - 1) : 0.f;
    a[r][1] = m.x; a[r][2] = m.y; a[r][3] = m.z; a[r][4] = m.w;
    a[r][5] = (X + 4 < g.pitch) ? __ldg(row + 4) : 0.f;
  }
}

// pixel (frame row gy, column x) is off the border ring of the u domain: the only place the reference's TV writes
__device__ __forceinline__ bool tv_inner(const Geom& g, int gy, int x) {
  return gy >= 1 && gy < g.M + g.K - 2 && x >= 1 && x < g.Wu - 1;
}

// TVut1 / TVut2 of the majoriser (= u at the start of an outer iteration), owned rows.
__global__ void __launch_bounds__(256)
k_tv_maps(Geom g, const State* __restrict__ st, const float* __restrict__ u, float eps, float* __restrict__ tv1, float* __restrict__ tv2) {
  if (st->stop) return;
  const int c = blockIdx.z, Y = g.own0 + blockIdx.y, X = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (X >= g.Wu) return;
  float a[3][6];
  tv_load_rows(g, u + size_t(c) * g.plane, Y, X, a);
  float o1[4], o2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float n[3][3] = {{a[0][i], a[0][i + 1], a[0][i + 2]}, {a[1][i], a[1][i + 1], a[1][i + 2]}, {a[2][i], a[2][i + 1], a[2][i + 2]}};
    const Tv2 t = tv_order2(n, eps);
    const bool in = tv_inner(g, g.row0 + Y, X + i);
    o1[i] = in ? t.tv1 : 0.f;
    o2[i] = in ? t.tv2 : 0.f;
  }
  const size_t off = size_t(c) * g.plane + size_t(Y) * g.pitch + X;
  *reinterpret_cast<float4*>(tv1 + off) = make_float4(o1[0], o1[1], o1[2], o1[3]);
  *reinterpret_cast<float4*>(tv2 + off) = make_float4(o2[0], o2[1], o2[2], o2[3]);
}

// T and the reductions max|G_c|, max|T_c|, max(image_c) into State::tvmax[slot]; `ut_is_u`: first inner step.
__global__ void __launch_bounds__(256)
k_tv_grad(Geom g, State* __restrict__ st, const float* __restrict__ u, const float* __restrict__ ut, const float* __restrict__ gbuf,
          const float* __restrict__ tvut1, const float* __restrict__ tvut2, const float* __restrict__ img, float lambd, float eps,
          float* __restrict__ tbuf, int slot, int ut_is_u) {
  if (st->stop) return;
  const int c = blockIdx.z, Y = g.own0 + blockIdx.y, X = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  float mG = 0.f, mT = 0.f, mI = -INFINITY;
  if (X < g.Wu) {
    float a[3][6];
    tv_load_rows(g, u + size_t(c) * g.plane, Y, X, a);
    const size_t off = size_t(c) * g.plane + size_t(Y) * g.pitch + X;
    const float4 tv4 = ut_is_u ? make_float4(a[1][1], a[1][2], a[1][3], a[1][4]) : *reinterpret_cast<const float4*>(ut + off);
    const float4 gv4 = *reinterpret_cast<const float4*>(gbuf + off);
    const float4 t1 = *reinterpret_cast<const float4*>(tvut1 + off);
    const float4 t2 = *reinterpret_cast<const float4*>(tvut2 + off);
    const float4 iv4 = *reinterpret_cast<const float4*>(img + off);
    const float tt[4] = {tv4.x, tv4.y, tv4.z, tv4.w}, gg[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
    const float p1[4] = {t1.x, t1.y, t1.z, t1.w}, p2[4] = {t2.x, t2.y, t2.z, t2.w}, ii[4] = {iv4.x, iv4.y, iv4.z, iv4.w};
    const int gy = g.row0 + Y;
    const bool rowin = (gy >= g.P) && (gy < g.P + g.M);
    float T[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float n[3][3] = {{a[0][i], a[0][i + 1], a[0][i + 2]}, {a[1][i], a[1][i + 1], a[1][i + 2]}, {a[2][i], a[2][i + 1], a[2][i + 2]}};
      const Tv2 t = tv_order2(n, eps);
      const float uu = a[1][i + 1];
      const bool in = tv_inner(g, gy, X + i);
      T[i] = in ? (t.div / t.tv1 / p1[i] / 2.f + t.div / t.tv2 / p2[i] / 2.f) : 0.f;                    // pyx:517 / :543
      const float G = in ? (T[i] + fmaf(lambd, gg[i], (uu - tt[i]) / 4.f)) : fmaf(lambd, gg[i], 0.5f * (uu - tt[i]));
      if (X + i < g.Wu) {
        mG = fmaxf(mG, fabsf(G));
        mT = fmaxf(mT, fabsf(T[i]));
        if (rowin && X + i >= g.P && X + i < g.P + g.N) mI = fmaxf(mI, ii[i]);
      }
    }
    *reinterpret_cast<float4*>(tbuf + off) = make_float4(T[0], T[1], T[2], T[3]);
  }
  __shared__ float red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mG = warp_max(mG); mT = warp_max(mT); mI = warp_max(mI);
  if (lane == 0) { red[0][warp] = mG; red[1][warp] = mT; red[2][warp] = mI; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float m = threadIdx.x == 2 ? -INFINITY : 0.f;
    for (int w = 0; w < 8; ++w) m = fmaxf(m, red[threadIdx.x][w]);
    atomicMax(&st->tvmax[slot][3 * threadIdx.x + c], f2ord(m));
  }
}

// PAM = true (mode "pam_ctv"): G = T + lambda g (no majoriser term), the image is left alone.
template <bool FIRST, bool PAM>
__global__ void __launch_bounds__(256)
k_update_tv(Geom g, State* __restrict__ st, float* __restrict__ u, const float* __restrict__ ut, float* __restrict__ ut_out,
            const float* __restrict__ gbuf, const float* __restrict__ tbuf, float* __restrict__ img, float step, float lambd,
            int blind, int sslot, int sreset, int tslot) {
  if (st->stop) return;
  const int c = blockIdx.z;
  const int Y = g.own0 + blockIdx.y;
  const int X = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  // sslot: statistics slot of max(u_c) (chain kernel: step parity, two-kernel gradient path: 0); tslot: TV statistics
  const float dt = step * ord2f(st->smax[sslot][c]) / (ord2f(st->tvmax[tslot][c]) + 1e-15f);                     // pyx:524
  const float dti = step * ord2f(st->tvmax[tslot][6 + c]) / (ord2f(st->tvmax[tslot][3 + c]) + 1e-15f);           // pyx:548
  if (X == 0 && blockIdx.y == 0) {
    st->dt[c] = dt;
    if (sreset >= 0) {
      st->smax[sreset][c] = ORD_LOWEST;
      st->smax[sreset][3 + c] = 0;
    }
    st->tvmax[tslot ^ 1][c] = 0;                       // the slot the NEXT inner step accumulates in
    st->tvmax[tslot ^ 1][3 + c] = 0;
    st->tvmax[tslot ^ 1][6 + c] = ORD_LOWEST;
  }
  if (X >= g.Wu) return;
  const size_t off = size_t(c) * g.plane + size_t(Y) * g.pitch + X;
  const float4 uv = *reinterpret_cast<const float4*>(u + off);
  const float4 tv = FIRST ? uv : *reinterpret_cast<const float4*>(ut + off);
  const float4 gv = *reinterpret_cast<const float4*>(gbuf + off);
  const float4 Tv = *reinterpret_cast<const float4*>(tbuf + off);
  const float4 iv = *reinterpret_cast<const float4*>(img + off);
  const float uu[4] = {uv.x, uv.y, uv.z, uv.w}, tt[4] = {tv.x, tv.y, tv.z, tv.w}, TT[4] = {Tv.x, Tv.y, Tv.z, Tv.w};
  const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, ii[4] = {iv.x, iv.y, iv.z, iv.w};
  const int gy = g.row0 + Y;
  const bool rowin = (gy >= g.P) && (gy < g.P + g.M);
  float o[4], io[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool in = tv_inner(g, gy, X + i);
    const float G = PAM ? fmaf(lambd, gg[i], TT[i])
                        : (in ? (TT[i] + fmaf(lambd, gg[i], (uu[i] - tt[i]) / 4.f)) : fmaf(lambd, gg[i], 0.5f * (uu[i] - tt[i])));
    float un = fmaf(-dt, G, uu[i]);
    io[i] = ii[i];
    if (rowin && (X + i) >= g.P && (X + i) < g.P + g.N) {
      const float d = (gg[i] - ii[i]) / (gg[i] + ii[i]);                 // DoF from the image BEFORE its update, pyx:499
      float dof = d * d;
      if (!blind) dof = dof / lambd;
      if (!PAM) io[i] = ii[i] - dti * TT[i] / lambd;                     // pyx:549
      un = (1.f - dof) * un + dof * io[i];                               // pyx:552
    }
    o[i] = ((X + i) < g.Wu) ? un : uu[i];
  }
  *reinterpret_cast<float4*>(u + off) = make_float4(o[0], o[1], o[2], o[3]);
  if (!PAM) *reinterpret_cast<float4*>(img + off) = make_float4(io[0], io[1], io[2], io[3]);
  if (FIRST) *reinterpret_cast<float4*>(ut_out + off) = uv;
}

// ------------------------------------------------------------------------------------------------------------------
// mode "pam_ctv" -- UNPINNED (no reference code exists: README.md:42-44, :113-117 only; definition in oracle/ctv_oracle.py)
// Forward-difference gradients of the three channels, the collaborative l^{inf,1,1} weight (per pixel and derivative
// direction only the channel with the largest |difference| carries the sub-gradient, floored at eps) and the divergence,
// in ONE pass over u:   T = -div p,  and the step statistic max|lambda g_c + T_c| of the PAM u-step.
// One thread = 4 adjacent pixels x 3 channels.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ctv_field(const float (&d)[3], float eps, float (&p)[3]) {
  const float a0 = fabsf(d[0]), a1 = fabsf(d[1]), a2 = fabsf(d[2]);
  const int cs = (a0 >= a1 && a0 >= a2) ? 0 : ((a1 >= a2) ? 1 : 2);          // first channel attaining the maximum
#pragma unroll
  for (int c = 0; c < 3; ++c) p[c] = (c == cs) ? d[c] / fmaxf(eps, fabsf(d[c])) : 0.f;
}

__global__ void __launch_bounds__(256)
k_ctv_grad(Geom g, State* __restrict__ st, const float* __restrict__ u, const float* __restrict__ gbuf, float lambd, float eps,
           float* __restrict__ tbuf, int slot) {
  if (st->stop) return;
  const int Y = g.own0 + blockIdx.y, X = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  float mG[3] = {0.f, 0.f, 0.f};
  if (X < g.Wu) {
    float a[3][3][6];                                   // [channel][row Y-1..Y+1][column X-1..X+4]
#pragma unroll
    for (int c = 0; c < 3; ++c) tv_load_rows(g, u + size_t(c) * g.plane, Y, X, a[c]);
    const int gy = g.row0 + Y, HuG = g.M + g.K - 1;
    float T[3][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = X + i;
      float dxc[3], dxl[3], dyc[3], dyu[3], pxc[3], pxl[3], pyc[3], pyu[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dxc[c] = (x + 1 < g.Wu) ? a[c][1][i + 2] - a[c][1][i + 1] : 0.f;          // D_x at (y, x)
        dxl[c] = (x >= 1 && x < g.Wu) ? a[c][1][i + 1] - a[c][1][i] : 0.f;         // D_x at (y, x-1)
        dyc[c] = (gy + 1 < HuG) ? a[c][2][i + 1] - a[c][1][i + 1] : 0.f;           // D_y at (y, x)
        dyu[c] = (gy >= 1) ? a[c][1][i + 1] - a[c][0][i + 1] : 0.f;                // D_y at (y-1, x)
      }
      ctv_field(dxc, eps, pxc);
      ctv_field(dxl, eps, pxl);
      ctv_field(dyc, eps, pyc);
      ctv_field(dyu, eps, pyu);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float px_l = (x >= 1) ? pxl[c] : 0.f, py_u = (gy >= 1) ? pyu[c] : 0.f;
        T[c][i] = -((pxc[c] - px_l) + (pyc[c] - py_u));
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t off = size_t(c) * g.plane + size_t(Y) * g.pitch + X;
      const float4 gv = *reinterpret_cast<const float4*>(gbuf + off);
      const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (X + i < g.Wu) mG[c] = fmaxf(mG[c], fabsf(fmaf(lambd, gg[i], T[c][i])));
      *reinterpret_cast<float4*>(tbuf + off) = make_float4(T[c][0], T[c][1], T[c][2], T[c][3]);
    }
  }
  __shared__ float red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float m = warp_max(mG[c]);
    if (lane == 0) red[c][warp] = m;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float m = 0.f;
    for (int w = 0; w < 8; ++w) m = fmaxf(m, red[threadIdx.x][w]);
    atomicMax(&st->tvmax[slot][threadIdx.x], f2ord(m));
  }
}

}  // namespace rltv
