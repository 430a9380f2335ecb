// Residual-whiteness stop statistic M_r (Almeida & Figueiredo), lib/deconvolution.pyx:623-654, on device.
//
//   E   = error[top:bottom, left:right, :]                           (h x w x 3 window of the residual)
//   t   = (E - mean(E)) / std(E);  t /= amax|t|                      == (E - mean) / max|E - mean|
//   A_c = convolve(t_c, rot180(t_c), "same")                         per-channel autocorrelation
//   M_r = mean(A^2 * W),  W = sqrt(outer(gw, gw)) / sum              (pyx:393-404)
//
// The reference evaluates the autocorrelation with a float32 FFT inside scipy; here it is an L x L
// (L = 2^k >= 1.5 * max(h,w)) complex FFT in DOUBLE precision held in shared memory -- the statistic
// gates control flow (the stop test compares consecutive M_r values that differ by ~1e-5 relative), so
// it is kept at the accuracy of the float64 oracle.  Cost is negligible: one 3 x 512 x 512 transform per
// OUTER iteration (5 inner steps) for the reference's 255-pixel window.
#pragma once
#include "rltv_common.cuh"

namespace rltv {

struct WhiteGeom {
  int top, left, h, w;   // window in IMAGE coordinates
  int L, log2L;          // FFT size
};

// S1a: per (window row, channel) partial sum / min / max of the residual window. block = 128 threads.
// The last block to finish reduces the per-row partials in row order (deterministic) into the State.
__global__ void __launch_bounds__(128)
k_win_rows(Geom g, State* __restrict__ st, const float* __restrict__ err, WhiteGeom wg,
           double* __restrict__ rowsum, float* __restrict__ rowmin, float* __restrict__ rowmax,
           unsigned* __restrict__ done_counter) {
  if (st->stop) return;
  const int y = blockIdx.x, c = blockIdx.y;
  const float* row = err + size_t(c) * g.plane + size_t(wg.top + g.P - g.row0 + y) * g.pitch + (wg.left + g.P);
  double s = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (int x = threadIdx.x; x < wg.w; x += blockDim.x) {
    const float v = row[x];
    s += double(v); mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  __shared__ double ss[4];
  __shared__ float smn[4], smx[4];
  s = warp_sum(s); mn = warp_min(mn); mx = warp_max(mx);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { ss[warp] = s; smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    rowsum[c * wg.h + y] = (ss[0] + ss[1]) + (ss[2] + ss[3]);
    rowmin[c * wg.h + y] = fminf(fminf(smn[0], smn[1]), fminf(smn[2], smn[3]));
    rowmax[c * wg.h + y] = fmaxf(fmaxf(smx[0], smx[1]), fmaxf(smx[2], smx[3]));
    __threadfence();
    last = (atomicAdd(done_counter, 1u) == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double s2 = 0.0;
  float mn2 = INFINITY, mx2 = -INFINITY;
  for (int i = threadIdx.x; i < 3 * wg.h; i += blockDim.x) {
    s2 += __ldcg(rowsum + i); mn2 = fminf(mn2, __ldcg(rowmin + i)); mx2 = fmaxf(mx2, __ldcg(rowmax + i));
  }
  s2 = warp_sum(s2); mn2 = warp_min(mn2); mx2 = warp_max(mx2);
  __syncthreads();
  if (lane == 0) { ss[warp] = s2; smn[warp] = mn2; smx[warp] = mx2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->win_sum = (ss[0] + ss[1]) + (ss[2] + ss[3]);
    st->win_min = fminf(fminf(smn[0], smn[1]), fminf(smn[2], smn[3]));
    st->win_max = fmaxf(fmaxf(smx[0], smx[1]), fmaxf(smx[2], smx[3]));
    *done_counter = 0u;
  }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place radix-2 DIT FFT of s[0..L) (already in bit-reversed order), L/2 threads cooperate.
// tw[k] = exp(-2 pi i k / L) (global or shared memory); inverse uses the conjugate.  Unnormalised.
__device__ __forceinline__ void fft_smem(double2* s, const double2* tw, int L, int log2L, bool inverse) {
  const int j = threadIdx.x;
  for (int stage = 1; stage <= log2L; ++stage) {
    const int half = 1 << (stage - 1);
    if (j < L / 2) {
      const int grp = j >> (stage - 1), k = j & (half - 1);
      const int i0 = (grp << stage) + k, i1 = i0 + half;
      double2 w = tw[k << (log2L - stage)];
      if (inverse) w.y = -w.y;
      const double2 t = cmul(w, s[i1]);
      const double2 a = s[i0];
      s[i0] = make_double2(a.x + t.x, a.y + t.y);
      s[i1] = make_double2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
}

// S2+S3: block (y, c): t row -> zero-padded length-L forward FFT along x -> Z[c][y][:].  threads = L/2.
__global__ void k_white_rows_fwd(Geom g, const State* __restrict__ st, const float* __restrict__ err, WhiteGeom wg,
                                 const double2* __restrict__ tw, double2* __restrict__ Z) {
  if (st->stop) return;
  extern __shared__ double2 sd[];
  const int y = blockIdx.x, c = blockIdx.y, L = wg.L;
  const double mean = st->win_sum / (3.0 * wg.h * wg.w);
  const double dev = fmax(double(st->win_max) - mean, mean - double(st->win_min));
  const double scale = 1.0 / dev;
  const float* row = err + size_t(c) * g.plane + size_t(wg.top + g.P - g.row0 + y) * g.pitch + (wg.left + g.P);
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const int r = __brev(unsigned(i)) >> (32 - wg.log2L);
    sd[r] = make_double2(i < wg.w ? (double(row[i]) - mean) * scale : 0.0, 0.0);
  }
  __syncthreads();
  fft_smem(sd, tw, L, wg.log2L, false);
  double2* out = Z + (size_t(c) * L + y) * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x) out[i] = sd[i];
}

// S4: block (kx, c): column forward FFT (rows >= h are zero), |.|^2, inverse FFT; result back into Z.
// Only kx in [0, L/2] is launched: the window is real, so its power spectrum P is real with P[ky][L-kx] =
// P[L-ky][kx], and the column-inverse transform obeys Z'[dy][L-kx] = conj(Z'[dy][kx]); k_white_rows_inv mirrors.
__global__ void k_white_cols(const State* __restrict__ st, WhiteGeom wg, const double2* __restrict__ tw_g,
                             double2* __restrict__ Z) {
  if (st->stop) return;
  extern __shared__ double2 sd[];
  const int kx = blockIdx.x, c = blockIdx.y, L = wg.L;
  double2* tw = sd + L;                                     // twiddles once per block in shared memory
  for (int i = threadIdx.x; i < L / 2; i += blockDim.x) tw[i] = tw_g[i];
  double2* col = Z + size_t(c) * L * L + kx;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const int r = __brev(unsigned(i)) >> (32 - wg.log2L);
    sd[r] = (i < wg.h) ? col[size_t(i) * L] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  fft_smem(sd, tw, L, wg.log2L, false);
  // power spectrum, written back in bit-reversed order for the inverse transform
  double2 p0 = make_double2(0, 0), p1 = make_double2(0, 0);
  const int i0 = threadIdx.x, i1 = threadIdx.x + L / 2;
  if (i0 < L / 2) {
    p0 = make_double2(sd[i0].x * sd[i0].x + sd[i0].y * sd[i0].y, 0.0);
    p1 = make_double2(sd[i1].x * sd[i1].x + sd[i1].y * sd[i1].y, 0.0);
  }
  __syncthreads();
  if (i0 < L / 2) {
    sd[__brev(unsigned(i0)) >> (32 - wg.log2L)] = p0;
    sd[__brev(unsigned(i1)) >> (32 - wg.log2L)] = p1;
  }
  __syncthreads();
  fft_smem(sd, tw, L, wg.log2L, true);
  for (int i = threadIdx.x; i < L; i += blockDim.x) col[size_t(i) * L] = sd[i];
}

// S5: block (n', c): inverse FFT along x of row dy(n'), then sum_m' (R/L^2)^2 * a[n'] * b[m'] -> rowacc[c][n'].
// "same" crop of the full autocorrelation (scipy 'same': offset (full - same)//2):
//   same[n'][m'] = R[dy][dx],  dy = (h-1) - (n' + (h-1)/2),  dx = (w-1) - (m' + (w-1)/2)   (indices mod L)
__global__ void k_white_rows_inv(const State* __restrict__ st, WhiteGeom wg, const double2* __restrict__ tw,
                                 const double2* __restrict__ Z, const double* __restrict__ wa,
                                 const double* __restrict__ wb, double* __restrict__ rowacc) {
  if (st->stop) return;
  extern __shared__ double2 sd[];
  __shared__ double red[32];
  const int n = blockIdx.x, c = blockIdx.y, L = wg.L;
  int dy = (wg.h - 1) - (n + (wg.h - 1) / 2);
  if (dy < 0) dy += L;
  const double2* in = Z + (size_t(c) * L + dy) * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    double2 v;
    if (i <= L / 2) {
      v = in[i];
    } else {                                                // columns kx > L/2 were not computed: conjugate mirror
      v = in[L - i];
      v.y = -v.y;
    }
    sd[__brev(unsigned(i)) >> (32 - wg.log2L)] = v;
  }
  __syncthreads();
  fft_smem(sd, tw, L, wg.log2L, true);
  const double inv = 1.0 / (double(L) * double(L));
  double acc = 0.0;
  for (int m = threadIdx.x; m < wg.w; m += blockDim.x) {
    int dx = (wg.w - 1) - (m + (wg.w - 1) / 2);
    if (dx < 0) dx += L;
    const double v = sd[dx].x * inv;
    acc += v * v * wb[m];
  }
  acc = warp_sum(acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    const int nw = (blockDim.x + 31) / 32;
    for (int i = 0; i < nw; ++i) t += red[i];
    rowacc[c * wg.h + n] = t * wa[n];
  }
}

// S6: one thread: M_r, stop rule (pyx:623-656).  advance == 0: only report (stage-level test entry point).
// owner == 0 (row band that does not hold the whiteness window): only the iteration counter advances; the
// stop flag arrives by an int32 MAX all-reduce from the owning band.
__global__ void k_outer_finalize(State* __restrict__ st, WhiteGeom wg, const double* __restrict__ rowacc,
                                 int blind, float tau, int advance, int owner, float* __restrict__ out) {
  if (st->stop) return;
  if (blockIdx.x != 0) return;
  if (!owner) {
    if (advance && threadIdx.x == 0) st->it = st->it + 1;
    return;
  }
  // one warp: lane-strided partial sums, fixed butterfly (deterministic)
  double t = 0.0;
  for (int i = threadIdx.x; i < 3 * wg.h; i += 32) t += rowacc[i];
  t = warp_sum(t);
  if (threadIdx.x != 0) return;
  const float M_r = float(t / (3.0 * double(wg.h) * double(wg.w)));   // np.mean(test), pyx:638
  if (out) *out = M_r;
  if (!advance) return;
  const int it = st->it;
  if (it > 0) st->M_r_prev = st->M_r;                                  // pyx:623-624
  st->M_r = M_r;
  if (it < 4096) st->hist[it] = M_r;
  st->n_hist = min(it + 1, 4096);
  if (it > 1) {                                                        // pyx:643
    const float prev = st->M_r_prev;
    // advance == 2: benchmark stepping -- the rule is evaluated but not acted upon, so that steady-state steps can
    // be timed on inputs whose rule would fire early (never used by the drop-in entry points)
    if (blind) {
      if (M_r > prev && advance == 1) st->stop = 1;                    // pyx:646-647
    } else {
      if ((M_r - prev) / (M_r + prev) > tau && advance == 1) st->stop = 1;   // pyx:652-653
    }
  }
  st->it = it + 1;                                                     // pyx:656
}

}  // namespace rltv
