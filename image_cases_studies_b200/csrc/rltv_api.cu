// Host side of the C ABI declared in include/rltv_b200.h: context, buffers, launch sequencing.
// One inner step of the reference (lib/deconvolution.pyx:473-591) is the kernel sequence
//   GRAD     : k_conv<K,fwd> -> k_conv<K,adj>                         (pyx:477-491, :519, :524)
//   UPDATE   : k_update [-> k_halo_push]                              (pyx:527-531, :499-502, :552)
//   PSF_GRAD : k_conv<K,fwd> -> k_gradk<K> -> reduce                  (pyx:557-571)
//   PSF_STEP : k_psf_update                                           (pyx:574-589)
// and one outer iteration (pyx:460-656) is  ut=u ; 5 inner steps ; whiteness statistic + stop rule.
// Nothing is read back between kernels: step sizes, the PSF, the statistic and the stop flag live in a
// device-resident State; once the flag is set every later kernel returns at its first instruction, so the
// host may run ahead (it polls the flag two outer iterations behind, without draining the stream).
// With row-band sharding (one context per GPU, rltv_band.cuh) halos, step scalars, PSF-gradient sums and the stop
// flag cross GPUs by peer stores from inside these same kernels; the NCCL-baseline mode instead steps phase by
// phase and lets the host-side caller all-reduce three tiny device buffers (step_max, gk_sum, stop).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rltv_b200.h"
#include "rltv_band.cuh"
#include "rltv_chain_fft.cuh"
#include "rltv_common.cuh"
#include "rltv_elementwise.cuh"
#include "rltv_fft.cuh"
#include "rltv_stencil.cuh"
#include "rltv_stencil_fft.cuh"
#include "rltv_tv.cuh"
#include "rltv_tvmode.cuh"
#include "rltv_whiteness.cuh"

using namespace rltv;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(RLTV_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
  } while (0)

constexpr int RLTV_MAX_DIRECT_MK = 31;

enum Family { F_CONV_FWD = 0, F_CONV_ADJ, F_UPDATE, F_GRADK, F_PSF, F_STATS, F_COPY, F_HALO, F_COUNT };
const char* kFamilyNames[F_COUNT] = {"conv_fwd", "conv_adj", "update", "gradk", "psf", "stats", "copy", "halo"};

}  // namespace

struct rltv_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  Geom g{};
  int row_lo = 0, row_hi = 0, own_lo = 0, own_hi = 0;   // frame coordinates (rows of u)
  int fwd0 = 0, fwd1 = 0;                               // local rows on which the residual is computed
  bool banded = false;
  float *u = nullptr, *ut = nullptr, *gbuf = nullptr, *img = nullptr, *err = nullptr;
  size_t u_alloc_bytes = 0, flag_offset = 0;            // u allocation also carries the two halo flags
  float* staging = nullptr;
  float *psf = nullptr, *psf_caller = nullptr, *psf_hwc = nullptr;
  State* st = nullptr;
  State* h_st = nullptr;        // pinned mirror
  int* h_poll = nullptr;        // pinned ring: {stop, it} per slot
  cudaEvent_t poll_ev[4]{};
  float* gk_partial = nullptr;
  int gk_nparts = 0;            // CTAs of the persistent k_gradk grid (one partial each)
  int num_sms = 148;
  // TMA descriptors (rank-3 tensors x: Wu, y: rows, c: 3 over the planar arrays; boxes per kernel)
  CUtensorMap tm_u_conv{}, tm_err_conv{}, tm_img_epi{}, tm_u_epi{}, tm_ut_epi{}, tm_u_gk{}, tm_err_gk{};
  // row-FFT hybrid stencils (9 <= K <= 17): input boxes 128 x (96+K-1)
  CUtensorMap tm_u_fft{}, tm_err_fft{}, tm_u_gkfft{}, tm_err_gkfft{};
  CUtensorMap tm_img_gkfft;     // image rows of the owned rows, same boxes as tm_err_gkfft (fused residual)
  float2* gkf_part = nullptr;   // k_gradk_fft: per-CTA frequency-domain sums
  float2* wspec = nullptr;      // tap spectra [2][3][K][128]
  // spectral chain kernel (11 <= K <= 17): forward blur + residual + adjoint in one pass (rltv_chain_fft.cuh)
  bool use_chain = false;
  CUtensorMap tm_u_chain{};
  ChainPiece* chain_pieces = nullptr;   // device
  int* chain_first = nullptr;           // device: [num_sms + 1] first piece of every CTA
  int chain_npieces = 0;
  float2* chain_ipk = nullptr;          // packed image spectra, one row of 128 bins per packed residual row of every piece
  size_t chain_ipk_rows = 0;
  bool chain_ipk_valid = false;
  int inner_count = 0;                  // inner steps enqueued since rltv_begin (statistics slot = parity)
  int inner_in_outer = 0;               // 0 right after ut = u
  // CUDA graphs of one outer iteration (launch-bound frames): [parity of the first inner step's statistics slot]
  cudaGraphExec_t outer_graph[2] = {nullptr, nullptr};
  int outer_graph_launches = 0;         // kernel launches one replay stands for
  bool use_graph = false;
  float *tvut1 = nullptr, *tvut2 = nullptr, *tbuf = nullptr;   // TV-alive mode: TV(ut) maps, T = TV gradient term
  float* gather_stage = nullptr;        // full-frame HWC staging buffer of THIS rank (device), IPC-exported
  float* gather_peer[MAXR] = {};        // other ranks' staging buffers (IPC-mapped)
  void* gather_peer_base[MAXR] = {};
  bool ut_is_u = false;                 // adjoint launches of the first inner step read u in place of ut
  bool fold_psf_step = false;           // PSF_GRAD phase: let the row-FFT finish kernel also do the PSF update + spectra
  bool psf_step_folded = false;         // ... it did: PSF_STEP has nothing left to launch
  bool use_fft = false;         // forward blur / adjoint through k_conv_fft
  bool use_fft_gradk = false;   // PSF gradient through k_gradk_fft
  bool fuse_residual = false;   // ... which also computes the residual of pyx:557-565 itself (no forward-blur launch)
  bool psf_grad_write_err = true;  // fused kernel stores the residual (needed by the whiteness statistic only)
  double* gk_sum = nullptr;
  // row bands: halo exchange + all-gathers through peer memory (rltv_band.cuh)
  HaloSide side[2]{};           // 0: band above, 1: band below
  void* peer_base[MAXR] = {};
  CommPeers peers{};            // peers.nranks == 1: whole frame or NCCL-baseline mode (no in-kernel exchange)
  int rank = 0, world = 1;
  int world_rank = 0;           // rank in the process group even when the kernels run with rank 0 (NCCL baseline)
  bool fused_comm = false;      // true: in-kernel all-gathers over NVLink; false: host all-reduces (NCCL baseline)
  unsigned* counters = nullptr; // "last CTA" tickets: [0] k_halo_push (unused since the update pushes), [1] adjoint / chain,
                                // [2] gradk, [3] whiteness, [4..6] k_update halo exchange (top side, bottom side, grid)
  int halo_seq = 0;             // exchange numbers: identical on every band, never reset
  int max_seq = 0, gk_seq = 0, stop_seq = 0;
  bool white_owner = true;
  bool ignore_stop = false;     // benchmark stepping only: evaluate the stop rule, do not act on it
  int outer_since_begin = 0;
  // whiteness
  WhiteGeom wg{};
  double2 *Z = nullptr, *tw = nullptr;
  double *wa = nullptr, *wb = nullptr, *rowacc = nullptr, *rowsum = nullptr;
  float *rowmin = nullptr, *rowmax = nullptr, *mr_out = nullptr;
  int wcap_L = 0, wcap_h = 0, wcap_w = 0;
  rltv_params_t params{};
  bool uploaded = false, begun = false;
  int outer_enqueued = 0;
  long launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // profiling
  bool prof = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[F_COUNT];
  size_t prof_used[F_COUNT]{};
  float prof_ms[F_COUNT]{};
  int prof_n[F_COUNT]{};
};

namespace {

struct ProfScope {
  rltv_ctx* c;
  int fam;
  cudaEvent_t stop = nullptr;
  ProfScope(rltv_ctx* ctx, int f) : c(ctx), fam(f) {
    c->launches++;
    if (!c->prof) return;
    auto& pool = c->prof_ev[fam];
    if (c->prof_used[fam] == pool.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      pool.emplace_back(a, b);
    }
    auto& p = pool[c->prof_used[fam]++];
    cudaEventRecord(p.first, c->stream);
    stop = p.second;
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, c->stream);
  }
};

void prof_collect(rltv_ctx* c) {
  for (int f = 0; f < F_COUNT; ++f) {
    for (size_t i = 0; i < c->prof_used[f]; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, c->prof_ev[f][i].first, c->prof_ev[f][i].second) == cudaSuccess) {
        c->prof_ms[f] += ms;
        c->prof_n[f] += 1;
      }
    }
    c->prof_used[f] = 0;
  }
}

// ---- K-dispatch ------------------------------------------------------------------------------------
// (-DRLTV_K_LIST="M(15)" restricts the instantiations: seconds instead of minutes for a kernel-development build)
#ifdef RLTV_K_LIST
#define RLTV_FOR_EACH_K(M) RLTV_K_LIST
#else
#define RLTV_FOR_EACH_K(M) \
  M(3) M(5) M(7) M(9) M(11) M(13) M(15) M(17) M(19) M(21) M(23) M(25) M(27) M(29) M(31) \
  M(33) M(35) M(37) M(39) M(41) M(43) M(45) M(47)
#endif

// ---- TMA descriptors ---------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// rank-3 view (x: Wu, y: rows, c: 3) of `rows` rows of one planar frame array starting at `base`,
// box = bw x bh x 1, zero fill out of bounds
int make_tmap(CUtensorMap* tm, float* base, const Geom& g, int rows, int bw, int bh) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) return fail(RLTV_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {cuuint64_t(g.Wu), cuuint64_t(rows), 3};
  cuuint64_t strides[2] = {cuuint64_t(g.pitch) * 4, cuuint64_t(g.plane) * 4};
  cuuint32_t box[3] = {cuuint32_t(bw), cuuint32_t(bh), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RLTV_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
  return RLTV_OK;
}

template <int K>
int make_maps_t(rltv_ctx* c) {
  const Geom& g = c->g;
  int rc;
  if constexpr (K <= RLTV_MAX_DIRECT_MK) {       // direct stencils exist up to 31 x 31
    using C = ConvCfg<K>;
    using G = GradkCfg<K>;
    if ((rc = make_tmap(&c->tm_u_conv, c->u, g, g.Hu, C::SP, C::SROWS))) return rc;
    if ((rc = make_tmap(&c->tm_err_conv, c->err, g, g.Hu, C::SP, C::SROWS))) return rc;
    if ((rc = make_tmap(&c->tm_img_epi, c->img, g, g.Hu, C::TW, C::TH))) return rc;
    if ((rc = make_tmap(&c->tm_u_epi, c->u, g, g.Hu, C::TW, C::TH))) return rc;
    if ((rc = make_tmap(&c->tm_ut_epi, c->ut, g, g.Hu, C::TW, C::TH))) return rc;
    if ((rc = make_tmap(&c->tm_u_gk, c->u, g, g.Hu, G::SP, G::SROWS))) return rc;
    // residual as seen by the PSF gradient: OWNED rows only (halo rows read as zero)
    if ((rc = make_tmap(&c->tm_err_gk, c->err + size_t(g.own0) * g.pitch, g, g.own1 - g.own0, G::TW, G::TH))) return rc;
  }
  if constexpr (K >= 9) {
    using F = FftCfg<K>;
    if ((rc = make_tmap(&c->tm_u_fft, c->u, g, g.Hu, F::INW, F::IN_ROWS))) return rc;
    if ((rc = make_tmap(&c->tm_err_fft, c->err, g, g.Hu, F::INW, F::IN_ROWS))) return rc;
  }
  if constexpr (K >= 9) {
    using GF = GradkFftCfg<K>;
    if ((rc = make_tmap(&c->tm_u_gkfft, c->u, g, g.Hu, GF::UW, GF::U_ROWS))) return rc;
    if ((rc = make_tmap(&c->tm_err_gkfft, c->err + size_t(g.own0) * g.pitch, g, g.own1 - g.own0, FFT_N, GF::TROWS))) return rc;
    if ((rc = make_tmap(&c->tm_img_gkfft, c->img + size_t(g.own0) * g.pitch, g, g.own1 - g.own0, FFT_N, GF::TROWS))) return rc;
  }
  return RLTV_OK;
}

template <int K, bool ADJ>
int launch_conv_fft_t(rltv_ctx* c, float lambd) {
  if constexpr (K >= 9) {
    using C = FftCfg<K>;
    constexpr int SMEM = C::SMEM_BYTES;
    CU(cudaFuncSetAttribute(k_conv_fft<K, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const int y0 = ADJ ? c->g.own0 : c->fwd0, y1 = ADJ ? c->g.own1 : c->fwd1;
    const int ntx = (c->g.Wu + C::TWO - 1) / C::TWO, nty = (y1 - y0 + C::TROWS - 1) / C::TROWS;
    int grid = C::CTAS_PER_SM * c->num_sms;
    if (grid > 3 * ntx * nty) grid = 3 * ntx * nty;
    ProfScope p(c, ADJ ? F_CONV_ADJ : F_CONV_FWD);
    if (ADJ) {
      if (c->peers.nranks > 1) c->max_seq += 1;
      k_conv_fft<K, true><<<grid, C::THREADS, SMEM, c->stream>>>(c->tm_err_fft, c->u, c->ut_is_u ? c->u : c->ut, c->g, c->st, c->wspec,
                                                               lambd, c->gbuf, ntx, nty, y0, y1, c->peers, c->max_seq, c->counters + 1);
    } else {
      k_conv_fft<K, false><<<grid, C::THREADS, SMEM, c->stream>>>(c->tm_u_fft, c->img, c->img, c->g, c->st, c->wspec,
                                                                lambd, c->err, ntx, nty, y0, y1, c->peers, 0, c->counters + 1);
    }
    return RLTV_OK;
  } else {
    return fail(RLTV_ERR_ARG, "row-FFT stencils exist for MK >= 9");
  }
}

// forward blur + residual on local rows [y0, y1) only (chain path, non-blind: the whiteness window, pyx:623-638)
template <int K>
int launch_conv_fwd_rows_t(rltv_ctx* c, int y0, int y1) {
  if constexpr (K >= 9) {
    using C = FftCfg<K>;
    constexpr int SMEM = C::SMEM_BYTES;
    CU(cudaFuncSetAttribute(k_conv_fft<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const int ntx = (c->g.Wu + C::TWO - 1) / C::TWO, nty = (y1 - y0 + C::TROWS - 1) / C::TROWS;
    int grid = C::CTAS_PER_SM * c->num_sms;
    if (grid > 3 * ntx * nty) grid = 3 * ntx * nty;
    ProfScope p(c, F_STATS);
    k_conv_fft<K, false><<<grid, C::THREADS, SMEM, c->stream>>>(c->tm_u_fft, c->img, c->img, c->g, c->st, c->wspec, 0.f, c->err,
                                                              ntx, nty, y0, y1, c->peers, 0, c->counters + 1);
    return RLTV_OK;
  } else {
    return fail(RLTV_ERR_ARG, "row-FFT stencils exist for MK >= 9");
  }
}

int launch_conv_fwd(rltv_ctx* c);
int launch_conv_fwd_window(rltv_ctx* c) {
  int y0 = c->wg.top + c->g.P - c->g.row0, y1 = y0 + c->wg.h;
  if (y0 < 0) y0 = 0;
  if (y1 > c->g.Hu) y1 = c->g.Hu;
  switch (c->g.K) {
    case 5:
    case 7: return launch_conv_fwd(c);       // direct forward stencil (whole frame; the row-FFT tile kernels start at K = 9)
    case 9: return launch_conv_fwd_rows_t<9>(c, y0, y1);
    case 11: return launch_conv_fwd_rows_t<11>(c, y0, y1);
    case 13: return launch_conv_fwd_rows_t<13>(c, y0, y1);
    case 15: return launch_conv_fwd_rows_t<15>(c, y0, y1);
    case 17: return launch_conv_fwd_rows_t<17>(c, y0, y1);
  }
  return fail(RLTV_ERR_ARG, "unsupported MK");
}

int launch_psf_spectrum(rltv_ctx* c) {
  if (!c->use_fft && !c->use_chain) return RLTV_OK;
  ProfScope p(c, F_PSF);
  k_psf_spectrum<<<dim3(c->g.K, 3, 2), FFT_N, 0, c->stream>>>(c->st, c->psf, c->g.K, c->wspec);
  return RLTV_OK;
}

// forward residual on local rows [fwd0, fwd1); adjoint on the owned rows
template <int K, bool ADJ>
int launch_conv_t(rltv_ctx* c, float lambd) {
  if (c->use_fft) return launch_conv_fft_t<K, ADJ>(c, lambd);
  if constexpr (K <= RLTV_MAX_DIRECT_MK) {
  using C = ConvCfg<K>;
  constexpr int SMEM = C::smem_bytes(ADJ);
  CU(cudaFuncSetAttribute(k_conv<K, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  const int y0 = ADJ ? c->g.own0 : c->fwd0, y1 = ADJ ? c->g.own1 : c->fwd1;
  const int ntx = (c->g.pitch + C::TW - 1) / C::TW, nty = (y1 - y0 + C::TH - 1) / C::TH;
  int grid = C::CTAS_PER_SM * c->num_sms;
  if (grid > 3 * ntx * nty) grid = 3 * ntx * nty;
  ProfScope p(c, ADJ ? F_CONV_ADJ : F_CONV_FWD);
  if (ADJ) {
    if (c->peers.nranks > 1) c->max_seq += 1;
    k_conv<K, true><<<grid, C::THREADS, SMEM, c->stream>>>(c->tm_err_conv, c->tm_u_epi, c->ut_is_u ? c->tm_u_epi : c->tm_ut_epi, c->g, c->st, c->psf,
                                                          lambd, c->gbuf, ntx, nty, y0, y1, c->peers, c->max_seq, c->counters + 1);
  } else {
    k_conv<K, false><<<grid, C::THREADS, SMEM, c->stream>>>(c->tm_u_conv, c->tm_img_epi, c->tm_img_epi, c->g, c->st, c->psf,
                                                           lambd, c->err, ntx, nty, y0, y1, c->peers, 0, c->counters + 1);
  }
  return RLTV_OK;
  } else {
    return fail(RLTV_ERR_ARG, "direct stencils exist for MK <= 31");
  }
}

template <int K>
int launch_gradk_fft_t(rltv_ctx* c) {
  if constexpr (K >= 9) {
    using C = GradkFftCfg<K>;
    CU(cudaFuncSetAttribute(k_gradk_fft<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    const int ntx = (c->g.Wu + C::TWO - 1) / C::TWO, nty = (c->g.own1 - c->g.own0 + C::TROWS - 1) / C::TROWS;
    if (c->peers.nranks > 1) c->gk_seq += 1;
    {
      ProfScope p(c, F_GRADK);
      bool fused = false;
      if constexpr (C::CAN_FUSE) {
        if (c->fuse_residual) {
          fused = true;
          CU(cudaFuncSetAttribute(k_gradk_fft<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES_FUSED));
          k_gradk_fft<K, true><<<c->gk_nparts, C::THREADS, C::SMEM_BYTES_FUSED, c->stream>>>(
              c->tm_u_gkfft, c->tm_img_gkfft, c->g, c->st, c->wspec, c->psf_grad_write_err ? c->err : nullptr, c->gkf_part, ntx, nty);
        }
      }
      if (!fused)
        k_gradk_fft<K, false><<<c->gk_nparts, C::THREADS, C::SMEM_BYTES, c->stream>>>(c->tm_u_gkfft, c->tm_err_gkfft, c->g, c->st,
                                                                                    nullptr, nullptr, c->gkf_part, ntx, nty);
    }
    {
      ProfScope p(c, F_GRADK);
      // whole PSF step in this launch unless the host has to all-reduce the sums in between (NCCL baseline) or a stage
      // entry point asked for the gradient only
      const int fold = (c->fold_psf_step && !(c->banded && !c->fused_comm)) ? 1 : 0;
      c->psf_step_folded = fold != 0;
      CU(cudaFuncSetAttribute(k_gradk_fft_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, int(6 * RLTV_MAX_MK * RLTV_MAX_MK * sizeof(float))));
      k_gradk_fft_finish<<<3 * K, 512, fold ? 6 * K * K * sizeof(float) : 0, c->stream>>>(
          c->st, c->gkf_part, c->gk_nparts, K, c->gk_sum, c->peers, c->gk_seq, c->counters + 2, fold, c->params.step_factor,
          c->params.correlation, c->psf, c->psf_caller, c->wspec);
    }
    return RLTV_OK;
  } else {
    return fail(RLTV_ERR_ARG, "row-FFT stencils exist for MK >= 9");
  }
}

template <int K>
int launch_gradk_t(rltv_ctx* c) {
  if (c->use_fft_gradk) return launch_gradk_fft_t<K>(c);
  if constexpr (K <= RLTV_MAX_DIRECT_MK) {
  using C = GradkCfg<K>;
  CU(cudaFuncSetAttribute(k_gradk<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const int ntx = (c->g.Wu + C::TW - 1) / C::TW, nty = (c->g.own1 - c->g.own0 + C::TH - 1) / C::TH;
  if (c->peers.nranks > 1) c->gk_seq += 1;
  ProfScope p(c, F_GRADK);
  k_gradk<K><<<c->gk_nparts, C::THREADS, C::SMEM_BYTES, c->stream>>>(c->tm_u_gk, c->tm_err_gk, c->g, c->st, c->gk_partial, ntx, nty,
                                                                     c->gk_sum, c->peers, c->gk_seq, c->counters + 2);
  return RLTV_OK;
  } else {
    return fail(RLTV_ERR_ARG, "direct stencils exist for MK <= 31");
  }
}

int make_maps(rltv_ctx* c) {
  switch (c->g.K) {
#define M(K_) case K_: return make_maps_t<K_>(c);
    RLTV_FOR_EACH_K(M)
#undef M
  }
  return fail(RLTV_ERR_ARG, "unsupported MK");
}
int launch_conv_fwd(rltv_ctx* c) {
  switch (c->g.K) {
#define M(K_) case K_: return launch_conv_t<K_, false>(c, 0.f);
    RLTV_FOR_EACH_K(M)
#undef M
  }
  return fail(RLTV_ERR_ARG, "unsupported MK");
}
int launch_conv_adj(rltv_ctx* c, float lambd) {
  switch (c->g.K) {
#define M(K_) case K_: return launch_conv_t<K_, true>(c, lambd);
    RLTV_FOR_EACH_K(M)
#undef M
  }
  return fail(RLTV_ERR_ARG, "unsupported MK");
}
int launch_gradk(rltv_ctx* c) {
  switch (c->g.K) {
#define M(K_) case K_: return launch_gradk_t<K_>(c);
    RLTV_FOR_EACH_K(M)
#undef M
  }
  return fail(RLTV_ERR_ARG, "unsupported MK");
}

// TV-alive mode: TV(u) stencils, the TV gradient term T and its step statistics (pyx:495-496, :517, :543); no-op otherwise
int launch_tv_grad(rltv_ctx* c) {
  if (c->params.mode == RLTV_MODE_PAM_CTV) {       // unpinned extension: collaborative TV sub-gradient, one pass
    dim3 grid((c->g.pitch / 4 + 255) / 256, c->g.own1 - c->g.own0, 1);
    ProfScope p(c, F_UPDATE);
    k_ctv_grad<<<grid, 256, 0, c->stream>>>(c->g, c->st, c->u, c->gbuf, c->params.lambd, 1e-3f, c->tbuf, c->inner_count & 1);
    return RLTV_OK;
  }
  if (c->params.mode != RLTV_MODE_MM_TV) return RLTV_OK;
  dim3 grid((c->g.pitch / 4 + 255) / 256, c->g.own1 - c->g.own0, 3);
  ProfScope p(c, F_UPDATE);
  k_tv_grad<<<grid, 256, 0, c->stream>>>(c->g, c->st, c->u, c->ut, c->gbuf, c->tvut1, c->tvut2, c->img, c->params.lambd,
                                         c->params.blind ? 1e-2f : 1e-6f, c->tbuf, c->inner_count & 1, c->inner_in_outer == 0 ? 1 : 0);
  return RLTV_OK;
}

int launch_update(rltv_ctx* c) {
  dim3 grid((c->g.pitch / 4 + 255) / 256, c->g.own1 - c->g.own0, 3);
  ProfScope p(c, F_UPDATE);
  // chain path: this step's statistics are in slot (inner_count & 1); the update resets the other slot for the next step
  const bool tvm = c->params.mode != RLTV_MODE_MM;
  const int slot = (c->use_chain || tvm) ? (c->inner_count & 1) : 0;
  const int reset_slot = c->use_chain ? (slot ^ 1) : -1;
  if (c->params.mode == RLTV_MODE_PAM_CTV) {
    if (c->inner_in_outer == 0)
      k_update_tv<true, true><<<grid, 256, 0, c->stream>>>(c->g, c->st, c->u, nullptr, c->ut, c->gbuf, c->tbuf, c->img, c->params.step_factor,
                                                           c->params.lambd, c->params.blind, c->use_chain ? slot : 0, reset_slot, slot);
    else
      k_update_tv<false, true><<<grid, 256, 0, c->stream>>>(c->g, c->st, c->u, c->ut, nullptr, c->gbuf, c->tbuf, c->img, c->params.step_factor,
                                                            c->params.lambd, c->params.blind, c->use_chain ? slot : 0, reset_slot, slot);
    return RLTV_OK;
  }
  if (c->params.mode == RLTV_MODE_MM_TV) {
    if (c->inner_in_outer == 0)
      k_update_tv<true, false><<<grid, 256, 0, c->stream>>>(c->g, c->st, c->u, nullptr, c->ut, c->gbuf, c->tbuf, c->img, c->params.step_factor,
                                                     c->params.lambd, c->params.blind, c->use_chain ? slot : 0, reset_slot, slot);
    else
      k_update_tv<false, false><<<grid, 256, 0, c->stream>>>(c->g, c->st, c->u, c->ut, nullptr, c->gbuf, c->tbuf, c->img, c->params.step_factor,
                                                      c->params.lambd, c->params.blind, c->use_chain ? slot : 0, reset_slot, slot);
    c->chain_ipk_valid = false;                            // the image changed: its packed spectra follow
    return RLTV_OK;
  }
  HaloPush hp{};
  if (c->side[0].peer_u || c->side[1].peer_u) {            // row band with at least one neighbour: push inside the update
    c->halo_seq += 1;
    const int* flags = reinterpret_cast<const Comm*>(reinterpret_cast<const char*>(c->u) + c->flag_offset)->halo_flag;
    hp.top = c->side[0];
    hp.bot = c->side[1];
    hp.counters = c->counters + 4;
    hp.flag_from_top = c->side[0].peer_u ? flags + 0 : nullptr;
    hp.flag_from_bot = c->side[1].peer_u ? flags + 1 : nullptr;
    hp.seq = c->halo_seq;
    hp.enabled = 1;
  }
  if (c->inner_in_outer == 0)
    k_update<true><<<grid, 256, 0, c->stream>>>(c->g, c->st, c->u, nullptr, c->ut, c->gbuf, c->img, c->params.step_factor,
                                                c->params.lambd, c->params.blind, slot, reset_slot, hp);
  else
    k_update<false><<<grid, 256, 0, c->stream>>>(c->g, c->st, c->u, c->ut, nullptr, c->gbuf, c->img, c->params.step_factor,
                                                 c->params.lambd, c->params.blind, slot, reset_slot, hp);
  return RLTV_OK;
}

// ---- spectral chain kernel: work partition, image spectra, launch -------------------------------------------
template <int K>
int chain_setup_t(rltv_ctx* c) {
  if constexpr (K >= 5 && K <= 17) {
    using C = ChainCfg<K>;
    const Geom& g = c->g;
    const int nseg = chain_nseg(g.Wu, C::V), nstrips = 3 * nseg, G = c->num_sms;
    const int y0 = g.own0, y1 = g.own1, R = y1 - y0;
    auto xfix = [&](int s) { const int X0 = C::V * s - (K - 1); return !(X0 >= 0 && X0 + FFT_N <= g.Wu); };
    // balanced split of the weighted row sequence (strip-major) into one contiguous range per CTA; a border strip
    // (masking path every step) counts 1.7x
    std::vector<double> wsum(nstrips + 1, 0.0);
    double fixw = 1.7;
    if (const char* e = getenv("RLTV_CHAIN_FIXW")) fixw = atof(e);
    for (int i = 0; i < nstrips; ++i) wsum[i + 1] = wsum[i] + (xfix(i % nseg) ? fixw : 1.0) * R;
    const double total = wsum[nstrips];
    auto pos_of = [&](double w, int& strip, int& row) {      // inverse of the cumulative weight
      strip = 0;
      while (strip + 1 < nstrips && wsum[strip + 1] <= w) ++strip;
      const double per = (wsum[strip + 1] - wsum[strip]) / R;
      row = int((w - wsum[strip]) / per + 0.5);
      if (row > R) row = R;
      if (row < 0) row = 0;
    };
    std::vector<ChainPiece> pieces;
    std::vector<int> first(G + 1, 0);
    size_t ipk_rows = 0;
    int ps = 0, pr = 0;                                       // current position (strip, row)
    for (int b = 0; b < G; ++b) {
      first[b] = int(pieces.size());
      int es, er;
      if (b == G - 1) { es = nstrips - 1; er = R; } else pos_of(total * (b + 1) / G, es, er);
      while (ps < es || (ps == es && pr < er)) {
        const int rend = (ps < es) ? R : er;
        if (rend > pr) {
          const int n = rend - pr;
          ChainPiece pc;
          pc.c = ps / nseg; pc.s = ps % nseg;
          pc.ya = y0 + pr; pc.L = (n + 1) / 2; pc.Lb = n - pc.L;
          pc.nsteps = (pc.L + 4 * C::P + C::S - 1) / C::S;
          pc.ipk_row0 = int(ipk_rows);
          pc.xfix = xfix(pc.s) ? 1 : 0;
          ipk_rows += size_t(pc.nsteps) * C::S;
          pieces.push_back(pc);
        }
        if (ps < es) { ++ps; pr = 0; } else pr = rend;
      }
    }
    first[G] = int(pieces.size());
    if (ipk_rows * FFT_N >= (size_t(1) << 31)) return fail(RLTV_ERR_ARG, "frame too large for the chain kernel's piece table");
    cudaFree(c->chain_pieces); cudaFree(c->chain_first); cudaFree(c->chain_ipk);
  cudaFree(c->tvut1); cudaFree(c->tvut2); cudaFree(c->tbuf);
    c->chain_pieces = nullptr; c->chain_first = nullptr; c->chain_ipk = nullptr;
    c->chain_npieces = int(pieces.size());
    c->chain_ipk_rows = ipk_rows;
    CU(cudaMalloc(&c->chain_pieces, pieces.size() * sizeof(ChainPiece)));
    CU(cudaMalloc(&c->chain_first, (G + 1) * sizeof(int)));
    CU(cudaMalloc(&c->chain_ipk, ipk_rows * FFT_N * sizeof(float2)));
    CU(cudaMemcpyAsync(c->chain_pieces, pieces.data(), pieces.size() * sizeof(ChainPiece), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->chain_first, first.data(), (G + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int rc;
    if ((rc = make_tmap(&c->tm_u_chain, c->u, g, g.Hu, C::INW, C::S))) return rc;
    CU(cudaFuncSetAttribute(k_chain_fft<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    c->chain_ipk_valid = false;
    return RLTV_OK;
  } else {
    return fail(RLTV_ERR_ARG, "the chain kernel exists for 5 <= MK <= 17");
  }
}

template <int K>
int chain_image_spectra_t(rltv_ctx* c) {
  if constexpr (K >= 5 && K <= 17) {
    k_chain_image_spectra<K><<<dim3(8, c->chain_npieces < 2048 ? c->chain_npieces : 2048), 256, 0, c->stream>>>(
        c->g, c->img, c->chain_pieces, c->chain_npieces, c->chain_ipk);
    c->launches++;
    c->chain_ipk_valid = true;
    return RLTV_OK;
  } else {
    return fail(RLTV_ERR_ARG, "the chain kernel exists for 5 <= MK <= 17");
  }
}

template <int K>
int launch_chain_t(rltv_ctx* c, float lambd) {
  if constexpr (K >= 5 && K <= 17) {
    using C = ChainCfg<K>;
    if (c->peers.nranks > 1) c->max_seq += 1;
    ProfScope p(c, F_CONV_ADJ);
    k_chain_fft<K><<<c->num_sms, C::THREADS, C::SMEM_BYTES, c->stream>>>(
        c->tm_u_chain, c->g, c->st, c->wspec, c->chain_ipk, c->chain_pieces, c->chain_first, c->u, c->ut, lambd, c->gbuf,
        c->inner_count & 1, c->inner_in_outer == 0 ? 1 : 0, c->peers, c->max_seq, c->counters + 1);
    return RLTV_OK;
  } else {
    return fail(RLTV_ERR_ARG, "the chain kernel exists for 5 <= MK <= 17");
  }
}

#define RLTV_CHAIN_DISPATCH(FN, ...)                                   \
  switch (c->g.K) {                                                    \
    case 5: return FN<5>(__VA_ARGS__);                                 \
    case 7: return FN<7>(__VA_ARGS__);                                 \
    case 9: return FN<9>(__VA_ARGS__);                                 \
    case 11: return FN<11>(__VA_ARGS__);                               \
    case 13: return FN<13>(__VA_ARGS__);                               \
    case 15: return FN<15>(__VA_ARGS__);                               \
    case 17: return FN<17>(__VA_ARGS__);                               \
  }                                                                    \
  return fail(RLTV_ERR_ARG, "the chain kernel exists for MK = 5..17");
int chain_setup(rltv_ctx* c) { RLTV_CHAIN_DISPATCH(chain_setup_t, c) }
int chain_image_spectra(rltv_ctx* c) { RLTV_CHAIN_DISPATCH(chain_image_spectra_t, c) }
int launch_chain(rltv_ctx* c, float lambd) { RLTV_CHAIN_DISPATCH(launch_chain_t, c, lambd) }

int launch_psf_update(rltv_ctx* c) {
  const int K = c->g.K;
  {
    ProfScope p2(c, F_PSF);
    CU(cudaFuncSetAttribute(k_psf_update, cudaFuncAttributeMaxDynamicSharedMemorySize, int(6 * RLTV_MAX_MK * RLTV_MAX_MK * sizeof(float))));
    k_psf_update<<<1, 256, 6 * K * K * sizeof(float), c->stream>>>(c->st, c->gk_sum, K, c->params.step_factor,
                                                                  c->params.correlation, c->psf, c->psf_caller,
                                                                  c->peers.peer[c->rank], c->peers.nranks, c->gk_seq);
  }
  return launch_psf_spectrum(c);
}

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// (Re)allocate the whiteness buffers for a window and upload twiddles / separable weights.
int setup_whiteness(rltv_ctx* c, int top, int bottom, int left, int right) {
  const int h = bottom - top, w = right - left;
  if (top < 0 || left < 0 || bottom > c->g.M || right > c->g.N || h < 2 || w < 2)
    return fail(RLTV_ERR_ARG, "whiteness window outside the image");
  if (c->white_owner) {
    const int l0 = top + c->g.P - c->g.row0, l1 = bottom + c->g.P - c->g.row0;
    const int r0 = c->fuse_residual ? c->g.own0 : c->fwd0, r1 = c->fuse_residual ? c->g.own1 : c->fwd1;
    if (l0 < r0 || l1 > r1)
      return fail(RLTV_ERR_ARG, "whiteness window is not inside the rows this band computes the residual on");
  }
  const int mx = h > w ? h : w;
  int L = next_pow2(mx + (mx + 1) / 2 + 1);
  if (L < 64) L = 64;
  if (L > 2048) return fail(RLTV_ERR_ARG, "whiteness window larger than 1364 pixels is not supported");
  int log2L = 0;
  while ((1 << log2L) < L) ++log2L;
  if (L > c->wcap_L) {
    cudaFree(c->Z); cudaFree(c->tw);
    c->Z = nullptr; c->tw = nullptr; c->wcap_L = 0;      // a failed re-allocation must not leave dangling pointers behind
    CU(cudaMalloc(&c->Z, size_t(3) * L * L * sizeof(double2)));
    CU(cudaMalloc(&c->tw, size_t(L / 2) * sizeof(double2)));
    c->wcap_L = L;
  }
  if (h > c->wcap_h || w > c->wcap_w) {
    cudaFree(c->wa); cudaFree(c->wb); cudaFree(c->rowacc); cudaFree(c->rowsum); cudaFree(c->rowmin); cudaFree(c->rowmax);
    c->wa = c->wb = c->rowacc = c->rowsum = nullptr; c->rowmin = c->rowmax = nullptr;
    const int keep_h = c->wcap_h, keep_w = c->wcap_w;
    c->wcap_h = c->wcap_w = 0;
    const int ch = h > keep_h ? h : keep_h, cw = w > keep_w ? w : keep_w;
    CU(cudaMalloc(&c->wa, ch * sizeof(double)));
    CU(cudaMalloc(&c->wb, cw * sizeof(double)));
    CU(cudaMalloc(&c->rowacc, 3 * ch * sizeof(double)));
    CU(cudaMalloc(&c->rowsum, 3 * ch * sizeof(double)));
    CU(cudaMalloc(&c->rowmin, 3 * ch * sizeof(float)));
    CU(cudaMalloc(&c->rowmax, 3 * ch * sizeof(float)));
    c->wcap_h = ch; c->wcap_w = cw;
  }
  if (!c->mr_out) CU(cudaMalloc(&c->mr_out, sizeof(float)));
  std::vector<double2> tw(L / 2);
  for (int k = 0; k < L / 2; ++k) {
    const double a = -2.0 * M_PI * double(k) / double(L);
    tw[k] = make_double2(cos(a), sin(a));
  }
  // separable weights: W[n][m] = sqrt(gw(x_n) gw(x_m)) / sum  (pyx:35-36, :393-404); x = float32 linspace(-1, 1)
  auto weights = [](int n) {
    std::vector<double> v(n);
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
      const double step = 2.0 / double(n - 1);
      const float x = float(-1.0 + step * i);
      const double gw = exp(-double(x) * double(x) / 2.0) / sqrt(2.0 * M_PI);
      v[i] = sqrt(gw);
      s += v[i];
    }
    for (auto& e : v) e /= s;
    return v;
  };
  std::vector<double> wa = weights(h), wb = weights(w);
  CU(cudaMemcpyAsync(c->tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->wa, wa.data(), h * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->wb, wb.data(), w * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));  // host vectors go out of scope
  c->wg = WhiteGeom{top, left, h, w, L, log2L};
  return RLTV_OK;
}

int launch_whiteness(rltv_ctx* c, int advance) {
  const WhiteGeom& wg = c->wg;
  if (!c->white_owner) {
    ProfScope p(c, F_STATS);
    k_outer_finalize<<<1, 32, 0, c->stream>>>(c->st, wg, c->rowacc, c->params.blind, c->params.tau, advance, 0, c->mr_out);
    return RLTV_OK;
  }
  const int L = wg.L;
  const int fft_threads = L / 2 < 32 ? 32 : L / 2;
  const size_t fft_smem = size_t(L) * sizeof(double2);
  CU(cudaFuncSetAttribute(k_white_rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 16));
  CU(cudaFuncSetAttribute(k_white_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 16 + 1024 * 16));
  CU(cudaFuncSetAttribute(k_white_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 16));
  { ProfScope p(c, F_STATS);
    k_win_rows<<<dim3(wg.h, 3), 128, 0, c->stream>>>(c->g, c->st, c->err, wg, c->rowsum, c->rowmin, c->rowmax, c->counters + 3); }
  { ProfScope p(c, F_STATS);
    k_white_rows_fwd<<<dim3(wg.h, 3), fft_threads, fft_smem, c->stream>>>(c->g, c->st, c->err, wg, c->tw, c->Z); }
  { ProfScope p(c, F_STATS);
    k_white_cols<<<dim3(L / 2 + 1, 3), fft_threads, fft_smem + size_t(L / 2) * sizeof(double2), c->stream>>>(c->st, wg, c->tw, c->Z); }
  { ProfScope p(c, F_STATS);
    k_white_rows_inv<<<dim3(wg.h, 3), fft_threads, fft_smem, c->stream>>>(c->st, wg, c->tw, c->Z, c->wa, c->wb, c->rowacc); }
  { ProfScope p(c, F_STATS);
    k_outer_finalize<<<1, 32, 0, c->stream>>>(c->st, wg, c->rowacc, c->params.blind, c->params.tau, advance, 1, c->mr_out); }
  return RLTV_OK;
}

int enqueue_phase(rltv_ctx* c, int phase) {
  int rc = RLTV_OK;
  switch (phase) {
    case RLTV_PH_OUTER_BEGIN:
      // ut[:] = u.copy() (pyx:462) costs nothing here: the first inner step reads u where it would read ut (the
      // adjoint kernels get u twice, the chain kernel a flag) and its update kernel writes ut <- u_old on the fly
      c->inner_in_outer = 0;
      if (c->params.mode == RLTV_MODE_MM_TV) {                     // TV(ut, ...) of pyx:464-465 (patched reference), ut == u here
        dim3 grid((c->g.pitch / 4 + 255) / 256, c->g.own1 - c->g.own0, 3);
        ProfScope p(c, F_STATS);
        k_tv_maps<<<grid, 256, 0, c->stream>>>(c->g, c->st, c->u, c->params.blind ? 1e-2f : 1e-6f, c->tvut1, c->tvut2);
      }
      return RLTV_OK;
    case RLTV_PH_GRAD:
      if (c->use_chain) {
        if (!c->chain_ipk_valid && (rc = chain_image_spectra(c)) != RLTV_OK) return rc;
        // the whiteness statistic reads the residual of the last inner step (pyx:623-638): in blind mode the PSF-gradient
        // kernel leaves it behind, otherwise compute it on the window rows only.  Launched BEFORE the chain kernel: like
        // every forward launch it resets statistics slot 0, which at this point is either already reset or unused.
        if (!c->params.blind && c->white_owner && c->inner_in_outer == RLTV_INNER_ITER - 1)
          if ((rc = launch_conv_fwd_window(c)) != RLTV_OK) return rc;
        if ((rc = launch_chain(c, c->params.lambd)) != RLTV_OK) return rc;    // pyx:477-491, :519, :524 in one pass
        return launch_tv_grad(c);
      }
      if ((rc = launch_conv_fwd(c)) != RLTV_OK) return rc;        // pyx:477-488
      c->ut_is_u = (c->inner_in_outer == 0);
      rc = launch_conv_adj(c, c->params.lambd);                   // pyx:490-491, :519, :524
      c->ut_is_u = false;
      if (rc != RLTV_OK) return rc;
      return launch_tv_grad(c);
    case RLTV_PH_UPDATE:
      if ((rc = launch_update(c)) != RLTV_OK) return rc;          // pyx:527-531, :499-502, :552
      c->inner_count += 1;
      c->inner_in_outer += 1;
      return RLTV_OK;                                             // (row bands: the halo exchange happened inside k_update)
    case RLTV_PH_PSF_GRAD:
      if (!c->fuse_residual)
        if ((rc = launch_conv_fwd(c)) != RLTV_OK) return rc;      // pyx:557-565 (else inside the PSF-gradient kernel)
      c->fold_psf_step = true;
      c->psf_step_folded = false;
      rc = launch_gradk(c);                                       // pyx:567-571 (+ pyx:574-589 when folded)
      c->fold_psf_step = false;
      return rc;
    case RLTV_PH_PSF_STEP:
      if (c->psf_step_folded) { c->psf_step_folded = false; return RLTV_OK; }
      return launch_psf_update(c);                                // pyx:574-589
    case RLTV_PH_OUTER_END:
      c->outer_since_begin += 1;
      if (c->peers.nranks > 1) {
        c->stop_seq += 1;
        if (!c->white_owner) {
          ProfScope p(c, F_STATS);
          k_stop_gather<<<1, 32, 0, c->stream>>>(c->st, c->peers.peer[c->rank], c->stop_seq);
          return RLTV_OK;
        }
        if ((rc = launch_whiteness(c, c->ignore_stop ? 2 : 1)) != RLTV_OK) return rc;  // pyx:623-656
        ProfScope p(c, F_STATS);
        k_stop_publish<<<1, 32, 0, c->stream>>>(c->st, c->peers, c->stop_seq, c->outer_since_begin);
        return RLTV_OK;
      }
      return launch_whiteness(c, c->ignore_stop ? 2 : 1);                              // pyx:623-656
  }
  return fail(RLTV_ERR_ARG, "unknown phase");
}

int enqueue_outer(rltv_ctx* c) {
  int rc;
  if ((rc = enqueue_phase(c, RLTV_PH_OUTER_BEGIN)) != RLTV_OK) return rc;
  for (int i = 0; i < RLTV_INNER_ITER; ++i) {
    if ((rc = enqueue_phase(c, RLTV_PH_GRAD)) != RLTV_OK) return rc;
    if ((rc = enqueue_phase(c, RLTV_PH_UPDATE)) != RLTV_OK) return rc;
    if (c->params.blind) {
      // only the last residual of an outer iteration is read again (by the whiteness statistic, pyx:623-638)
      c->psf_grad_write_err = (i == RLTV_INNER_ITER - 1);
      rc = enqueue_phase(c, RLTV_PH_PSF_GRAD);
      c->psf_grad_write_err = true;
      if (rc != RLTV_OK) return rc;
      if ((rc = enqueue_phase(c, RLTV_PH_PSF_STEP)) != RLTV_OK) return rc;
    }
  }
  if ((rc = enqueue_phase(c, RLTV_PH_OUTER_END)) != RLTV_OK) return rc;
  CU(cudaGetLastError());
  return RLTV_OK;
}

void drop_graphs(rltv_ctx* c) {
  for (auto& g : c->outer_graph)
    if (g) { cudaGraphExecDestroy(g); g = nullptr; }
}

// One outer iteration, replayed from a CUDA graph when the frame is small enough to be launch-bound (SURVEY.md section 7
// step 7; the pyramid's 255-px crops, BASELINE config 1): ~20 launches of 8-17 us each collapse into one graph launch.
// Everything a launch depends on is either constant within a solve or periodic in the parity of the inner-step counter
// (the statistics slot), hence two graphs.  Row bands keep plain launches (their exchange numbers grow every step).
int enqueue_outer_fast(rltv_ctx* c) {
  if (!c->use_graph || c->prof || c->banded || c->outer_enqueued < 2) return enqueue_outer(c);
  const int par = c->inner_count & 1;
  if (!c->outer_graph[par]) {
    cudaGraph_t graph = nullptr;
    const long before = c->launches;
    const int ic = c->inner_count, osb = c->outer_since_begin;
    CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_outer(c);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != RLTV_OK || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      c->use_graph = false;                                   // capture not possible here: plain launches from now on
      c->inner_count = ic; c->outer_since_begin = osb; c->launches = before;
      return enqueue_outer(c);
    }
    c->outer_graph_launches = int(c->launches - before);
    e = cudaGraphInstantiate(&c->outer_graph[par], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(RLTV_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    // the capture only recorded the work (and advanced the host-side counters): run it now
    CU(cudaGraphLaunch(c->outer_graph[par], c->stream));
    return RLTV_OK;
  }
  CU(cudaGraphLaunch(c->outer_graph[par], c->stream));
  c->inner_count += RLTV_INNER_ITER;
  c->inner_in_outer = RLTV_INNER_ITER;
  c->outer_since_begin += 1;
  c->launches += c->outer_graph_launches;
  return RLTV_OK;
}

void fill_stats(rltv_ctx* c, rltv_stats_t* s, float ms) {
  if (!s) return;
  const State& h = *c->h_st;
  s->iterations_executed = h.it;
  s->stopped = h.stop;
  s->M_r = h.M_r;
  s->M_r_prev = h.M_r_prev;
  for (int i = 0; i < 3; ++i) s->dt[i] = h.dt[i];
  s->dtpsf = h.dtpsf;
  s->solve_ms = ms;
  s->kernel_launches = int(c->launches);
  s->n_history = h.n_hist;
  std::memcpy(s->M_r_history, h.hist, sizeof(float) * (h.n_hist > 0 ? h.n_hist : 0));
}

int check_ctx(rltv_ctx* c) {
  if (!c) return fail(RLTV_ERR_ARG, "null context");
  CU(cudaSetDevice(c->device));
  return RLTV_OK;
}

int hwc_grid(int cols) {
  const int b = (cols * 3 + 255) / 256;
  return b > 64 ? 64 : b;
}

}  // namespace

// =====================================================================================================
extern "C" {

int rltv_abi_version(void) { return RLTV_ABI_VERSION; }
const char* rltv_last_error(void) { return g_err.c_str(); }

int rltv_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
    cudaGetLastError();
    return 0;
  }
  if (e != cudaSuccess) return fail(RLTV_ERR_CUDA, cudaGetErrorString(e));
  return n;
}

int rltv_create_band(rltv_ctx** out, int32_t device, int32_t M, int32_t N, int32_t MK, const rltv_band_t* band, void* stream) {
  if (!out) return fail(RLTV_ERR_ARG, "null out pointer");
  *out = nullptr;
  if (MK < 3 || (MK % 2) == 0 || MK > RLTV_MAX_MK) return fail(RLTV_ERR_ARG, "MK must be odd and in [3, 47]");
  if (M < 1 || N < 1) return fail(RLTV_ERR_ARG, "empty image");
  const int HuG = M + MK - 1, P = MK / 2;
  rltv_band_t b = band ? *band : rltv_band_t{0, HuG, 0, HuG};
  if (b.row_lo < 0 || b.row_hi > HuG || b.own_lo < b.row_lo || b.own_hi > b.row_hi || b.own_lo >= b.own_hi)
    return fail(RLTV_ERR_ARG, "bad row band");
  if ((b.own_lo > 0 && b.own_lo - b.row_lo < 2 * P) || (b.own_hi < HuG && b.row_hi - b.own_hi < 2 * P))
    return fail(RLTV_ERR_ARG, "row band needs a halo of 2*(MK/2) rows on every interior side");
  if (b.row_hi - b.row_lo > 65535) return fail(RLTV_ERR_ARG, "more than 65535 padded rows");
  int ndev = rltv_device_count();
  if (ndev <= 0) return fail(RLTV_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(RLTV_ERR_ARG, "bad device index");
  CU(cudaSetDevice(device));
  rltv_ctx* c = new rltv_ctx();
  c->device = device;
  if (stream) {
    c->stream = reinterpret_cast<cudaStream_t>(stream);
  } else {
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  c->row_lo = b.row_lo; c->row_hi = b.row_hi; c->own_lo = b.own_lo; c->own_hi = b.own_hi;
  c->banded = band != nullptr && !(b.row_lo == 0 && b.row_hi == HuG);
  Geom& g = c->g;
  g.M = M; g.N = N; g.K = MK; g.P = P;
  g.Hu = b.row_hi - b.row_lo; g.Wu = N + MK - 1;
  g.pitch = (g.Wu + 3) & ~3;
  g.plane = size_t(g.Hu) * g.pitch;
  g.row0 = b.row_lo; g.own0 = b.own_lo - b.row_lo; g.own1 = b.own_hi - b.row_lo;
  // the residual is needed P rows beyond the owned rows (adjoint support); clip to the rows whose own
  // support (P more rows) is held, or to the frame border where zero fill is the true boundary
  c->fwd0 = (b.row_lo == 0) ? 0 : g.own0 - P;
  c->fwd1 = (b.row_hi == HuG) ? g.Hu : g.own1 + P;
  const size_t pb = 3 * g.plane * sizeof(float);
  c->flag_offset = (pb + 255) & ~size_t(255);
  c->u_alloc_bytes = c->flag_offset + sizeof(Comm);
  if (cudaMalloc(&c->u, c->u_alloc_bytes) != cudaSuccess) { rltv_destroy(c); return fail(RLTV_ERR_ALLOC, "cudaMalloc of the estimate failed"); }
  CU(cudaMemsetAsync(c->u, 0, c->u_alloc_bytes, c->stream));
  float** planes[4] = {&c->ut, &c->gbuf, &c->img, &c->err};
  for (auto p : planes) {
    if (cudaMalloc(p, pb) != cudaSuccess) { rltv_destroy(c); return fail(RLTV_ERR_ALLOC, "cudaMalloc of a frame plane set failed"); }
    CU(cudaMemsetAsync(*p, 0, pb, c->stream));
  }
  CU(cudaMalloc(&c->staging, size_t(g.Hu) * g.Wu * 3 * sizeof(float)));
  const size_t kb = size_t(3) * MK * MK * sizeof(float);
  CU(cudaMalloc(&c->psf, kb));
  CU(cudaMalloc(&c->psf_caller, kb));
  CU(cudaMalloc(&c->psf_hwc, kb));
  CU(cudaMalloc(&c->st, sizeof(State)));
  CU(cudaMemsetAsync(c->st, 0, sizeof(State), c->stream));
  CU(cudaMalloc(&c->counters, 8 * sizeof(unsigned)));
  CU(cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned), c->stream));
  c->peers.nranks = 1;
  c->peers.rank = 0;
  c->peers.peer[0] = reinterpret_cast<Comm*>(reinterpret_cast<char*>(c->u) + c->flag_offset);
  CU(cudaMallocHost(&c->h_st, sizeof(State)));
  CU(cudaMallocHost(&c->h_poll, 4 * 2 * sizeof(int)));
  for (auto& e : c->poll_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CU(cudaEventCreate(&c->ev0));
  CU(cudaEventCreate(&c->ev1));
  {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
  }
  c->gk_nparts = c->num_sms;   // persistent k_gradk: one CTA per SM, one partial per CTA
  CU(cudaMalloc(&c->gk_partial, size_t(3) * c->gk_nparts * MK * MK * sizeof(float)));
  CU(cudaMalloc(&c->gk_sum, size_t(3) * MK * MK * sizeof(double)));
  CU(cudaMalloc(&c->wspec, size_t(2) * 3 * MK * FFT_N * sizeof(float2)));
  {
    CU(cudaMalloc(&c->gkf_part, size_t(c->gk_nparts) * 3 * MK * FFT_N * sizeof(float2)));
  }
  {
    // Row-FFT hybrid stencils (csrc/rltv_stencil_fft.cuh) for the FP32-bound sizes; RLTV_CONV=direct|fft overrides
    const char* e = getenv("RLTV_CONV");
    // K = 9: the chain kernel + fused PSF gradient win on megapixel frames since the five-role chain kernel (C2, 2 MP:
    // 13.2 vs 12.2 GPix*iter/s); small crops keep the direct stencils (no pipeline fill, no piece priming)
    size_t minpix = size_t(1) << 20;                             // RLTV_CHAIN_MINPIX: tests drive the small-K chain path on small frames
    if (const char* mp = getenv("RLTV_CHAIN_MINPIX")) minpix = size_t(atoll(mp));
    c->use_fft = (MK >= 11) || (MK == 9 && size_t(M) * N >= minpix);
    if (e && !strcmp(e, "direct") && MK <= RLTV_MAX_DIRECT_MK) c->use_fft = false;
    if (e && !strcmp(e, "fft") && MK >= 9) c->use_fft = true;
    c->use_fft_gradk = c->use_fft;
    c->fuse_residual = c->use_fft_gradk && MK <= 17;             // GradkFftCfg<K>::CAN_FUSE
    if (const char* e = getenv("RLTV_FUSE")) c->fuse_residual = c->fuse_residual && atoi(e) != 0;
    // spectral chain kernel: default for the sizes it exists for; RLTV_CHAIN=0 keeps the two-kernel gradient path
    c->use_chain = c->use_fft && MK >= 9 && MK <= 17;
    // K = 5, 7 on megapixel frames: the chain kernel does the gradient (it is bound by its FFT / epilogue roles there, not by
    // the 2K complex MACs), everything else stays on the direct stencils (the row-FFT tile kernels start at K = 9)
    if ((MK == 5 || MK == 7) && size_t(M) * N >= minpix && !(e && !strcmp(e, "direct"))) c->use_chain = true;
    if (const char* e = getenv("RLTV_CHAIN")) c->use_chain = c->use_chain && atoi(e) != 0;
  }
  {
    int rc = make_maps(c);
    if (rc) { std::string keep = g_err; rltv_destroy(c); g_err = keep; return rc; }
  }
  if (c->use_chain) {
    int rc = chain_setup(c);
    if (rc) { std::string keep = g_err; rltv_destroy(c); g_err = keep; return rc; }
  }
  CU(cudaStreamSynchronize(c->stream));
  *out = c;
  return RLTV_OK;
}

int rltv_create(rltv_ctx** out, int32_t device, int32_t M, int32_t N, int32_t MK, void* stream) {
  return rltv_create_band(out, device, M, N, MK, nullptr, stream);
}

int rltv_destroy(rltv_ctx* c) {
  if (!c) return RLTV_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  drop_graphs(c);
  for (auto& pb : c->peer_base) if (pb) cudaIpcCloseMemHandle(pb);
  for (auto& pb : c->gather_peer_base) if (pb) cudaIpcCloseMemHandle(pb);
  cudaFree(c->gather_stage);
  float* f[] = {c->u, c->ut, c->gbuf, c->img, c->err, c->staging, c->psf, c->psf_caller, c->psf_hwc,
                c->gk_partial, c->rowmin, c->rowmax, c->mr_out};
  for (auto p : f) cudaFree(p);
  double* d[] = {c->gk_sum, c->wa, c->wb, c->rowacc, c->rowsum};
  for (auto p : d) cudaFree(p);
  cudaFree(c->Z); cudaFree(c->tw); cudaFree(c->st); cudaFree(c->counters); cudaFree(c->wspec); cudaFree(c->gkf_part);
  cudaFree(c->chain_pieces); cudaFree(c->chain_first); cudaFree(c->chain_ipk);
  cudaFree(c->tvut1); cudaFree(c->tvut2); cudaFree(c->tbuf);
  if (c->h_st) cudaFreeHost(c->h_st);
  if (c->h_poll) cudaFreeHost(c->h_poll);
  for (auto& e : c->poll_ev) if (e) cudaEventDestroy(e);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (auto& pool : c->prof_ev) for (auto& p : pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return RLTV_OK;
}

void* rltv_stream(rltv_ctx* c) { return c ? reinterpret_cast<void*>(c->stream) : nullptr; }

int rltv_upload_band(rltv_ctx* c, const float* image, size_t image_rs, int32_t image_row0, int32_t n_image_rows,
                     const float* u, size_t u_rs, const float* psf) {
  int rc = check_ctx(c);
  if (rc) return rc;
  const Geom& g = c->g;
  if (image && n_image_rows > 0) {
    if (image_rs < size_t(g.N) * 12) return fail(RLTV_ERR_ARG, "image row stride smaller than a packed row");
    const int dst = image_row0 + g.P - g.row0;     // local row of the first provided image row
    if (image_row0 < 0 || image_row0 + n_image_rows > g.M || dst < 0 || dst + n_image_rows > g.Hu)
      return fail(RLTV_ERR_ARG, "image rows outside this band");
    CU(cudaMemcpy2DAsync(c->staging, size_t(g.N) * 12, image, image_rs, size_t(g.N) * 12, n_image_rows, cudaMemcpyHostToDevice, c->stream));
    k_hwc_to_planar<<<dim3(hwc_grid(g.N), n_image_rows), 256, 0, c->stream>>>(c->staging, size_t(g.N) * 3, n_image_rows, g.N, c->img, g, dst, g.P);
    c->launches++;
    c->chain_ipk_valid = false;                    // the chain kernel's packed image spectra follow the image
  }
  if (u) {
    if (u_rs < size_t(g.Wu) * 12) return fail(RLTV_ERR_ARG, "u row stride smaller than a packed row");
    CU(cudaMemcpy2DAsync(c->staging, size_t(g.Wu) * 12, u, u_rs, size_t(g.Wu) * 12, g.Hu, cudaMemcpyHostToDevice, c->stream));
    k_hwc_to_planar<<<dim3(hwc_grid(g.Wu), g.Hu), 256, 0, c->stream>>>(c->staging, size_t(g.Wu) * 3, g.Hu, g.Wu, c->u, g, 0, 0);
    c->launches++;
  }
  if (psf) {
    const int KK2 = g.K * g.K;
    CU(cudaMemcpyAsync(c->psf_hwc, psf, size_t(3) * KK2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    k_psf_pack<<<(3 * KK2 + 255) / 256, 256, 0, c->stream>>>(c->psf_hwc, c->psf, KK2, 1);
    CU(cudaMemcpyAsync(c->psf_caller, c->psf, size_t(3) * KK2 * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    c->launches++;
    if (c->use_fft || c->use_chain) {
      CU(cudaMemsetAsync(c->st, 0, sizeof(int), c->stream));   // a stop flag left by an earlier solve must not skip this
      int rc2 = launch_psf_spectrum(c);
      if (rc2) return rc2;
    }
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));   // the host arrays may be pageable views: do not outlive the call
  c->uploaded = true;
  return RLTV_OK;
}

int rltv_upload(rltv_ctx* c, const float* image, size_t image_rs, const float* u, size_t u_rs, const float* psf) {
  if (!c) return fail(RLTV_ERR_ARG, "null context");
  if (c->banded) return fail(RLTV_ERR_STATE, "use rltv_upload_band on a row-band context");
  return rltv_upload_band(c, image, image_rs, 0, image ? c->g.M : 0, u, u_rs, psf);
}

int rltv_download_rows(rltv_ctx* c, float* u, size_t u_rs, int32_t row0, int32_t nrows, float* psf_caller, float* psf_refined) {
  int rc = check_ctx(c);
  if (rc) return rc;
  const Geom& g = c->g;
  if (u && nrows > 0) {
    if (u_rs < size_t(g.Wu) * 12) return fail(RLTV_ERR_ARG, "u row stride smaller than a packed row");
    const int l0 = row0 - g.row0;
    if (l0 < 0 || l0 + nrows > g.Hu) return fail(RLTV_ERR_ARG, "rows not held by this band");
    k_planar_to_hwc<<<dim3(hwc_grid(g.Wu), nrows), 256, 0, c->stream>>>(c->u, g, l0, 0, nrows, g.Wu, c->staging, size_t(g.Wu) * 3);
    c->launches++;
    CU(cudaMemcpy2DAsync(u, u_rs, c->staging, size_t(g.Wu) * 12, size_t(g.Wu) * 12, nrows, cudaMemcpyDeviceToHost, c->stream));
  }
  const int KK2 = g.K * g.K;
  float* dsts[2] = {psf_caller, psf_refined};
  float* srcs[2] = {c->psf_caller, c->psf};
  for (int i = 0; i < 2; ++i) {
    if (!dsts[i]) continue;
    k_psf_pack<<<(3 * KK2 + 255) / 256, 256, 0, c->stream>>>(c->psf_hwc, srcs[i], KK2, 0);
    c->launches++;
    CU(cudaMemcpyAsync(dsts[i], c->psf_hwc, size_t(3) * KK2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  return RLTV_OK;
}

int rltv_download(rltv_ctx* c, float* u, size_t u_rs, float* psf_caller, float* psf_refined) {
  if (!c) return fail(RLTV_ERR_ARG, "null context");
  return rltv_download_rows(c, u, u_rs, c->row_lo, u ? c->g.Hu : 0, psf_caller, psf_refined);
}

int rltv_begin(rltv_ctx* c, const rltv_params_t* p) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!p) return fail(RLTV_ERR_ARG, "null params");
  if (!c->uploaded) return fail(RLTV_ERR_STATE, "rltv_upload must precede rltv_begin/rltv_solve");
  if (p->iterations < 0) return fail(RLTV_ERR_ARG, "negative iteration count");
  if (p->mode != RLTV_MODE_MM && p->mode != RLTV_MODE_MM_TV && p->mode != RLTV_MODE_PAM_CTV) return fail(RLTV_ERR_ARG, "unknown solver mode");
  if (p->mode != RLTV_MODE_MM) {
    if (c->banded) return fail(RLTV_ERR_STATE, "the TV-alive modes need a whole-frame context (the denoised image is not exchanged between row bands)");
    const size_t pb = 3 * c->g.plane * sizeof(float);
    if (!c->tvut1) {
      CU(cudaMalloc(&c->tvut1, pb));
      CU(cudaMalloc(&c->tvut2, pb));
      CU(cudaMalloc(&c->tbuf, pb));
    }
  }
  c->params = *p;
  drop_graphs(c);                                  // kernel arguments of the previous solve are baked into them
  {
    const char* e = getenv("RLTV_GRAPH");
    c->use_graph = !c->banded && size_t(c->g.M) * c->g.N <= (size_t(1) << 22);     // <= 4 MP: launch-bound
    if (e) c->use_graph = atoi(e) != 0 && !c->banded;
  }
  rc = setup_whiteness(c, p->top, p->bottom, p->left, p->right);
  if (rc) return rc;
  CU(cudaMemsetAsync(c->st, 0, sizeof(State), c->stream));
  k_state_init<<<1, 32, 0, c->stream>>>(c->st);
  c->inner_count = 0;
  c->inner_in_outer = 0;
  c->outer_enqueued = 0;
  c->outer_since_begin = 0;
  c->launches = 0;
  for (int f = 0; f < F_COUNT; ++f) { c->prof_ms[f] = 0.f; c->prof_n[f] = 0; c->prof_used[f] = 0; }
  c->begun = true;
  CU(cudaEventRecord(c->ev0, c->stream));
  return RLTV_OK;
}

int rltv_enqueue_phase(rltv_ctx* c, int32_t phase) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!c->begun) return fail(RLTV_ERR_STATE, "rltv_begin must precede rltv_enqueue_phase");
  return enqueue_phase(c, phase);
}

int rltv_enqueue_outer(rltv_ctx* c, int32_t n_outer) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!c->begun) return fail(RLTV_ERR_STATE, "rltv_begin must precede rltv_enqueue_outer");
  if (c->banded && !c->fused_comm) return fail(RLTV_ERR_STATE, "NCCL-baseline row bands are stepped phase by phase (rltv_enqueue_phase)");
  for (int i = 0; i < n_outer; ++i) {
    if ((rc = enqueue_outer_fast(c)) != RLTV_OK) return rc;
    c->outer_enqueued++;
  }
  return RLTV_OK;
}

int rltv_poll_record(rltv_ctx* c, int32_t it) {
  int rc = check_ctx(c);
  if (rc) return rc;
  const int slot = it & 3;
  CU(cudaMemcpyAsync(c->h_poll + 2 * slot, c->st, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaEventRecord(c->poll_ev[slot], c->stream));
  return RLTV_OK;
}

int rltv_poll_wait(rltv_ctx* c, int32_t it, int32_t* stop) {
  int rc = check_ctx(c);
  if (rc) return rc;
  const int slot = it & 3;
  CU(cudaEventSynchronize(c->poll_ev[slot]));
  if (stop) *stop = c->h_poll[2 * slot];
  return RLTV_OK;
}

int rltv_finish(rltv_ctx* c, rltv_stats_t* stats) {
  int rc = check_ctx(c);
  if (rc) return rc;
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaMemcpyAsync(c->h_st, c->st, sizeof(State), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  if (c->prof) prof_collect(c);
  fill_stats(c, stats, ms);
  return RLTV_OK;
}

int rltv_solve(rltv_ctx* c, const rltv_params_t* p, rltv_stats_t* stats) {
  if (c && c->banded && !c->fused_comm) return fail(RLTV_ERR_STATE, "NCCL-baseline row bands are driven by the distributed host loop");
  int rc = rltv_begin(c, p);
  if (rc) return rc;
  // Host runs at most two outer iterations ahead of the device-side stop flag.
  for (int it = 0; it < p->iterations; ++it) {
    if (it >= 2) {
      int stop = 0;
      if ((rc = rltv_poll_wait(c, it - 2, &stop)) != RLTV_OK) return rc;
      if (stop) break;
    }
    if ((rc = enqueue_outer_fast(c)) != RLTV_OK) return rc;
    c->outer_enqueued++;
    if ((rc = rltv_poll_record(c, it)) != RLTV_OK) return rc;
  }
  return rltv_finish(c, stats);
}

int rltv_profile_enable(rltv_ctx* c, int32_t on) {
  if (!c) return fail(RLTV_ERR_ARG, "null context");
  c->prof = on != 0;
  return RLTV_OK;
}

int rltv_profile_get(rltv_ctx* c, const char* family, float* total_ms, int32_t* launches) {
  if (!c || !family) return fail(RLTV_ERR_ARG, "null argument");
  for (int f = 0; f < F_COUNT; ++f)
    if (std::strcmp(family, kFamilyNames[f]) == 0) {
      if (total_ms) *total_ms = c->prof_ms[f];
      if (launches) *launches = c->prof_n[f];
      return RLTV_OK;
    }
  return fail(RLTV_ERR_ARG, "unknown kernel family");
}

// ---- row bands -----------------------------------------------------------------------------------------
int rltv_ipc_export(rltv_ctx* c, void* handle64) {
  int rc = check_ctx(c);
  if (rc) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, c->u));
  std::memcpy(handle64, &h, 64);
  return RLTV_OK;
}

int rltv_set_rank(rltv_ctx* c, int32_t rank, int32_t world, int32_t fused) {
  if (!c) return fail(RLTV_ERR_ARG, "null context");
  if (world < 1 || world > MAXR || rank < 0 || rank >= world) return fail(RLTV_ERR_ARG, "bad rank/world (at most 8 bands)");
  c->rank = rank;
  c->world_rank = rank;
  c->world = world;
  c->fused_comm = fused != 0;
  Comm* mine = reinterpret_cast<Comm*>(reinterpret_cast<char*>(c->u) + c->flag_offset);
  for (auto& p : c->peers.peer) p = nullptr;
  c->peers.peer[c->fused_comm ? rank : 0] = mine;
  c->peers.rank = c->fused_comm ? rank : 0;
  c->peers.nranks = c->fused_comm ? world : 1;
  if (!c->fused_comm) c->rank = 0;   // kernels index peers.peer[c->rank]; the NCCL baseline never exchanges in-kernel
  if (!c->fused_comm && world > 1) c->use_chain = false;   // the host all-reduces statistics slot 0 between two kernels
  return RLTV_OK;
}

int rltv_ipc_attach(rltv_ctx* c, int32_t peer_rank, const void* handle64, int32_t peer_row_lo, int32_t peer_row_hi) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (peer_rank < 0 || peer_rank >= MAXR || !handle64) return fail(RLTV_ERR_ARG, "bad peer rank/handle");
  const int my_rank = c->fused_comm ? c->rank : c->peers.rank;
  (void)my_rank;
  const int P2 = 2 * c->g.P;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* base = nullptr;
  CU(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  c->peer_base[peer_rank] = base;
  const size_t peer_plane = size_t(peer_row_hi - peer_row_lo) * c->g.pitch;
  const size_t peer_flag_off = (3 * peer_plane * sizeof(float) + 255) & ~size_t(255);
  Comm* pc = reinterpret_cast<Comm*>(reinterpret_cast<char*>(base) + peer_flag_off);
  if (c->fused_comm) c->peers.peer[peer_rank] = pc;
  // neighbours also exchange halos: the band above has row_hi inside my band, the band below has row_lo inside
  // (by RANK: comparing the clipped held-row ranges misses a neighbour when both bands' halos are clipped by the frame
  // border, e.g. a first band that owns exactly 2P rows)
  int side = -1;
  if (peer_rank == c->world_rank - 1) side = 0;        // band above
  else if (peer_rank == c->world_rank + 1) side = 1;   // band below
  if (side < 0) return RLTV_OK;
  HaloSide& s = c->side[side];
  s.peer_u = reinterpret_cast<float*>(base);
  s.peer_plane = peer_plane;
  s.nrows = P2;
  if (side == 0) {
    // band above: its bottom halo = frame rows [own_lo, own_lo + 2P) = my first owned rows; I am ITS lower neighbour
    s.src_row = c->own_lo - c->row_lo;
    s.dst_row = c->own_lo - peer_row_lo;
    s.peer_flag = &pc->halo_flag[1];
  } else {
    // band below: its top halo = frame rows [own_hi - 2P, own_hi) = my last owned rows; I am ITS upper neighbour
    s.src_row = c->own_hi - P2 - c->row_lo;
    s.dst_row = c->own_hi - P2 - peer_row_lo;
    s.peer_flag = &pc->halo_flag[0];
  }
  if (s.src_row < c->g.own0 || s.src_row + s.nrows > c->g.own1 || s.dst_row < 0 || s.dst_row + s.nrows > peer_row_hi - peer_row_lo)
    return fail(RLTV_ERR_ARG, "halo rows do not fit: every band must own at least 2*(MK/2) rows");
  return RLTV_OK;
}

int rltv_download_image(rltv_ctx* c, float* image, size_t image_rs) {
  int rc = check_ctx(c);
  if (rc) return rc;
  const Geom& g = c->g;
  if (!c->uploaded || c->banded || !image) return fail(RLTV_ERR_STATE, "needs an uploaded whole-frame context");
  if (image_rs < size_t(g.N) * 12) return fail(RLTV_ERR_ARG, "image row stride smaller than a packed row");
  k_planar_to_hwc<<<dim3(hwc_grid(g.N), g.M), 256, 0, c->stream>>>(c->img, g, g.P, g.P, g.M, g.N, c->staging, size_t(g.N) * 3);
  c->launches++;
  CU(cudaMemcpy2DAsync(image, image_rs, c->staging, size_t(g.N) * 12, size_t(g.N) * 12, g.M, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return RLTV_OK;
}

// ---- device-side gather of the result bands ------------------------------------------------------------
int rltv_gather_alloc(rltv_ctx* c, void* handle64_out) {
  int rc = check_ctx(c);
  if (rc) return rc;
  const Geom& g = c->g;
  const size_t HuG = size_t(g.M) + g.K - 1;
  if (!c->gather_stage) CU(cudaMalloc(&c->gather_stage, HuG * g.Wu * 3 * sizeof(float)));
  if (handle64_out) {
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->gather_stage));
    std::memcpy(handle64_out, &h, 64);
  }
  return RLTV_OK;
}

int rltv_gather_attach(rltv_ctx* c, int32_t peer_rank, const void* handle64) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (peer_rank < 0 || peer_rank >= MAXR || !handle64) return fail(RLTV_ERR_ARG, "bad peer rank/handle");
  if (c->gather_peer_base[peer_rank]) return RLTV_OK;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* base = nullptr;
  CU(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  c->gather_peer_base[peer_rank] = base;
  c->gather_peer[peer_rank] = reinterpret_cast<float*>(base);
  return RLTV_OK;
}

int rltv_gather_push(rltv_ctx* c, uint32_t dst_rank_mask) {
  int rc = check_ctx(c);
  if (rc) return rc;
  const Geom& g = c->g;
  const int nrows = c->own_hi - c->own_lo;
  for (int r = 0; r < MAXR; ++r) {
    if (!(dst_rank_mask & (1u << r))) continue;
    float* dst = (r == c->world_rank) ? c->gather_stage : c->gather_peer[r];
    if (!dst) return fail(RLTV_ERR_STATE, "gather buffer of the destination rank is not allocated / attached");
    // owned rows, planar -> packed HWC rows of the FULL frame at frame row own_lo
    k_planar_to_hwc<<<dim3(hwc_grid(g.Wu), nrows), 256, 0, c->stream>>>(c->u, g, g.own0, 0, nrows, g.Wu,
                                                                       dst + size_t(c->own_lo) * g.Wu * 3, size_t(g.Wu) * 3);
    c->launches++;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));          // the caller's barrier may follow immediately
  return RLTV_OK;
}

int rltv_gather_download(rltv_ctx* c, float* u, size_t u_rs) {
  int rc = check_ctx(c);
  if (rc) return rc;
  const Geom& g = c->g;
  if (!c->gather_stage || !u) return fail(RLTV_ERR_STATE, "no gather buffer / destination");
  if (u_rs < size_t(g.Wu) * 12) return fail(RLTV_ERR_ARG, "u row stride smaller than a packed row");
  const size_t HuG = size_t(g.M) + g.K - 1;
  CU(cudaMemcpy2DAsync(u, u_rs, c->gather_stage, size_t(g.Wu) * 12, size_t(g.Wu) * 12, HuG, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RLTV_OK;
}

int rltv_set_ignore_stop(rltv_ctx* c, int32_t on) {
  if (!c) return fail(RLTV_ERR_ARG, "null context");
  c->ignore_stop = on != 0;
  return RLTV_OK;
}

int rltv_set_whiteness_owner(rltv_ctx* c, int32_t owner) {
  if (!c) return fail(RLTV_ERR_ARG, "null context");
  c->white_owner = owner != 0;
  return RLTV_OK;
}

void* rltv_device_ptr(rltv_ctx* c, const char* name, size_t* nbytes) {
  if (!c || !name) return nullptr;
  size_t n = 0;
  void* p = nullptr;
  if (!std::strcmp(name, "step_max")) { p = c->st->smax[0]; n = 6 * sizeof(int); }
  else if (!std::strcmp(name, "gk_sum")) { p = c->gk_sum; n = size_t(3) * c->g.K * c->g.K * sizeof(double); }
  else if (!std::strcmp(name, "stop")) { p = &c->st->stop; n = sizeof(int); }
  if (nbytes) *nbytes = n;
  return p;
}

// ---- stage-level entry points ------------------------------------------------------------------------
int rltv_stage_residual(rltv_ctx* c, float* err_out) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!c->uploaded) return fail(RLTV_ERR_STATE, "upload first");
  if (c->banded) return fail(RLTV_ERR_STATE, "stage entry points need a whole-frame context");
  const Geom& g = c->g;
  CU(cudaMemsetAsync(c->st, 0, sizeof(State), c->stream));
  if ((rc = launch_conv_fwd(c)) != RLTV_OK) return rc;
  if (err_out) {
    k_planar_to_hwc<<<dim3(8, g.M), 256, 0, c->stream>>>(c->err, g, g.P, g.P, g.M, g.N, c->staging, size_t(g.N) * 3);
    CU(cudaMemcpyAsync(err_out, c->staging, size_t(g.M) * g.N * 12, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return RLTV_OK;
}

int rltv_stage_adjoint(rltv_ctx* c, float* g_out) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!c->uploaded) return fail(RLTV_ERR_STATE, "upload first");
  if (c->banded) return fail(RLTV_ERR_STATE, "stage entry points need a whole-frame context");
  const Geom& g = c->g;
  CU(cudaMemcpyAsync(c->ut, c->u, 3 * g.plane * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = launch_conv_adj(c, 1.0f)) != RLTV_OK) return rc;
  if (g_out) {
    k_planar_to_hwc<<<dim3(8, g.Hu), 256, 0, c->stream>>>(c->gbuf, g, 0, 0, g.Hu, g.Wu, c->staging, size_t(g.Wu) * 3);
    CU(cudaMemcpyAsync(g_out, c->staging, size_t(g.Hu) * g.Wu * 12, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return RLTV_OK;
}

int rltv_stage_chain(rltv_ctx* c, float* g_out, float* max6) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!c->uploaded) return fail(RLTV_ERR_STATE, "upload first");
  if (c->banded) return fail(RLTV_ERR_STATE, "stage entry points need a whole-frame context");
  if (!c->use_chain) return fail(RLTV_ERR_STATE, "this context does not use the chain kernel (MK outside 11..17 or RLTV_CHAIN=0)");
  const Geom& g = c->g;
  CU(cudaMemsetAsync(c->st, 0, sizeof(State), c->stream));
  k_state_init<<<1, 32, 0, c->stream>>>(c->st);
  c->inner_count = 0;
  c->inner_in_outer = 0;                           // ut = u
  if (!c->chain_ipk_valid && (rc = chain_image_spectra(c)) != RLTV_OK) return rc;
  if ((rc = launch_chain(c, 1.0f)) != RLTV_OK) return rc;
  if (g_out) {
    k_planar_to_hwc<<<dim3(8, g.Hu), 256, 0, c->stream>>>(c->gbuf, g, 0, 0, g.Hu, g.Wu, c->staging, size_t(g.Wu) * 3);
    CU(cudaMemcpyAsync(g_out, c->staging, size_t(g.Hu) * g.Wu * 12, cudaMemcpyDeviceToHost, c->stream));
  }
  int h[6];
  CU(cudaMemcpyAsync(h, c->st->smax[0], sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  if (max6)
    for (int i = 0; i < 6; ++i) {
      const int o = h[i];
      const int b = o >= 0 ? o : (o ^ 0x7fffffff);
      std::memcpy(&max6[i], &b, 4);
    }
  return RLTV_OK;
}

int rltv_debug_download_err(rltv_ctx* c, float* err_out) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!c->uploaded || c->banded || !err_out) return fail(RLTV_ERR_STATE, "needs an uploaded whole-frame context");
  const Geom& g = c->g;
  k_planar_to_hwc<<<dim3(8, g.M), 256, 0, c->stream>>>(c->err, g, g.P, g.P, g.M, g.N, c->staging, size_t(g.N) * 3);
  CU(cudaMemcpyAsync(err_out, c->staging, size_t(g.M) * g.N * 12, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return RLTV_OK;
}

int rltv_stage_gradk(rltv_ctx* c, float* gk_out) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!c->uploaded) return fail(RLTV_ERR_STATE, "upload first");
  if (c->banded) return fail(RLTV_ERR_STATE, "stage entry points need a whole-frame context");
  const int K = c->g.K, KK2 = K * K;
  if ((rc = launch_gradk(c)) != RLTV_OK) return rc;
  std::vector<double> s(size_t(3) * KK2);
  CU(cudaMemcpyAsync(s.data(), c->gk_sum, s.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  // the K-1-q index flip, exactly as k_psf_update does it
  for (int ch = 0; ch < 3; ++ch)
    for (int q = 0; q < KK2; ++q) {
      const int qy = q / K, qx = q % K;
      gk_out[q * 3 + ch] = float(s[size_t(ch) * KK2 + (K - 1 - qy) * K + (K - 1 - qx)]);
    }
  return RLTV_OK;
}

int rltv_stage_whiteness(rltv_ctx* c, int32_t top, int32_t bottom, int32_t left, int32_t right, float* M_r) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!c->uploaded) return fail(RLTV_ERR_STATE, "upload first");
  if (c->banded) return fail(RLTV_ERR_STATE, "stage entry points need a whole-frame context");
  if ((rc = setup_whiteness(c, top, bottom, left, right)) != RLTV_OK) return rc;
  if ((rc = launch_whiteness(c, 0)) != RLTV_OK) return rc;
  float v = 0.f;
  CU(cudaMemcpyAsync(&v, c->mr_out, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  if (M_r) *M_r = v;
  return RLTV_OK;
}

// TV(u, out, ..., epsilon, order, norm, div) of pyx:137-239 on the estimate currently on the device.
// `out` goes to the g buffer, `div` to the err buffer (both are scratch between solves).
int rltv_stage_tv(rltv_ctx* c, int32_t order, int32_t norm, float epsilon, float* out_hwc, float* div_hwc, float* ms) {
  int rc = check_ctx(c);
  if (rc) return rc;
  if (!c->uploaded) return fail(RLTV_ERR_STATE, "upload first");
  if (c->banded) return fail(RLTV_ERR_STATE, "stage entry points need a whole-frame context");
  if ((order != 1 && order != 2) || (norm != 1 && norm != 2)) return fail(RLTV_ERR_ARG, "order and norm must be 1 or 2");
  const Geom& g = c->g;
  dim3 grid((g.pitch / 4 + 255) / 256, g.Hu, 3);
  CU(cudaEventRecord(c->ev0, c->stream));
  if (order == 2 && norm == 1) k_tv<2, 1><<<grid, 256, 0, c->stream>>>(g, c->u, epsilon, c->gbuf, c->err);
  if (order == 2 && norm == 2) k_tv<2, 2><<<grid, 256, 0, c->stream>>>(g, c->u, epsilon, c->gbuf, c->err);
  if (order == 1 && norm == 1) k_tv<1, 1><<<grid, 256, 0, c->stream>>>(g, c->u, epsilon, c->gbuf, c->err);
  if (order == 1 && norm == 2) k_tv<1, 2><<<grid, 256, 0, c->stream>>>(g, c->u, epsilon, c->gbuf, c->err);
  CU(cudaEventRecord(c->ev1, c->stream));
  float* srcs[2] = {c->gbuf, c->err};
  float* dsts[2] = {out_hwc, div_hwc};
  for (int i = 0; i < 2; ++i) {
    if (!dsts[i]) continue;
    k_planar_to_hwc<<<dim3(hwc_grid(g.Wu), g.Hu), 256, 0, c->stream>>>(srcs[i], g, 0, 0, g.Hu, g.Wu, c->staging, size_t(g.Wu) * 3);
    CU(cudaMemcpyAsync(dsts[i], c->staging, size_t(g.Hu) * g.Wu * 12, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  if (ms) cudaEventElapsedTime(ms, c->ev0, c->ev1);
  // the residual buffer was used as scratch: restore its zero ring for the next solve
  CU(cudaMemsetAsync(c->err, 0, 3 * g.plane * sizeof(float), c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RLTV_OK;
}

// Debug: skip roles of the chain kernel (bit mask, see g_chain_skip_roles) -- timing experiments only.
int rltv_debug_chain_roles(int32_t mask) {
  CU(cudaMemcpyToSymbol(g_chain_skip_roles, &mask, sizeof(int)));
  return RLTV_OK;
}

// Debug: per-phase cycle totals of k_conv_fft since the last call (thread 0 of every CTA), then reset.
int rltv_debug_phase_cycles(uint64_t* out8) {
  unsigned long long h[8] = {0};
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(h, g_fft_phase_cycles, sizeof(h)));
  unsigned long long z[8] = {0};
  CU(cudaMemcpyToSymbol(g_fft_phase_cycles, z, sizeof(z)));
  for (int i = 0; i < 8; ++i) out8[i] = h[i];
  return RLTV_OK;
}

// Test entry point of the shared-memory FFT engine (csrc/rltv_fft.cuh): nrows x 128 complex (interleaved re,im).
int rltv_debug_fft128(const float* in, float* out, int32_t nrows, int32_t inverse, int32_t device) {
  if (!in || !out || nrows < 1) return fail(RLTV_ERR_ARG, "bad arguments");
  if (rltv_device_count() <= 0) return fail(RLTV_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  CU(cudaSetDevice(device));
  float2 *d_in = nullptr, *d_out = nullptr;
  const size_t bytes = size_t(nrows) * FFT_N * sizeof(float2);
  CU(cudaMalloc(&d_in, bytes));
  CU(cudaMalloc(&d_out, bytes));
  CU(cudaMemcpy(d_in, in, bytes, cudaMemcpyHostToDevice));
  if (inverse) k_fft128_debug<true><<<(nrows + 15) / 16, 128>>>(d_in, d_out, nrows);
  else k_fft128_debug<false><<<(nrows + 15) / 16, 128>>>(d_in, d_out, nrows);
  cudaError_t e = cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost);
  cudaFree(d_in);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(RLTV_ERR_CUDA, cudaGetErrorString(e));
  return RLTV_OK;
}

// ---- one-shot drop-ins -------------------------------------------------------------------------------
int rltv_richardson_lucy_mm(const float* image, size_t image_rs, float* u, size_t u_rs, float* psf, int32_t M,
                            int32_t N, int32_t MK, const rltv_params_t* params, rltv_stats_t* stats,
                            float* psf_refined, int32_t device) {
  if (!image || !u || !psf || !params) return fail(RLTV_ERR_ARG, "null argument");
  rltv_ctx* c = nullptr;
  int rc = rltv_create(&c, device, M, N, MK, nullptr);
  if (rc) return rc;
  rc = rltv_upload(c, image, image_rs, u, u_rs, psf);
  if (!rc) rc = rltv_solve(c, params, stats);
  if (!rc) rc = rltv_download(c, u, u_rs, psf, psf_refined);
  std::string keep = g_err;
  rltv_destroy(c);
  g_err = keep;
  return rc;
}

int rltv_normalize_kernel(float* kern, int32_t MK, int32_t device) {
  if (!kern || MK < 1) return fail(RLTV_ERR_ARG, "bad kernel");
  if (rltv_device_count() <= 0) return fail(RLTV_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  CU(cudaSetDevice(device));
  const int KK2 = MK * MK;
  float *d_hwc = nullptr, *d_pl = nullptr;
  CU(cudaMalloc(&d_hwc, size_t(3) * KK2 * sizeof(float)));
  CU(cudaMalloc(&d_pl, size_t(3) * KK2 * sizeof(float)));
  CU(cudaMemcpy(d_hwc, kern, size_t(3) * KK2 * sizeof(float), cudaMemcpyHostToDevice));
  k_psf_pack<<<(3 * KK2 + 255) / 256, 256>>>(d_hwc, d_pl, KK2, 1);
  k_normalize_kernel<<<1, 128>>>(d_pl, KK2);
  k_psf_pack<<<(3 * KK2 + 255) / 256, 256>>>(d_hwc, d_pl, KK2, 0);
  cudaError_t e = cudaMemcpy(kern, d_hwc, size_t(3) * KK2 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d_hwc);
  cudaFree(d_pl);
  if (e != cudaSuccess) return fail(RLTV_ERR_CUDA, cudaGetErrorString(e));
  return RLTV_OK;
}

}  // extern "C"
