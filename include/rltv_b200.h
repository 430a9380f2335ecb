/*
 * rltv_b200.h -- C ABI of the B200-native Richardson-Lucy / MM deconvolution hot path.
 *
 * Drop-in boundary for the solver of aurelienpierre/Image-Cases-Studies:
 *   lib/deconvolution.pyx:341-342  cpdef richardson_lucy_MM(image, u, psf, top, bottom, left, right, tau,
 *                                        M, N, C, MK, iterations, step_factor, lambd, blind, correlation, ...)
 *   lib/deconvolution.pyx:73-75    cpdef normalize_kernel(kern, MK)
 * The reference has no FFI of its own (it is a Cython module called from Python,
 * deconvolve.py:21,:250,:277,:291,:304); the entry points below are what a ctypes/cffi binding of that
 * module binds instead.  Plain pointers, sizes and byte strides only.
 *
 * Conventions
 *   - all pixel data is float32; host arrays are HWC with packed RGB pixels (channel stride 4 bytes,
 *     pixel stride 12 bytes) and an arbitrary ROW stride in bytes (the reference's callers pass row-sliced
 *     views, deconvolve.py:277-286);
 *   - `u` has shape (M + MK - 1, N + MK - 1, 3): the estimate with its pad = MK/2 ring (pyx:376);
 *   - `psf` has shape (MK, MK, 3), packed;
 *   - every function returns 0 on success or a negative rltv_status; rltv_last_error() gives the text.
 *     No exceptions, no torch types cross this boundary.  There is no CPU fallback: without a CUDA device
 *     every compute entry point returns RLTV_ERR_CUDA.
 */
#ifndef RLTV_B200_H
#define RLTV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLTV_ABI_VERSION 2
#define RLTV_INNER_ITER 5       /* pyx:375 */
#define RLTV_MAX_MK 47          /* odd MK in [3, 47]: direct stencils up to 9, chain kernel 11..17, row-FFT stencils above
                                 * (the reference itself has no cap; its own job list goes up to 45, deconvolve.py:409) */
#define RLTV_MAX_HISTORY 4096   /* outer iterations whose M_r is kept in rltv_stats_t.M_r_history */

typedef enum {
  RLTV_OK = 0,
  RLTV_ERR_ARG = -1,      /* bad shape / parameter: the Python wrapper raises ValueError */
  RLTV_ERR_CUDA = -2,     /* CUDA runtime error (including "no device") */
  RLTV_ERR_ALLOC = -3,
  RLTV_ERR_STATE = -4     /* call order (e.g. solve before upload) */
} rltv_status;

/* Solver parameters: the scalar arguments of richardson_lucy_MM (pyx:341-342). */
typedef struct {
  int32_t top, bottom, left, right; /* whiteness window in image coordinates (pyx:627) */
  float tau;                        /* non-blind stop threshold (pyx:652) */
  int32_t iterations;               /* max OUTER iterations; each runs RLTV_INNER_ITER inner steps */
  float step_factor;                /* pyx:524, :574 */
  float lambd;                      /* pyx:502, :519 */
  int32_t blind;                    /* pyx:555 */
  int32_t correlation;              /* pyx:584-585 (channel-mean PSF) */
  int32_t mode;                     /* RLTV_MODE_MM: the reference's shipped arithmetic (TV term dead) -- the
                                     * pinned drop-in.  RLTV_MODE_MM_TV: the TV branches of pyx:516-517 / :542-549 alive, as in
                                     * the "patched reference" of oracle/build_ref_tv.py; the blurry image is denoised in place
                                     * (read it back with rltv_download_image).  Whole-frame contexts only. */
} rltv_params_t;
#define RLTV_MODE_MM 0
#define RLTV_MODE_MM_TV 1
#define RLTV_MODE_PAM_CTV 2      /* UNPINNED extension (no reference code exists for it, README.md:42-44, :113-117 only): PAM u-step
                                  * u -= dt (lambda g - div p) with the collaborative l-inf,1,1 TV sub-gradient p over the three
                                  * channels (definition: oracle/ctv_oracle.py); DoF blend and PSF step as in the reference */

typedef struct {
  int32_t iterations_executed;      /* outer iterations run (pyx:656) */
  int32_t stopped;                  /* stop_flag (pyx:647,:653) */
  float M_r, M_r_prev;              /* whiteness statistic of the last two outer iterations (pyx:638) */
  float dt[3];                      /* last image step sizes (pyx:524) */
  float dtpsf;                      /* last PSF step size (pyx:574) */
  float solve_ms;                   /* device time of the solve (CUDA events on the solver stream) */
  int32_t kernel_launches;          /* kernels launched by this solve */
  int32_t n_history;
  float M_r_history[RLTV_MAX_HISTORY];
} rltv_stats_t;

typedef struct rltv_ctx rltv_ctx;

int rltv_abi_version(void);
const char* rltv_last_error(void);
int rltv_device_count(void);      /* number of CUDA devices, 0 if none, <0 on error */

/* ---- one-shot host entry points (exact drop-ins) -------------------------------------------------- */

/* richardson_lucy_MM (pyx:341-675).  `u` and `psf` are updated IN PLACE (pyx:531,:552,:581); the caller's
 * result is the view u[pad:pad+M, pad:pad+N] (pyx:675).  With `correlation` set, `psf` receives what the
 * reference leaves in the caller's array: one un-normalised step (pyx:581), because pyx:585 rebinds the
 * local name; the solver's own refined PSF is returned through `psf_refined` when it is not NULL. */
int rltv_richardson_lucy_mm(const float* image, size_t image_row_stride_bytes,
                            float* u, size_t u_row_stride_bytes,
                            float* psf,
                            int32_t M, int32_t N, int32_t MK,
                            const rltv_params_t* params, rltv_stats_t* stats,
                            float* psf_refined, int32_t device);

/* normalize_kernel (pyx:73-75 -> :47-70): clip negatives, divide each channel by its sum; in place. */
int rltv_normalize_kernel(float* kern, int32_t MK, int32_t device);

/* ---- persistent context (device-resident state; what bench.py times and the multi-GPU driver uses) -- */

/* `stream` is a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL for a private one.
 * The context owns planar device copies of image/u/ut/g/err for an (M,N,MK) problem. */
int rltv_create(rltv_ctx** ctx, int32_t device, int32_t M, int32_t N, int32_t MK, void* stream);
int rltv_destroy(rltv_ctx* ctx);
int rltv_upload(rltv_ctx* ctx, const float* image, size_t image_row_stride_bytes,
                const float* u, size_t u_row_stride_bytes, const float* psf);
int rltv_download(rltv_ctx* ctx, float* u, size_t u_row_stride_bytes, float* psf_caller, float* psf_refined);
/* Runs up to params->iterations outer iterations on the device-resident state (resets the iteration
 * counter and the stop flag first).  Synchronises the stream before returning. */
int rltv_solve(rltv_ctx* ctx, const rltv_params_t* params, rltv_stats_t* stats);
/* Enqueue `n_outer` outer iterations without synchronising and without touching the counters
 * (used by bench.py to time steady-state steps with CUDA events on the same stream). */
int rltv_begin(rltv_ctx* ctx, const rltv_params_t* params);
int rltv_enqueue_outer(rltv_ctx* ctx, int32_t n_outer);
int rltv_finish(rltv_ctx* ctx, rltv_stats_t* stats);
void* rltv_stream(rltv_ctx* ctx);
/* Time (ms, CUDA events on the context's stream) and launch count of each kernel family since the last
 * rltv_begin, for bench.py's roofline.  names: "conv_fwd","conv_adj","update","gradk","psf","stats","copy". */
/* Benchmark stepping only: the stop rule (pyx:643-654) is still evaluated every outer iteration but not acted
 * upon, so that steady-state iterations can be timed on inputs whose rule fires after 3 iterations. */
int rltv_set_ignore_stop(rltv_ctx* ctx, int32_t on);
int rltv_profile_enable(rltv_ctx* ctx, int32_t on);
int rltv_profile_get(rltv_ctx* ctx, const char* family, float* total_ms, int32_t* launches);

/* ---- row-band sharding: one context per GPU holds a band of rows of ONE frame ----------- */
/* Rows are rows of the padded estimate u (0 .. M+MK-2).  A band HOLDS rows [row_lo, row_hi) and OWNS
 * [own_lo, own_hi) of them; the rest are halos of width 2*(MK/2) copied from the neighbouring bands.  The host
 * side (image_cases_studies_b200/distributed.py) plans the bands, exchanges IPC handles and issues the
 * NCCL all-reduces between phases; this library does the halo exchange itself with peer stores over NVLink. */
typedef struct {
  int32_t row_lo, row_hi;   /* rows held */
  int32_t own_lo, own_hi;   /* rows owned (updated here, counted in reductions) */
} rltv_band_t;

int rltv_create_band(rltv_ctx** ctx, int32_t device, int32_t M, int32_t N, int32_t MK, const rltv_band_t* band,
                     void* stream);
/* image_rows: `n_image_rows` rows of the blurry image starting at image row `image_row0`; u_rows: rows
 * row_lo .. row_hi-1 of u. */
int rltv_upload_band(rltv_ctx* ctx, const float* image_rows, size_t image_row_stride_bytes, int32_t image_row0,
                     int32_t n_image_rows, const float* u_rows, size_t u_row_stride_bytes, const float* psf);
/* copies u rows [row0, row0+nrows) (frame coordinates, must be held by this band) to the host */
int rltv_download_rows(rltv_ctx* ctx, float* u_rows, size_t u_row_stride_bytes, int32_t row0, int32_t nrows,
                       float* psf_caller, float* psf_refined);
/* rank / world of this band; fused != 0: step scalars, PSF-gradient sums and the stop flag are exchanged by the
 * kernels themselves through peer memory (no NCCL, rltv_enqueue_outer / rltv_solve work on the band);
 * fused == 0: the host all-reduces rltv_device_ptr buffers between phases (NCCL baseline).  Call before attach. */
int rltv_set_rank(rltv_ctx* ctx, int32_t rank, int32_t world, int32_t fused);
/* CUDA IPC: `handle` is a 64-byte cudaIpcMemHandle_t of a band's u allocation (which also carries its halo flags
 * and all-gather slots).  Attach every other band (fused) or at least the two neighbours (NCCL baseline). */
int rltv_ipc_export(rltv_ctx* ctx, void* handle64);
int rltv_ipc_attach(rltv_ctx* ctx, int32_t peer_rank, const void* handle64, int32_t peer_row_lo, int32_t peer_row_hi);
/* exactly one band evaluates the whiteness statistic / stop rule (the one holding the window rows) */
int rltv_set_whiteness_owner(rltv_ctx* ctx, int32_t owner);
/* one outer iteration = BEGIN, 5 x (GRAD, [all-reduce MAX step_max], UPDATE, blind: PSF_GRAD,
 * [all-reduce SUM gk_sum], PSF_STEP), END, [all-reduce MAX stop] */
enum { RLTV_PH_OUTER_BEGIN = 0, RLTV_PH_GRAD = 1, RLTV_PH_UPDATE = 2, RLTV_PH_PSF_GRAD = 3, RLTV_PH_PSF_STEP = 4,
       RLTV_PH_OUTER_END = 5 };
int rltv_enqueue_phase(rltv_ctx* ctx, int32_t phase);
/* device pointers the host all-reduces in place: "step_max" (6 x int32), "gk_sum" (3*MK*MK x double),
 * "stop" (1 x int32) */
void* rltv_device_ptr(rltv_ctx* ctx, const char* name, size_t* nbytes);
/* deterministic host-side stop polling: record after outer iteration `it`, wait for iteration `it` */
int rltv_poll_record(rltv_ctx* ctx, int32_t it);
int rltv_poll_wait(rltv_ctx* ctx, int32_t it, int32_t* stop);

/* the blurry image as it is on the device (the TV-alive mode denoises it in place, pyx:547-549) -> packed HWC (M,N,3) rows */
int rltv_download_image(rltv_ctx* ctx, float* image, size_t image_row_stride_bytes);

/* ---- gathering the result bands on the device ----------------------------------------------------------------
 * The destination rank allocates a full-frame HWC staging buffer (rltv_gather_alloc) and exports it over CUDA IPC;
 * every band converts its OWNED rows planar -> HWC straight into the mapped buffer(s) (peer stores over NVLink,
 * rltv_gather_push), and after a barrier of the caller's choice the destination copies the whole frame to the host in
 * one transfer (rltv_gather_download).  Replaces a host round trip per band (download, re-upload, NCCL send/recv). */
int rltv_gather_alloc(rltv_ctx* ctx, void* handle64_out);
int rltv_gather_attach(rltv_ctx* ctx, int32_t peer_rank, const void* handle64);
int rltv_gather_push(rltv_ctx* ctx, uint32_t dst_rank_mask /* bit r: band r's staging buffer (own rank: the local one) */);
int rltv_gather_download(rltv_ctx* ctx, float* u, size_t u_row_stride_bytes);

/* ---- stage-level entry points (parity tests drive each kernel against the oracle) ------------------ */
/* out[M][N][3] = valid-conv(u, psf) - image   (pyx:477-488) */
int rltv_stage_residual(rltv_ctx* ctx, float* err_out /* packed HWC (M,N,3) */);
/* g = full-conv(err, rot180(psf)) on the u domain (pyx:490-491), using the residual currently on device */
int rltv_stage_adjoint(rltv_ctx* ctx, float* g_out /* packed HWC (M+MK-1, N+MK-1, 3) */);
/* Spectral chain kernel (5 <= MK <= 17, csrc/rltv_chain_fft.cuh; default for MK >= 11, and for MK = 5, 7, 9 on megapixel frames): g = full-conv(valid-conv(u, psf) - image, rot180(psf))
 * (pyx:477-491) in one pass from u, psf and the image on the device; also returns the step statistics it reduces,
 * max(u_c) and max|g_c| per channel (pyx:524 with lambda = 1 and ut = u).  RLTV_ERR_STATE if the context does not use it. */
int rltv_stage_chain(rltv_ctx* ctx, float* g_out /* packed HWC (M+MK-1, N+MK-1, 3) */, float* max6 /* 6 floats or NULL */);
/* debug: skip roles of the chain kernel (bit 0 forward FFT, 1 role E, 2 inverse FFT, 3 TMA loads, 4 role G, 5 epilogue); timing only */
int rltv_debug_chain_roles(int32_t mask);
/* gk = valid-conv(rot180(u), err) (pyx:567-571).  Direct kernels / MK > 17: uses the residual currently on device
 * (call rltv_stage_residual first).  Row-FFT kernels, MK <= 17: the kernel computes the residual of pyx:557-565
 * itself from u, psf and the image and leaves it on the device (read it with rltv_debug_download_err). */
int rltv_stage_gradk(rltv_ctx* ctx, float* gk_out /* packed (MK,MK,3) */);
/* debug: the residual buffer currently on the device, packed HWC (M, N, 3) */
int rltv_debug_download_err(rltv_ctx* ctx, float* err_out);
/* TV(u, out, M, N, epsilon, order, norm, div) (pyx:137-239) of the estimate on the device; order, norm in {1,2};
 * out/div: packed HWC (M+MK-1, N+MK-1, 3), zero on the border ring; *ms = device time of the stencil kernel */
int rltv_stage_tv(rltv_ctx* ctx, int32_t order, int32_t norm, float epsilon, float* out, float* div, float* ms);
/* debug: cycles spent per phase of the row-FFT stencil kernel since the last call (8 counters), then reset;
 * all zero unless the library was built with -DRLTV_PHASE_PROBE */
int rltv_debug_phase_cycles(uint64_t* out8);
/* test entry point of the shared-memory FFT engine used by the row-FFT stencils: nrows x 128 complex values
 * (interleaved re, im), forward (exp(-i..)) or unnormalised inverse */
int rltv_debug_fft128(const float* in, float* out, int32_t nrows, int32_t inverse, int32_t device);
/* whiteness statistic of the residual window (pyx:627-638) */
int rltv_stage_whiteness(rltv_ctx* ctx, int32_t top, int32_t bottom, int32_t left, int32_t right, float* M_r);

#ifdef __cplusplus
}
#endif
#endif /* RLTV_B200_H */
