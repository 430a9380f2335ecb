// Micro-benchmarks that pin the roofline denominators bench.py cannot take from MEASURED_PEAKS.json:
// the non-tensor FP32 FMA peak (scalar FFMA and packed fma.rn.f32x2), shared-memory LDS.128 bandwidth,
// and a float4 streaming copy.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
// Output: one JSON object on stdout.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b, int iters) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b, int iters) {
  unsigned long long acc[ILP];
  unsigned long long av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    float x = threadIdx.x * 1e-3f + i;
    asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x));
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  if (s == 123.456f) out[0] = s;
}

// stencil-like mix: per 16 FFMA one LDS.128 (broadcast) -- what the conv kernels issue
__global__ void __launch_bounds__(256) k_ffma_lds(float* out, int iters) {
  __shared__ float4 w[64];
  if (threadIdx.x < 64) w[threadIdx.x] = make_float4(1.0001f, 0.9999f, 1.0002f, 0.9998f);
  __syncthreads();
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float4 c = w[(it + r) & 63];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[4 * i + 0] = fmaf(acc[4 * i + 0], c.x, c.y);
        acc[4 * i + 1] = fmaf(acc[4 * i + 1], c.y, c.z);
        acc[4 * i + 2] = fmaf(acc[4 * i + 2], c.z, c.w);
        acc[4 * i + 3] = fmaf(acc[4 * i + 3], c.w, c.x);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_lds128(float* out, int iters) {
  extern __shared__ float4 sm4[];
  for (int i = threadIdx.x; i < 2048; i += 256) sm4[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float4 a = make_float4(0, 0, 0, 0);
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float4 v = sm4[(idx + r * 256) & 2047];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    idx = (idx + 32) & 2047;
  }
  if (a.x + a.y + a.z + a.w == 123.456f) out[0] = a.x;
}

__global__ void __launch_bounds__(256) k_copy4(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) out[i] = in[i];
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  float* d; CK(cudaMalloc(&d, 1024));
  const int iters = 4096, blocks = sms * 8;
  // FLOP = blocks*256 threads * iters*8*ILP fma * 2
  auto tf = [&](double ilp, float ms, double per) { return blocks * 256.0 * iters * 8.0 * ilp * per / (ms * 1e-3) / 1e12; };
  float t8 = time_ms([&] { k_ffma<8><<<blocks, 256>>>(d, 1.0001f, 1e-6f, iters); });
  float t16 = time_ms([&] { k_ffma<16><<<blocks, 256>>>(d, 1.0001f, 1e-6f, iters); });
  float t2_8 = time_ms([&] { k_ffma2<8><<<blocks, 256>>>(d, 1.0001f, 1e-6f, iters); });
  float t2_16 = time_ms([&] { k_ffma2<16><<<blocks, 256>>>(d, 1.0001f, 1e-6f, iters); });
  float tl = time_ms([&] { k_ffma_lds<<<blocks, 256>>>(d, iters); });
  CK(cudaFuncSetAttribute(k_lds128, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  float ts = time_ms([&] { k_lds128<<<sms * 4, 256, 32768>>>(d, iters); });
  const double lds_bytes = double(sms) * 4 * 256 * iters * 8.0 * 16.0;
  const size_t n4 = size_t(1) << 28;  // 4 GiB each way
  float4 *a, *b; CK(cudaMalloc(&a, n4 * 16)); CK(cudaMalloc(&b, n4 * 16));
  CK(cudaMemset(a, 1, n4 * 16));
  float tc = time_ms([&] { k_copy4<<<sms * 16, 256>>>(a, b, n4); }, 5);
  float tm = time_ms([&] { cudaMemcpyAsync(b, a, n4 * 16, cudaMemcpyDeviceToDevice); }, 5);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"clock_khz_attr\": %d, "
         "\"ffma_tflops_ilp8\": %.2f, \"ffma_tflops_ilp16\": %.2f, \"ffma2_tflops_ilp8\": %.2f, \"ffma2_tflops_ilp16\": %.2f, "
         "\"ffma_with_lds128_tflops\": %.2f, \"lds128_tb_s\": %.2f, \"lds128_bytes_per_clk_per_sm_at_attr_clock\": %.1f, "
         "\"copy_kernel_gb_s\": %.1f, \"memcpy_d2d_gb_s\": %.1f}\n",
         p.name, sms, p.l2CacheSize, clk, tf(8, t8, 2), tf(16, t16, 2), tf(8, t2_8, 4), tf(16, t2_16, 4),
         blocks * 256.0 * iters * 8.0 * 16 * 2 / (tl * 1e-3) / 1e12, lds_bytes / (ts * 1e-3) / 1e12,
         lds_bytes / (ts * 1e-3) / (double(clk) * 1e3) / sms, 2.0 * n4 * 16 / (tc * 1e-3) / 1e9, 2.0 * n4 * 16 / (tm * 1e-3) / 1e9);
  return 0;
}
