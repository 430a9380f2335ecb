// Stand-alone probe of the TMA box-load semantics the stencil kernels rely on (negative start coordinates,
// boxes larger than the tensor, zero fill, full-box transaction count).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../image_cases_studies_b200/csrc/rltv_tma.cuh"
using namespace rltv;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_probe(const __grid_constant__ CUtensorMap tm, int x, int y, int c, int bw, int bh, float* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
  float* tile = reinterpret_cast<float*>(sm);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + ((bw * bh * 4 + 127) & ~127));
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive_expect_tx(bar, bw * bh * 4); tma_load_3d(tile, &tm, x, y, c, bar); }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1; int idx = -1;
  const int W = 58, H = 50, pitch = 60;
  size_t plane = size_t(H) * pitch;
  std::vector<float> h(3 * plane);
  for (int c = 0; c < 3; ++c) for (int y = 0; y < H; ++y) for (int x = 0; x < pitch; ++x) h[c * plane + y * pitch + x] = x < W ? c * 10000 + y * 100 + x : -1.f;
  float *d, *o; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&o, 256 * 256 * 4));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN enc = (PFN)p;
  struct Case { int bw, bh, x, y, c; } cases[] = {{16, 8, 4, 4, 1}, {16, 8, 3, 2, 0}, {16, 8, 5, 0, 0}, {16, 8, 0, -2, 0}, {16, 8, -4, 0, 0}, {16, 8, -4, -2, 0}, {132, 36, -4, -1, 2}, {132, 36, 0, 40, 1}, {16, 8, 50, 45, 2}, {16, 8, -3, -2, 0}};
  for (auto cs : cases) {
    ++idx; if (only >= 0 && idx != only) continue;
    CUtensorMap tm;
    cuuint64_t dims[3] = {W, H, 3}; cuuint64_t str[2] = {pitch * 4ull, plane * 4ull};
    cuuint32_t box[3] = {(cuuint32_t)cs.bw, (cuuint32_t)cs.bh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("case box %dx%d at (%d,%d,%d): encode=%d ", cs.bw, cs.bh, cs.x, cs.y, cs.c, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); continue; }
    int smem = ((cs.bw * cs.bh * 4 + 127) & ~127) + 64 + 128;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_probe<<<1, 128, smem>>>(tm, cs.x, cs.y, cs.c, cs.bw, cs.bh, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> res(cs.bw * cs.bh);
    CK(cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r2 = 0; r2 < cs.bh; ++r2) for (int c2 = 0; c2 < cs.bw; ++c2) {
      int yy = cs.y + r2, xx = cs.x + c2;
      float exp = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? cs.c * 10000 + yy * 100 + xx : 0.f;
      if (res[r2 * cs.bw + c2] != exp) ++bad;
    }
    printf("mismatches=%d\n", bad);
  }
  return 0;
}
