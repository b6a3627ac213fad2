// Micro-benchmark: TMA tiled-load throughput as a function of the innermost box row (16-byte rows of an 8-channel chunk vs merged
// rows), boxes in flight per CTA and CTAs per SM.  The source tensor (2 MB) stays in L2, so this is the rate of the TMA unit itself.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../dffinthewild_b200/csrc -o tma_row_bench tma_row_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "tc_common.cuh"
namespace dff { void set_error(const std::string&) {} int fail(int c, const std::string&) { return c; } int check_cuda(cudaError_t e, const char*) { return e != cudaSuccess; } }
using namespace dff;
using namespace dff::tc;

// each CTA: `iters` rounds of `nbox` boxes (all on one mbarrier), one issuing thread
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tm, int nbox, int box_bytes, int iters, int rank4, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_expect_tx(smem_u32(&bar), (uint32_t)(nbox * box_bytes));
      for (int b = 0; b < nbox; ++b) {
        const int y0 = ((blockIdx.x * 7 + i * 3 + b * 11) % 40) * 4;
        tma_load_4d(s0 + (uint32_t)(b * 3072), &tm, smem_u32(&bar), 0, rank4 ? (b % 4) : 0, y0, 0);
      }
      mbar_wait(smem_u32(&bar), ph);
      ph ^= 1;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)f;
  const int W = 256, H = 512;   // (8 ch, W, H) bf16 = 2 MB
  void* d; cudaMalloc(&d, (size_t)8 * W * H * 2); cudaMemset(d, 0, (size_t)8 * W * H * 2);
  long long* o; cudaMalloc(&o, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  for (int merged = 0; merged < 2; ++merged) {
    const int RX = 10, RY = 18;
    CUtensorMap tm;
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (!merged) {   // rows of 16 bytes: (8, W, H, 1), box (8, RX, RY, 1)
      cuuint64_t dims[4] = {8, (cuuint64_t)W, (cuuint64_t)H, 1}, st[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)W * H * 16};
      cuuint32_t box[4] = {8, RX, RY, 1};
      enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {         // merged rows: (8*W, 1, H, 1), box (8*RX, 1, RY, 1)
      cuuint64_t dims[4] = {8ull * W, 1, (cuuint64_t)H, 1}, st[3] = {(cuuint64_t)W * 16, (cuuint64_t)W * 16, (cuuint64_t)W * H * 16};
      cuuint32_t box[4] = {8 * RX, 1, RY, 1};
      enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    const int box_bytes = RX * RY * 16;
    for (int cps : {1, 2, 4})
      for (int nbox : {1, 4, 12}) {
        const int iters = 400;
        k<<<148 * cps, 128, 48 * 1024>>>(tm, nbox, box_bytes, iters, 0, o);
        cudaError_t e = cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, o, 8, cudaMemcpyDeviceToHost);
        const double clk_per_round = (double)h / iters;
        const double rows = (merged ? RY : RX * RY) * (double)nbox * cps;   // innermost rows per round and SM
        printf("%s rows, %d CTA/SM, %d boxes in flight per CTA: %.0f clk per round; per SM %.2f rows/clk, %.1f B/clk %s\n", merged ? "merged (160 B)" : "16-byte", cps, nbox,
               clk_per_round, rows / clk_per_round, box_bytes * (double)nbox * cps / clk_per_round, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  }
  return 0;
}
