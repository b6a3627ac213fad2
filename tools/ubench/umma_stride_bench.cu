// Micro-benchmark: SS-form tcgen05.mma (M=128, K=16, N=32, no swizzle) as a function of the A operand's row-group stride (SBO) and
// K-half distance (LBO): the row-folded first layer multiplies planes whose 8-row groups are 8 plane rows = 3072 bytes apart.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../dffinthewild_b200/csrc -o umma_stride_bench umma_stride_bench.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
namespace dff { void set_error(const std::string&) {} int fail(int c, const std::string&) { return c; } int check_cuda(cudaError_t e, const char*) { return e != cudaSuccess; } }
using namespace dff;
using namespace dff::tc;

__global__ void __launch_bounds__(128, 1) k(int N, int sbo, int lbo, int step16, int iters, long long* out, int nacc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 150 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x < 32) {   // whole warp runs the loop on uniform values, one elected lane issues (the kernels' issue form: 39-40 clk per MMA)
    uint32_t elected;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(elected));
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sa = s0, sb = s0 + 128 * 1024;
    const uint32_t a_lo = (sa >> 4) | (((uint32_t)lbo >> 4) << 16), a_hi = ((uint32_t)sbo >> 4) | (1u << 14);
    const uint32_t b_lo = (sb >> 4) | (((uint32_t)N * 16 >> 4) << 16), b_hi = (128u >> 4) | (1u << 14);
    const long long t0 = clock64();
    const uint32_t d0 = tb, d1 = nacc > 1 ? tb + N : tb;   // nacc == 1: every MMA accumulates into the SAME columns (dependent chain)
    const uint64_t ad = ((uint64_t)a_hi << 32) | a_lo, bd = ((uint64_t)b_hi << 32) | b_lo;
    const uint64_t st = (uint64_t)step16;
#pragma unroll 1
    for (int i = 0; i < iters; i += 8) {
      if (elected) {
        umma(d0, ad, bd, idesc, 1); umma(d1, ad + st, bd, idesc, 1); umma(d0, ad + 2 * st, bd, idesc, 1); umma(d1, ad + 3 * st, bd, idesc, 1);
        umma(d0, ad + 4 * st, bd, idesc, 1); umma(d1, ad + 5 * st, bd, idesc, 1); umma(d0, ad + 6 * st, bd, idesc, 1); umma(d1, ad + 7 * st, bd, idesc, 1);
      }
      __syncwarp();
    }
    if (elected) umma_commit(smem_u32(&bar));
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  fence_before(); __syncthreads();
  if (threadIdx.x < 32) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory"); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int sbos[] = {128, 3072};
  for (int nacc : {1, 2})
  for (int N : {16, 32, 48, 64})
    for (int lbo : {64})
      for (int sbo : sbos)
        for (int step16 : {4}) {
          const int iters = 2000;
          k<<<148, 128, 200 * 1024>>>(N, sbo, lbo, step16, iters, d, nacc);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("nacc=%d N=%2d LBO=%4d SBO=%4d tap step %3d B: issue %.1f clk/MMA, complete %.1f clk/MMA %s\n", nacc, N, lbo, sbo, 16 * step16, h[0] / (double)iters,
                 h[1] / (double)iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
