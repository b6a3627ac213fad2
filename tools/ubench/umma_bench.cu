// Micro-benchmark: tcgen05.mma throughput (M=128, K=16, kind::f16) vs N, smem layout and operand alignment.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../dffinthewild_b200/csrc -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
namespace dff { void set_error(const std::string&) {} int fail(int c, const std::string&) { return c; } int check_cuda(cudaError_t e, const char*) { return e != cudaSuccess; } }
using namespace dff;
using namespace dff::tc;

// mode 0: no-swizzle, A SBO=128 (aligned, contiguous core matrices); 1: no-swizzle SBO=160, start +16 (unaligned)
// mode 2: SW128 K-major (SBO 1024); 3: SW32 (SBO 256)
__global__ void __launch_bounds__(128, 1) k(int N, int mode, int iters, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t a_lo, a_hi, b_lo, b_hi;
    const uint32_t sa = s0, sb = s0 + 24 * 1024;
    if (mode == 0)      { a_lo = (sa >> 4) | ((2048u >> 4) << 16); a_hi = (128u >> 4) | (1u << 14); b_lo = (sb >> 4) | (((uint32_t)N * 16 >> 4) << 16); b_hi = (128u >> 4) | (1u << 14); }
    else if (mode == 1) { a_lo = ((sa + 16) >> 4) | ((2880u >> 4) << 16); a_hi = (160u >> 4) | (1u << 14); b_lo = (sb >> 4) | (((uint32_t)N * 16 >> 4) << 16); b_hi = (128u >> 4) | (1u << 14); }
    else if (mode == 2) { a_lo = (sa >> 4) | (1u << 16); a_hi = (1024u >> 4) | (1u << 14) | (2u << 29); b_lo = (sb >> 4) | (1u << 16); b_hi = (1024u >> 4) | (1u << 14) | (2u << 29); }
    else                { a_lo = (sa >> 4) | (1u << 16); a_hi = (256u >> 4) | (1u << 14) | (6u << 29); b_lo = (sb >> 4) | (1u << 16); b_hi = (256u >> 4) | (1u << 14) | (6u << 29); }
    const long long t0 = clock64();
    const uint32_t d0 = tb, d1 = tb + (nacc - 1) * N;
    const uint64_t ad = ((uint64_t)a_hi << 32) | a_lo, bd = ((uint64_t)b_hi << 32) | b_lo;
#pragma unroll 1
    for (int i = 0; i < iters; i += 8) {
      umma(d0, ad, bd, idesc, 1); umma(d1, ad + 2, bd, idesc, 1); umma(d0, ad + 4, bd, idesc, 1); umma(d1, ad + 6, bd, idesc, 1);
      umma(d0, ad + 8, bd, idesc, 1); umma(d1, ad + 10, bd, idesc, 1); umma(d0, ad + 12, bd, idesc, 1); umma(d1, ad + 14, bd, idesc, 1);
    }
    umma_commit(smem_u32(&bar));
    const long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  fence_before(); __syncthreads();
  if (threadIdx.x < 32) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory"); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[4] = {"noswz aligned", "noswz SBO160+16", "SW128", "SW32"};
  for (int grid : {1, 148})
    for (int mode = 0; mode < 4; ++mode)
      for (int N : {16, 32, 64, 128})
        for (int nacc : {1, 2}) {
          const int iters = 2000;
          k<<<grid, 128, 64 * 1024>>>(N, mode, iters, nacc, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("grid %3d %-16s N=%3d nacc=%d: issue %.1f clk/MMA, complete %.1f clk/MMA %s\n", grid, names[mode], N, nacc,
                 h[0] / (double)iters, h[1] / (double)iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
