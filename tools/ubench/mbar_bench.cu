// Micro-benchmark: latency of the synchronisation hops the conv kernels are built from.
//   0: mbarrier ping-pong between two warps (lane 0 arrives, whole warp waits): 2 hops per iteration
//   1: same, all 32 lanes arrive (count 32)
//   2: tcgen05.commit -> mbarrier -> wait by the same warp (no MMA outstanding)
//   3: tcgen05.commit (one MMA N=16 outstanding) -> wait
//   4: ping-pong where the waiter spins on ld.shared of a flag instead of mbarrier.try_wait
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../dffinthewild_b200/csrc -o mbar_bench mbar_bench.cu
#include <cstdio>
#include "tc_common.cuh"
namespace dff { void set_error(const std::string&) {} int fail(int c, const std::string&) { return c; } int check_cuda(cudaError_t e, const char*) { return e != cudaSuccess; } }
using namespace dff;
using namespace dff::tc;

__global__ void __launch_bounds__(128, 1) k(int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int flag[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), mode == 1 ? 32 : 1);
    mbar_init(smem_u32(&bar[1]), mode == 1 ? 32 : 1);
    flag[0] = flag[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t b0 = smem_u32(&bar[0]), b1 = smem_u32(&bar[1]);
  const long long t0 = clock64();
  if (mode <= 1) {
    if (warp == 0) {
      for (int i = 0; i < iters; ++i) {
        if (mode == 1 || lane == 0) mbar_arrive(b0);
        mbar_wait(b1, i & 1);
      }
    } else if (warp == 1) {
      for (int i = 0; i < iters; ++i) {
        mbar_wait(b0, i & 1);
        if (mode == 1 || lane == 0) mbar_arrive(b1);
      }
    }
  } else if (mode == 2 || mode == 3) {
    if (warp == 0) {
      const bool leader = elect_one();
      const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t ad = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((s0 >> 4) | ((2048u >> 4) << 16));
      const uint64_t bd = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | (((s0 + 8192) >> 4) | ((256u >> 4) << 16));
      for (int i = 0; i < iters; ++i) {
        if (leader) {
          if (mode == 3) umma(tb, ad, bd, idesc, 0);
          umma_commit(b0);
        }
        __syncwarp();
        mbar_wait(b0, i & 1);
      }
    }
  } else {
    if (warp == 0) {
      for (int i = 1; i <= iters; ++i) {
        if (lane == 0) flag[0] = i;
        while (flag[1] != i) {}
      }
    } else if (warp == 1) {
      for (int i = 1; i <= iters; ++i) {
        while (flag[0] != i) {}
        if (lane == 0) flag[1] = i;
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  fence_before(); __syncthreads();
  if (warp == 0) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32) : "memory"); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  const char* names[5] = {"mbarrier ping-pong (1 arrive)", "mbarrier ping-pong (32 arrive)", "tcgen05.commit -> wait (idle pipe)", "mma + commit -> wait", "smem flag ping-pong"};
  for (int grid : {1, 148})
    for (int mode = 0; mode < 5; ++mode) {
      const int iters = 2000;
      k<<<grid, 128, 32 * 1024>>>(mode, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("grid %3d %-36s %.1f clk/iteration %s\n", grid, names[mode], h / (double)iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
