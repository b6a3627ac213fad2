// Micro-benchmark + correctness probe: tcgen05.mma with the A operand in TMEM (staged by tcgen05.cp from shared memory)
// versus the SS form (A from a shared-memory descriptor).  Answers two questions for the convolution kernels:
//   1. does tcgen05.cp.128x256b of a no-swizzle K-major [128 x 16] bf16 tile produce the A layout the TS-form MMA expects?
//   2. what are the issue/complete rates of cp and of TS-form MMAs for small N?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../dffinthewild_b200/csrc -o tmem_a_bench tmem_a_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
namespace dff { void set_error(const std::string&) {} int fail(int c, const std::string&) { return c; } int check_cuda(cudaError_t e, const char*) { return e != cudaSuccess; } }
using namespace dff;
using namespace dff::tc;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

// A: 128 rows x 16 bf16 as two 8-channel planes [row] x 16 B (plane stride 2048 B + `shift` start offset in 16 B units)
// B: N rows x 16 bf16 likewise (plane stride N*16)
// out[0..128*N): SS result, out[128*N .. 2*128*N): TS result
__global__ void __launch_bounds__(128, 1) probe(int N, int shift, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g0 = smem_raw + (s0 - smem_u32(smem_raw));
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(g0);            // 2 planes x (128+16) rows x 8
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(g0 + 8192);     // 2 planes x N rows x 8
  const int arows = 128 + 16;
  for (int i = threadIdx.x; i < 2 * arows * 8; i += 128) {
    const int j = i & 7, r = (i >> 3) % arows, pl = i / (8 * arows);
    A[i] = __float2bfloat16((float)(((r * 7 + j * 3 + pl * 5) % 13) - 6));
  }
  for (int i = threadIdx.x; i < 2 * N * 8; i += 128) {
    const int j = i & 7, r = (i >> 3) % N, pl = i / (8 * N);
    B[i] = __float2bfloat16((float)(((r * 5 + j * 11 + pl * 3) % 7) - 3));
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint32_t sa = s0 + shift * 16, sb = s0 + 8192;
    const uint64_t ad = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((sa >> 4) | (((uint32_t)arows * 16 >> 4) << 16));
    const uint64_t bd = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((sb >> 4) | (((uint32_t)N * 16 >> 4) << 16));
    umma(tb, ad, bd, idesc, 0);                 // SS  -> columns [0, N)
    tmem_cp_128x256b(tb + 256, ad);             // A   -> columns [256, 264)
    umma_ts(tb + 128, tb + 256, bd, idesc, 0);  // TS  -> columns [128, 128+N)
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = warp * 32 + lane;
  for (int half = 0; half < 2; ++half)
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + half * 128 + c0, v);
      for (int j = 0; j < 16; ++j) out[(size_t)half * 128 * N + row * N + c0 + j] = __uint_as_float(v[j]);
    }
  fence_before(); __syncthreads();
  if (threadIdx.x < 32) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory"); }
}

// MODE 0: TS MMAs only (A resident in TMEM); 1: cp only; 2: 1 cp + PER TS MMAs into PER accumulators; 3: SS MMAs
template <int MODE, int PER>
__global__ void __launch_bounds__(128, 4) rate(int N, int ncols, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 30 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sa = s0 + 16, sb = s0 + 24 * 1024;
    const uint64_t ad = ((uint64_t)((160u >> 4) | (1u << 14)) << 32) | ((sa >> 4) | ((2880u >> 4) << 16));
    const uint64_t bd = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((sb >> 4) | (((uint32_t)N * 16 >> 4) << 16));
    const uint32_t abase = tb + ncols - 32;  // 4 rotating A slots of 8 columns
    uint32_t dacc[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) dacc[k] = tb + k * N;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t aslot = abase + u * 8;
        if (MODE == 1 || MODE == 2) tmem_cp_128x256b(aslot, ad + u);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          if (MODE == 0 || MODE == 2) umma_ts(dacc[k], aslot, bd, idesc, 1);
          if (MODE == 3) umma(dacc[k], ad + ((u + k) & 7), bd, idesc, 1);
        }
      }
    }
    umma_commit(smem_u32(&bar));
    const long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  fence_before(); __syncthreads();
  if (threadIdx.x < 32) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(ncols) : "memory"); }
}

// Multi-warp issue: `nw` warps of ONE CTA each issue PER-accumulator MMA streams (MODE 0 TS / 3 SS / 2 cp+TS) concurrently.
// Whole warp runs the loop (uniform values), the MMA is issued by one elected lane.
template <int MODE, int PER>
__global__ void __launch_bounds__(128, 1) rate_mw(int N, int nw, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ long long tmax[4];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 30 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tmem_base_s;
  const int warp = threadIdx.x >> 5;
  tmax[warp] = 0;
  if (warp < nw) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sa = s0 + 16, sb = s0 + 24 * 1024;
    const uint64_t ad = ((uint64_t)((160u >> 4) | (1u << 14)) << 32) | ((sa >> 4) | ((2880u >> 4) << 16));
    const uint64_t bd = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((sb >> 4) | (((uint32_t)N * 16 >> 4) << 16));
    const uint32_t abase = tb + 384 + warp * 32;  // 4 rotating A slots of 8 columns per warp
    uint32_t dacc[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) dacc[k] = tb + (warp * PER + k) * N;
    uint32_t elected;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(elected));
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 4) {
      if (elected) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t aslot = abase + u * 8;
          if (MODE == 2) tmem_cp_128x256b(aslot, ad + u);
#pragma unroll
          for (int k = 0; k < PER; ++k) {
            if (MODE == 0 || MODE == 2) umma_ts(dacc[k], aslot, bd, idesc, 1);
            if (MODE == 3) umma(dacc[k], ad + ((u + k) & 7), bd, idesc, 1);
          }
        }
      }
      __syncwarp();
    }
    if (elected) umma_commit(smem_u32(&bar[warp]));
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0);
    tmax[warp] = clock64() - t0;
  }
  fence_before(); __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) { long long m = 0; for (int i = 0; i < 4; ++i) m = tmax[i] > m ? tmax[i] : m; out[0] = m; }
  if (threadIdx.x < 32) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory"); }
}
template <int MODE, int PER>
static void run_mw(int N, int nw, long long* d, const char* name) {
  if (4 * PER * N > 384) return;
  const int iters = 1024;
  cudaFuncSetAttribute(rate_mw<MODE, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  rate_mw<MODE, PER><<<148, 128, 32 * 1024>>>(N, nw, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("1 CTA/SM, %d issuing warps, N=%2d %-18s per=%d: %.1f clk/iter/warp -> %.1f clk/MMA/SM %s\n", nw, N, name, PER,
         h[0] / (double)iters, h[0] / (double)iters / PER / nw, e == cudaSuccess ? "" : cudaGetErrorString(e));
  if (e != cudaSuccess) exit(1);
}

template <int MODE, int PER>
static void run_rate(int grid, int N, int ncols, long long* d, const char* name) {
  if (PER * N > ncols - 32) return;
  const int iters = 1024;
  cudaFuncSetAttribute(rate<MODE, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  rate<MODE, PER><<<grid, 128, 32 * 1024>>>(N, ncols, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("grid %3d cols %3d N=%2d %-18s per=%d: issue %.1f clk/iter, complete %.1f clk/iter (%.1f clk/MMA) %s\n", grid, ncols, N, name, PER,
         h[0] / (double)iters, h[1] / (double)iters, MODE == 1 ? 0.0 : h[1] / (double)iters / PER, e == cudaSuccess ? "" : cudaGetErrorString(e));
  if (e != cudaSuccess) exit(1);
}

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {16, 32, 64})
    for (int shift : {0, 3}) {
      float* d; cudaMalloc(&d, 2 * 128 * N * 4);
      cudaMemset(d, 0xff, 2 * 128 * N * 4);
      probe<<<1, 128, 64 * 1024>>>(N, shift, d);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> h(2 * 128 * N);
      cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
      // host reference
      int bad_ss = 0, bad_ts = 0;
      const int arows = 128 + 16;
      for (int r = 0; r < 128; ++r)
        for (int c = 0; c < N; ++c) {
          float acc = 0;
          for (int k = 0; k < 16; ++k) {
            const int pl = k >> 3, j = k & 7, ar = r + shift;
            const float a = (float)((((ar % arows) * 7 + j * 3 + pl * 5) % 13) - 6);
            const float b = (float)(((c * 5 + j * 11 + pl * 3) % 7) - 3);
            acc += a * b;
          }
          if (h[r * N + c] != acc) ++bad_ss;
          if (h[128 * N + r * N + c] != acc) ++bad_ts;
        }
      printf("probe N=%2d shift=%d: %s  SS mismatches %d, TS(A via tcgen05.cp) mismatches %d of %d\n", N, shift,
             e == cudaSuccess ? "ok" : cudaGetErrorString(e), bad_ss, bad_ts, 128 * N);
      cudaFree(d);
      if (e != cudaSuccess) return 1;
    }
  long long* d; cudaMalloc(&d, 16);
  for (int nw : {1, 2, 4})
    for (int N : {16, 32, 64}) {
      run_mw<0, 1>(N, nw, d, "TS mma");
      run_mw<0, 3>(N, nw, d, "TS mma");
      run_mw<3, 1>(N, nw, d, "SS mma");
      run_mw<3, 3>(N, nw, d, "SS mma");
      run_mw<2, 3>(N, nw, d, "cp + 3 TS mma");
    }
  for (int cfg = 0; cfg < 0; ++cfg) {
    const int grid = cfg == 0 ? 1 : cfg == 1 ? 148 : 592, ncols = cfg == 2 ? 128 : 512;
    for (int N : {16, 32, 48, 64, 128}) {
      run_rate<0, 1>(grid, N, ncols, d, "TS mma only");
      run_rate<0, 3>(grid, N, ncols, d, "TS mma only");
      run_rate<0, 9>(grid, N, ncols, d, "TS mma only");
      run_rate<1, 1>(grid, N, ncols, d, "cp only");
      run_rate<2, 1>(grid, N, ncols, d, "cp + per x TS mma");
      run_rate<2, 3>(grid, N, ncols, d, "cp + per x TS mma");
      run_rate<2, 9>(grid, N, ncols, d, "cp + per x TS mma");
      run_rate<3, 1>(grid, N, ncols, d, "SS mma");
      run_rate<3, 3>(grid, N, ncols, d, "SS mma");
      run_rate<3, 9>(grid, N, ncols, d, "SS mma");
    }
  }
  return 0;
}
