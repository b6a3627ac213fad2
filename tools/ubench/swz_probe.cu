// Probe: SS-form tcgen05.mma with a SWIZZLED K-major A operand written by TMA (pixel-major rows of 32 / 64 / 128 bytes = 16 / 32 / 64
// bf16 channels) and read through a descriptor whose start address is an ARBITRARY row of the tile (a convolution tap) and an arbitrary
// 32-byte K step inside the row.  Which base-offset rule makes the hardware read the rows TMA wrote?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../dffinthewild_b200/csrc -o swz_probe swz_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include "tc_common.cuh"
namespace dff { void set_error(const std::string&) {} int fail(int c, const std::string&) { return c; } int check_cuda(cudaError_t e, const char*) { return e != cudaSuccess; } }
using namespace dff;
using namespace dff::tc;

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, int C, int P, int r0, int kstep, int bo_mode, int pad, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, tbar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t s0 = ((smem_u32(smem_raw) + 1023u) & ~1023u) + (uint32_t)pad;   // pad: tile base deliberately off the 1024-byte grid (multiples of 128)
  uint8_t* g0 = smem_raw + (s0 - smem_u32(smem_raw));
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(g0 + 48 * 1024);
  const int N = 16;
  for (int i = threadIdx.x; i < 2 * N * 8; i += 128) {
    const int j = i & 7, n = (i >> 3) % N, half = i / (8 * N);
    B[i] = __float2bfloat16((half * 8 + j) == n ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&tbar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tb = tmem_base_s;
  const int rowbytes = 2 * C;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&tbar), (uint32_t)(P * rowbytes));
    tma_load_2d(s0, &tm, smem_u32(&tbar), 0, 0);
    mbar_wait(smem_u32(&tbar), 0);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t layout = C == 16 ? 6u : (C == 32 ? 4u : 2u);
    const uint32_t sa = s0 + (uint32_t)(r0 * rowbytes + kstep * 32), sb = s0 + 48 * 1024;
    uint32_t bo = 0;
    if (bo_mode == 1) bo = (sa >> 7) & 7u;
    if (bo_mode == 2) bo = ((s0 + (uint32_t)(r0 * rowbytes)) >> 7) & 7u;
    uint64_t ad = make_desc(sa, 16, 8 * rowbytes, layout) | ((uint64_t)bo << 49);
    const uint64_t bd = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | ((sb >> 4) | (((uint32_t)N * 16 >> 4) << 16));
    umma(tb, ad, bd, idesc, 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = warp * 32 + lane;
  uint32_t v[16];
  tmem_ld16(tb + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < 16; ++j) out[row * 16 + j] = __uint_as_float(v[j]);
  fence_before(); __syncthreads();
  if (threadIdx.x < 32) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32) : "memory"); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)f;
  const int P = 160;
  float* dout; cudaMalloc(&dout, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int C : {16, 32, 64}) {
    std::vector<__nv_bfloat16> hx((size_t)P * C);
    for (int p = 0; p < P; ++p) for (int c = 0; c < C; ++c) hx[(size_t)p * C + c] = __float2bfloat16((float)(((p * C + c) * 7) % 251 - 125));
    void* dx; cudaMalloc(&dx, hx.size() * 2); cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)P}, st[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)P}, es[2] = {1, 1};
    const CUtensorMapSwizzle sw = C == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int pad : {0, 128, 384})
    for (int bo_mode = 0; bo_mode < 3; ++bo_mode) {
      printf("C=%2d tile base +%3d, base-offset rule %d:", C, pad, bo_mode);
      for (int r0 : {0, 1, 2, 3, 5, 8, 13, 21}) {
        int bad = 0;
        for (int kstep = 0; kstep < C / 16; ++kstep) {
          probe<<<1, 128, 64 * 1024>>>(tm, C, P, r0, kstep, bo_mode, pad, dout);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf(" CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          std::vector<float> h(128 * 16);
          cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
          for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n)
            if (h[m * 16 + n] != __bfloat162float(hx[(size_t)(r0 + m) * C + kstep * 16 + n])) ++bad;
        }
        printf(" r0=%d:%s", r0, bad ? "BAD" : "ok");
      }
      printf("\n");
    }
    cudaFree(dx);
  }
  return 0;
}
