// SRD channel-attention branch as ONE bandwidth pass on the warp-level tensor-core path (mma.sync, bf16 in / fp32 accumulate).
//
//   out = F + relu( W1 . relu( conv3x1x1(F; W0) ) )      (reference train_codes/Depth_Estimation_Network.py:399-407; no BN, no bias)
//
// The two convolutions have 4-12 FLOP per byte: as two implicit-GEMM launches they cost two read+write passes over a full-,
// half- or quarter-resolution tensor, and as scalar FMAs (srd_attention_kernel, elementwise.cu) the C = 8 case is bound by
// the FP32 pipe (256 FMA per voxel) at 3.5x the HBM time.  Here a warp owns 16 consecutive pixels and walks their S slices:
//   * the A fragment of m16n8k16 wants, per thread, two adjacent K values of rows g and g+8 — with channels-last storage that
//     is a plain 4-byte (C = 8) or 8-byte (C >= 16, K permuted so a thread's four channels are contiguous) load of pixel g /
//     g+8; a quad reads one pixel's 16/32 contiguous bytes, a warp-wide load covers 128/256 contiguous bytes;
//   * the 3 focal taps are 3 K blocks fed from a rotating register window (slices z-1, z, z+1; two more in flight), so F is
//     read from HBM exactly once;
//   * the accumulator fragment of the first product IS the A fragment of the second (rows g/g+8, columns 2t/2t+1), so the
//     intermediate never leaves registers (rounded to bf16 like every activation of the bf16 mode);
//   * the output columns of the second product are permuted (weights packed accordingly) so a thread ends up with exactly the
//     channels it loaded: the residual F comes from the register window and the store is the mirror image of the load.
// HBM traffic: C*2 bytes in + C*2 bytes out per voxel; 3 (C = 8), 8 (C = 16) or 32 (C = 32) MMAs per 16 voxels.
#include <cstdlib>

#include "common.cuh"

namespace dff {

__device__ __forceinline__ void mma_16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_1688(float* d, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t relu_pack(float lo, float hi) { return pack_bf2(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }
// relu(a) + F for two adjacent channels; f holds F as bf16x2
__device__ __forceinline__ uint32_t relu_add_pack(float a_lo, float a_hi, uint32_t f) {
  return pack_bf2(fmaxf(a_lo, 0.f) + __uint_as_float(f << 16), fmaxf(a_hi, 0.f) + __uint_as_float(f & 0xffff0000u));
}
__device__ __forceinline__ uint32_t ldg_nc_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ldg_nc_u64(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

constexpr int kAttnThreads = 256;

// ---- C = 8: K block 1 = [slice z-1 | slice z] (k16), K block 2 = slice z+1 (k8); second product k8 --------------------------
__global__ void __launch_bounds__(kAttnThreads) srd_attention_mma8_kernel(const __nv_bfloat16* __restrict__ F, const float* __restrict__ w0,
                                                                          const float* __restrict__ w1, __nv_bfloat16* __restrict__ out,
                                                                          int S, size_t plane, size_t ntiles) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  pdl_trigger();
  // B fragments (k = 2t, 2t+1 ; n = g) straight from the fp32 packs w0 [dz][ci][co], w1 [c][co]
  const uint32_t bm = pack_bf2(__ldg(w0 + (0 * 8 + 2 * t) * 8 + g), __ldg(w0 + (0 * 8 + 2 * t + 1) * 8 + g));
  const uint32_t bc = pack_bf2(__ldg(w0 + (1 * 8 + 2 * t) * 8 + g), __ldg(w0 + (1 * 8 + 2 * t + 1) * 8 + g));
  const uint32_t bp = pack_bf2(__ldg(w0 + (2 * 8 + 2 * t) * 8 + g), __ldg(w0 + (2 * 8 + 2 * t + 1) * 8 + g));
  const uint32_t b2 = pack_bf2(__ldg(w1 + (2 * t) * 8 + g), __ldg(w1 + (2 * t + 1) * 8 + g));
  pdl_wait();   // F is the previous layer's output
  const size_t zs = plane * 8;   // elements between slices
  const size_t warp0 = (size_t)blockIdx.x * (kAttnThreads / 32) + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * (kAttnThreads / 32);
  for (size_t tile = warp0; tile < ntiles; tile += nwarps) {
    const size_t pix0 = tile * 16, b = pix0 / plane, r = pix0 % plane;
    const size_t base = (b * S * plane + r + g) * 8 + 2 * t;
    const __nv_bfloat16* src = F + base;
    __nv_bfloat16* dst = out + base;
    uint32_t p0 = 0, p1 = 0;                                  // slice z-1 (rows g, g+8)
    uint32_t c0 = ldg_nc_u32(src), c1 = ldg_nc_u32(src + 64);  // slice z
    uint32_t n0 = 0, n1 = 0, m0 = 0, m1 = 0;                  // slices z+1, z+2
    if (S > 1) { n0 = ldg_nc_u32(src + zs); n1 = ldg_nc_u32(src + zs + 64); }
    for (int z = 0; z < S; ++z) {
      if (z + 2 < S) { m0 = ldg_nc_u32(src + (size_t)(z + 2) * zs); m1 = ldg_nc_u32(src + (size_t)(z + 2) * zs + 64); }
      else { m0 = 0; m1 = 0; }
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      mma_16816(a, p0, p1, c0, c1, bm, bc);
      mma_1688(a, n0, n1, bp);
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      mma_1688(o, relu_pack(a[0], a[1]), relu_pack(a[2], a[3]), b2);
      *reinterpret_cast<uint32_t*>(dst + (size_t)z * zs) = relu_add_pack(o[0], o[1], c0);
      *reinterpret_cast<uint32_t*>(dst + (size_t)z * zs + 64) = relu_add_pack(o[2], o[3], c1);
      p0 = c0; p1 = c1; c0 = n0; c1 = n1; n0 = m0; n1 = m1;
    }
  }
}

// ---- C = 16, 32: one k16 block per (focal tap, 16-channel chunk); a thread holds channels 16i + 4t .. 4t+3 of its two pixels ----
template <int C>
__global__ void __launch_bounds__(kAttnThreads) srd_attention_mma_kernel(const __nv_bfloat16* __restrict__ F, const float* __restrict__ w0,
                                                                         const float* __restrict__ w1, __nv_bfloat16* __restrict__ out,
                                                                         int S, size_t plane, size_t ntiles) {
  constexpr int KC = C / 16, NT = C / 8;
  __shared__ uint2 sB1[3 * KC * NT * 32], sB2[KC * NT * 32];
  pdl_trigger();
  // B fragments in MMA order.  First product: logical k = 2t, 2t+1 | 2t+8, 2t+9 of block (dz, i) <-> input channels 16i+4t, +1 | +2, +3;
  // n = 8j + g natural.  Second product: k natural (the first product's accumulator fragment), output column g of tile j = 2i'+h
  // <-> channel 16i' + 4(g >> 1) + 2h + (g & 1), so the thread's accumulators are channels 16i'+4t .. 4t+3.
  for (int e = threadIdx.x; e < 3 * KC * NT * 32; e += kAttnThreads) {
    const int ln = e & 31, j = (e >> 5) % NT, i = (e / (32 * NT)) % KC, dz = e / (32 * NT * KC);
    const int gg = ln >> 2, tt = ln & 3, co = 8 * j + gg, ci = 16 * i + 4 * tt;
    const float* w = w0 + ((size_t)dz * C + ci) * C + co;
    sB1[e] = make_uint2(pack_bf2(__ldg(w), __ldg(w + C)), pack_bf2(__ldg(w + 2 * C), __ldg(w + 3 * C)));
  }
  for (int e = threadIdx.x; e < KC * NT * 32; e += kAttnThreads) {
    const int ln = e & 31, j = (e >> 5) % NT, i = e / (32 * NT);
    const int gg = ln >> 2, tt = ln & 3;
    const int co = 16 * (j >> 1) + 4 * (gg >> 1) + 2 * (j & 1) + (gg & 1), c = 16 * i + 2 * tt;
    const float* w = w1 + (size_t)c * C + co;
    sB2[e] = make_uint2(pack_bf2(__ldg(w), __ldg(w + C)), pack_bf2(__ldg(w + 8 * C), __ldg(w + 9 * C)));
  }
  __syncthreads();
  pdl_wait();   // F is the previous layer's output
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t zs = plane * C;
  const size_t warp0 = (size_t)blockIdx.x * (kAttnThreads / 32) + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * (kAttnThreads / 32);
  for (size_t tile = warp0; tile < ntiles; tile += nwarps) {
    const size_t pix0 = tile * 16, b = pix0 / plane, r = pix0 % plane;
    const size_t base = (b * S * plane + r + g) * C + 4 * t;
    const __nv_bfloat16* src = F + base;
    __nv_bfloat16* dst = out + base;
    uint2 pv[2][KC], cv[2][KC], nv[2][KC], mv[2][KC];   // slices z-1, z, z+1, z+2 ; [row g | g+8][chunk]
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < KC; ++i) {
        pv[h][i] = make_uint2(0u, 0u);
        cv[h][i] = ldg_nc_u64(src + h * 8 * C + 16 * i);
        nv[h][i] = S > 1 ? ldg_nc_u64(src + zs + h * 8 * C + 16 * i) : make_uint2(0u, 0u);
      }
    for (int z = 0; z < S; ++z) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < KC; ++i)
          mv[h][i] = z + 2 < S ? ldg_nc_u64(src + (size_t)(z + 2) * zs + h * 8 * C + 16 * i) : make_uint2(0u, 0u);
      float a[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) a[j][0] = a[j][1] = a[j][2] = a[j][3] = 0.f;
#pragma unroll
      for (int dz = 0; dz < 3; ++dz)
#pragma unroll
        for (int i = 0; i < KC; ++i) {
          const uint2 x0 = dz == 0 ? pv[0][i] : (dz == 1 ? cv[0][i] : nv[0][i]);
          const uint2 x1 = dz == 0 ? pv[1][i] : (dz == 1 ? cv[1][i] : nv[1][i]);
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const uint2 bw = sB1[((dz * KC + i) * NT + j) * 32 + lane];
            mma_16816(a[j], x0.x, x1.x, x0.y, x1.y, bw.x, bw.y);
          }
        }
      float o[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
      for (int i = 0; i < KC; ++i) {
        const uint32_t h0 = relu_pack(a[2 * i][0], a[2 * i][1]), h1 = relu_pack(a[2 * i][2], a[2 * i][3]);
        const uint32_t h2 = relu_pack(a[2 * i + 1][0], a[2 * i + 1][1]), h3 = relu_pack(a[2 * i + 1][2], a[2 * i + 1][3]);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const uint2 bw = sB2[(i * NT + j) * 32 + lane];
          mma_16816(o[j], h0, h1, h2, h3, bw.x, bw.y);
        }
      }
#pragma unroll
      for (int i = 0; i < KC; ++i) {
        uint2 r0, r1;
        r0.x = relu_add_pack(o[2 * i][0], o[2 * i][1], cv[0][i].x);
        r0.y = relu_add_pack(o[2 * i + 1][0], o[2 * i + 1][1], cv[0][i].y);
        r1.x = relu_add_pack(o[2 * i][2], o[2 * i][3], cv[1][i].x);
        r1.y = relu_add_pack(o[2 * i + 1][2], o[2 * i + 1][3], cv[1][i].y);
        *reinterpret_cast<uint2*>(dst + (size_t)z * zs + 16 * i) = r0;
        *reinterpret_cast<uint2*>(dst + (size_t)z * zs + 8 * C + 16 * i) = r1;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < KC; ++i) { pv[h][i] = cv[h][i]; cv[h][i] = nv[h][i]; nv[h][i] = mv[h][i]; }
    }
  }
}

// F, out: (B,S,H,W,C) bf16 ; w0: [3][C][C] fp32 ([dz][ci][co]) ; w1: [C][C] fp32 ([c][co]).  C = 8, 16 or 32; H*W % 16 == 0.
int launch_srd_attention_mma(const void* F, const float* w0, const float* w1, void* out, int B, int S, int H, int W, int C, int num_sms,
                             cudaStream_t st) {
  const size_t plane = (size_t)H * W;
  if (plane % 16) return fail(-5, "srd_attention: H*W must be a multiple of 16");
  const size_t ntiles = (size_t)B * plane / 16;
  const size_t want = (ntiles + kAttnThreads / 32 - 1) / (kAttnThreads / 32);
  const size_t cap = (size_t)(num_sms > 0 ? num_sms : 148) * 8 * (C == 32 ? 1 : 4);   // C = 32 rebuilds an 8 KB table per block: stay persistent
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const __nv_bfloat16* f = (const __nv_bfloat16*)F;
  __nv_bfloat16* o = (__nv_bfloat16*)out;
  if (C == 8) DFF_CUDA(launch_pdl(srd_attention_mma8_kernel, dim3(grid), dim3(kAttnThreads), 0, st, f, w0, w1, o, S, plane, ntiles));
  else if (C == 16) DFF_CUDA(launch_pdl(srd_attention_mma_kernel<16>, dim3(grid), dim3(kAttnThreads), 0, st, f, w0, w1, o, S, plane, ntiles));
  else if (C == 32) DFF_CUDA(launch_pdl(srd_attention_mma_kernel<32>, dim3(grid), dim3(kAttnThreads), 0, st, f, w0, w1, o, S, plane, ntiles));
  else return fail(-5, "srd_attention: C must be 8, 16 or 32");
  return 0;
}

}  // namespace dff
