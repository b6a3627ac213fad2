// Bandwidth kernels of the depth-from-focus path: layout conversion, pooling, the depth head, the FOV warp and the
// weight/BatchNorm packing.  All are coalesced, vectorised where alignment allows, and read every byte once.
#include <cstdlib>

#include "common.cuh"

namespace dff {

// ------------------------------------------------------------------------------------------------------------
// reference layout (B,C,S,H,W) fp32  ->  channels-last (B,S,H,W,Cp), zero padded channels
// (first thing DFF_net.forward needs: FS arrives as (B,3,S,H,W), reference :79-81)
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void to_cl_kernel(const float* __restrict__ src, T* __restrict__ dst, int B, int C, int S, int H, int W, int Cp) {
  const size_t npix = (size_t)B * S * H * W;
  const size_t plane = (size_t)S * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / plane, r = i % plane;
    const float* s = src + b * C * plane + r;
    T* d = dst + i * Cp;
    for (int c = 0; c < Cp; c += 4) {
      float4 v;
      v.x = c < C ? __ldg(s + (size_t)c * plane) : 0.f;
      v.y = c + 1 < C ? __ldg(s + (size_t)(c + 1) * plane) : 0.f;
      v.z = c + 2 < C ? __ldg(s + (size_t)(c + 2) * plane) : 0.f;
      v.w = c + 3 < C ? __ldg(s + (size_t)(c + 3) * plane) : 0.f;
      Elem<T>::store4(d + c, v);
    }
  }
}

template <typename T>
__global__ void from_cl_kernel(const T* __restrict__ src, float* __restrict__ dst, int B, int C, int S, int H, int W, int Cp) {
  const size_t npix = (size_t)B * S * H * W;
  const size_t plane = (size_t)S * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / plane, r = i % plane;
    for (int c = 0; c < C; ++c) dst[(b * C + c) * plane + r] = Elem<T>::load(src + i * Cp + c);
  }
}

// First-layer input of the tensor-core path: (B,S,H,W+2,8) bf16, column c = [R,G,B of pixel c-2 | R,G,B of pixel c | 0 0] (zeros
// outside the image).  The network's first convolution is 1x9x9 with dilation 2 on 3 channels (reference :144); with a pixel's
// dilated right neighbour already in its channel vector, two horizontal taps are ONE 8-channel tap: 9 x 5 taps instead of 9 x 9 —
// the K dimension of the implicit GEMM shrinks from 81 x 8 to 45 x 8 stored channels for the same 16 bytes per pixel.  The two
// margin columns make the pairs whose left pixel is outside the image (but whose right pixel is inside) addressable.
__global__ void to_cl_pair_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int S, int H, int W) {
  const int Wp = W + 2;
  const size_t npix = (size_t)B * S * H * Wp;
  const size_t plane = (size_t)S * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Wp);
    const size_t row = i / Wp;                 // (b*S + s)*H + y
    const size_t b = row / ((size_t)S * H), sy = row % ((size_t)S * H);
    const float* s = src + b * 3 * plane + sy * W;
    const bool lo = c >= 2, hi = c < W;
    float4 v0, v1;
    v0.x = lo ? __ldg(s + c - 2) : 0.f; v0.y = lo ? __ldg(s + plane + c - 2) : 0.f; v0.z = lo ? __ldg(s + 2 * plane + c - 2) : 0.f;
    v0.w = hi ? __ldg(s + c) : 0.f;
    v1.x = hi ? __ldg(s + plane + c) : 0.f;
    v1.y = hi ? __ldg(s + 2 * plane + c) : 0.f;
    v1.z = 0.f; v1.w = 0.f;
    Elem<__nv_bfloat16>::store4(dst + i * 8, v0);
    Elem<__nv_bfloat16>::store4(dst + i * 8 + 4, v1);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Input staging from the datasets' own format: uint8 focal stacks (B,S,H0,W0,3) exactly as the reference's
// dataloaders read them (Depth_Estimation_Test/test_Dataloader.py:122-147: hdf5 stack S x H x W x C uint8 -> `/127.5 - 1.0` in
// fp32 -> -1 padding of H, W to multiples of 32 -> transpose to C x S x H x W).  One pass does all of it and writes the layout the
// first convolution reads, so a quarter of the bytes cross PCIe and the fp32 (B,3,S,H,W) tensor never exists.
//   v = fl32(fl32(u8) / 127.5f) - 1.0f   (IEEE division and subtraction: bit-identical to numpy's float32 arithmetic)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float u8_norm(unsigned char u) { return __fsub_rn(__fdiv_rn((float)u, 127.5f), 1.0f); }

// -> pair-packed bf16 (B,S,H,W+2,8): the u8 twin of to_cl_pair_kernel (pixels with y >= H0 or x >= W0 are the -1 padding)
__global__ void u8_to_cl_pair_kernel(const unsigned char* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int S, int H0, int W0,
                                     int H, int W) {
  const int Wp = W + 2;
  const size_t npix = (size_t)B * S * H * Wp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Wp);
    const size_t row = i / Wp;                 // (b*S + s)*H + y
    const int y = (int)(row % H);
    const size_t bs = row / H;
    const bool yin = y < H0;
    const unsigned char* s = src + ((bs * H0 + (yin ? y : 0)) * (size_t)W0) * 3;
    float l[3], r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int xl = c - 2;
      l[k] = xl < 0 ? 0.f : ((yin && xl < W0) ? u8_norm(__ldg(s + (size_t)xl * 3 + k)) : -1.f);
      r[k] = c >= W ? 0.f : ((yin && c < W0) ? u8_norm(__ldg(s + (size_t)c * 3 + k)) : -1.f);
    }
    Elem<__nv_bfloat16>::store4(dst + i * 8, make_float4(l[0], l[1], l[2], r[0]));
    Elem<__nv_bfloat16>::store4(dst + i * 8 + 4, make_float4(r[1], r[2], 0.f, 0.f));
  }
}

// -> channels-last (B,S,H,W,Cp) fp32 | bf16 (fp32 parity mode, FFMA kernels), and -> the reference layout (B,3,S,H,W) fp32
template <typename T>
__global__ void u8_to_cl_kernel(const unsigned char* __restrict__ src, T* __restrict__ dst, int B, int S, int H0, int W0, int H, int W,
                                int Cp) {
  const size_t npix = (size_t)B * S * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const size_t bs = i / ((size_t)W * H);
    const bool in = y < H0 && x < W0;
    const unsigned char* s = src + ((bs * H0 + (in ? y : 0)) * (size_t)W0 + (in ? x : 0)) * 3;
    float4 v;
    v.x = in ? u8_norm(__ldg(s)) : -1.f; v.y = in ? u8_norm(__ldg(s + 1)) : -1.f; v.z = in ? u8_norm(__ldg(s + 2)) : -1.f; v.w = 0.f;
    Elem<T>::store4(dst + i * Cp, v);
    for (int c = 4; c < Cp; c += 4) Elem<T>::store4(dst + i * Cp + c, make_float4(0.f, 0.f, 0.f, 0.f));
  }
}
__global__ void u8_to_planar_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int B, int S, int H0, int W0, int H,
                                    int W) {
  const size_t plane = (size_t)S * H * W, npix = (size_t)B * plane;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const size_t bs = i / ((size_t)W * H), b = i / plane, r = i % plane;
    const bool in = y < H0 && x < W0;
    const unsigned char* s = src + ((bs * H0 + (in ? y : 0)) * (size_t)W0 + (in ? x : 0)) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) dst[(b * 3 + k) * plane + r] = in ? u8_norm(__ldg(s + k)) : -1.f;
  }
}

// W' (Cout, 8, 1, 9, 5) fp32 for the paired input: W'[co][ci][ky][j] = W[co][ci][ky][2j] (ci < 3), W[co][ci-3][ky][2j+1] (3 <= ci < 6,
// 2j+1 < 9), 0 otherwise.  W: (Cout, 3, 1, 9, 9).
__global__ void pair_weight_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout) {
  const int n = Cout * 8 * 9 * 5;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = i % 5, ky = (i / 5) % 9, ci = (i / 45) % 8, co = i / 360;
    float v = 0.f;
    if (ci < 3) v = w[((co * 3 + ci) * 9 + ky) * 9 + 2 * j];
    else if (ci < 6 && 2 * j + 1 < 9) v = w[((co * 3 + ci - 3) * 9 + ky) * 9 + 2 * j + 1];
    dst[i] = v;
  }
}

// The stride-2 (1,2,2) 3x3 layer on a source read as pixel pairs (.., W/2, 2C): dst (Cout, 2C, kd, 3, 2) with pair tap q = 0, 1 <-> pair
// offset -1, 0 and channel g*C + ci of a pair = pixel g of it: fine offset dx = 2*(q-1) + g in {-1, 0, 1} (dx = -2: zero).
__global__ void xpair_weight_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin, int kd) {
  const int n = Cout * 2 * Cin * kd * 6;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int q = i % 2, kb = (i / 2) % 3, ka = (i / 6) % kd, cp = (i / (6 * kd)) % (2 * Cin), co = i / (6 * kd * 2 * Cin);
    const int g = cp / Cin, ci = cp - g * Cin, dx = 2 * (q - 1) + g;
    dst[i] = dx < -1 ? 0.f : w[((((size_t)co * Cin + ci) * kd + ka) * 3 + kb) * 3 + (dx + 1)];
  }
}

static inline int grid_for(size_t n, int threads, int cap = 148 * 16) {
  size_t g = (n + threads - 1) / threads;
  return (int)(g < (size_t)cap ? (g ? g : 1) : cap);
}

int launch_to_cl(const float* src, int B, int C, int S, int H, int W, void* dst, int Cp, bool bf16, cudaStream_t st) {
  if (Cp % 4 || Cp < C) return fail(-1, "to_channels_last: Cp must be a multiple of 4 and >= C");
  const size_t n = (size_t)B * S * H * W;
  if (bf16) to_cl_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, (__nv_bfloat16*)dst, B, C, S, H, W, Cp);
  else to_cl_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, (float*)dst, B, C, S, H, W, Cp);
  DFF_LAUNCH_CHECK("to_cl");
  return 0;
}
int launch_to_cl_pair(const float* src, int B, int S, int H, int W, void* dst, cudaStream_t st) {
  const size_t n = (size_t)B * S * H * (W + 2);
  to_cl_pair_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, (__nv_bfloat16*)dst, B, S, H, W);
  DFF_LAUNCH_CHECK("to_cl_pair");
  return 0;
}
int launch_u8_to_cl_pair(const unsigned char* src, int B, int S, int H0, int W0, int H, int W, void* dst, cudaStream_t st) {
  const size_t n = (size_t)B * S * H * (W + 2);
  u8_to_cl_pair_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, (__nv_bfloat16*)dst, B, S, H0, W0, H, W);
  DFF_LAUNCH_CHECK("u8_to_cl_pair");
  return 0;
}
int launch_u8_to_cl(const unsigned char* src, int B, int S, int H0, int W0, int H, int W, void* dst, int Cp, bool bf16, cudaStream_t st) {
  if (Cp % 4 || Cp < 4) return fail(-1, "u8_to_channels_last: Cp must be a multiple of 4");
  const size_t n = (size_t)B * S * H * W;
  if (bf16) u8_to_cl_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, (__nv_bfloat16*)dst, B, S, H0, W0, H, W, Cp);
  else u8_to_cl_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, (float*)dst, B, S, H0, W0, H, W, Cp);
  DFF_LAUNCH_CHECK("u8_to_cl");
  return 0;
}
int launch_u8_to_planar(const unsigned char* src, int B, int S, int H0, int W0, int H, int W, float* dst, cudaStream_t st) {
  const size_t n = (size_t)B * S * H * W;
  u8_to_planar_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, dst, B, S, H0, W0, H, W);
  DFF_LAUNCH_CHECK("u8_to_planar");
  return 0;
}
int launch_pair_weight(const float* w, float* dst, int Cout, cudaStream_t st) {
  pair_weight_kernel<<<cdiv(Cout * 360, 256), 256, 0, st>>>(w, dst, Cout);
  DFF_LAUNCH_CHECK("pair_weight");
  return 0;
}
// x-grouped forms: the source (.., W, C) is read as groups of P pixels (.., W/P, P*C) and G = P / s adjacent outputs of a stride-s layer
// are the G*Cout channels of one GEMM row.  dst (G*Cout, P*C, kd, 3, nq) = the equivalent stride-1 convolution over groups with nq
// group taps (group offset q - 1): output g of a group reads the fine offsets s*g + dx, dx in {-1, 0, 1}, i.e. pixel p of the group
// q - 1 with P*(q - 1) + p = s*g + dx — a banded weight, zero elsewhere.
__global__ void xgroup_weight_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin, int kd, int s, int G, int P,
                                     int nq) {
  const int n = G * Cout * P * Cin * kd * 3 * nq;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int r = i;
    const int q = r % nq; r /= nq;
    const int kb = r % 3; r /= 3;
    const int ka = r % kd; r /= kd;
    const int cp = r % (P * Cin); r /= P * Cin;
    const int p = cp / Cin, ci = cp - p * Cin;
    const int g = r / Cout, co = r - g * Cout;
    const int dx = P * (q - 1) + p - s * g;
    dst[i] = (dx < -1 || dx > 1) ? 0.f : w[((((size_t)co * Cin + ci) * kd + ka) * 3 + kb) * 3 + (dx + 1)];
  }
}
int launch_xgroup_weight(const float* w, float* dst, int Cout, int Cin, int kd, int s, int G, int P, int nq, cudaStream_t st) {
  xgroup_weight_kernel<<<cdiv(G * Cout * P * Cin * kd * 3 * nq, 256), 256, 0, st>>>(w, dst, Cout, Cin, kd, s, G, P, nq);
  DFF_LAUNCH_CHECK("xgroup_weight");
  return 0;
}
int launch_xpair_weight(const float* w, float* dst, int Cout, int Cin, int kd, cudaStream_t st) {
  xpair_weight_kernel<<<cdiv(Cout * 2 * Cin * kd * 6, 256), 256, 0, st>>>(w, dst, Cout, Cin, kd);
  DFF_LAUNCH_CHECK("xpair_weight");
  return 0;
}
int launch_from_cl(const void* src, int B, int C, int S, int H, int W, int Cp, bool bf16, float* dst, cudaStream_t st) {
  const size_t n = (size_t)B * S * H * W;
  if (bf16) from_cl_kernel<<<grid_for(n, 256), 256, 0, st>>>((const __nv_bfloat16*)src, dst, B, C, S, H, W, Cp);
  else from_cl_kernel<<<grid_for(n, 256), 256, 0, st>>>((const float*)src, dst, B, C, S, H, W, Cp);
  DFF_LAUNCH_CHECK("from_cl");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// (1,k,k) pooling on channels-last volumes: MaxPool3d((1,2,2)) of EFD (reference :387) and the AvgPool3d
// pyramid of hourglassup (reference :183-187, 248-250).  One thread = one output pixel x 4 channels.
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool kMax>
__global__ void pool_kernel(const T* __restrict__ src, T* __restrict__ dst, int BS, int H, int W, int C, int k) {
  const int OH = H / k, OW = W / k, C4 = C / 4;
  const size_t n = (size_t)BS * OH * OW * C4;
  const float inv = 1.f / (float)(k * k);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    size_t r = i / C4;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const size_t bs = r / OH;
    const T* p = src + ((bs * H + (size_t)oy * k) * W + (size_t)ox * k) * C + 4 * c4;
    float4 acc = kMax ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        const float4 v = Elem<T>::load4(p + ((size_t)dy * W + dx) * C);
        if (kMax) {
          acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
        } else {
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    if (!kMax) { acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv; }
    Elem<T>::store4(dst + i * 4, acc);
  }
}

// MaxPool3d((1,2,2)) on bf16 channels-last volumes, 8 channels (16 bytes) per thread: four 16-byte loads, packed bf16x2 maxima (exact),
// one 16-byte store — the generic kernel above spends its time on 8-byte accesses and fp32 conversions.
__global__ void __launch_bounds__(256) maxpool2_bf16x8_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int BS, int H, int W, int C8) {
  const int OH = H / 2, OW = W / 2;
  const size_t n = (size_t)BS * OH * OW * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    size_t r = i / C8;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const size_t bs = r / OH;
    const uint4* p = src + ((bs * H + 2 * oy) * W + 2 * ox) * C8 + c8;
    const uint4 a = __ldg(p), b = __ldg(p + C8), c = __ldg(p + (size_t)W * C8), d = __ldg(p + (size_t)W * C8 + C8);
    auto mx = [](uint32_t x, uint32_t y) {
      const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&x), *reinterpret_cast<const __nv_bfloat162*>(&y));
      return *reinterpret_cast<const uint32_t*>(&m);
    };
    uint4 o;
    o.x = mx(mx(a.x, b.x), mx(c.x, d.x)); o.y = mx(mx(a.y, b.y), mx(c.y, d.y));
    o.z = mx(mx(a.z, b.z), mx(c.z, d.z)); o.w = mx(mx(a.w, b.w), mx(c.w, d.w));
    dst[i] = o;
  }
}

// The three average pools of hourglassup (AvgPool3d (1,2,2), (1,4,4), (1,8,8) of the same volume, reference :183-187, 248-250) in ONE
// pass over the bf16 channels-last input: a lane owns one 2x2 block x 8 channels (four 16-byte loads), 16 lanes cover an 8x8 block,
// the 4x4 and 8x8 means come from the same fp32 sums through two rounds of lane shuffles.  Reads the volume once instead of three times.
__global__ void __launch_bounds__(256) avgpool_pyramid_bf16_kernel(const uint4* __restrict__ src, uint4* __restrict__ d2, uint4* __restrict__ d4,
                                                                   uint4* __restrict__ d8, int BS, int H, int W, int C8) {
  const int H8 = H / 8, W8 = W / 8;
  const size_t nblk = (size_t)BS * H8 * W8 * C8;               // (8x8 block, 8-channel chunk) pairs, 16 lanes each
  const size_t gi = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 4;
  if (gi >= nblk) return;                                      // (whole 16-lane groups leave together: nblk is counted in groups)
  const int l = threadIdx.x & 15, bx = l & 3, by = l >> 2;     // 2x2 block inside the 8x8 block
  const int c8 = (int)(gi % C8);
  size_t r = gi / C8;
  const int X8 = (int)(r % W8); r /= W8;
  const int Y8 = (int)(r % H8);
  const size_t bs = r / H8;
  const int y = 8 * Y8 + 2 * by, x = 8 * X8 + 2 * bx;
  const uint4* p = src + ((bs * H + y) * W + x) * C8 + c8;
  const uint4 v[4] = {__ldg(p), __ldg(p + C8), __ldg(p + (size_t)W * C8), __ldg(p + (size_t)W * C8 + C8)};
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t u[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] += __uint_as_float(u[j] << 16);
      acc[2 * j + 1] += __uint_as_float(u[j] & 0xffff0000u);
    }
  }
  auto pack = [&](float sc) {
    uint4 o;
    uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * j] * sc, acc[2 * j + 1] * sc);
      ou[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return o;
  };
  d2[((bs * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)) * C8 + c8] = pack(0.25f);
  // 4x4: the four 2x2 blocks with equal (bx >> 1, by >> 1): lane bits 0 and 2
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
  }
  if ((bx & 1) == 0 && (by & 1) == 0) d4[((bs * (H / 4) + (y >> 2)) * (W / 4) + (x >> 2)) * C8 + c8] = pack(1.f / 16.f);
  // 8x8: lane bits 1 and 3
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
  }
  if (l == 0) d8[((bs * H8 + Y8) * W8 + X8) * C8 + c8] = pack(1.f / 64.f);
}

int launch_avgpool_pyramid(const void* src, void* d2, void* d4, void* d8, int BS, int H, int W, int C, cudaStream_t st) {
  if (C % 8 || H % 8 || W % 8) return fail(-1, "avgpool_pyramid: C, H, W must be multiples of 8");
  const size_t nthreads = (size_t)BS * (H / 8) * (W / 8) * (C / 8) * 16;
  avgpool_pyramid_bf16_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>((const uint4*)src, (uint4*)d2, (uint4*)d4, (uint4*)d8, BS, H, W,
                                                                                 C / 8);
  DFF_LAUNCH_CHECK("avgpool_pyramid");
  return 0;
}

int launch_pool(const void* src, void* dst, int BS, int H, int W, int C, int k, bool is_max, bool bf16, cudaStream_t st) {
  if (C % 4 || H % k || W % k) return fail(-1, "pool: C % 4, H % k, W % k must be 0");
  if (bf16 && is_max && k == 2 && C % 8 == 0) {
    const size_t n8 = (size_t)BS * (H / 2) * (W / 2) * (C / 8);
    maxpool2_bf16x8_kernel<<<grid_for(n8, 256, 148 * 32), 256, 0, st>>>((const uint4*)src, (uint4*)dst, BS, H, W, C / 8);
    DFF_LAUNCH_CHECK("maxpool2");
    return 0;
  }
  const size_t n = (size_t)BS * (H / k) * (W / k) * (C / 4);
  const int g = grid_for(n, 256);
  if (bf16) {
    if (is_max) pool_kernel<__nv_bfloat16, true><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, BS, H, W, C, k);
    else pool_kernel<__nv_bfloat16, false><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, BS, H, W, C, k);
  } else {
    if (is_max) pool_kernel<float, true><<<g, 256, 0, st>>>((const float*)src, (float*)dst, BS, H, W, C, k);
    else pool_kernel<float, false><<<g, 256, 0, st>>>((const float*)src, (float*)dst, BS, H, W, C, k);
  }
  DFF_LAUNCH_CHECK("pool");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Depth head (reference :92-98, 118-136): bilinear upsample (align_corners=False) of the per-slice cost,
// softplus(beta=1, threshold=20) + 1e-6, normalise over the S slices, expectation of the focus distance.
// One thread per output pixel, slices walked in order; cost rows are re-used through L1/L2, focus_dists and the
// output are touched exactly once.  num/den form: depth = (sum fd*p) / (sum p).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_ref(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__global__ void depth_head_kernel(const float* __restrict__ cost, int h, int w, const float* __restrict__ fd, long long sb,
                                  long long ss, long long sy, long long sx, int B, int S, int H, int W,
                                  float* __restrict__ depth) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= W) return;
  // source coordinates exactly as ATen's upsample_bilinear2d: src = scale*(dst+0.5)-0.5, clamped at 0
  const float ry = (float)h / (float)H, rx = (float)w / (float)W;
  float fy = ry * ((float)y + 0.5f) - 0.5f, fx = rx * ((float)x + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const bool same = (h == H) && (w == W);
  float num = 0.f, den = 0.f;
  const float* fp = fd + b * sb + y * sy + x * sx;
  for (int s = 0; s < S; ++s) {
    const float* c = cost + ((size_t)b * S + s) * h * w;
    float v;
    if (same) v = __ldg(c + (size_t)y * w + x);
    else
      v = ly0 * (lx0 * __ldg(c + (size_t)y0 * w + x0) + lx1 * __ldg(c + (size_t)y0 * w + x1)) +
          ly1 * (lx0 * __ldg(c + (size_t)y1 * w + x0) + lx1 * __ldg(c + (size_t)y1 * w + x1));
    const float p = softplus_ref(v) + 1e-6f;
    den += p;
    num = fmaf(__ldg(fp + s * ss), p, num);
  }
  depth[((size_t)b * H + y) * W + x] = num / den;
}

// All four heads of DFF_net in one pass: `focus_dists` (the largest operand, S*H*W fp32) is read once instead of four times.
struct Head4 {
  const float* cost[4];
  int h[4], w[4];
  float* depth[4];
};
// FAST (bf16 mode): exp/log through the SFU (relative error ~2e-7 on p = softplus + 1e-6, far inside that mode's tolerance).
template <bool FAST>
__device__ __forceinline__ float softplus_p(float v) {
  if (FAST) return (v > 20.f ? v : __logf(1.f + __expf(v))) + 1e-6f;
  return softplus_ref(v) + 1e-6f;
}
// One thread = one output pixel x one PAIR of heads (blockIdx.z parity): half the registers per thread, twice the threads —
// the kernel is latency-bound, occupancy is what it needs (the second read of focus_dists hits L1/L2).
template <bool FAST, int HP>
__global__ void __launch_bounds__(128, 10) depth_head4_kernel(const __grid_constant__ Head4 a, const float* __restrict__ fd, long long sb,
                                                              long long ss, long long sy, long long sx, int B, int S, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z / (4 / HP), hp = blockIdx.z % (4 / HP);
  if (x >= W) return;
  // per head of the pair: the four bilinear tap offsets inside a slice (align_corners=False) and the interpolation weights;
  // a full-resolution head degenerates to one tap with weight 1
  int o00[HP], o01[HP], o10[HP], o11[HP], sl[HP];
  float ly1[HP], lx1[HP];
  const float* c[HP];
#pragma unroll
  for (int k = 0; k < HP; ++k) {
    const int hd = HP * hp + k, h = a.h[hd], w = a.w[hd];
    const float ry = (float)h / (float)H, rx = (float)w / (float)W;
    float fy = ry * ((float)y + 0.5f) - 0.5f, fx = rx * ((float)x + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    ly1[k] = fy - (float)y0; lx1[k] = fx - (float)x0;
    o00[k] = y0 * w + x0; o01[k] = y0 * w + x1; o10[k] = y1 * w + x0; o11[k] = y1 * w + x1;
    sl[k] = h * w;
    c[k] = a.cost[hd] + (size_t)b * S * sl[k];
  }
  float num[HP], den[HP];
#pragma unroll
  for (int k = 0; k < HP; ++k) num[k] = den[k] = 0.f;
  const float* fp = fd + b * sb + y * sy + x * sx;
#pragma unroll 2
  for (int s = 0; s < S; ++s) {
    const float f = __ldg(fp + s * ss);
#pragma unroll
    for (int k = 0; k < HP; ++k) {
      const float* p = c[k];
      const float v = (1.f - ly1[k]) * ((1.f - lx1[k]) * __ldg(p + o00[k]) + lx1[k] * __ldg(p + o01[k])) +
                      ly1[k] * ((1.f - lx1[k]) * __ldg(p + o10[k]) + lx1[k] * __ldg(p + o11[k]));
      c[k] += sl[k];
      const float pr = softplus_p<FAST>(v);
      den[k] += pr;
      num[k] = fmaf(f, pr, num[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < HP; ++k) a.depth[HP * hp + k][((size_t)b * H + y) * W + x] = num[k] / den[k];
}


// ---- the four heads of DFF_net, four pixels per thread (bf16 throughput mode) ------------------------------------------------------
// The one-pixel-per-thread kernel above is instruction-bound (~60 instructions per softplus evaluation: four scalar taps with 64-bit
// address arithmetic and a separate bilinear blend per pixel, head and slice).  The heads' resolutions are fixed — 1/8, 1/4, 1/2, 1/1
// (reference :92-98, 118-121) — so for four horizontally adjacent output pixels the source columns and blend weights are a fixed
// pattern per ratio: a thread loads the 2-4 source columns of its two source rows once per slice, blends vertically per column, then
// horizontally per pixel (2 FMAs each), and evaluates softplus with two SFU operations.  Edge behaviour (ATen's clamped source index)
// = replicate-edge column loads with the blend weight forced to 0 where ATen's taps coincide.  focus_dists is read once for all four
// heads: as float4 when it is contiguous in x, as one broadcast scalar per slice when it is the S focus distances of the stack.
template <int R> struct QuadTaps {
  static constexpr int NC = R == 1 ? 4 : (R == 2 ? 4 : (R == 4 ? 3 : 2));   // source columns touched by 4 adjacent output pixels
  int col[NC];        // clamped source columns
  int o0, o1;         // source row offsets (elements)
  float ly, lx[4];
  __device__ __forceinline__ void init(int t, int y, int h, int w, int H) {
    // rows: ATen's upsample_bilinear2d (align_corners=False): src = scale*(dst+0.5)-0.5, clamped at 0 (scale = 1/R exactly)
    float fy = (1.f / (float)R) * ((float)y + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    const int y0 = (int)fy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
    ly = fy - (float)y0;
    o0 = y0 * w; o1 = y1 * w;
    const int x = 4 * t;
    const int c0 = R == 1 ? x : (int)floorf(((float)x + 0.5f) * (1.f / (float)R) - 0.5f);
#pragma unroll
    for (int j = 0; j < NC; ++j) col[j] = min(max(c0 + j, 0), w - 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float fx = ((float)(x + k) + 0.5f) * (1.f / (float)R) - 0.5f;     // exact: R is a power of two
      const float x0 = floorf(fx);
      lx[k] = (R == 1 || x0 < 0.f || x0 >= (float)(w - 1)) ? 0.f : fx - x0;
    }
  }
  // column index (relative to col[0]) of the left tap of output pixel k
  static __device__ __forceinline__ constexpr int idx(int k) { return R == 1 ? k : (R == 2 ? (k + 1) / 2 : (R == 4 ? k / 2 : 0)); }
  __device__ __forceinline__ void eval(const float* __restrict__ c, float* v) const {
    if (R == 1) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(c + o0 + col[0]));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      return;
    }
    float cv[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const float a = __ldg(c + o0 + col[j]), b = __ldg(c + o1 + col[j]);
      cv[j] = fmaf(ly, b - a, a);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = idx(k);
      v[k] = fmaf(lx[k], cv[i + 1 < NC ? i + 1 : NC - 1] - cv[i], cv[i]);
    }
  }
};

// p = softplus(v) + 1e-6 with two SFU operations and fp32-grade accuracy everywhere: max(v,0) + log1p(exp(-|v|)), where log1p of a
// small argument is its series (the plain log(1 + e) loses e's low bits below ~1e-3 — visible when every slice of a pixel is
// strongly negative and the normalisation divides two sums of such terms).
// (returns softplus(v) WITHOUT the +1e-6: the caller adds S*1e-6 to the denominator and 1e-6*sum(fd) to the numerator once per pixel)
__device__ __forceinline__ float softplus_sfu(float v) {
  float e, l;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(v) * -1.4426950408889634f));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.f + e));
  l *= 0.6931471805599453f;
  const float ser = e * fmaf(e, -0.5f, 1.f);      // log1p(e) for small e (error e^3/3 < 2e-8)
  return fmaxf(v, 0.f) + (e < 3.9e-3f ? ser : l);
}
__device__ __forceinline__ void cp_async16_ca(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
constexpr int kHeadRing = 8;   // slices of full-resolution operands in flight per thread (cp.async ring in shared memory)

// The full-resolution operands (cost3 and, when it is tiled, focus_dists: 4 + 4 of the ~9.3 bytes per pixel and slice) are what has
// to come from DRAM; with plain loads a thread has one slice in flight and 512 threads per SM cover a fifth of the latency-bandwidth
// product.  Each thread therefore runs an 8-deep cp.async ring of its own 16-byte pieces (no barrier: a thread only reads what it
// copied itself); the low-resolution heads' taps are L1/L2 hits.
template <int FDMODE>   // 0: generic strides, 1: contiguous in x (float4), 2: one scalar per slice
__global__ void __launch_bounds__(128, 4) depth_head4_quad_kernel(const __grid_constant__ Head4 a, const float* __restrict__ fd, long long sb,
                                                                  long long ss, long long sy, long long sx, int B, int S, int H, int W) {
  __shared__ float4 ring_c[kHeadRing][128];
  __shared__ float4 ring_f[FDMODE == 1 ? kHeadRing : 1][128];
  // block = 32 quads (128 pixels) x 4 rows; grid = (row segments, H / 4, B)
  const int W4 = W >> 2;
  int t = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 5), b = blockIdx.z;
  const bool live = t < W4;
  if (!live) t = 0;   // (keeps the thread's loads in range; it stores nothing)
  QuadTaps<8> t0; QuadTaps<4> t1; QuadTaps<2> t2;
  t0.init(t, y, a.h[0], a.w[0], H); t1.init(t, y, a.h[1], a.w[1], H); t2.init(t, y, a.h[2], a.w[2], H);
  const int sl0 = a.h[0] * a.w[0], sl1 = a.h[1] * a.w[1], sl2 = a.h[2] * a.w[2];
  const size_t sl3 = (size_t)H * W;
  const float* c0 = a.cost[0] + (size_t)b * S * sl0;
  const float* c1 = a.cost[1] + (size_t)b * S * sl1;
  const float* c2 = a.cost[2] + (size_t)b * S * sl2;
  const float* c3 = a.cost[3] + (size_t)b * S * sl3 + (size_t)y * W + 4 * t;
  const float* fp = fd + b * sb + y * sy + (long long)(4 * t) * sx;
#pragma unroll
  for (int d = 0; d < kHeadRing; ++d) {
    if (d < S) {
      cp_async16_ca(&ring_c[d][threadIdx.x], c3 + (size_t)d * sl3);
      if (FDMODE == 1) cp_async16_ca(&ring_f[d][threadIdx.x], fp + d * ss);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float num[4][4], den[4][4], fsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int hd = 0; hd < 4; ++hd)
#pragma unroll
    for (int k = 0; k < 4; ++k) num[hd][k] = den[hd][k] = 0.f;
  for (int s = 0; s < S; ++s) {
    float v[4][4];
    t0.eval(c0, v[0]); t1.eval(c1, v[1]); t2.eval(c2, v[2]);
    c0 += sl0; c1 += sl1; c2 += sl2;
    float f[4];
    if (FDMODE == 2) { f[0] = f[1] = f[2] = f[3] = __ldg(fp + s * ss); }
    else if (FDMODE == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = __ldg(fp + s * ss + k * sx);
    }
    asm volatile("cp.async.wait_group %0;" ::"n"(kHeadRing - 1) : "memory");
    const int slot = s % kHeadRing;
    {
      const float4 c = ring_c[slot][threadIdx.x];
      v[3][0] = c.x; v[3][1] = c.y; v[3][2] = c.z; v[3][3] = c.w;
      if (FDMODE == 1) { const float4 g = ring_f[slot][threadIdx.x]; f[0] = g.x; f[1] = g.y; f[2] = g.z; f[3] = g.w; }
    }
    if (s + kHeadRing < S) {
      cp_async16_ca(&ring_c[slot][threadIdx.x], c3 + (size_t)(s + kHeadRing) * sl3);
      if (FDMODE == 1) cp_async16_ca(&ring_f[slot][threadIdx.x], fp + (long long)(s + kHeadRing) * ss);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 4; ++k) fsum[k] += f[k];
#pragma unroll
    for (int hd = 0; hd < 4; ++hd)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float pr = softplus_sfu(v[hd][k]);
        den[hd][k] += pr;
        num[hd][k] = fmaf(f[k], pr, num[hd][k]);
      }
  }
  if (!live) return;
  // p = softplus + 1e-6 (reference :93): the constant enters the two sums once per pixel
  const float eps_den = 1e-6f * (float)S;
  const size_t o = ((size_t)b * H + y) * W + 4 * t;
#pragma unroll
  for (int hd = 0; hd < 4; ++hd) {
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = __fdividef(fmaf(1e-6f, fsum[k], num[hd][k]), den[hd][k] + eps_den);
    *reinterpret_cast<float4*>(a.depth[hd] + o) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// cost[0..2]: upsampled heads (any resolution dividing H, W) ; cost[3]: the full-resolution head.  fast: SFU exp/log.
int launch_depth_head4(const float* const cost[4], const int h[4], const int w[4], const float* fd, const int64_t* st4, int B, int S,
                       int H, int W, float* const depth[4], bool fast, cudaStream_t st) {
  Head4 a;
  for (int k = 0; k < 4; ++k) {
    if (h[k] <= 0 || w[k] <= 0 || H % h[k] || W % w[k]) return fail(-1, "depth_head: H,W must be multiples of the cost resolution");
    a.cost[k] = cost[k]; a.h[k] = h[k]; a.w[k] = w[k]; a.depth[k] = depth[k];
  }
  static const int hp = getenv("DFF_HEAD_HP") ? atoi(getenv("DFF_HEAD_HP")) : 2;
  static const bool no_quad = getenv("DFF_HEAD_NO_QUAD") != nullptr;
  const bool pyramid = H == 8 * h[0] && W == 8 * w[0] && H == 4 * h[1] && W == 4 * w[1] && H == 2 * h[2] && W == 2 * w[2] && H == h[3] &&
                       W == w[3] && w[0] >= 2 && h[0] >= 1;
  bool aligned = (reinterpret_cast<uintptr_t>(cost[3]) % 16 == 0);
  for (int k = 0; k < 4; ++k) aligned = aligned && (reinterpret_cast<uintptr_t>(depth[k]) % 16 == 0);
  if (fast && pyramid && aligned && !no_quad && B <= 65535) {
    const dim3 g(cdiv(W / 4, 32), H / 4, B);
    const bool fd_scalar = st4[2] == 0 && st4[3] == 0;
    const bool fd_vec = st4[3] == 1 && st4[2] % 4 == 0 && st4[1] % 4 == 0 && st4[0] % 4 == 0 && reinterpret_cast<uintptr_t>(fd) % 16 == 0;
    if (fd_scalar) depth_head4_quad_kernel<2><<<g, 128, 0, st>>>(a, fd, st4[0], st4[1], st4[2], st4[3], B, S, H, W);
    else if (fd_vec) depth_head4_quad_kernel<1><<<g, 128, 0, st>>>(a, fd, st4[0], st4[1], st4[2], st4[3], B, S, H, W);
    else depth_head4_quad_kernel<0><<<g, 128, 0, st>>>(a, fd, st4[0], st4[1], st4[2], st4[3], B, S, H, W);
    DFF_LAUNCH_CHECK("depth_head4_quad");
    return 0;
  }
  dim3 grid(cdiv(W, 128), H, B * (4 / hp));
  if (fast) { if (hp == 1) depth_head4_kernel<true, 1><<<grid, 128, 0, st>>>(a, fd, st4[0], st4[1], st4[2], st4[3], B, S, H, W); else depth_head4_kernel<true, 2><<<grid, 128, 0, st>>>(a, fd, st4[0], st4[1], st4[2], st4[3], B, S, H, W); }
  else depth_head4_kernel<false, 2><<<grid, 128, 0, st>>>(a, fd, st4[0], st4[1], st4[2], st4[3], B, S, H, W);
  DFF_LAUNCH_CHECK("depth_head4");
  return 0;
}

int launch_depth_head(const float* cost, int h, int w, const float* fd, const int64_t* st4, int B, int S, int H, int W,
                      float* depth, cudaStream_t st) {
  if (h <= 0 || w <= 0 || H % h || W % w) return fail(-1, "depth_head: H,W must be multiples of the cost resolution");
  dim3 grid(cdiv(W, 128), H, B);
  depth_head_kernel<<<grid, 128, 0, st>>>(cost, h, w, fd, st4[0], st4[1], st4[2], st4[3], B, S, H, W, depth);
  DFF_LAUNCH_CHECK("depth_head");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// FOV warp (reference End_to_End/End_to_End.py:106-134) with analytic sampling coordinates — no grid tensor.
//   f = fov[b,s] + a0 ; flow_x = (W//2)(f-1)*lin(-1,1,W)[x] + a1 ; flow_y likewise ; sample at (x-flow_x, y-flow_y)
//   through grid_sample's normalise/un-normalise round trip (align_corners=True), bilinear, zero padding.
// The scale correction a0 is taken from sample 0 (the reference's batch>1 broadcast quirk).
// ------------------------------------------------------------------------------------------------------------
__global__ void fov_warp_kernel(const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ fov,
                                int B, int C, int S, int H, int W, float* __restrict__ out, float* __restrict__ flow) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y;
  const int bs = blockIdx.z, b = bs / S, s = bs % S;
  if (px >= W) return;
  const float a0 = alpha ? __ldg(alpha + (0 * 3 + 0) * S + s) : 0.f;  // sample 0 on purpose
  const float a1 = alpha ? __ldg(alpha + ((size_t)b * 3 + 1) * S + s) : 0.f;
  const float a2 = alpha ? __ldg(alpha + ((size_t)b * 3 + 2) * S + s) : 0.f;
  const float f = a0 + __ldg(fov + (size_t)b * S + s);
  // torch.linspace(-1, 1, n): start + i*step for i < n/2, end - (n-1-i)*step otherwise
  auto lin = [](int i, int n) -> float {
    if (n == 1) return -1.f;
    const float step = 2.f / (float)(n - 1);
    return i < n / 2 ? -1.f + step * (float)i : 1.f - step * (float)(n - 1 - i);
  };
  const float flx = (float)(W / 2) * (f - 1.f) * lin(px, W) + a1;
  const float fly = (float)(H / 2) * (f - 1.f) * lin(py, H) + a2;
  if (flow) {
    const size_t plane = (size_t)S * H * W, o = ((size_t)s * H + py) * W + px;
    flow[((size_t)b * 2 + 0) * plane + o] = flx;
    flow[((size_t)b * 2 + 1) * plane + o] = fly;
  }
  // normalise (as the reference does) then un-normalise (as grid_sample does, align_corners=True)
  const float gx = 2.0f * ((float)px - flx) / (float)max(W - 1, 1) - 1.0f;
  const float gy = 2.0f * ((float)py - fly) / (float)max(H - 1, 1) - 1.0f;
  const float ix = (gx + 1.f) * 0.5f * (float)(W - 1);
  const float iy = (gy + 1.f) * 0.5f * (float)(H - 1);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  const float tx = ix - fx0, ty = iy - fy0;
  // trilinear weights degenerate to bilinear: the z coordinate is the slice itself
  const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
  for (int c = 0; c < C; ++c) {
    const float* p = x + (((size_t)b * C + c) * S + s) * H * W;
    float v = 0.f;
    if (vy0 && vx0) v += __ldg(p + (size_t)y0 * W + x0) * w00;
    if (vy0 && vx1) v += __ldg(p + (size_t)y0 * W + x1) * w01;
    if (vy1 && vx0) v += __ldg(p + (size_t)y1 * W + x0) * w10;
    if (vy1 && vx1) v += __ldg(p + (size_t)y1 * W + x1) * w11;
    out[((((size_t)b * C + c) * S + s) * H + py) * W + px] = v;
  }
}

// Four horizontally adjacent output pixels per thread: the row geometry is computed once, the four column geometries are
// independent instruction streams, every channel contributes 16 tap loads in flight per thread and the results leave as one
// 16-byte store per channel (the one-pixel kernel above is latency-bound at a fifth of the HBM bandwidth).  Same arithmetic, same
// order, same rounding as fov_warp_kernel.
__device__ __forceinline__ float fov_lin(int i, int n, float step) {   // torch.linspace(-1, 1, n)[i], step = 2/(n-1) in fp32
  if (n == 1) return -1.f;
  return i < n / 2 ? -1.f + step * (float)i : 1.f - step * (float)(n - 1 - i);
}
template <bool FLOW>
__global__ void __launch_bounds__(128) fov_warp_quad_kernel(const float* __restrict__ x, const float* __restrict__ alpha,
                                                            const float* __restrict__ fov, int B, int C, int S, int H, int W,
                                                            float stepW, float stepH, float* __restrict__ out,
                                                            float* __restrict__ flow) {
  const int px0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int py = blockIdx.y;
  const int bs = blockIdx.z, b = bs / S, s = bs % S;
  if (px0 >= W) return;
  const float a0 = alpha ? __ldg(alpha + (0 * 3 + 0) * S + s) : 0.f;  // sample 0 on purpose (reference broadcast quirk)
  const float a1 = alpha ? __ldg(alpha + ((size_t)b * 3 + 1) * S + s) : 0.f;
  const float a2 = alpha ? __ldg(alpha + ((size_t)b * 3 + 2) * S + s) : 0.f;
  const float f = a0 + __ldg(fov + (size_t)b * S + s);
  const float fly = (float)(H / 2) * (f - 1.f) * fov_lin(py, H, stepH) + a2;
  const float gy = 2.0f * ((float)py - fly) / (float)max(H - 1, 1) - 1.0f;
  const float iy = (gy + 1.f) * 0.5f * (float)(H - 1);
  const float fy0 = floorf(iy);
  const int y0 = (int)fy0, y1 = y0 + 1;
  const float ty = iy - fy0;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
  int x0[4];
  float w00[4], w01[4], w10[4], w11[4], flx[4];
  bool vx0[4], vx1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int px = px0 + k;
    flx[k] = (float)(W / 2) * (f - 1.f) * fov_lin(px, W, stepW) + a1;
    const float gx = 2.0f * ((float)px - flx[k]) / (float)max(W - 1, 1) - 1.0f;
    const float ix = (gx + 1.f) * 0.5f * (float)(W - 1);
    const float fx0 = floorf(ix);
    x0[k] = (int)fx0;
    const float tx = ix - fx0;
    w00[k] = (1.f - tx) * (1.f - ty); w01[k] = tx * (1.f - ty); w10[k] = (1.f - tx) * ty; w11[k] = tx * ty;
    vx0[k] = x0[k] >= 0 && x0[k] < W; vx1[k] = x0[k] + 1 >= 0 && x0[k] + 1 < W;
  }
  const size_t plane = (size_t)H * W;
  if (FLOW) {
    const size_t fo = (((size_t)b * 2 * S + s) * H + py) * W + px0;
    *reinterpret_cast<float4*>(flow + fo) = make_float4(flx[0], flx[1], flx[2], flx[3]);
    *reinterpret_cast<float4*>(flow + fo + (size_t)S * plane) = make_float4(fly, fly, fly, fly);
  }
  const float* p = x + ((size_t)b * C * S + s) * plane;
  float* o = out + (((size_t)b * C * S + s) * H + py) * W + px0;
  for (int c = 0; c < C; ++c, p += (size_t)S * plane, o += (size_t)S * plane) {
    const float* r0 = p + (size_t)y0 * W;
    const float* r1 = r0 + W;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float acc = 0.f;
      if (vy0 && vx0[k]) acc += __ldg(r0 + x0[k]) * w00[k];
      if (vy0 && vx1[k]) acc += __ldg(r0 + x0[k] + 1) * w01[k];
      if (vy1 && vx0[k]) acc += __ldg(r1 + x0[k]) * w10[k];
      if (vy1 && vx1[k]) acc += __ldg(r1 + x0[k] + 1) * w11[k];
      v[k] = acc;
    }
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

int launch_fov_warp(const float* x, const float* alpha, const float* fov, int B, int C, int S, int H, int W, float* out,
                    float* flow, cudaStream_t st) {
  const bool vec = W % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 && (!flow || reinterpret_cast<uintptr_t>(flow) % 16 == 0);
  static const bool no_quad = getenv("DFF_FOV_NO_QUAD") != nullptr;
  if (vec && !no_quad) {
    // (the linspace steps are per-call constants: computed here in fp32 exactly as the device would, 2.f / (float)(n - 1))
    const float stepW = W > 1 ? 2.f / (float)(W - 1) : 0.f, stepH = H > 1 ? 2.f / (float)(H - 1) : 0.f;
    const int bt = (W / 4) % 128 == 0 ? 128 : ((W / 4) % 64 == 0 ? 64 : ((W / 4) % 96 == 0 ? 96 : ((W / 4) <= 64 ? 64 : 128)));
    dim3 grid(cdiv(W / 4, bt), H, B * S);
    if (flow) fov_warp_quad_kernel<true><<<grid, bt, 0, st>>>(x, alpha, fov, B, C, S, H, W, stepW, stepH, out, flow);
    else fov_warp_quad_kernel<false><<<grid, bt, 0, st>>>(x, alpha, fov, B, C, S, H, W, stepW, stepH, out, flow);
    DFF_LAUNCH_CHECK("fov_warp_quad");
    return 0;
  }
  dim3 grid(cdiv(W, 128), H, B * S);
  fov_warp_kernel<<<grid, 128, 0, st>>>(x, alpha, fov, B, C, S, H, W, out, flow);
  DFF_LAUNCH_CHECK("fov_warp");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// SRD channel-attention branch in ONE pass (reference Feature_Extraction / SRD, train_codes/Depth_Estimation_Network.py:399-407):
//   out = F + relu( W1 . relu( conv3x1x1(F; W0) ) )          (no BatchNorm, no bias)
// Two tiny convolutions (3 taps along the focal dimension, then 1x1x1) with arithmetic intensity 4-12 FLOP/B: as two
// implicit-GEMM launches they cost two full read+write passes over a full-resolution tensor.  Here a thread owns one pixel
// and walks its S slices with a 3-slice window in registers: F is read once, out written once, the intermediate never
// leaves registers (fp32).  Weights ([dz][ci][co] and [c][co] fp32, the FFMA pack) are shared-memory broadcasts.
// ------------------------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void load_px(const __nv_bfloat16* p, float* v) {
#pragma unroll
  for (int j = 0; j < C; j += 8) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p + j));
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[j + 2 * k] = __uint_as_float(u[k] << 16);
      v[j + 2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
    }
  }
}

template <int C>
__global__ void __launch_bounds__(128) srd_attention_kernel(const __nv_bfloat16* __restrict__ F, const float* __restrict__ w0,
                                                            const float* __restrict__ w1, __nv_bfloat16* __restrict__ out, int B,
                                                            int S, size_t plane /* H*W */) {
  __shared__ __align__(16) float sw0[3 * C * C], sw1[C * C];
  for (int i = threadIdx.x; i < 3 * C * C; i += 128) sw0[i] = __ldg(w0 + i);
  for (int i = threadIdx.x; i < C * C; i += 128) sw1[i] = __ldg(w1 + i);
  __syncthreads();
  const size_t pix = blockIdx.x * (size_t)128 + threadIdx.x;
  if (pix >= (size_t)B * plane) return;
  const size_t b = pix / plane, r = pix % plane;
  const __nv_bfloat16* src = F + (b * S * plane + r) * C;
  __nv_bfloat16* dst = out + (b * S * plane + r) * C;
  const size_t zs = plane * C;
  float prev[C], cur[C], nxt[C];
#pragma unroll
  for (int j = 0; j < C; ++j) prev[j] = 0.f;
  load_px<C>(src, cur);
  for (int z = 0; z < S; ++z) {
    if (z + 1 < S) load_px<C>(src + (size_t)(z + 1) * zs, nxt);
    else {
#pragma unroll
      for (int j = 0; j < C; ++j) nxt[j] = 0.f;
    }
    float a[C];
#pragma unroll
    for (int j = 0; j < C; ++j) a[j] = 0.f;
#pragma unroll
    for (int ci = 0; ci < C; ++ci) {
#pragma unroll
      for (int c4 = 0; c4 < C; c4 += 4) {
        const float4 wa = *reinterpret_cast<const float4*>(sw0 + (0 * C + ci) * C + c4);
        const float4 wb = *reinterpret_cast<const float4*>(sw0 + (1 * C + ci) * C + c4);
        const float4 wc = *reinterpret_cast<const float4*>(sw0 + (2 * C + ci) * C + c4);
        a[c4] = fmaf(prev[ci], wa.x, fmaf(cur[ci], wb.x, fmaf(nxt[ci], wc.x, a[c4])));
        a[c4 + 1] = fmaf(prev[ci], wa.y, fmaf(cur[ci], wb.y, fmaf(nxt[ci], wc.y, a[c4 + 1])));
        a[c4 + 2] = fmaf(prev[ci], wa.z, fmaf(cur[ci], wb.z, fmaf(nxt[ci], wc.z, a[c4 + 2])));
        a[c4 + 3] = fmaf(prev[ci], wa.w, fmaf(cur[ci], wb.w, fmaf(nxt[ci], wc.w, a[c4 + 3])));
      }
    }
    float o[C];
#pragma unroll
    for (int j = 0; j < C; ++j) o[j] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float av = fmaxf(a[c], 0.f);
#pragma unroll
      for (int c4 = 0; c4 < C; c4 += 4) {
        const float4 w = *reinterpret_cast<const float4*>(sw1 + c * C + c4);
        o[c4] = fmaf(av, w.x, o[c4]); o[c4 + 1] = fmaf(av, w.y, o[c4 + 1]);
        o[c4 + 2] = fmaf(av, w.z, o[c4 + 2]); o[c4 + 3] = fmaf(av, w.w, o[c4 + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < C; j += 8) {
      uint4 w;
      uint32_t* u = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(o[j + 2 * k], 0.f) + cur[j + 2 * k], fmaxf(o[j + 2 * k + 1], 0.f) + cur[j + 2 * k + 1]);
        u[k] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(dst + (size_t)z * zs + j) = w;
    }
#pragma unroll
    for (int j = 0; j < C; ++j) { prev[j] = cur[j]; cur[j] = nxt[j]; }
  }
}

// F, out: (B,S,H,W,C) bf16 ; w0: [3][C][C] fp32 ([dz][ci][co]) ; w1: [C][C] fp32 ([c][co]).  C = 8 or 16.
int launch_srd_attention(const void* F, const float* w0, const float* w1, void* out, int B, int S, int H, int W, int C, cudaStream_t st) {
  const size_t plane = (size_t)H * W, npix = (size_t)B * plane;
  const unsigned grid = (unsigned)((npix + 127) / 128);
  if (C == 8) srd_attention_kernel<8><<<grid, 128, 0, st>>>((const __nv_bfloat16*)F, w0, w1, (__nv_bfloat16*)out, B, S, plane);
  else if (C == 16) srd_attention_kernel<16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)F, w0, w1, (__nv_bfloat16*)out, B, S, plane);
  else return fail(-5, "srd_attention: C must be 8 or 16");
  DFF_LAUNCH_CHECK("srd_attention");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// End-to-End alignment network (reference End_to_End/End_to_End.py:63-134) on channels-last feature volumes
// ------------------------------------------------------------------------------------------------------------
// Sampling geometry of FOV_warp for output pixel (px, py) of slice s: flow (in pixels) and the bilinear taps.
struct WarpGeom {
  float flx, fly, w00, w01, w10, w11;
  int x0, y0;
  bool vx0, vx1, vy0, vy1;
};
__device__ __forceinline__ WarpGeom warp_geom(const float* alpha, const float* fov, int b, int s, int S, int H, int W, int px, int py) {
  WarpGeom g;
  const float a0 = alpha ? __ldg(alpha + (0 * 3 + 0) * S + s) : 0.f;  // sample 0 on purpose (reference broadcast quirk)
  const float a1 = alpha ? __ldg(alpha + ((size_t)b * 3 + 1) * S + s) : 0.f;
  const float a2 = alpha ? __ldg(alpha + ((size_t)b * 3 + 2) * S + s) : 0.f;
  const float f = a0 + __ldg(fov + (size_t)b * S + s);
  auto lin = [](int i, int n) -> float {
    if (n == 1) return -1.f;
    const float step = 2.f / (float)(n - 1);
    return i < n / 2 ? -1.f + step * (float)i : 1.f - step * (float)(n - 1 - i);
  };
  g.flx = (float)(W / 2) * (f - 1.f) * lin(px, W) + a1;
  g.fly = (float)(H / 2) * (f - 1.f) * lin(py, H) + a2;
  const float gx = 2.0f * ((float)px - g.flx) / (float)max(W - 1, 1) - 1.0f;
  const float gy = 2.0f * ((float)py - g.fly) / (float)max(H - 1, 1) - 1.0f;
  const float ix = (gx + 1.f) * 0.5f * (float)(W - 1);
  const float iy = (gy + 1.f) * 0.5f * (float)(H - 1);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  g.x0 = (int)fx0; g.y0 = (int)fy0;
  const float tx = ix - fx0, ty = iy - fy0;
  g.w00 = (1.f - tx) * (1.f - ty); g.w01 = tx * (1.f - ty); g.w10 = (1.f - tx) * ty; g.w11 = tx * ty;
  g.vx0 = g.x0 >= 0 && g.x0 < W; g.vx1 = g.x0 + 1 >= 0 && g.x0 + 1 < W;
  g.vy0 = g.y0 >= 0 && g.y0 < H; g.vy1 = g.y0 + 1 >= 0 && g.y0 + 1 < H;
  return g;
}

// FOV_warp of a channels-last volume x (B,S,H,W,C) -> out (same layout); C % 4 == 0
template <typename T>
__global__ void fov_warp_cl_kernel(const T* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ fov, int B, int C,
                                   int S, int H, int W, T* __restrict__ out) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y;
  const int bs = blockIdx.z, b = bs / S, s = bs % S;
  if (px >= W) return;
  const WarpGeom g = warp_geom(alpha, fov, b, s, S, H, W, px, py);
  const T* base = x + (size_t)bs * H * W * C;
  T* o = out + (((size_t)bs * H + py) * W + px) * C;
  for (int c = 0; c < C; c += 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    auto tap = [&](int yy, int xx, float w) {
      const float4 t = Elem<T>::load4(base + ((size_t)yy * W + xx) * C + c);
      v.x = fmaf(t.x, w, v.x); v.y = fmaf(t.y, w, v.y); v.z = fmaf(t.z, w, v.z); v.w = fmaf(t.w, w, v.w);
    };
    if (g.vy0 && g.vx0) tap(g.y0, g.x0, g.w00);
    if (g.vy0 && g.vx1) tap(g.y0, g.x0 + 1, g.w01);
    if (g.vy1 && g.vx0) tap(g.y0 + 1, g.x0, g.w10);
    if (g.vy1 && g.vx1) tap(g.y0 + 1, g.x0 + 1, g.w11);
    Elem<T>::store4(o + c, v);
  }
}

// The alignment head's input (reference :71-76, 81-86, 92-97):  out (B,S,H,W,Cs) = [ feat[b, S-1] | feat[b, s] | flow_x, flow_y, 0 ... ]
// with Cs >= 2C+8 stored channels (Cs = 2C+16 keeps the 8-channel chunk count even for the tensor-core kernels)
template <typename T>
__global__ void pair_volume_kernel(const T* __restrict__ feat, const float* __restrict__ alpha, const float* __restrict__ fov, int B,
                                   int C, int S, int H, int W, T* __restrict__ out, int Cs) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y;
  const int bs = blockIdx.z, b = bs / S, s = bs % S;
  if (px >= W) return;
  const WarpGeom g = warp_geom(alpha, fov, b, s, S, H, W, px, py);
  const size_t pin = ((size_t)py * W + px) * C;
  const T* cur = feat + (size_t)bs * H * W * C + pin;
  const T* last = feat + ((size_t)b * S + (S - 1)) * H * W * C + pin;
  T* o = out + (((size_t)bs * H + py) * W + px) * Cs;
  for (int c = 0; c < C; c += 4) {
    Elem<T>::store4(o + c, Elem<T>::load4(last + c));
    Elem<T>::store4(o + C + c, Elem<T>::load4(cur + c));
  }
  Elem<T>::store4(o + 2 * C, make_float4(g.flx, g.fly, 0.f, 0.f));
  for (int c = 2 * C + 4; c < Cs; c += 4) Elem<T>::store4(o + c, make_float4(0.f, 0.f, 0.f, 0.f));
}

// bf16, C % 8 == 0: one 16-byte piece of the output per thread — consecutive lanes write consecutive pieces of a pixel (coalesced
// 16-byte accesses instead of 8-byte stores 2*Cs bytes apart); only the flow piece evaluates the warp geometry.  Same values.
__global__ void __launch_bounds__(256) pair_volume_cl8_kernel(const uint4* __restrict__ feat, const float* __restrict__ alpha,
                                                              const float* __restrict__ fov, int B, int C8, int S, int H, int W,
                                                              uint4* __restrict__ out, int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // (pixel of the row, piece)
  if (i >= W * P) return;
  const int px = i / P, j = i - px * P, py = blockIdx.y;
  const int bs = blockIdx.z, b = bs / S, s = bs % S;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (j < 2 * C8) {
    const int ss = j < C8 ? S - 1 : s, c8 = j < C8 ? j : j - C8;
    o = __ldg(feat + ((((size_t)b * S + ss) * H + py) * W + px) * C8 + c8);
  } else if (j == 2 * C8) {
    const WarpGeom g = warp_geom(alpha, fov, b, s, S, H, W, px, py);
    const __nv_bfloat162 h = __floats2bfloat162_rn(g.flx, g.fly);
    o.x = *reinterpret_cast<const uint32_t*>(&h);
  }
  out[(((size_t)bs * H + py) * W + px) * P + j] = o;
}

// alpha_out[b][c][s] = (alpha_in ? alpha_in[b][c][s] : 0) + scale[c] * mean over (y, x) of x[b,s,y,x,c]     (c < 3; x fp32, Cs stored)
// = AdaptiveAvgPool3d((S,1,1)) of the head's last conv + the 0.001 factor on the scale term + the running sum (reference :78-79, 88-90)
__global__ void spatial_mean_accum_kernel(const float* __restrict__ x, int Cs, int S, int H, int W, const float* __restrict__ alpha_in,
                                          float s0, float s1, float s2, float* __restrict__ alpha_out) {
  const int bs = blockIdx.x, b = bs / S, s = bs % S;
  const size_t n = (size_t)H * W;
  const float* p = x + (size_t)bs * n * Cs;
  double a0 = 0, a1 = 0, a2 = 0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    a0 += p[i * Cs]; a1 += p[i * Cs + 1]; a2 += p[i * Cs + 2];
  }
  __shared__ double sh[3][256];
  sh[0][threadIdx.x] = a0; sh[1][threadIdx.x] = a1; sh[2][threadIdx.x] = a2;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k)
      for (int c = 0; c < 3; ++c) sh[c][threadIdx.x] += sh[c][threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const size_t o = ((size_t)b * 3 + c) * S + s;
    alpha_out[o] = (alpha_in ? alpha_in[o] : 0.f) + sc * (float)(sh[c][0] / (double)n);
  }
}

// Two-stage and deterministic: a fixed split of every slice (fixed partition, fixed tree), then one small block per slice.
// (One block per slice — 10 blocks for a 10-slice stack — took 767 us at 512x768: 38 % of the alignment network's time at the C4 shape.)
__global__ void __launch_bounds__(256) spatial_mean_partial_kernel(const float* __restrict__ x, int Cs, size_t n, double* __restrict__ partial) {
  const int bs = blockIdx.y, nb = gridDim.x;
  const float* p = x + (size_t)bs * n * Cs;
  const size_t per = (n + nb - 1) / nb, lo = (size_t)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  float f0 = 0, f1 = 0, f2 = 0;      // (<= a few hundred addends per thread; the cross-thread tree is double)
  if (Cs == 4) {
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
      f0 += v.x; f1 += v.y; f2 += v.z;
    }
  } else {
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      f0 += p[i * Cs]; f1 += p[i * Cs + 1]; f2 += p[i * Cs + 2];
    }
  }
  __shared__ double sh[3][256];
  sh[0][threadIdx.x] = f0; sh[1][threadIdx.x] = f1; sh[2][threadIdx.x] = f2;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k)
      for (int c = 0; c < 3; ++c) sh[c][threadIdx.x] += sh[c][threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x < 3) partial[((size_t)bs * nb + blockIdx.x) * 3 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void __launch_bounds__(96) spatial_mean_finish_kernel(const double* __restrict__ partial, int nb, int S, double n,
                                                                const float* __restrict__ alpha_in, float s0, float s1, float s2,
                                                                float* __restrict__ alpha_out) {
  const int bs = blockIdx.x, b = bs / S, s = bs % S;
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;   // one warp per channel
  double a = 0;
  for (int i = lane; i < nb; i += 32) a += partial[((size_t)bs * nb + i) * 3 + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if (lane == 0) {
    const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const size_t o = ((size_t)b * 3 + c) * S + s;
    alpha_out[o] = (alpha_in ? alpha_in[o] : 0.f) + sc * (float)(a / n);
  }
}

// bf16 channels-last volumes, one 16-byte piece (8 channels) per thread: consecutive lanes take consecutive pieces of a pixel, so a
// warp's tap load is one contiguous run per pixel; packed bf16x2 blends in fp32.
__global__ void __launch_bounds__(256) fov_warp_cl8_kernel(const uint4* __restrict__ x, const float* __restrict__ alpha,
                                                           const float* __restrict__ fov, int B, int C8, int S, int H, int W,
                                                           uint4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // (pixel of the row, chunk)
  if (i >= W * C8) return;
  const int px = i / C8, c8 = i - px * C8, py = blockIdx.y;
  const int bs = blockIdx.z, b = bs / S, s = bs % S;
  const WarpGeom g = warp_geom(alpha, fov, b, s, S, H, W, px, py);
  const uint4* base = x + (size_t)bs * H * W * C8 + c8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  auto tap = [&](int yy, int xx, float w) {
    const uint4 r = __ldg(base + ((size_t)yy * W + xx) * C8);
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] = fmaf(__uint_as_float(u[j] << 16), w, acc[2 * j]);
      acc[2 * j + 1] = fmaf(__uint_as_float(u[j] & 0xffff0000u), w, acc[2 * j + 1]);
    }
  };
  if (g.vy0 && g.vx0) tap(g.y0, g.x0, g.w00);
  if (g.vy0 && g.vx1) tap(g.y0, g.x0 + 1, g.w01);
  if (g.vy1 && g.vx0) tap(g.y0 + 1, g.x0, g.w10);
  if (g.vy1 && g.vx1) tap(g.y0 + 1, g.x0 + 1, g.w11);
  uint4 o;
  uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    ou[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
  out[(((size_t)bs * H + py) * W + px) * C8 + c8] = o;
}

int launch_fov_warp_cl(const void* x, const float* alpha, const float* fov, int B, int C, int S, int H, int W, void* out, bool bf16,
                       cudaStream_t st) {
  if (C % 4) return fail(-1, "fov_warp_cl: C must be a multiple of 4");
  if (bf16 && C % 8 == 0) {
    dim3 grid(cdiv(W * (C / 8), 256), H, B * S);
    fov_warp_cl8_kernel<<<grid, 256, 0, st>>>((const uint4*)x, alpha, fov, B, C / 8, S, H, W, (uint4*)out);
    DFF_LAUNCH_CHECK("fov_warp_cl8");
    return 0;
  }
  dim3 grid(cdiv(W, 128), H, B * S);
  if (bf16) fov_warp_cl_kernel<<<grid, 128, 0, st>>>((const __nv_bfloat16*)x, alpha, fov, B, C, S, H, W, (__nv_bfloat16*)out);
  else fov_warp_cl_kernel<<<grid, 128, 0, st>>>((const float*)x, alpha, fov, B, C, S, H, W, (float*)out);
  DFF_LAUNCH_CHECK("fov_warp_cl");
  return 0;
}
int launch_pair_volume(const void* feat, const float* alpha, const float* fov, int B, int C, int S, int H, int W, void* out, bool bf16,
                       cudaStream_t st, int Cs) {
  if (C % 4) return fail(-1, "pair_volume: C must be a multiple of 4");
  if (Cs == 0) Cs = 2 * C + 8;
  if (Cs % 4 || Cs < 2 * C + 4) return fail(-1, "pair_volume: stored channels must be a multiple of 4 and >= 2C+4");
  if (bf16 && C % 8 == 0 && Cs % 8 == 0 && Cs >= 2 * C + 8) {
    dim3 g8(cdiv(W * (Cs / 8), 256), H, B * S);
    pair_volume_cl8_kernel<<<g8, 256, 0, st>>>((const uint4*)feat, alpha, fov, B, C / 8, S, H, W, (uint4*)out, Cs / 8);
    DFF_LAUNCH_CHECK("pair_volume_cl8");
    return 0;
  }
  dim3 grid(cdiv(W, 128), H, B * S);
  if (bf16) pair_volume_kernel<<<grid, 128, 0, st>>>((const __nv_bfloat16*)feat, alpha, fov, B, C, S, H, W, (__nv_bfloat16*)out, Cs);
  else pair_volume_kernel<<<grid, 128, 0, st>>>((const float*)feat, alpha, fov, B, C, S, H, W, (float*)out, Cs);
  DFF_LAUNCH_CHECK("pair_volume");
  return 0;
}
// blocks per slice of the first stage (a function of the shape only: the result does not depend on the device)
int spatial_mean_blocks(int B, int S, int H, int W) {
  const size_t n = (size_t)H * W;
  int nb = (int)(n / 4096);                                  // >= 16 pixels per thread
  const int cap = (2368 + B * S - 1) / (B * S);              // ~16 blocks per SM in total
  if (nb > cap) nb = cap;
  if (nb > 256) nb = 256;
  return nb < 1 ? 1 : nb;
}
size_t spatial_mean_scratch_bytes(int B, int S, int H, int W) { return (size_t)B * S * spatial_mean_blocks(B, S, H, W) * 3 * sizeof(double); }

int launch_spatial_mean_accum(const float* x, int Cs, int B, int S, int H, int W, const float* alpha_in, float s0, float s1, float s2,
                              float* alpha_out, void* scratch, cudaStream_t st) {
  if (Cs < 3) return fail(-1, "spatial_mean_accum: needs at least 3 stored channels");
  if (B * S > 65535) return fail(-1, "spatial_mean_accum: more than 65535 slices");
  if (!scratch) {   // (single-operator entry point without scratch: one block per slice)
    spatial_mean_accum_kernel<<<B * S, 256, 0, st>>>(x, Cs, S, H, W, alpha_in, s0, s1, s2, alpha_out);
    DFF_LAUNCH_CHECK("spatial_mean_accum");
    return 0;
  }
  const int nb = spatial_mean_blocks(B, S, H, W);
  spatial_mean_partial_kernel<<<dim3(nb, B * S), 256, 0, st>>>(x, Cs, (size_t)H * W, (double*)scratch);
  DFF_LAUNCH_CHECK("spatial_mean_partial");
  spatial_mean_finish_kernel<<<B * S, 96, 0, st>>>((const double*)scratch, nb, S, (double)H * W, alpha_in, s0, s1, s2, alpha_out);
  DFF_LAUNCH_CHECK("spatial_mean_finish");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// weight packing: reference layouts -> [tap][CinP][CoutP] fp32 (zero padded), BatchNorm(eval) -> scale/shift
// ------------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin, int ntaps,
                                   int CinP, int CoutP, int transposed) {
  const int n = ntaps * CinP * CoutP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int co = i % CoutP, ci = (i / CoutP) % CinP, t = i / (CoutP * CinP);
    float v = 0.f;
    if (co < Cout && ci < Cin)
      v = transposed ? w[((size_t)ci * Cout + co) * ntaps + t] : w[((size_t)co * Cin + ci) * ntaps + t];
    dst[i] = v;
  }
}

// scale = gamma / sqrt(var + eps), shift = beta - mean * scale   (nn.BatchNorm3d eval, eps 1e-5, reference :355)
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, const float* __restrict__ bias, float* __restrict__ scale,
                               float* __restrict__ shift, int C, int CP) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= CP) return;
  float sc = 1.f, sh = 0.f;
  if (c < C) {
    if (gamma) {
      const double inv = 1.0 / sqrt((double)var[c] + 1e-5);
      const double s = (double)gamma[c] * inv;
      sc = (float)s;
      sh = (float)((double)beta[c] - (double)mean[c] * s);
    } else if (bias) {
      sh = bias[c];
    }
  }
  scale[c] = sc;
  shift[c] = sh;
}

// scale/shift of the x-folded form: dst[g*C + c] = scale[c], dst[G*C + g*C + c] = shift[c]
__global__ void replicate_ss_kernel(const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ dst, int C, int G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * C) return;
  dst[i] = scale[i % C];
  dst[G * C + i] = shift[i % C];
}
int launch_replicate_ss(const float* scale, const float* shift, float* dst, int C, int G, cudaStream_t st) {
  replicate_ss_kernel<<<cdiv(G * C, 128), 128, 0, st>>>(scale, shift, dst, C, G);
  DFF_LAUNCH_CHECK("replicate_ss");
  return 0;
}

int launch_pack_weight(const float* w, float* dst, int Cout, int Cin, int ntaps, int CinP, int CoutP, int transposed,
                       cudaStream_t st) {
  const int n = ntaps * CinP * CoutP;
  pack_weight_kernel<<<grid_for(n, 256, 256), 256, 0, st>>>(w, dst, Cout, Cin, ntaps, CinP, CoutP, transposed);
  DFF_LAUNCH_CHECK("pack_weight");
  return 0;
}
int launch_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* bias,
                   float* scale, float* shift, int C, int CP, cudaStream_t st) {
  bn_fold_kernel<<<cdiv(CP, 128), 128, 0, st>>>(gamma, beta, mean, var, bias, scale, shift, C, CP);
  DFF_LAUNCH_CHECK("bn_fold");
  return 0;
}

}  // namespace dff
