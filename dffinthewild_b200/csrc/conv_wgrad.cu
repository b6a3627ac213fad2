// Weight gradient of the generic 3-D convolution (train path; autograd of `nn.Conv3d` / `nn.ConvTranspose3d`, invoked by
// `Total.backward()` at reference train_codes/train_code_Defocus.py:167 — ATen's convolution_backward, 74 % of the reference's
// CPU training time).
//
//   dw[widx_t][ci][co] = sum over output positions o of  in[o*is + d_t][ci] * dy[o*os + oo][co]
//
// in the same "strided gather over a tap table" form as the forward kernels (conv_ffma.cu), so ordinary, strided and
// (per output-parity phase) transposed convolutions and the two-source concat all go through one kernel.  One CTA owns a
// 32 x TY tile of output positions of one (batch, slice): it stages the input region (halo included) channel-planar for CK
// input channels and the dy tile for all output channels in shared memory; a thread owns one (tap, ci) and a block of COB output
// channels, sums over all positions of the tile in registers (one activation load feeds COB FMAs, the dy row is a shared-memory
// broadcast), and adds its partial sums to the fp32 gradient in the reference's weight layout with red.global.add.f32.
// fp32 accumulate; the order of the global adds is not fixed, which moves results by ~1e-7 relative (gate: gradient cosine).
#include "common.cuh"

namespace dff {

constexpr int kWgThreads = 256;
constexpr int kWgTY = 8;   // tile rows (tile cols = 32)

struct WgradArgs {
  ConvArgs a;        // geometry + tap table + in0/in1 (a.out / a.w unused)
  const void* dy;    // (B,S,OH,OW,CoS) channels-last, CoS stored channels
  int CoS, Cout, Cin;
  float* dw;         // reference layout: conv (Cout,Cin,ntaps_total) ; transposed (Cin,Cout,ntaps_total)
  int ntaps_total, wt_transposed;
  int ci_base;       // channel offset of in0 inside the layer's Cin (two-source layers: in1 starts at ci_base + C0)
};

template <typename T, int CK, int COB>
__global__ void __launch_bounds__(kWgThreads, 2) conv_wgrad_kernel(const __grid_constant__ WgradArgs g) {
  extern __shared__ __align__(16) float smem[];
  const ConvArgs& a = g.a;
  const int REG = a.RZ * a.RY * a.RXP;
  const int NP = 32 * a.TY;              // positions per tile
  const int CoP = (g.Cout + 3) & ~3;
  float* in_s = smem;                     // [CK][REG]
  float* dy_s = smem + CK * REG;          // [NP][CoP]
  const int tid = threadIdx.x;
  const int bs = blockIdx.z, b = bs / a.S, s = bs % a.S;
  const int ty0 = blockIdx.y * a.TY, tx0 = blockIdx.x * 32;
  const int gy0 = ty0 * a.isy + a.dymin, gx0 = tx0 * a.isx + a.dxmin, gz0 = s + a.dzmin;

  // ---- dy tile (zero outside the phase grid) -----------------------------------------------------------------------
  const T* dy = reinterpret_cast<const T*>(g.dy);
  for (int i = tid; i < NP * (CoP / 4); i += kWgThreads) {
    const int q = i % (CoP / 4), pos = i / (CoP / 4);
    const int ox = tx0 + (pos & 31), oy = ty0 + (pos >> 5);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ox < a.OWt && oy < a.OHt) {
      const size_t pix = (((size_t)b * a.S + s) * a.OH + (oy * a.osy + a.ooy)) * a.OW + (ox * a.osx + a.oox);
      const T* p = dy + pix * g.CoS + 4 * q;
      if (4 * q + 3 < g.CoS) v = Elem<T>::load4(p);
      else {
        v.x = Elem<T>::load(p);
        if (4 * q + 1 < g.CoS) v.y = Elem<T>::load(p + 1);
        if (4 * q + 2 < g.CoS) v.z = Elem<T>::load(p + 2);
      }
    }
    *reinterpret_cast<float4*>(dy_s + pos * CoP + 4 * q) = v;
  }

  const int Ctot = a.C0 + a.C1;
  const int npos = a.RZ * a.RY * a.RX;
  constexpr int NQ = CK / 4;
  const int cosplit = CoP / COB;                       // output-channel blocks per (tap, ci)
  const int nitems = a.taps.n * CK * cosplit;          // work items of one channel chunk: (tap, ci, co block)

  for (int c0 = 0; c0 < Ctot; c0 += CK) {
    const bool second = c0 >= a.C0;
    const T* src = reinterpret_cast<const T*>(second ? a.in1 : a.in0);
    const int Csrc = second ? a.C1 : a.C0;
    const int cb = second ? c0 - a.C0 : c0;
    __syncthreads();
    for (int i = tid; i < npos * NQ; i += kWgThreads) {
      const int q = i % NQ, pos = i / NQ;
      const int x = pos % a.RX, y = (pos / a.RX) % a.RY, z = pos / (a.RX * a.RY);
      const int gz = gz0 + z, gy = gy0 + y, gx = gx0 + x;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gz >= 0 && gz < a.S && gy >= 0 && gy < a.IH && gx >= 0 && gx < a.IW)
        v = Elem<T>::load4(src + ((((size_t)b * a.S + gz) * a.IH + gy) * a.IW + gx) * Csrc + cb + 4 * q);
      float* d = in_s + (4 * q) * REG + (z * a.RY + y) * a.RXP + x;
      d[0] = v.x; d[REG] = v.y; d[2 * REG] = v.z; d[3 * REG] = v.w;
    }
    __syncthreads();
    // a thread owns (tap, ci) x COB output channels and sums over all positions of the tile: one activation load feeds COB FMAs,
    // the dy row is a shared-memory broadcast for the warp
    for (int it = tid; it < nitems; it += kWgThreads) {
      const int cs = it % cosplit, ci = (it / cosplit) % CK, t = it / (cosplit * CK);
      const int cig = g.ci_base + c0 + ci;   // channel inside the layer's Cin (stored channels beyond Cin are padding)
      if (cig >= g.Cin) continue;
      const float* ip = in_s + ci * REG + ((a.taps.dz[t] - a.dzmin) * a.RY + (a.taps.dy[t] - a.dymin)) * a.RXP +
                        (a.taps.dx[t] - a.dxmin);
      const float* dp = dy_s + cs * COB;
      float acc[COB];
#pragma unroll
      for (int j = 0; j < COB; ++j) acc[j] = 0.f;
      for (int py = 0; py < a.TY; ++py) {
        const float* irow = ip + py * a.isy * a.RXP;
        const float* drow = dp + (py * 32) * CoP;
#pragma unroll 4
        for (int px = 0; px < 32; ++px) {
          const float av = irow[px * a.isx];
#pragma unroll
          for (int j4 = 0; j4 < COB / 4; ++j4) {
            const float4 d4 = *reinterpret_cast<const float4*>(drow + px * CoP + 4 * j4);
            acc[4 * j4] = fmaf(av, d4.x, acc[4 * j4]);
            acc[4 * j4 + 1] = fmaf(av, d4.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(av, d4.z, acc[4 * j4 + 2]);
            acc[4 * j4 + 3] = fmaf(av, d4.w, acc[4 * j4 + 3]);
          }
        }
      }
      const int wi = a.taps.widx[t];
#pragma unroll
      for (int j = 0; j < COB; ++j) {
        const int co = cs * COB + j;
        if (co >= g.Cout) continue;
        const size_t o = g.wt_transposed ? ((size_t)cig * g.Cout + co) * g.ntaps_total + wi
                                         : ((size_t)co * g.Cin + cig) * g.ntaps_total + wi;
        atomicAdd(g.dw + o, acc[j]);
      }
    }
  }
}

static size_t wg_plan(ConvArgs& a, int CK, int Cout) {
  int dzmax = -100, dymax = -100, dxmax = -100;
  a.dzmin = a.dymin = a.dxmin = 100;
  for (int t = 0; t < a.taps.n; ++t) {
    a.dzmin = a.taps.dz[t] < a.dzmin ? a.taps.dz[t] : a.dzmin;
    a.dymin = a.taps.dy[t] < a.dymin ? a.taps.dy[t] : a.dymin;
    a.dxmin = a.taps.dx[t] < a.dxmin ? a.taps.dx[t] : a.dxmin;
    dzmax = a.taps.dz[t] > dzmax ? a.taps.dz[t] : dzmax;
    dymax = a.taps.dy[t] > dymax ? a.taps.dy[t] : dymax;
    dxmax = a.taps.dx[t] > dxmax ? a.taps.dx[t] : dxmax;
  }
  const int CoP = (Cout + 3) & ~3;
  a.TY = CoP <= 32 ? kWgTY : (CoP <= 64 ? kWgTY / 2 : kWgTY / 4);   // keep the dy tile <= 32 KB
  a.RZ = dzmax - a.dzmin + 1;
  a.RY = (a.TY - 1) * a.isy + (dymax - a.dymin) + 1;
  a.RX = 31 * a.isx + (dxmax - a.dxmin) + 1;
  a.RXP = a.RX | 1;
  return ((size_t)CK * a.RZ * a.RY * a.RXP + (size_t)32 * a.TY * CoP) * sizeof(float);
}

// `a`: geometry/taps/in0/in1 as for the forward launch of the same (phase of the) layer.
int launch_conv_wgrad(ConvArgs a, const void* dy, int CoS, int Cout, int Cin, int ci_base, float* dw, int ntaps_total,
                      int wt_transposed, bool bf16, cudaStream_t st) {
  if (a.C0 % 4 || a.C1 % 4) return fail(-1, "conv_wgrad: stored input channels must be multiples of 4");
  WgradArgs g{};
  int CK = (a.C0 % 8 == 0 && a.C1 % 8 == 0) ? 8 : 4;
  size_t smem = wg_plan(a, CK, Cout);
  if (smem > 100 * 1024 && CK == 8) { CK = 4; smem = wg_plan(a, CK, Cout); }
  if (smem > 227 * 1024) return fail(-5, "conv_wgrad: tile does not fit shared memory");
  g.a = a; g.dy = dy; g.CoS = CoS; g.Cout = Cout; g.Cin = Cin; g.dw = dw; g.ntaps_total = ntaps_total;
  g.wt_transposed = wt_transposed; g.ci_base = ci_base;
  dim3 grid(cdiv(a.OWt, 32), cdiv(a.OHt, a.TY), a.B * a.S);
  const int CoP = (Cout + 3) & ~3;
  const int COB = CoP >= 32 ? 32 : (CoP >= 16 ? 16 : (CoP >= 8 ? 8 : 4));
  if (CoP % COB) return fail(-5, "conv_wgrad: unsupported output channel count");
#define DFF_WG(T, CK_, COB_)                                                                                        \
  do {                                                                                                              \
    auto k = conv_wgrad_kernel<T, CK_, COB_>;                                                                       \
    DFF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
    k<<<grid, kWgThreads, smem, st>>>(g);                                                                           \
  } while (0)
#define DFF_WG_T(T)                                                               \
  do {                                                                            \
    if (CK == 8) {                                                                \
      if (COB == 32) DFF_WG(T, 8, 32); else if (COB == 16) DFF_WG(T, 8, 16);      \
      else if (COB == 8) DFF_WG(T, 8, 8); else DFF_WG(T, 8, 4);                   \
    } else {                                                                      \
      if (COB == 32) DFF_WG(T, 4, 32); else if (COB == 16) DFF_WG(T, 4, 16);      \
      else if (COB == 8) DFF_WG(T, 4, 8); else DFF_WG(T, 4, 4);                   \
    }                                                                             \
  } while (0)
  if (bf16) DFF_WG_T(__nv_bfloat16);
  else DFF_WG_T(float);
#undef DFF_WG_T
#undef DFF_WG
  DFF_LAUNCH_CHECK("conv_wgrad");
  return 0;
}

}  // namespace dff
