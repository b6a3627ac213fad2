// Weight gradient of the generic 3-D convolution (train path; autograd of `nn.Conv3d` / `nn.ConvTranspose3d`, invoked by
// `Total.backward()` at reference train_codes/train_code_Defocus.py:167 — ATen's convolution_backward, 74 % of the reference's
// CPU training time).
//
//   dw[widx_t][ci][co] = sum over output positions o of  in[o*is + d_t][ci] * dy[o*os + oo][co]
//
// in the same "strided gather over a tap table" form as the forward kernels (conv_ffma.cu), so ordinary, strided and
// (per output-parity phase) transposed convolutions and the two-source concat all go through one kernel.  One CTA owns a
// 32 x TY tile of output positions of one (batch, slice): it stages the input region (halo included) channel-planar for CK
// input channels and the dy tile for all output channels in shared memory, then every thread owns one (ci, co) pair (pixel
// range split across thread groups when there are fewer pairs than threads), sums over the tile's positions in registers tap
// by tap, and adds its partial sums to the fp32 gradient in the reference's weight layout with red.global.add.f32.
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

template <typename T, int CK>
__global__ void __launch_bounds__(kWgThreads, 2) conv_wgrad_kernel(const __grid_constant__ WgradArgs g) {
  extern __shared__ __align__(16) float smem[];
  const ConvArgs& a = g.a;
  const int REG = a.RZ * a.RY * a.RXP;
  const int NP = 32 * kWgTY;             // positions per tile
  const int CoP = (g.Cout + 3) & ~3;
  float* in_s = smem;                     // [CK][REG]
  float* dy_s = smem + CK * REG;          // [NP][CoP]
  const int tid = threadIdx.x;
  const int bs = blockIdx.z, b = bs / a.S, s = bs % a.S;
  const int ty0 = blockIdx.y * kWgTY, tx0 = blockIdx.x * 32;
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

  // thread -> (ci, co, pixel split)
  const int npairs = CK * g.Cout;
  const int PS = max(1, kWgThreads / npairs);         // thread groups splitting the tile's positions
  const int Ctot = a.C0 + a.C1;
  const int npos = a.RZ * a.RY * a.RX;
  constexpr int NQ = CK / 4;

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
    // every thread walks the pairs it owns: pair index pi = (tid % (npairs or 256)) + k*256 ; position split by tid / npairs
    for (int pi = tid % min(npairs, kWgThreads); pi < npairs; pi += kWgThreads) {
      const int ci = pi % CK, co = pi / CK;
      const int ps = npairs >= kWgThreads ? 0 : tid / npairs;
      if (ps >= PS) continue;
      const int cig = g.ci_base + c0 + ci;   // channel inside the layer's Cin (stored channels beyond Cin are padding)
      if (cig >= g.Cin) continue;
      const float* ip0 = in_s + ci * REG;
      for (int t = 0; t < a.taps.n; ++t) {
        const float* ip = ip0 + ((a.taps.dz[t] - a.dzmin) * a.RY + (a.taps.dy[t] - a.dymin)) * a.RXP + (a.taps.dx[t] - a.dxmin);
        float acc = 0.f;
        for (int pos = ps; pos < NP; pos += PS) {
          const int px = pos & 31, py = pos >> 5;
          acc = fmaf(ip[py * a.isy * a.RXP + px * a.isx], dy_s[pos * CoP + co], acc);
        }
        const int wi = a.taps.widx[t];
        const size_t o = g.wt_transposed ? ((size_t)cig * g.Cout + co) * g.ntaps_total + wi
                                         : ((size_t)co * g.Cin + cig) * g.ntaps_total + wi;
        atomicAdd(g.dw + o, acc);
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
  a.TY = kWgTY;
  a.RZ = dzmax - a.dzmin + 1;
  a.RY = (kWgTY - 1) * a.isy + (dymax - a.dymin) + 1;
  a.RX = 31 * a.isx + (dxmax - a.dxmin) + 1;
  a.RXP = a.RX | 1;
  const int CoP = (Cout + 3) & ~3;
  return ((size_t)CK * a.RZ * a.RY * a.RXP + (size_t)32 * kWgTY * CoP) * sizeof(float);
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
  dim3 grid(cdiv(a.OWt, 32), cdiv(a.OHt, kWgTY), a.B * a.S);
#define DFF_WG(T, CK_)                                                                                              \
  do {                                                                                                              \
    auto k = conv_wgrad_kernel<T, CK_>;                                                                             \
    DFF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
    k<<<grid, kWgThreads, smem, st>>>(g);                                                                           \
  } while (0)
  if (bf16) { if (CK == 8) DFF_WG(__nv_bfloat16, 8); else DFF_WG(__nv_bfloat16, 4); }
  else      { if (CK == 8) DFF_WG(float, 8); else DFF_WG(float, 4); }
#undef DFF_WG
  DFF_LAUNCH_CHECK("conv_wgrad");
  return 0;
}

}  // namespace dff
