// Generic 3-D convolution on CUDA cores (FFMA, fp32 accumulate) — the fp32 PARITY path and the fallback for
// layers the tcgen05 path does not cover.  Replaces the cuDNN conv3d / conv_transpose3d calls behind
// `convbn_3d`, `nn.Conv3d` and `nn.ConvTranspose3d` (reference train_codes/Depth_Estimation_Network.py:352-355,
// 43-50) with BatchNorm(eval)/bias, residual adds and ReLU fused into the epilogue.
//
// One CTA computes a 32 x TY tile of output positions of one (batch, slice) for ALL output channels:
//   * the input region the tile's taps touch (halo included, zero-filled outside the tensor = conv padding) is
//     staged in shared memory channel-planar for CK input channels at a time, so every global byte of the
//     activation is read once per CTA instead of once per tap;
//   * the [tap][CK][Cout] slice of the packed weights is staged next to it;
//   * each thread owns P=4 output rows x Q (8|16) output channels in registers; a warp shares one channel group,
//     so weight reads are shared-memory broadcasts and activation reads are conflict-free 128 B rows.
// Transposed convolutions run as 4 output-parity phases (3/6/6/12 taps) through the same kernel with an output
// stride of 2; strided convolutions use an input stride of 2; the virtual channel concat of the hourglass
// `conv0` layers (torch.cat, reference :103,109,114,261,266) is a second source pointer.
#include "common.cuh"

namespace dff {

constexpr int kThreads = 256;
constexpr int kP = 4;

template <typename T, int Q, int CK>
__global__ void __launch_bounds__(kThreads, 2) conv_ffma_kernel(const __grid_constant__ ConvArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int REG = a.RZ * a.RY * a.RXP;        // floats per staged channel plane
  float* in_s = smem;                          // [CK][REG]
  float* w_s = smem + CK * REG;                // [ntaps][CK][CoutP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NG = a.CoutP / Q;                  // channel groups
  const int WR = (kThreads / 32) / NG;         // warps per channel group
  const int cg = warp / WR, wr = warp % WR;
  const int bs = blockIdx.z, b = bs / a.S, s = bs % a.S;
  const int ty0 = blockIdx.y * a.TY, tx0 = blockIdx.x * 32;
  const int gy0 = ty0 * a.isy + a.dymin, gx0 = tx0 * a.isx + a.dxmin, gz0 = s + a.dzmin;

  float acc[kP][Q];
#pragma unroll
  for (int p = 0; p < kP; ++p)
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[p][q] = 0.f;

  const int Ctot = a.C0 + a.C1;
  const int npos = a.RZ * a.RY * a.RX;
  constexpr int NQ = CK / 4;
  const int wrow = CK * a.CoutP;  // floats per tap in w_s
  const int ntaps = a.taps.n;

  for (int c0 = 0; c0 < Ctot; c0 += CK) {
    const bool second = c0 >= a.C0;
    const T* src = reinterpret_cast<const T*>(second ? a.in1 : a.in0);
    const int Csrc = second ? a.C1 : a.C0;
    const int cb = second ? c0 - a.C0 : c0;
    __syncthreads();
    // ---- stage the input region (zero fill = padding) -------------------------------------------------------
    for (int i = tid; i < npos * NQ; i += kThreads) {
      const int q = i % NQ, pos = i / NQ;
      const int x = pos % a.RX, y = (pos / a.RX) % a.RY, z = pos / (a.RX * a.RY);
      const int gz = gz0 + z, gy = gy0 + y, gx = gx0 + x;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gz >= 0 && gz < a.S && gy >= 0 && gy < a.IH && gx >= 0 && gx < a.IW)
        v = Elem<T>::load4(src + ((((size_t)b * a.S + gz) * a.IH + gy) * a.IW + gx) * Csrc + cb + 4 * q);
      float* d = in_s + (4 * q) * REG + (z * a.RY + y) * a.RXP + x;
      d[0] = v.x;
      d[REG] = v.y;
      d[2 * REG] = v.z;
      d[3 * REG] = v.w;
    }
    // ---- stage the weight slice: rows (tap, c0..c0+CK) of CoutP floats ---------------------------------------
    {
      const int row4 = wrow / 4;  // float4 per tap
      for (int i = tid; i < ntaps * row4; i += kThreads) {
        const int t = i / row4, r = i % row4;
        const float4* g = reinterpret_cast<const float4*>(a.w + ((size_t)a.taps.widx[t] * a.CinP + c0) * a.CoutP);
        reinterpret_cast<float4*>(w_s + t * wrow)[r] = __ldg(g + r);
      }
    }
    __syncthreads();
    // ---- accumulate ---------------------------------------------------------------------------------------------
    const float* ibase = in_s + (wr * kP * a.isy) * a.RXP + lane * a.isx;
    const int pstride = a.isy * a.RXP;
    for (int t = 0; t < ntaps; ++t) {
      const float* ip = ibase + ((a.taps.dz[t] - a.dzmin) * a.RY + (a.taps.dy[t] - a.dymin)) * a.RXP +
                        (a.taps.dx[t] - a.dxmin);
      const float* wp = w_s + t * wrow + cg * Q;
#pragma unroll
      for (int ck = 0; ck < CK; ++ck) {
        float av[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) av[p] = ip[ck * REG + p * pstride];
        float wv[Q];
#pragma unroll
        for (int q4 = 0; q4 < Q / 4; ++q4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + ck * a.CoutP + 4 * q4);
          wv[4 * q4] = w4.x;
          wv[4 * q4 + 1] = w4.y;
          wv[4 * q4 + 2] = w4.z;
          wv[4 * q4 + 3] = w4.w;
        }
#pragma unroll
        for (int p = 0; p < kP; ++p)
#pragma unroll
          for (int q = 0; q < Q; ++q) acc[p][q] = fmaf(av[p], wv[q], acc[p][q]);
      }
    }
  }

  // ---- epilogue: BN scale/shift (or bias), residual(s), ReLU, store -----------------------------------------------
  const int ox = tx0 + lane;
  if (ox >= a.OWt) return;
  float sc[Q], sh[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    sc[q] = a.scale ? __ldg(a.scale + cg * Q + q) : 1.f;
    sh[q] = a.shift ? __ldg(a.shift + cg * Q + q) : 0.f;
  }
  T* out = reinterpret_cast<T*>(a.out);
  const T* rpre = reinterpret_cast<const T*>(a.res_pre);
  const T* rpost = reinterpret_cast<const T*>(a.res_post);
  T* oaux = reinterpret_cast<T*>(a.out_aux);
  const T* aadd = reinterpret_cast<const T*>(a.aux_add);
  const bool vec = (a.Cout == a.CoutP) && !a.out_f32;
#pragma unroll
  for (int p = 0; p < kP; ++p) {
    const int oy = ty0 + wr * kP + p;
    if (oy >= a.OHt) continue;
    const size_t pix = (((size_t)b * a.S + s) * a.OH + (oy * a.osy + a.ooy)) * a.OW + (ox * a.osx + a.oox);
    const size_t o = pix * a.Cout + cg * Q;
    if (vec) {
#pragma unroll
      for (int q4 = 0; q4 < Q / 4; ++q4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaf(acc[p][4 * q4 + j], sc[4 * q4 + j], sh[4 * q4 + j]);
        if (rpre) {
          const float4 r = Elem<T>::load4(rpre + o + 4 * q4);
          v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (rpost) {
          const float4 r = Elem<T>::load4(rpost + o + 4 * q4);
          v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        Elem<T>::store4(out + o + 4 * q4, make_float4(v[0], v[1], v[2], v[3]));
        if (oaux) {
          const float4 r = Elem<T>::load4(aadd + o + 4 * q4);
          Elem<T>::store4(oaux + o + 4 * q4, make_float4(v[0] + r.x, v[1] + r.y, v[2] + r.z, v[3] + r.w));
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        if (cg * Q + q >= a.Cout) continue;
        float v = fmaf(acc[p][q], sc[q], sh[q]);
        if (rpre) v += Elem<T>::load(rpre + o + q);
        if (a.relu) v = fmaxf(v, 0.f);
        if (rpost) v += Elem<T>::load(rpost + o + q);
        if (a.out_f32) reinterpret_cast<float*>(a.out)[o + q] = v;
        else Elem<T>::store(out + o + q, v);
        if (oaux) Elem<T>::store(oaux + o + q, v + Elem<T>::load(aadd + o + q));
      }
    }
  }
}

// smem bytes for a configuration; also fills the tile geometry of `a`.
static size_t plan_tile(ConvArgs& a, int Q, int CK) {
  const int NG = a.CoutP / Q;
  const int WR = (kThreads / 32) / NG;
  a.TY = WR * kP;
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
  a.RZ = dzmax - a.dzmin + 1;
  a.RY = (a.TY - 1) * a.isy + (dymax - a.dymin) + 1;
  a.RX = 31 * a.isx + (dxmax - a.dxmin) + 1;
  a.RXP = a.RX | 1;  // odd pitch: staging writes of neighbouring rows land in different banks
  return ((size_t)CK * a.RZ * a.RY * a.RXP + (size_t)a.taps.n * CK * a.CoutP) * sizeof(float);
}

template <typename T, int Q, int CK>
static int launch_one(ConvArgs& a, cudaStream_t st) {
  const size_t smem = plan_tile(a, Q, CK);
  if (smem > 227 * 1024) return fail(-5, "conv_ffma: tile does not fit shared memory");
  auto k = conv_ffma_kernel<T, Q, CK>;
  DFF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(a.OWt, 32), cdiv(a.OHt, a.TY), a.B * a.S);
  k<<<grid, kThreads, smem, st>>>(a);
  DFF_LAUNCH_CHECK("conv_ffma");
  return 0;
}

// Picks (Q, CK): Q = 8 for <= 16 output channels else 16; CK = 8 when it divides both sources and two CTAs still fit
// an SM, else 4.
int launch_conv_ffma(ConvArgs a, bool bf16, cudaStream_t st) {
  if (a.C0 % 4 || a.C1 % 4 || a.CoutP % 8 || a.CoutP > 128) return fail(-1, "conv_ffma: channel counts must be padded (4 in / 8 out)");
  const int Q = a.CoutP <= 16 ? 8 : 16;
  if (a.CoutP % Q || 8 % (a.CoutP / Q)) return fail(-1, "conv_ffma: unsupported output channel count");
  int CK = 4;
  if (a.C0 % 8 == 0 && a.C1 % 8 == 0) {
    ConvArgs tmp = a;
    if (plan_tile(tmp, Q, 8) <= 110 * 1024) CK = 8;
  }
#define DFF_GO(T, Q_, CK_) return launch_one<T, Q_, CK_>(a, st)
  if (!bf16) {
    if (Q == 8 && CK == 4) DFF_GO(float, 8, 4);
    if (Q == 8 && CK == 8) DFF_GO(float, 8, 8);
    if (Q == 16 && CK == 4) DFF_GO(float, 16, 4);
    DFF_GO(float, 16, 8);
  } else {
    if (Q == 8 && CK == 4) DFF_GO(__nv_bfloat16, 8, 4);
    if (Q == 8 && CK == 8) DFF_GO(__nv_bfloat16, 8, 8);
    if (Q == 16 && CK == 4) DFF_GO(__nv_bfloat16, 16, 4);
    DFF_GO(__nv_bfloat16, 16, 8);
  }
#undef DFF_GO
}

}  // namespace dff
