// Train-mode building blocks around the convolutions: BatchNorm3d in batch-statistics mode
// (reference `convbn_3d`, train_codes/Depth_Estimation_Network.py:352-355, under model.train() at train_code_Defocus.py:145),
// its backward fused with the ReLU / residual backward, pooling backward (reference :183-187, 387), the depth-head
// backward (reference :92-98, 118-136) and the skip add.  Activations are channels-last (pixels, C) in fp32 or bf16; statistics,
// parameter gradients and all reductions are fp32/fp64.  Reductions are two-stage with a fixed grid, so they are deterministic.
#include "common.cuh"

namespace dff {

constexpr int kRedBlocks = 592;   // 4 CTAs per SM
constexpr int kRedThreads = 256;

// ------------------------------------------------------------------------------------------------------------------
// per-channel partial sums over pixels.  MODE 0: (sum x, sum x^2).  MODE 1: (sum g, sum g*xhat) with g = dy * mask.
// thread -> (channel quad, pixel lane); partial[block][c] as double2.
// ------------------------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(kRedThreads) bn_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                const T* __restrict__ y, const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, size_t npix, int C,
                                                                double2* __restrict__ partial) {
  const int C4 = C / 4;
  const int lanes = kRedThreads / C4;          // pixel lanes per block (threads beyond lanes*C4 idle)
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  float4 mu = s0, is = s0;
  if (MODE == 1 && mean) {
    mu = *reinterpret_cast<const float4*>(mean + 4 * c4);
    is = *reinterpret_cast<const float4*>(invstd + 4 * c4);
  }
  if (pl < lanes) {
    for (size_t p = (size_t)blockIdx.x * lanes + pl; p < npix; p += (size_t)gridDim.x * lanes) {
      const size_t o = p * C + 4 * c4;
      const float4 v = Elem<T>::load4(x + o);
      if (MODE == 0) {
        s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
        s1.x = fmaf(v.x, v.x, s1.x); s1.y = fmaf(v.y, v.y, s1.y); s1.z = fmaf(v.z, v.z, s1.z); s1.w = fmaf(v.w, v.w, s1.w);
      } else {
        float4 g = Elem<T>::load4(dy + o);
        if (y) {  // ReLU backward: the stored output is positive exactly where the pre-activation was
          const float4 yy = Elem<T>::load4(y + o);
          g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f; g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
        }
        s0.x += g.x; s0.y += g.y; s0.z += g.z; s0.w += g.w;
        s1.x = fmaf(g.x, (v.x - mu.x) * is.x, s1.x); s1.y = fmaf(g.y, (v.y - mu.y) * is.y, s1.y);
        s1.z = fmaf(g.z, (v.z - mu.z) * is.z, s1.z); s1.w = fmaf(g.w, (v.w - mu.w) * is.w, s1.w);
      }
    }
  }
  // block reduction over the pixel lanes (fixed order)
  __shared__ float4 sh0[kRedThreads], sh1[kRedThreads];
  sh0[threadIdx.x] = s0;
  sh1[threadIdx.x] = s1;
  __syncthreads();
  if (pl == 0) {
    double a[4] = {0, 0, 0, 0}, b2[4] = {0, 0, 0, 0};
    for (int l = 0; l < lanes; ++l) {
      const float4 u = sh0[l * C4 + c4], w = sh1[l * C4 + c4];
      a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w;
      b2[0] += w.x; b2[1] += w.y; b2[2] += w.z; b2[3] += w.w;
    }
    for (int j = 0; j < 4; ++j) partial[(size_t)blockIdx.x * C + 4 * c4 + j] = make_double2(a[j], b2[j]);
  }
}

// MODE 0: mean / biased variance -> scale, shift (and mean, invstd for the backward); running statistics updated in place
//         exactly as nn.BatchNorm3d does (momentum 0.1, unbiased variance; reference :355).
// MODE 1: dgamma, dbeta.
// One warp per channel: the up to 592 per-block partials are summed by 32 lanes (fixed lane order, fixed shuffle tree: deterministic)
// instead of serially by one thread — that serial loop was 78 us per launch, 22 % of a training step over its 348 launches.
__global__ void __launch_bounds__(32) bn_finalize_kernel(const double2* __restrict__ partial, int nblocks, int C, double npix, int mode,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps, float* __restrict__ out0,
                                   float* __restrict__ out1, float* __restrict__ out2, float* __restrict__ out3) {
  const int c = blockIdx.x;
  if (c >= C) return;
  double a = 0, b = 0;
  for (int i = threadIdx.x; i < nblocks; i += 32) {
    const double2 v = partial[(size_t)i * C + c];
    a += v.x;
    b += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  if (threadIdx.x != 0) return;
  if (mode == 0) {
    const double m = a / npix;
    double var = b / npix - m * m;
    if (var < 0) var = 0;
    const double inv = 1.0 / sqrt(var + (double)eps);
    const double g = gamma ? (double)gamma[c] : 1.0, bt = beta ? (double)beta[c] : 0.0;
    out0[c] = (float)(g * inv);              // scale
    out1[c] = (float)(bt - m * g * inv);     // shift
    out2[c] = (float)m;                      // mean
    out3[c] = (float)inv;                    // invstd
    if (running_mean) {
      const double unbiased = npix > 1 ? var * npix / (npix - 1.0) : var;
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
  } else {
    out0[c] = (float)b;  // dgamma = sum g * xhat
    out1[c] = (float)a;  // dbeta  = sum g
  }
}

// BatchNorm3d with its RUNNING statistics inside a differentiable (taped) forward — frozen-BN fine-tuning, gradients w.r.t. the
// input of an eval-mode network: mean = running_mean, invstd = 1/sqrt(running_var + eps); nothing is updated.
__global__ void bn_eval_stats_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rm,
                                     const float* __restrict__ rv, float eps, int C, float* __restrict__ scale, float* __restrict__ shift,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double inv = 1.0 / sqrt((double)rv[c] + (double)eps);
  const double g = gamma ? (double)gamma[c] : 1.0, bt = beta ? (double)beta[c] : 0.0, m = rm[c];
  scale[c] = (float)(g * inv);
  shift[c] = (float)(bt - m * g * inv);
  mean[c] = (float)m;
  invstd[c] = (float)inv;
}

// y = x*scale[c] + shift[c] (+ res_pre) ; ReLU ; (+ res_post)        (scale/shift null: identity)
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                const T* __restrict__ res_pre, const T* __restrict__ res_post, int relu, size_t n4, int C,
                                T* __restrict__ out) {
  const int C4 = C / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    float4 v = Elem<T>::load4(x + 4 * i);
    if (scale) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4), sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    if (res_pre) {
      const float4 r = Elem<T>::load4(res_pre + 4 * i);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (res_post) {
      const float4 r = Elem<T>::load4(res_post + 4 * i);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    Elem<T>::store4(out + 4 * i, v);
  }
}

// g = dy * (y > 0 if relu mask given) ; dx = gamma*invstd*(g - dbeta/M - xhat*dgamma/M)   (no BN: dx = g) ; optional g_out = g
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x,
                                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, float inv_m, size_t n4, int C,
                                    T* __restrict__ dx, T* __restrict__ g_out) {
  const int C4 = C / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    float4 g = Elem<T>::load4(dy + 4 * i);
    if (y) {
      const float4 yy = Elem<T>::load4(y + 4 * i);
      g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f; g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
    }
    if (g_out) Elem<T>::store4(g_out + 4 * i, g);
    if (!dx) continue;
    float4 d = g;
    if (mean) {
      const float4 v = Elem<T>::load4(x + 4 * i);
      const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c4), is = __ldg(reinterpret_cast<const float4*>(invstd) + c4);
      const float4 dg = __ldg(reinterpret_cast<const float4*>(dgamma) + c4), db = __ldg(reinterpret_cast<const float4*>(dbeta) + c4);
      float4 ga = make_float4(1.f, 1.f, 1.f, 1.f);
      if (gamma) ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      d.x = ga.x * is.x * (g.x - db.x * inv_m - (v.x - mu.x) * is.x * dg.x * inv_m);
      d.y = ga.y * is.y * (g.y - db.y * inv_m - (v.y - mu.y) * is.y * dg.y * inv_m);
      d.z = ga.z * is.z * (g.z - db.z * inv_m - (v.z - mu.z) * is.z * dg.z * inv_m);
      d.w = ga.w * is.w * (g.w - db.w * inv_m - (v.w - mu.w) * is.w * dg.w * inv_m);
    }
    Elem<T>::store4(dx + 4 * i, d);
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, size_t n4, T* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 u = Elem<T>::load4(a + 4 * i), v = Elem<T>::load4(b + 4 * i);
    Elem<T>::store4(out + 4 * i, make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w));
  }
}

// (1,k,k) pooling backward.  Max: the gradient goes to the FIRST maximum of the window in row-major order (ATen's rule:
// a later element replaces the running maximum only if strictly greater) ; average: dy / k^2 to every element.
template <typename T, bool kMax>
__global__ void pool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int BS, int H, int W, int C,
                                int k) {
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
    const size_t base = ((bs * H + (size_t)oy * k) * W + (size_t)ox * k) * C + 4 * c4;
    const float4 g = Elem<T>::load4(dy + i * 4);
    if (!kMax) {
      const float4 v = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
      for (int dyy = 0; dyy < k; ++dyy)
        for (int dxx = 0; dxx < k; ++dxx) Elem<T>::store4(dx + base + ((size_t)dyy * W + dxx) * C, v);
    } else {
      float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      int ax = 0, ay = 0, az = 0, aw = 0;
      for (int j = 0; j < k * k; ++j) {
        const float4 v = Elem<T>::load4(x + base + ((size_t)(j / k) * W + (j % k)) * C);
        if (v.x > m.x) { m.x = v.x; ax = j; }
        if (v.y > m.y) { m.y = v.y; ay = j; }
        if (v.z > m.z) { m.z = v.z; az = j; }
        if (v.w > m.w) { m.w = v.w; aw = j; }
      }
      for (int j = 0; j < k * k; ++j) {
        const float4 v = make_float4(j == ax ? g.x : 0.f, j == ay ? g.y : 0.f, j == az ? g.z : 0.f, j == aw ? g.w : 0.f);
        Elem<T>::store4(dx + base + ((size_t)(j / k) * W + (j % k)) * C, v);
      }
    }
  }
}

// Depth head backward: d cost[b,s,yl,xl] += w(y,x -> yl,xl) * d depth[b,y,x] * (fd_s - depth) / den * sigmoid'(v_s)
// One thread per output pixel recomputes the forward (bilinear taps, softplus, normalisation) and scatters with fp32 atomics
// (dcost zero-initialised by the caller of the launcher).
__global__ void depth_head_bwd_kernel(const float* __restrict__ cost, int h, int w, const float* __restrict__ fd, long long sb,
                                      long long ss, long long sy, long long sx, int B, int S, int H, int W,
                                      const float* __restrict__ ddepth, float* __restrict__ dcost) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const float ry = (float)h / (float)H, rx = (float)w / (float)W;
  float fy = ry * ((float)y + 0.5f) - 0.5f, fx = rx * ((float)x + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const bool same = (h == H) && (w == W);
  const float* fp = fd + b * sb + y * sy + x * sx;
  float num = 0.f, den = 0.f;
  for (int s = 0; s < S; ++s) {
    const float* c = cost + ((size_t)b * S + s) * h * w;
    const float v = same ? __ldg(c + (size_t)y * w + x)
                         : ly0 * (lx0 * __ldg(c + (size_t)y0 * w + x0) + lx1 * __ldg(c + (size_t)y0 * w + x1)) +
                               ly1 * (lx0 * __ldg(c + (size_t)y1 * w + x0) + lx1 * __ldg(c + (size_t)y1 * w + x1));
    const float p = (v > 20.f ? v : log1pf(expf(v))) + 1e-6f;
    den += p;
    num = fmaf(__ldg(fp + s * ss), p, num);
  }
  const float depth = num / den;
  const float go = __ldg(ddepth + ((size_t)b * H + y) * W + x) / den;
  for (int s = 0; s < S; ++s) {
    const float* c = cost + ((size_t)b * S + s) * h * w;
    float* dc = dcost + ((size_t)b * S + s) * h * w;
    const float v = same ? __ldg(c + (size_t)y * w + x)
                         : ly0 * (lx0 * __ldg(c + (size_t)y0 * w + x0) + lx1 * __ldg(c + (size_t)y0 * w + x1)) +
                               ly1 * (lx0 * __ldg(c + (size_t)y1 * w + x0) + lx1 * __ldg(c + (size_t)y1 * w + x1));
    const float sig = v > 20.f ? 1.f : 1.f / (1.f + expf(-v));   // d softplus / dv
    const float gv = go * (__ldg(fp + s * ss) - depth) * sig;
    if (same) {
      dc[(size_t)y * w + x] = gv;
    } else {
      atomicAdd(dc + (size_t)y0 * w + x0, gv * ly0 * lx0);
      atomicAdd(dc + (size_t)y0 * w + x1, gv * ly0 * lx1);
      atomicAdd(dc + (size_t)y1 * w + x0, gv * ly1 * lx0);
      atomicAdd(dc + (size_t)y1 * w + x1, gv * ly1 * lx1);
    }
  }
}

// dgrad weight pack: packed[t][a][b] (fp32, [ntaps][CaP][CbP]) with a = dy channel (the layer's Cout), b = dx channel inside
// [ci0, ci0+nci) of the layer's Cin.  conv weight (Cout,Cin,k): W[a][ci0+b][t] ; transposed (Cin,Cout,k): W[ci0+b][a][t].
__global__ void pack_weight_dgrad_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin, int ntaps, int ci0,
                                         int nci, int CaP, int CbP, int transposed) {
  const int n = ntaps * CaP * CbP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = i % CbP, a = (i / CbP) % CaP, t = i / (CbP * CaP);
    float v = 0.f;
    if (a < Cout && b < nci) {
      const int ci = ci0 + b;
      v = transposed ? w[((size_t)ci * Cout + a) * ntaps + t] : w[((size_t)a * Cin + ci) * ntaps + t];
    }
    dst[i] = v;
  }
}

// Weight of the ADJOINT convolution in the reference layout, so that a data gradient is one more call of the forward operator
// (tensor-core path).  a = dx channel inside [ci0, ci0+nci), b = dy channel (zero beyond Cout, i.e. for the stored padding of dy):
//   mode 0  layer = stride-1 conv     : adjoint = stride-1 conv,       weight (nci, CoS, k)  = W[b][ci0+a][flipped tap]
//   mode 1  layer = stride-2 conv     : adjoint = transposed conv,     weight (CoS, nci, k)  = W[b][ci0+a][tap]
//   mode 2  layer = transposed conv   : adjoint = stride-2 conv,       weight (nci, CoS, k)  = Wt[ci0+a][b][tap]
__global__ void adjoint_weight_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin, int ntaps, int ci0, int nci,
                                      int CoS, int mode) {
  const int n = nci * CoS * ntaps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = i % ntaps;
    int a, b;
    if (mode == 1) { a = (i / ntaps) % nci; b = i / (ntaps * nci); }
    else { b = (i / ntaps) % CoS; a = i / (ntaps * CoS); }
    float v = 0.f;
    if (b < Cout && ci0 + a < Cin) {
      if (mode == 0) v = w[((size_t)b * Cin + ci0 + a) * ntaps + (ntaps - 1 - t)];
      else if (mode == 1) v = w[((size_t)b * Cin + ci0 + a) * ntaps + t];
      else v = w[((size_t)(ci0 + a) * Cout + b) * ntaps + t];
    }
    dst[i] = v;
  }
}
int launch_adjoint_weight(const float* w, float* dst, int Cout, int Cin, int ntaps, int ci0, int nci, int CoS, int mode, cudaStream_t st) {
  const int n = nci * CoS * ntaps;
  int g = cdiv(n, 256);
  if (g > 512) g = 512;
  adjoint_weight_kernel<<<g, 256, 0, st>>>(w, dst, Cout, Cin, ntaps, ci0, nci, CoS, mode);
  DFF_LAUNCH_CHECK("adjoint_weight");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------------
static inline int ew_grid(size_t n, int threads) {
  size_t g = (n + threads - 1) / threads;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

size_t bn_partial_bytes(int C) { return (size_t)kRedBlocks * C * sizeof(double2); }

// stats -> scale/shift/mean/invstd (+ running statistics).  `partial`: >= bn_partial_bytes(C) of scratch.
int launch_bn_stats(const void* x, size_t npix, int C, bool bf16, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, float* scale, float* shift, float* mean, float* invstd,
                    void* partial, cudaStream_t st) {
  if (C % 4 || C > 4 * kRedThreads) return fail(-1, "bn_stats: C must be a multiple of 4 (<= 1024)");
  const int lanes = kRedThreads / (C / 4);
  int nb = (int)((npix + lanes - 1) / lanes);
  if (nb > kRedBlocks) nb = kRedBlocks;
  if (nb < 1) nb = 1;
  if (bf16) bn_reduce_kernel<__nv_bfloat16, 0><<<nb, kRedThreads, 0, st>>>((const __nv_bfloat16*)x, nullptr, nullptr, nullptr, nullptr, npix, C, (double2*)partial);
  else bn_reduce_kernel<float, 0><<<nb, kRedThreads, 0, st>>>((const float*)x, nullptr, nullptr, nullptr, nullptr, npix, C, (double2*)partial);
  DFF_LAUNCH_CHECK("bn_reduce");
  bn_finalize_kernel<<<C, 32, 0, st>>>((const double2*)partial, nb, C, (double)npix, 0, gamma, beta, running_mean, running_var,
                                                   momentum, eps, scale, shift, mean, invstd);
  DFF_LAUNCH_CHECK("bn_finalize");
  return 0;
}

int launch_bn_eval_stats(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C, float* scale,
                         float* shift, float* mean, float* invstd, cudaStream_t st) {
  bn_eval_stats_kernel<<<cdiv(C, 128), 128, 0, st>>>(gamma, beta, rm, rv, eps, C, scale, shift, mean, invstd);
  DFF_LAUNCH_CHECK("bn_eval_stats");
  return 0;
}

int launch_bn_apply(const void* x, const float* scale, const float* shift, const void* res_pre, const void* res_post, int relu,
                    size_t npix, int C, bool bf16, void* out, cudaStream_t st) {
  if (C % 4) return fail(-1, "bn_apply: C must be a multiple of 4");
  const size_t n4 = npix * C / 4;
  if (bf16) bn_apply_kernel<<<ew_grid(n4, 256), 256, 0, st>>>((const __nv_bfloat16*)x, scale, shift, (const __nv_bfloat16*)res_pre, (const __nv_bfloat16*)res_post, relu, n4, C, (__nv_bfloat16*)out);
  else bn_apply_kernel<<<ew_grid(n4, 256), 256, 0, st>>>((const float*)x, scale, shift, (const float*)res_pre, (const float*)res_post, relu, n4, C, (float*)out);
  DFF_LAUNCH_CHECK("bn_apply");
  return 0;
}

// Backward of  out = [relu]( BN(x) + res_pre ) (+ res_post is handled by the caller: its gradient is dy itself).
// y != null: ReLU mask from the stored (pre-res_post) output.  mean == null: no BatchNorm (dx = g).
// Writes dx (may be null), g_out (may be null: gradient of res_pre), dgamma/dbeta (with BN).
// fixed_stats: the statistics were constants of the forward (running statistics): dx = gamma*invstd*g, no batch-mean terms.
int launch_bn_backward(const void* dy, const void* y, const void* x, const float* mean, const float* invstd, const float* gamma,
                       size_t npix, int C, bool bf16, void* dx, void* g_out, float* dgamma, float* dbeta, void* partial,
                       cudaStream_t st, bool fixed_stats) {
  if (C % 4 || C > 4 * kRedThreads) return fail(-1, "bn_backward: C must be a multiple of 4 (<= 1024)");
  const size_t n4 = npix * C / 4;
  if (mean) {
    const int lanes = kRedThreads / (C / 4);
    int nb = (int)((npix + lanes - 1) / lanes);
    if (nb > kRedBlocks) nb = kRedBlocks;
    if (nb < 1) nb = 1;
    if (bf16) bn_reduce_kernel<__nv_bfloat16, 1><<<nb, kRedThreads, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, mean, invstd, npix, C, (double2*)partial);
    else bn_reduce_kernel<float, 1><<<nb, kRedThreads, 0, st>>>((const float*)x, (const float*)dy, (const float*)y, mean, invstd, npix, C, (double2*)partial);
    DFF_LAUNCH_CHECK("bn_bwd_reduce");
    bn_finalize_kernel<<<C, 32, 0, st>>>((const double2*)partial, nb, C, (double)npix, 1, nullptr, nullptr, nullptr, nullptr, 0.f,
                                                     0.f, dgamma, dbeta, nullptr, nullptr);
    DFF_LAUNCH_CHECK("bn_bwd_finalize");
  }
  const float inv_m = fixed_stats ? 0.f : 1.f / (float)npix;
  if (bf16) bn_bwd_apply_kernel<<<ew_grid(n4, 256), 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (const __nv_bfloat16*)x, mean, invstd, gamma, dgamma, dbeta, inv_m, n4, C, (__nv_bfloat16*)dx, (__nv_bfloat16*)g_out);
  else bn_bwd_apply_kernel<<<ew_grid(n4, 256), 256, 0, st>>>((const float*)dy, (const float*)y, (const float*)x, mean, invstd, gamma, dgamma, dbeta, inv_m, n4, C, (float*)dx, (float*)g_out);
  DFF_LAUNCH_CHECK("bn_bwd_apply");
  return 0;
}


// ------------------------------------------------------------------------------------------------------------------
// Loss + optimizer of the reference's training loop (train_codes/train_code_Defocus.py:17-19, 67, 160-168).
//   Total = sum_k w_k * mean_{mask} (pred_k - gt)^2 over the four heads.  The reference gathers `pred[mask]` / `gt[mask]` four times
//   (nonzero + index, and an index_put_ with a radix sort in the backward); here one pass reduces the four masked sums and the count,
//   a one-block finalise turns them into the loss, and one pass writes d Total / d pred_k = 2 w_k (pred_k - gt) mask / count.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kLossBlocks = 592;
struct Ptr4 { const float* p[4]; };
struct MPtr4 { float* p[4]; };

__global__ void __launch_bounds__(256) masked_mse_reduce_kernel(Ptr4 pred, const float* __restrict__ gt, const unsigned char* __restrict__ mask,
                                                                size_t n, double* __restrict__ partial) {
  double acc[5] = {0, 0, 0, 0, 0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (!mask[i]) continue;
    const float g = __ldg(gt + i);
    acc[4] += 1.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = __ldg(pred.p[k] + i) - g;
      acc[k] += (double)d * (double)d;
    }
  }
  __shared__ double sh[5][256];
#pragma unroll
  for (int k = 0; k < 5; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int k = 0; k < 5; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 5) partial[(size_t)blockIdx.x * 5 + threadIdx.x] = sh[threadIdx.x][0];
}
// stats: [0] valid count, [1] Total, [2..5] per-head masked MSE, [6] 1/count (0 when there is no valid pixel: zero loss, zero gradients)
__global__ void masked_mse_finalize_kernel(const double* __restrict__ partial, int nblocks, float w0, float w1, float w2, float w3,
                                           float* __restrict__ stats) {
  __shared__ double tot[5];
  if (threadIdx.x < 5) {
    double a = 0;
    for (int i = 0; i < nblocks; ++i) a += partial[(size_t)i * 5 + threadIdx.x];
    tot[threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double cnt = tot[4], inv = cnt > 0 ? 1.0 / cnt : 0.0;
    const double w[4] = {w0, w1, w2, w3};
    double total = 0;
    for (int k = 0; k < 4; ++k) {
      stats[2 + k] = (float)(tot[k] * inv);
      total += w[k] * tot[k] * inv;
    }
    stats[0] = (float)cnt;
    stats[1] = (float)total;
    stats[6] = (float)inv;
  }
}
__global__ void __launch_bounds__(256) masked_mse_grad_kernel(Ptr4 pred, const float* __restrict__ gt, const unsigned char* __restrict__ mask,
                                                              size_t n, float w0, float w1, float w2, float w3,
                                                              const float* __restrict__ stats, MPtr4 grad) {
  const float inv = stats[6];
  const float w[4] = {2.f * w0 * inv, 2.f * w1 * inv, 2.f * w2 * inv, 2.f * w3 * inv};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const bool m = mask[i] != 0;
    const float g = __ldg(gt + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) grad.p[k][i] = m ? w[k] * (__ldg(pred.p[k] + i) - g) : 0.f;
  }
}

int launch_masked_mse(const float* const pred[4], const float* gt, const unsigned char* mask, size_t n, const float w[4],
                      float* const grad[4], float* stats, double* scratch, cudaStream_t st) {
  Ptr4 pp; MPtr4 gp;
  for (int k = 0; k < 4; ++k) { pp.p[k] = pred[k]; gp.p[k] = grad[k]; }
  int nb = (int)((n + 255) / 256);
  if (nb > kLossBlocks) nb = kLossBlocks;
  if (nb < 1) nb = 1;
  masked_mse_reduce_kernel<<<nb, 256, 0, st>>>(pp, gt, mask, n, scratch);
  DFF_LAUNCH_CHECK("masked_mse_reduce");
  masked_mse_finalize_kernel<<<1, 32, 0, st>>>(scratch, nb, w[0], w[1], w[2], w[3], stats);
  DFF_LAUNCH_CHECK("masked_mse_finalize");
  masked_mse_grad_kernel<<<ew_grid(n, 256), 256, 0, st>>>(pp, gt, mask, n, w[0], w[1], w[2], w[3], stats, gp);
  DFF_LAUNCH_CHECK("masked_mse_grad");
  return 0;
}

// Adam over ONE flat fp32 buffer (parameters, gradients, exp_avg, exp_avg_sq all flat and aligned): the element-wise arithmetic of
// torch.optim.Adam's default CUDA (foreach) implementation in the same order and rounding —
//   m = lerp(m, g, 1-b1) ; v = v*b2 + (1-b2)*(g*g) ; denom = sqrt(v)/sqrt(1-b2^t) + eps ; p = p + (-lr/(1-b1^t)) * (m/denom)
// — so optimizer states stay interchangeable with the reference's `torch.optim.Adam(model.parameters(), betas=(0.9, 0.99))`.
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, size_t n, float one_minus_b1, float b2, float one_minus_b2,
                                                        float bc2_sqrt, float eps, float neg_step, const float* __restrict__ gscale) {
  const float gs = gscale ? __ldg(gscale) : 1.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gr = g[i];
    if (gscale) gr *= gs;
    float mi = m[i], vi = v[i];
    mi = mi + one_minus_b1 * (gr - mi);                 // lerp, weight < 0.5
    vi = __fmul_rn(vi, b2);
    vi = fmaf(one_minus_b2, __fmul_rn(gr, gr), vi);     // _foreach_addcmul_: self + value * (t1 * t2)
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
    p[i] = p[i] + neg_step * __fdiv_rn(mi, denom);      // addcdiv: self + value * (t1 / t2)
    m[i] = mi;
    v[i] = vi;
  }
}

int launch_adam_flat(float* p, const float* g, float* m, float* v, size_t n, double lr, double b1, double b2, double eps, int step,
                     const float* gscale, cudaStream_t st) {
  // scalars are computed in double on the host exactly as the Python optimizer does, then narrowed to fp32 as ATen narrows them
  const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
  const float one_minus_b1 = (float)(1.0 - b1), one_minus_b2 = (float)(1.0 - b2);
  const float neg_step = (float)(-(lr / bc1)), bc2_sqrt = (float)sqrt(bc2);
  adam_flat_kernel<<<ew_grid(n, 256), 256, 0, st>>>(p, g, m, v, n, one_minus_b1, (float)b2, one_minus_b2, bc2_sqrt, (float)eps, neg_step, gscale);
  DFF_LAUNCH_CHECK("adam_flat");
  return 0;
}

int launch_add(const void* a, const void* b, size_t n, bool bf16, void* out, cudaStream_t st) {
  if (n % 4) return fail(-1, "add: element count must be a multiple of 4");
  if (bf16) add_kernel<<<ew_grid(n / 4, 256), 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, n / 4, (__nv_bfloat16*)out);
  else add_kernel<<<ew_grid(n / 4, 256), 256, 0, st>>>((const float*)a, (const float*)b, n / 4, (float*)out);
  DFF_LAUNCH_CHECK("add");
  return 0;
}

int launch_pool_bwd(const void* x, const void* dy, void* dx, int BS, int H, int W, int C, int k, bool is_max, bool bf16, cudaStream_t st) {
  if (C % 4 || H % k || W % k) return fail(-1, "pool_bwd: C % 4, H % k, W % k must be 0");
  const size_t n = (size_t)BS * (H / k) * (W / k) * (C / 4);
  const int g = ew_grid(n, 256);
  if (bf16) {
    if (is_max) pool_bwd_kernel<__nv_bfloat16, true><<<g, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, BS, H, W, C, k);
    else pool_bwd_kernel<__nv_bfloat16, false><<<g, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, BS, H, W, C, k);
  } else {
    if (is_max) pool_bwd_kernel<float, true><<<g, 256, 0, st>>>((const float*)x, (const float*)dy, (float*)dx, BS, H, W, C, k);
    else pool_bwd_kernel<float, false><<<g, 256, 0, st>>>((const float*)x, (const float*)dy, (float*)dx, BS, H, W, C, k);
  }
  DFF_LAUNCH_CHECK("pool_bwd");
  return 0;
}

int launch_depth_head_bwd(const float* cost, int h, int w, const float* fd, const int64_t* st4, int B, int S, int H, int W,
                          const float* ddepth, float* dcost, cudaStream_t st) {
  if (h <= 0 || w <= 0 || H % h || W % w) return fail(-1, "depth_head_bwd: H,W must be multiples of the cost resolution");
  if (!(h == H && w == W)) DFF_CUDA(cudaMemsetAsync(dcost, 0, (size_t)B * S * h * w * sizeof(float), st));
  dim3 grid(cdiv(W, 128), H, B);
  depth_head_bwd_kernel<<<grid, 128, 0, st>>>(cost, h, w, fd, st4[0], st4[1], st4[2], st4[3], B, S, H, W, ddepth, dcost);
  DFF_LAUNCH_CHECK("depth_head_bwd");
  return 0;
}

int launch_pack_weight_dgrad(const float* w, float* dst, int Cout, int Cin, int ntaps, int ci0, int nci, int CaP, int CbP,
                             int transposed, cudaStream_t st) {
  const int n = ntaps * CaP * CbP;
  int g = cdiv(n, 256);
  if (g > 512) g = 512;
  pack_weight_dgrad_kernel<<<g, 256, 0, st>>>(w, dst, Cout, Cin, ntaps, ci0, nci, CaP, CbP, transposed);
  DFF_LAUNCH_CHECK("pack_weight_dgrad");
  return 0;
}

}  // namespace dff
