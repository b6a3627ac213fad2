// On-device evaluation tail of the reference's test loop: the masked depth metrics of metrics.py:90-133 and the
// crop + colour-mapped output image of Depth_Estimation_Test/test.py:123-140, so an eval loop needs no per-stack device->host
// copy of the fp32 depth map (test.py:121 `test_pred3.data.cpu().numpy()` is a 0.9 MB synchronising read per stack).
//
// dff_depth_metrics: one pass over est / gt / mask (/ conf) per stack, fp64 two-stage deterministic reduction, 11 figures per stack:
//   abs_rel, sq_rel, mse, mae, rmse, rmse_log, accuracy_1..3 (delta < 1.25^k), mse_w_conf, mae_w_conf.
// dff_depth_to_jet:  pred[:Hc,:Wc] -> (pred - lo)/(hi - lo) -> matplotlib's 'jet' look-up table -> uint8 RGB (Hc,Wc,3).
#include "common.cuh"

namespace dff {

constexpr int kMetBlocks = 64;    // partial blocks per stack
constexpr int kMetTerms = 12;

// partial[(b*kMetBlocks + blk)*kMetTerms + k]
__global__ void __launch_bounds__(256) depth_metrics_reduce_kernel(const float* __restrict__ est, const float* __restrict__ gt,
                                                                   const unsigned char* __restrict__ mask, const float* __restrict__ conf,
                                                                   int H, int W, int Hc, int Wc, double* __restrict__ partial) {
  const int b = blockIdx.y;
  const size_t n = (size_t)Hc * Wc;
  const float* e = est + (size_t)b * H * W;        // padded (H, W) maps, metrics over the [:Hc, :Wc] crop (test.py:125)
  const float* g = gt + (size_t)b * n;
  const unsigned char* m = mask ? mask + (size_t)b * n : nullptr;
  const float* c = conf ? conf + (size_t)b * n : nullptr;
  double acc[kMetTerms];
#pragma unroll
  for (int k = 0; k < kMetTerms; ++k) acc[k] = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (m && !m[i]) continue;
    const int y = (int)(i / Wc), x = (int)(i - (size_t)y * Wc);
    const double ev = (double)__ldg(e + (size_t)y * W + x), gv = (double)__ldg(g + i);
    const double d = gv - ev, ad = fabs(d);
    acc[0] += 1.0;
    acc[1] += ad / gv;                       // mask_abs_rel  (metrics.py:90-91)
    acc[2] += d * d / gv;                    // mask_sq_rel   (:93-94)
    acc[3] += d * d;                         // mask_mse / mask_rmse (:96-97, 102-103)
    acc[4] += ad;                            // mask_mae      (:99-100)
    const double ld = log(gv) - log(ev);
    acc[5] += ld * ld;                       // mask_rmse_log (:105-109)
    const double th = fmax(ev / gv, gv / ev);   // mask_accuracy_k (:112-121)
    acc[6] += th < 1.25 ? 1.0 : 0.0;
    acc[7] += th < 1.25 * 1.25 ? 1.0 : 0.0;
    acc[8] += th < 1.25 * 1.25 * 1.25 ? 1.0 : 0.0;
    if (c) {
      const double cv = (double)__ldg(c + i);
      acc[9] += cv * d * d;                  // mask_mse_w_conf (:123-124)
      acc[10] += cv * ad;                    // mask_mae_w_conf (:126-127)
      acc[11] += cv;
    }
  }
  __shared__ double sh[256];
  for (int k = 0; k < kMetTerms; ++k) {
    sh[threadIdx.x] = acc[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[((size_t)b * gridDim.x + blockIdx.x) * kMetTerms + k] = sh[0];
    __syncthreads();
  }
}

__global__ void depth_metrics_finalize_kernel(const double* __restrict__ partial, int nblocks, int has_conf, float* __restrict__ out) {
  const int b = blockIdx.x;
  __shared__ double tot[kMetTerms];
  if (threadIdx.x < kMetTerms) {
    double a = 0;
    for (int i = 0; i < nblocks; ++i) a += partial[((size_t)b * nblocks + i) * kMetTerms + threadIdx.x];
    tot[threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = tot[0];
    float* o = out + (size_t)b * 12;
    o[0] = (float)(tot[1] / n); o[1] = (float)(tot[2] / n); o[2] = (float)(tot[3] / n); o[3] = (float)(tot[4] / n);
    o[4] = (float)sqrt(tot[3] / n); o[5] = (float)sqrt(tot[5] / n);
    o[6] = (float)(tot[6] / n); o[7] = (float)(tot[7] / n); o[8] = (float)(tot[8] / n);
    o[9] = has_conf ? (float)(tot[9] / tot[11]) : 0.f;
    o[10] = has_conf ? (float)(tot[10] / tot[11]) : 0.f;
    o[11] = (float)n;
  }
}

// matplotlib's 'jet' (LinearSegmentedColormap, 256-entry look-up table): piecewise-linear segment data
__device__ __forceinline__ float jet_seg(float x, const float* xs, const float* ys, int n) {
  for (int i = 1; i < n; ++i)
    if (x <= xs[i]) return ys[i - 1] + (ys[i] - ys[i - 1]) * (x - xs[i - 1]) / (xs[i] - xs[i - 1]);
  return ys[n - 1];
}
__global__ void jet_lut_kernel(unsigned char* __restrict__ lut) {
  const int i = threadIdx.x;   // 256 entries
  const float x = (float)i / 255.f;
  const float rx[5] = {0.f, 0.35f, 0.66f, 0.89f, 1.f}, ry[5] = {0.f, 0.f, 1.f, 1.f, 0.5f};
  const float gx[6] = {0.f, 0.125f, 0.375f, 0.64f, 0.91f, 1.f}, gy[6] = {0.f, 0.f, 1.f, 1.f, 0.f, 0.f};
  const float bx[5] = {0.f, 0.11f, 0.34f, 0.65f, 1.f}, by[5] = {0.5f, 1.f, 1.f, 0.f, 0.f};
  lut[3 * i] = (unsigned char)(jet_seg(x, rx, ry, 5) * 255.f + 0.5f);
  lut[3 * i + 1] = (unsigned char)(jet_seg(x, gx, gy, 6) * 255.f + 0.5f);
  lut[3 * i + 2] = (unsigned char)(jet_seg(x, bx, by, 5) * 255.f + 0.5f);
}
__global__ void __launch_bounds__(256) depth_to_jet_kernel(const float* __restrict__ est, int H, int W, int Hc, int Wc, float lo, float hi,
                                                           const unsigned char* __restrict__ lut, unsigned char* __restrict__ out) {
  __shared__ unsigned char sl[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) sl[i] = lut[i];
  __syncthreads();
  const int b = blockIdx.y;
  const size_t n = (size_t)Hc * Wc;
  const float inv = 1.f / (hi - lo);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / Wc), x = (int)(i - (size_t)y * Wc);
    const float v = (__ldg(est + ((size_t)b * H + y) * W + x) - lo) * inv;
    // Colormap.__call__: floats in [0, 1) -> int(v * N); v == 1 -> N - 1; below / above range clip to the end colours
    int k = v >= 1.f ? 255 : (v < 0.f ? 0 : (int)(v * 256.f));
    k = k > 255 ? 255 : k;
    unsigned char* o = out + ((size_t)b * n + i) * 3;
    o[0] = sl[3 * k]; o[1] = sl[3 * k + 1]; o[2] = sl[3 * k + 2];
  }
}

size_t depth_metrics_scratch_bytes(int B) { return (size_t)B * kMetBlocks * kMetTerms * sizeof(double); }

int launch_depth_metrics(const float* est, const float* gt, const unsigned char* mask, const float* conf, int B, int H, int W, int Hc,
                         int Wc, float* out, double* scratch, cudaStream_t st) {
  dim3 grid(kMetBlocks, B);
  depth_metrics_reduce_kernel<<<grid, 256, 0, st>>>(est, gt, mask, conf, H, W, Hc, Wc, scratch);
  DFF_LAUNCH_CHECK("depth_metrics_reduce");
  depth_metrics_finalize_kernel<<<B, 32, 0, st>>>(scratch, kMetBlocks, conf ? 1 : 0, out);
  DFF_LAUNCH_CHECK("depth_metrics_finalize");
  return 0;
}

int launch_depth_to_jet(const float* est, int B, int H, int W, int Hc, int Wc, float lo, float hi, unsigned char* lut768,
                        unsigned char* out, cudaStream_t st) {
  jet_lut_kernel<<<1, 256, 0, st>>>(lut768);
  DFF_LAUNCH_CHECK("jet_lut");
  dim3 grid(cdiv(Hc * Wc, 256 * 4), B);
  depth_to_jet_kernel<<<grid, 256, 0, st>>>(est, H, W, Hc, Wc, lo, hi, lut768, out);
  DFF_LAUNCH_CHECK("depth_to_jet");
  return 0;
}

}  // namespace dff
