// Host side of the dff_b200 library: the layer table of DFF_net, the raw-parameter handshake, weight packing,
// the forward schedule (reference train_codes/Depth_Estimation_Network.py:77-137) and the C-ABI of
// include/dff_b200.h.  No device memory is allocated here; activations live in the caller's workspace.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dff_b200.h"
#include "common.cuh"

namespace dff {

// ---- launchers implemented in the other translation units --------------------------------------------------
int launch_conv_ffma(ConvArgs a, bool bf16, cudaStream_t st);
int launch_to_cl(const float* src, int B, int C, int S, int H, int W, void* dst, int Cp, bool bf16, cudaStream_t st);
int launch_to_cl_pair(const float* src, int B, int S, int H, int W, void* dst, cudaStream_t st);
int launch_pair_weight(const float* w, float* dst, int Cout, cudaStream_t st);
int launch_u8_to_cl_pair(const unsigned char* src, int B, int S, int H0, int W0, int H, int W, void* dst, cudaStream_t st);
int launch_u8_to_cl(const unsigned char* src, int B, int S, int H0, int W0, int H, int W, void* dst, int Cp, bool bf16, cudaStream_t st);
int launch_u8_to_planar(const unsigned char* src, int B, int S, int H0, int W0, int H, int W, float* dst, cudaStream_t st);
int launch_from_cl(const void* src, int B, int C, int S, int H, int W, int Cp, bool bf16, float* dst, cudaStream_t st);
int launch_pool(const void* src, void* dst, int BS, int H, int W, int C, int k, bool is_max, bool bf16, cudaStream_t st);
int launch_avgpool_pyramid(const void* src, void* d2, void* d4, void* d8, int BS, int H, int W, int C, cudaStream_t st);
int launch_depth_head(const float* cost, int h, int w, const float* fd, const int64_t* st4, int B, int S, int H, int W,
                      float* depth, cudaStream_t st);
int launch_depth_head4(const float* const cost[4], const int h[4], const int w[4], const float* fd, const int64_t* st4, int B, int S,
                       int H, int W, float* const depth[4], bool fast, cudaStream_t st);
int launch_srd_attention(const void* F, const float* w0, const float* w1, void* out, int B, int S, int H, int W, int C, cudaStream_t st);
int launch_srd_attention_mma(const void* F, const float* w0, const float* w1, void* out, int B, int S, int H, int W, int C, int num_sms,
                             cudaStream_t st);
int launch_fov_warp_cl(const void* x, const float* alpha, const float* fov, int B, int C, int S, int H, int W, void* out, bool bf16,
                       cudaStream_t st);
int launch_xpair_weight(const float* w, float* dst, int Cout, int Cin, int kd, cudaStream_t st);
int launch_xgroup_weight(const float* w, float* dst, int Cout, int Cin, int kd, int s, int G, int P, int nq, cudaStream_t st);
int launch_pair_volume(const void* feat, const float* alpha, const float* fov, int B, int C, int S, int H, int W, void* out, bool bf16,
                       cudaStream_t st, int Cs = 0);
int launch_spatial_mean_accum(const float* x, int Cs, int B, int S, int H, int W, const float* alpha_in, float s0, float s1, float s2,
                              float* alpha_out, void* scratch, cudaStream_t st);
size_t spatial_mean_scratch_bytes(int B, int S, int H, int W);
int launch_fov_warp(const float* x, const float* alpha, const float* fov, int B, int C, int S, int H, int W, float* out,
                    float* flow, cudaStream_t st);
int launch_pack_weight(const float* w, float* dst, int Cout, int Cin, int ntaps, int CinP, int CoutP, int transposed,
                       cudaStream_t st);
bool conv_tc_supported(const ConvArgs& a, int Ntc);
int launch_conv_tc(const ConvArgs& a, const void* wtc, int ntaps_total, int Ntc, int num_sms, cudaStream_t st);
int launch_pack_weight_tc(const float* w, void* dst, int Cout, int Cin, int CinP, int ntaps, int Ntc, int transposed,
                          cudaStream_t st);
bool conv_slab_supported(const ConvArgs& a, const TapTable* ptaps, int nph, int Ntc);
int launch_conv_slab(const ConvArgs& a, const TapTable* ptaps, int nph, const void* wslab, int Ntc, int num_sms, cudaStream_t st);
int launch_pack_weight_slab(const float* w, void* dst, int Cout, int Cin, int CinP, int ntaps, int Ntc, int transposed,
                            cudaStream_t st);
int launch_pack_weight_slab_deconv_fold(const float* w, void* dst, int Cout, int Cin, int CinP, cudaStream_t st);
int launch_pack_weight_slab_zmerge(const float* w, void* dst, int Cout, int Cin, int CinP, int Ntc, cudaStream_t st);
int launch_replicate_ss(const float* scale, const float* shift, float* dst, int C, int G, cudaStream_t st);
int launch_pack_weight_slab_fold(const float* w, void* dst, int Cout, int Cin, int CinP, int kd, int kh, int kw, int G, cudaStream_t st,
                                 int foldy = 0);
int launch_pack_weight_slab_rowfold(const float* wpair, void* dst, int Cout, cudaStream_t st);
int launch_conv_wgrad(ConvArgs a, const void* dy, int CoS, int Cout, int Cin, int ci_base, float* dw, int ntaps_total,
                      int wt_transposed, bool bf16, cudaStream_t st);
int launch_conv_wgrad_tma(const ConvArgs& a, const TapTable* pt, int nph, const void* dy, int CoS, int Cout, int Cin, int ci_base,
                          float* dw, int ntaps_total, int wt_transposed, cudaStream_t st, bool* handled);
size_t bn_partial_bytes(int C);
int launch_bn_stats(const void* x, size_t npix, int C, bool bf16, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, float* scale, float* shift, float* mean, float* invstd,
                    void* partial, cudaStream_t st);
int launch_bn_apply(const void* x, const float* scale, const float* shift, const void* res_pre, const void* res_post, int relu,
                    size_t npix, int C, bool bf16, void* out, cudaStream_t st);
int launch_bn_backward(const void* dy, const void* y, const void* x, const float* mean, const float* invstd, const float* gamma,
                       size_t npix, int C, bool bf16, void* dx, void* g_out, float* dgamma, float* dbeta, void* partial,
                       cudaStream_t st, bool fixed_stats = false);
int launch_bn_eval_stats(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C, float* scale,
                         float* shift, float* mean, float* invstd, cudaStream_t st);
int launch_add(const void* a, const void* b, size_t n, bool bf16, void* out, cudaStream_t st);
int launch_adjoint_weight(const float* w, float* dst, int Cout, int Cin, int ntaps, int ci0, int nci, int CoS, int mode, cudaStream_t st);
size_t depth_metrics_scratch_bytes(int B);
int launch_depth_metrics(const float* est, const float* gt, const unsigned char* mask, const float* conf, int B, int H, int W, int Hc,
                         int Wc, float* out, double* scratch, cudaStream_t st);
int launch_depth_to_jet(const float* est, int B, int H, int W, int Hc, int Wc, float lo, float hi, unsigned char* lut768,
                        unsigned char* out, cudaStream_t st);
int launch_masked_mse(const float* const pred[4], const float* gt, const unsigned char* mask, size_t n, const float w[4],
                      float* const grad[4], float* stats, double* scratch, cudaStream_t st);
int launch_adam_flat(float* p, const float* g, float* m, float* v, size_t n, double lr, double b1, double b2, double eps, int step,
                     const float* gscale, cudaStream_t st);
int launch_pool_bwd(const void* x, const void* dy, void* dx, int BS, int H, int W, int C, int k, bool is_max, bool bf16, cudaStream_t st);
int launch_depth_head_bwd(const float* cost, int h, int w, const float* fd, const int64_t* st4, int B, int S, int H, int W,
                          const float* ddepth, float* dcost, cudaStream_t st);
int launch_pack_weight_dgrad(const float* w, float* dst, int Cout, int Cin, int ntaps, int ci0, int nci, int CaP, int CbP,
                             int transposed, cudaStream_t st);
bool conv_row_supported(const ConvArgs& a, int Ntc);
int launch_conv_row(const ConvArgs& a, const void* wslab, int Ntc, int num_sms, cudaStream_t st);
int launch_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* bias,
                   float* scale, float* shift, int C, int CP, cudaStream_t st);

// ---- errors ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
// Programmatic dependent launch pays while per-launch fixed costs matter (measured on DDFF stacks: -10 % at 2 stacks per call, -5 % at
// 4, +1 % at 16, +2 % at 64), so dff_forward switches it per call on the amount of work; DFF_B200_PDL=0/1 forces it off/on.
static thread_local bool g_pdl_call = true;
bool pdl_enabled() {
  static const int forced = getenv("DFF_B200_PDL") ? atoi(getenv("DFF_B200_PDL")) : (getenv("DFF_B200_NO_PDL") ? 0 : -1);
  return forced >= 0 ? forced != 0 : g_pdl_call;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
  cudaGetLastError();  // clear the (non-sticky) error so the next call starts clean
  return DFF_E_CUDA;
}

// ---- layer table -------------------------------------------------------------------------------------------
struct Layer {
  std::string name;  // state_dict prefix of the conv (key = name + ".weight")
  std::string bn;    // state_dict prefix of its BatchNorm3d ("" = none)
  int cin, cout, kd, kh, kw, stride, dil;
  bool transposed, bias;
  int gfold = 1;         // x-fold factor of the tensor-core path: G adjacent output pixels = G*Cout channels of one GEMM row (N' = 32)
  bool pair_x = false;   // tensor-core path of the first layer: input pixels carry their dilated right neighbour (9 x 5 paired taps)
  // derived
  int CinP, CoutP, ntaps;
  int CinT, Ntc;  // tensor-core path: stored input channels (multiple of 8) and MMA N (multiple of 16, >= 16)
  int64_t raw_w, raw_gamma, raw_beta, raw_mean, raw_var, raw_bias;  // element offsets in the raw buffer (-1: none)
  size_t pk_w, pk_scale, pk_shift, pk_wtc, pk_wslab, pk_proj, pk_pair, pk_wfold, pk_ssfold;                  // byte offsets in the packed buffer
  size_t pk_wfoldy = 0;  // y-folded form of the 8-output-channel x-folded layers (8-channel sources: merged TMA rows, no x-stride views)
  bool has_foldy = false;
  // x-paired form of the stride-2 layers with 8 / 16-channel sources: the source (.., W, C) is read as (.., W/2, 2C) — a stride-1 walk in x
  // over pixel PAIRS with two taps (pair -1: its second pixel = dx -1; pair 0: dx 0 and +1) — so the planes have two (row-parity) views
  // of 2C-channel pixels instead of four views of C-channel ones: half the TMA rows, twice as long (wide-row layout for 2C >= 16)
  size_t pk_xpw = 0, pk_wxp = 0;   // the equivalent (Cout, 2C, kd, 3, 2) convolution weight (fp32) and its slab pack
  bool has_xpair = false;
  // x-grouped forms of the layers with 8-channel sources whose planes were strided views of 16-byte TMA rows (and of the 1x3x3 16 -> 16
  // layers, P = 2): the source read as groups of P = 4 pixels (64-byte rows, wide-row layout), G = P / stride adjacent outputs per GEMM row, banded group-tap weights, all-zero K steps
  // not issued (ConvArgs::tap_kmask).  Stride 2 (8 -> <= 16): G = 2, group taps -1, 0.  Stride-1 1x3x3 8 -> 8: G = 4, group taps -1, 0, +1.
  bool has_xgroup = false;
  int xg_G = 0, xg_P = 0, xg_nq = 0;
  size_t pk_xgw = 0, pk_wxg = 0, pk_ssxg = 0;
  size_t pk_wz = 0;      // focal-merged streaming layout (3x3x3 layers with >= 64 stored input channels, whose weights are streamed)
  bool has_wz = false;
};
// Where each kernel layout of one layer lives inside a packed buffer (byte offsets from its start); `packed_bytes` is advanced.
// Shared by the network's layer table and the single-operator entry point (dff_conv3d with the forward's plan).
static void layout_layer(Layer& l, size_t& packed_bytes) {
  l.pk_w = packed_bytes;
  packed_bytes += align_up((size_t)l.ntaps * l.CinP * l.CoutP * sizeof(float), 256);
  l.pk_scale = packed_bytes;
  packed_bytes += align_up(l.CoutP * sizeof(float), 256);
  l.pk_shift = packed_bytes;
  packed_bytes += align_up(l.CoutP * sizeof(float), 256);
  l.pk_wtc = packed_bytes;
  packed_bytes += align_up((size_t)(l.ntaps + 1) * l.Ntc * l.CinT * 2, 256);
  l.pk_wslab = packed_bytes;
  packed_bytes += align_up((size_t)l.ntaps * l.Ntc * l.CinT * 2, 256);
  // x-folded form (see pack_weight_slab_fold_kernel): small-Cout stride-1 layers are bound by the A-operand fetch of the MMA
  // (39 clk for any N <= 32), so N' = G*Cout = 32 output channels per row cost the same as 8 or 16
  l.pk_wfold = l.pk_ssfold = packed_bytes;
  // (measured per layer class, profiles/: the fold pays where the folded kernel keeps its occupancy — Cout = 8 with G = 4 (Cin = 8)
  // or G = 2 (Cin >= 16), Cout = 16 only for Cin >= 32)
  const int cin = l.cin, cout = l.cout, kd = l.kd, kh = l.kh, kw = l.kw, stride = l.stride, dil = l.dil;
  const bool transposed = l.transposed;
  const int gsel = cout == 8 ? (cin == 8 ? 4 : 2) : (cout == 16 && cin >= 32 ? 2 : 1);
  if (!transposed && stride == 1 && dil == 1 && kw == 3 && gsel > 1 && cin % 8 == 0) {
    l.gfold = gsel;
    packed_bytes += align_up((size_t)kd * kh * (kw + l.gfold - 1) * l.CinT * l.gfold * cout * 2, 256);
    l.pk_ssfold = packed_bytes;
    packed_bytes += align_up((size_t)2 * l.gfold * cout * sizeof(float), 256);
    // (measured: 3x3x3 layers gain — dres4.conv0 2.11 -> 1.73 ms; the 1x3x3 G = 4 layers lose 6 %: their epilogue then writes four
    // 16-byte pieces per thread to four rows instead of 64 contiguous bytes)
    if (cout == 8 && kd == 3 && kh == 3 && (cin == 8 || cin == 16)) {   // (sources of 8 channels each: checked per call)
      l.has_foldy = true;
      l.pk_wfoldy = packed_bytes;
      packed_bytes += align_up((size_t)kd * (kh + l.gfold - 1) * kw * l.CinT * l.gfold * cout * 2, 256);
    }
  }
  if (transposed && cout <= 32 && cout % 8 == 0 && cin % 8 == 0) {   // x-folded transposed conv: the two column phases in one GEMM row
    l.gfold = 2;
    l.pk_wfold = packed_bytes;
    packed_bytes += align_up((size_t)18 * l.CinT * 2 * cout * 2, 256);
    l.pk_ssfold = packed_bytes;
    packed_bytes += align_up((size_t)2 * 2 * cout * sizeof(float), 256);
  }
  l.pk_pair = packed_bytes;   // paired-tap fp32 weights (Cout, 8, 1, 9, 5) of the first layer
  if (cin == 3 && kh == 9 && kw == 9 && dil == 2 && kd == 1) {
    l.pair_x = true;
    packed_bytes += align_up((size_t)cout * 360 * sizeof(float), 256);
    if (cout == 8) {   // row-folded form (pack_weight_slab_rowfold_kernel): 60 taps x 32 output channels; scale/shift replicated 4x
      l.pk_wfold = packed_bytes;
      packed_bytes += align_up((size_t)60 * 32 * 8 * 2, 256);
      l.pk_ssfold = packed_bytes;
      packed_bytes += align_up((size_t)2 * 32 * sizeof(float), 256);
    }
  }
  l.pk_proj = packed_bytes;   // C -> 1 projections (classifiers): contiguous fp32 weights for the fused epilogue
  if (cout == 1 && l.ntaps == 1) packed_bytes += align_up((size_t)l.CinT * sizeof(float), 256);
  l.pk_wz = packed_bytes;
  if (!transposed && kd == 3 && kh == 3 && kw == 3 && dil == 1 && l.CinT % 16 == 0 && l.CinT * l.Ntc >= 32 * 64) {
    l.has_wz = true;
    packed_bytes += align_up((size_t)27 * l.CinT * l.Ntc * 2, 256);
  }
  const bool xg8 = cin == 8 && l.CinT == 8 && cout % 8 == 0 && ((stride == 2 && cout <= 16) || (stride == 1 && kd == 1 && cout == 8));
  const bool xg16 = cin == 16 && l.CinT == 16 && cout == 16 && stride == 1 && kd == 1;   // 1x3x3 16 -> 16: pixel pairs, G = 2
  if (!transposed && dil == 1 && kh == 3 && kw == 3 && (xg8 || xg16)) {
    l.has_xgroup = true;
    l.xg_P = xg16 ? 2 : 4; l.xg_G = l.xg_P / stride; l.xg_nq = stride == 2 ? 2 : 3;
    const int Ng = (int)align_up(l.xg_G * cout, 16);
    l.pk_xgw = packed_bytes;
    packed_bytes += align_up((size_t)l.xg_G * cout * l.xg_P * cin * kd * 3 * l.xg_nq * sizeof(float), 256);
    l.pk_wxg = packed_bytes;
    packed_bytes += align_up((size_t)kd * 3 * l.xg_nq * Ng * l.xg_P * cin * 2, 256);
    l.pk_ssxg = packed_bytes;
    packed_bytes += align_up((size_t)2 * Ng * sizeof(float), 256);
  }
  if (!transposed && stride == 2 && dil == 1 && kh == 3 && kw == 3 && (cin == 8 || cin == 16) && l.CinT == cin) {
    l.has_xpair = true;
    l.pk_xpw = packed_bytes;
    packed_bytes += align_up((size_t)cout * 2 * cin * kd * 6 * sizeof(float), 256);
    l.pk_wxp = packed_bytes;
    packed_bytes += align_up((size_t)kd * 6 * l.Ntc * 2 * cin * 2, 256);
  }
}
struct Param {
  std::string name;
  int64_t numel, offset;
};
struct Net {
  std::vector<Layer> layers;
  std::map<std::string, int> index;
  std::vector<Param> params;
  int64_t raw_numel = 0;
  size_t packed_bytes = 0;

  // cin_store > 0: the input tensor stores that many channels in both precisions (zero weights beyond cin)
  void add(const std::string& name, const std::string& bn, int cin, int cout, int kd, int kh, int kw, int stride, int dil,
           bool transposed = false, bool bias = false, int cin_store = 0) {
    Layer l{name, bn, cin, cout, kd, kh, kw, stride, dil, transposed, bias};
    l.CinP = cin_store ? cin_store : (int)align_up(cin, 4);
    l.CoutP = (int)align_up(cout, 8);
    l.ntaps = kd * kh * kw;
    l.CinT = cin_store ? cin_store : (int)align_up(cin, 8);
    l.Ntc = (int)align_up(cout, 16);
    auto reg = [&](const std::string& n, int64_t numel) {
      params.push_back({n, numel, raw_numel});
      raw_numel += numel;
      return params.back().offset;
    };
    l.raw_w = reg(name + ".weight", (int64_t)cin * cout * l.ntaps);
    l.raw_bias = bias ? reg(name + ".bias", cout) : -1;
    l.raw_gamma = l.raw_beta = l.raw_mean = l.raw_var = -1;
    if (!bn.empty()) {
      l.raw_gamma = reg(bn + ".weight", cout);
      l.raw_beta = reg(bn + ".bias", cout);
      l.raw_mean = reg(bn + ".running_mean", cout);
      l.raw_var = reg(bn + ".running_var", cout);
    }
    layout_layer(l, packed_bytes);
    index[name] = (int)layers.size();
    layers.push_back(l);
  }
  // convbn_3d: `p.0` conv + `p.1` BN
  void cbn(const std::string& p, int cin, int cout, int kd, int kh, int kw, int stride = 1, int dil = 1) {
    add(p + ".0", p + ".1", cin, cout, kd, kh, kw, stride, dil);
  }
  void up(const std::string& p, int cin, int cout) { add(p + ".0", p + ".1", cin, cout, 3, 3, 3, 2, 1, true); }
  void srd(const std::string& p, int c) {
    cbn(p + ".Focus_Measure.conv.0", c, c, 1, 3, 3);
    cbn(p + ".Focus_Measure.conv.2", c, c, 1, 3, 3);
    add(p + ".N_ch_attention.0", "", c, c, 3, 1, 1, 1, 1);
    add(p + ".N_ch_attention.2", "", c, c, 1, 1, 1, 1, 1);
  }
  void efd(const std::string& p, int cin, int cout) {
    cbn(p + ".stride_conv", cin, cout, 3, 3, 3, 2);
    cbn(p + ".max_pooling.1", cin, cout, 3, 3, 3, 1);
  }
  void hourglass(const std::string& p, int c) {
    cbn(p + ".conv0.0", 2 * c, c, 3, 3, 3);
    cbn(p + ".conv1.0", c, 2 * c, 3, 3, 3, 2);
    cbn(p + ".conv2", 2 * c, 2 * c, 3, 3, 3);
    cbn(p + ".conv3.0", 2 * c, 2 * c, 3, 3, 3, 2);
    cbn(p + ".conv4.0", 2 * c, 2 * c, 3, 3, 3);
    up(p + ".conv5", 2 * c, 2 * c);
    up(p + ".conv6", 2 * c, c);
  }
};

// DFF_net (reference train_codes/Depth_Estimation_Network.py:17-57); only executed layers are listed: `redir3` and
// `pre_conv` exist in the state_dict but never run (reference :244, :285-286).
static Net build_dff() {
  Net n;
  n.cbn("FM_measure.Focus_extraction.0", 3, 8, 1, 9, 9, 1, 2);
  n.srd("FM_measure.Focus_extraction.2", 8);
  n.efd("FM_conv1.0", 8, 16);
  n.srd("FM_conv1.1", 16);
  n.efd("FM_conv2.0", 16, 32);
  n.srd("FM_conv2.1", 32);
  const std::string sp = "SPP_module.";
  const int c = 32;
  auto tower = [&](const std::string& a, const std::string& b, int ci, int co) {
    n.cbn(sp + a + ".0", ci, co, 3, 3, 3);
    n.cbn(sp + a + ".2", co, co, 3, 3, 3);
    n.cbn(sp + b + ".0", co, co, 3, 3, 3);
    n.cbn(sp + b + ".2", co, co, 3, 3, 3);
  };
  tower("dres8_0", "dres8_1", c, c);
  tower("dres16_0", "dres16_1", c, 2 * c);
  tower("dres32_0", "dres32_1", c, 2 * c);
  n.add(sp + "conv1", "", c, 2 * c, 3, 3, 3, 2, 1);
  n.cbn(sp + "conv2.0", 2 * c, 2 * c, 3, 3, 3);
  n.add(sp + "conv3", "", 2 * c, 4 * c, 3, 3, 3, 2, 1);
  n.cbn(sp + "conv4.0", 4 * c, 4 * c, 3, 3, 3);
  n.up(sp + "conv8", 4 * c, 2 * c);
  n.up(sp + "conv9", 2 * c, c);
  n.cbn(sp + "combine1.0", 4 * c, 2 * c, 3, 3, 3);
  n.cbn(sp + "combine2.0", 6 * c, 4 * c, 3, 3, 3);
  n.cbn(sp + "redir1", c, c, 1, 1, 1);
  n.cbn(sp + "redir2", 2 * c, 2 * c, 1, 1, 1);
  n.cbn("confidence.0", 32, 32, 3, 3, 3);
  n.add("confidence.2", "", 32, 1, 3, 3, 3, 1, 1);
  n.cbn("dres0.0", 32, 64, 3, 3, 3);
  n.cbn("dres0.2", 64, 64, 3, 3, 3);
  n.up("deconv_1", 64, 32);
  n.hourglass("dres2", 32);
  n.up("deconv_2", 32, 16);
  n.hourglass("dres3", 16);
  n.up("deconv_3", 16, 8);
  n.hourglass("dres4", 8);
  n.add("classif1.0", "", 32, 1, 1, 1, 1, 1, 1);
  n.add("classif2.0", "", 16, 1, 1, 1, 1, 1, 1);
  n.add("classif3.0", "", 8, 1, 1, 1, 1, 1, 1);
  return n;
}

// FlowNetwork(8), the End-to-End alignment network (reference End_to_End/End_to_End.py:18-61); keys relative to
// `optical_flow_aggregation.`.  Every layer is per-slice 2-D: 1x3x3 or 1x1x1.
static Net build_flow() {
  Net n;
  auto block = [&](const std::string& p, int cin, int cout, int stride) {   // resnet_block_2d_OF (reference :135-145)
    n.cbn(p + ".conv.0", cin, cout, 1, 3, 3, stride);
    n.cbn(p + ".conv.2", cout, cout, 1, 3, 3);
    n.add(p + ".feature", "", cin, cout, 1, 1, 1, stride, 1);
  };
  block("OF_feature.0", 3, 8, 1);
  block("OF_feature.1", 8, 8, 1);
  block("OF_feature1.0", 8, 16, 2);
  block("OF_feature1.1", 16, 16, 1);
  block("OF_feature2.0", 16, 32, 2);
  block("OF_feature2.1", 32, 32, 1);
  auto head = [&](const std::string& p, int c) {   // reference :31-61: input = [last slice | slice | flow] = 2c+2 channels, stored as 2c+16
    n.add(p + ".0.0", p + ".0.1", 2 * c + 2, 2 * c, 1, 3, 3, 1, 1, false, false, 2 * c + 16);
    n.cbn(p + ".2", 2 * c, 2 * c, 1, 3, 3);
    n.cbn(p + ".4", 2 * c, 2 * c, 1, 3, 3);
    n.add(p + ".6", "", 2 * c, 3, 1, 3, 3, 1, 1, false, true);
  };
  head("conv1", 32);
  head("conv2", 16);
  head("conv3", 8);
  return n;
}

static const Net& net_of(int which) {
  static const Net dffnet = build_dff();
  static const Net flownet = build_flow();
  return which == DFF_NET_FLOW ? flownet : dffnet;
}

// ---- tap tables --------------------------------------------------------------------------------------------
// Ordinary convolution: tap (kd,kh,kw) reads in[s + kd - pd, oy*stride + kh*dil - ph, ...] with "same"-style padding
// p = dil*(k-1)/2 (every conv of the network uses it: reference :144, 352-355, 361-367, 383-403).
static void conv_taps(const Layer& l, TapTable& t) {
  t.n = 0;
  const int pd = (l.kd - 1) / 2, ph = l.dil * (l.kh - 1) / 2, pw = l.dil * (l.kw - 1) / 2;
  for (int a = 0; a < l.kd; ++a)
    for (int b = 0; b < l.kh; ++b)
      for (int c = 0; c < l.kw; ++c) {
        t.dz[t.n] = (int8_t)(a - pd);
        t.dy[t.n] = (int8_t)(b * l.dil - ph);
        t.dx[t.n] = (int8_t)(c * l.dil - pw);
        t.widx[t.n] = (uint8_t)((a * l.kh + b) * l.kw + c);
        ++t.n;
      }
}
// Transposed convolution k=3, stride (1,2,2), pad 1, output_padding (0,1,1) (reference :43-50), output parity phase
// (py,px):  out[d, 2i+py, 2j+px] = sum_{kd} sum_{kh in K(py)} sum_{kw in K(px)} in[d+1-kd, i+dy(kh), j+dx(kw)] * w[kd,kh,kw]
// with K(0) = {1} (offset 0) and K(1) = {0 (offset +1), 2 (offset 0)}: 3 / 6 / 6 / 12 taps.
static void deconv_taps(int py, int px, TapTable& t) {
  t.n = 0;
  const int ky[2][2] = {{1, -1}, {0, 2}}, off[2][2] = {{0, 0}, {1, 0}};
  for (int kd = 0; kd < 3; ++kd)
    for (int iy = 0; iy < (py ? 2 : 1); ++iy)
      for (int ix = 0; ix < (px ? 2 : 1); ++ix) {
        const int kh = ky[py][iy], kw = ky[px][ix];
        t.dz[t.n] = (int8_t)(1 - kd);
        t.dy[t.n] = (int8_t)off[py][iy];
        t.dx[t.n] = (int8_t)off[px][ix];
        t.widx[t.n] = (uint8_t)((kd * 3 + kh) * 3 + kw);
        ++t.n;
      }
}

// ---- activation tensors in the workspace -------------------------------------------------------------------
struct Ten {
  void* p = nullptr;
  int B = 0, S = 0, H = 0, W = 0, C = 0;
  bool f32 = false;  // stored as fp32 regardless of the mode
};

struct EpiOpt {
  const Ten* in1 = nullptr;       // second source (virtual concat)
  const Ten* res_pre = nullptr;   // added before ReLU
  const Ten* res_post = nullptr;  // added after ReLU
  bool relu = false;
  const Ten* aux_add = nullptr;   // second output = out + aux_add
  Ten* aux_out = nullptr;
  bool out_f32 = false;
  const Layer* proj = nullptr;    // fused 1x1x1 classifier (tensor-core path): cost = proj(out or aux_out), fp32
  Ten* proj_out = nullptr;
  bool proj_aux = false, skip_out = false;
};

// Runs one conv layer (all phases) given packed weights.
static int num_sms_of_current_device() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  static int cache[64] = {0};
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

// Single-operator calls pack their weights on the fly: the per-tap TMA layout is only needed by the few shapes the slab kernel
// rejects, so it is packed when (and only when) run_conv falls through to that kernel.
struct LazyTcPack {
  const float* w;
  void* dst;
  int Cout, Cin, CinT, ntaps, Ntc, transposed;
  bool done;
};
static int ensure_tc_pack(LazyTcPack* z, cudaStream_t st) {
  if (!z || z->done) return 0;
  z->done = true;
  return launch_pack_weight_tc(z->w, z->dst, z->Cout, z->Cin, z->CinT, z->ntaps, z->Ntc, z->transposed, st);
}

// `wtc` != null selects the tcgen05 path (bf16 only); otherwise the FFMA kernel runs.
static int run_conv(const Layer& l, const float* w, const float* scale, const float* shift, const Ten& in, const EpiOpt& e,
                    Ten& out, bool bf16, cudaStream_t st, const void* wtc = nullptr, const void* wslab = nullptr,
                    int* nlaunch = nullptr, bool count_only = false, bool use_row = true, const char* packed_base = nullptr,
                    bool use_fold = false, LazyTcPack* lazy = nullptr) {
  int dummy = 0;
  if (!nlaunch) nlaunch = &dummy;
  *nlaunch = 0;
  ConvArgs a{};
  a.in0 = in.p;
  a.C0 = in.C;
  a.in1 = e.in1 ? e.in1->p : nullptr;
  a.C1 = e.in1 ? e.in1->C : 0;
  if (a.C0 + a.C1 != (wtc ? l.CinT : l.CinP))
    return fail(DFF_E_ARG, "conv " + l.name + ": input channels do not match the layer");
  const int nsm = wtc ? num_sms_of_current_device() : 0;
  a.B = in.B; a.S = in.S; a.IH = in.H; a.IW = in.W;
  a.OH = out.H; a.OW = out.W;
  a.w = w; a.CinP = l.CinP; a.CoutP = l.CoutP;
  a.scale = scale; a.shift = shift;
  a.res_pre = e.res_pre ? e.res_pre->p : nullptr;
  a.res_post = e.res_post ? e.res_post->p : nullptr;
  a.relu = e.relu ? 1 : 0;
  a.out = out.p;
  a.out_aux = e.aux_out ? e.aux_out->p : nullptr;
  a.aux_add = e.aux_add ? e.aux_add->p : nullptr;
  a.Cout = out.C;
  a.out_f32 = e.out_f32 ? 1 : 0;
  a.wz = (wtc && packed_base && l.has_wz) ? packed_base + l.pk_wz : nullptr;
  if (e.proj && wtc && packed_base) {
    a.proj_w = (const float*)(packed_base + e.proj->pk_proj);
    a.proj_out = (float*)e.proj_out->p;
    a.proj_src = e.proj_aux ? 1 : 0;
    a.skip_out = e.skip_out ? 1 : 0;
  }
  if (!l.transposed) {
    static const bool no_rowfold = getenv("DFF_B200_NO_ROWFOLD") != nullptr;
    if (l.pair_x && wtc && use_fold && packed_base && l.pk_wfold != l.pk_ssfold && !no_rowfold && !a.proj_w && a.Cout == 8 && out.H % 8 == 0) {
      // Row-folded first layer: with dilation 2 the output rows y, y+2, y+4, y+6 share 12 of their 4 x 9 tap rows, so they become the
      // 32 channels of ONE GEMM row (3x fewer MMAs, whose cost does not depend on N <= 32).  GEMM row (R, x) of phase p = 0, 1 holds the
      // output rows 8R + p + 2g: tile rows are 8 input rows apart (the A descriptor's row-group stride), the two row parities are the two
      // phases of the launch and share the staged plane; the epilogue stores group g two rows below group g-1.
      ConvArgs f = a;
      TapTable pt[2];
      for (int ph = 0; ph < 2; ++ph) {
        TapTable& t = pt[ph];
        t.n = 0;
        for (int q = 0; q < 12; ++q)
          for (int c = 0; c < 5; ++c) {
            t.dz[t.n] = 0;
            t.dy[t.n] = (int8_t)(2 * q - 8 + ph);
            t.dx[t.n] = (int8_t)(4 * c - 8 + 2);
            t.widx[t.n] = (uint8_t)(q * 5 + c);
            ++t.n;
          }
      }
      f.taps = pt[0];
      f.isy = f.isx = 1; f.row_step = 8; f.grp_rows = 2;
      f.osy = 8; f.osx = 1; f.ooy = f.oox = 0;
      f.OHt = out.H / 8; f.OWt = out.W;
      f.Cout = 32;
      f.scale = (const float*)(packed_base + l.pk_ssfold);
      f.shift = f.scale + 32;
      if (conv_slab_supported(f, pt, 2, 32)) {
        *nlaunch = 1;
        return count_only ? 0 : launch_conv_slab(f, pt, 2, packed_base + l.pk_wfold, 32, nsm, st);
      }
    }
    if (l.pair_x && wtc) {   // 9 (dy, dilation 2) x 5 (paired dx, step 4) taps on the pair-packed input
      a.taps.n = 0;
      for (int b = 0; b < 9; ++b)
        for (int c = 0; c < 5; ++c) {
          a.taps.dz[a.taps.n] = 0;
          a.taps.dy[a.taps.n] = (int8_t)(2 * b - 8);
          a.taps.dx[a.taps.n] = (int8_t)(4 * c - 8 + 2);   // +2: the pair-packed input has two margin columns on the left
          a.taps.widx[a.taps.n] = (uint8_t)(b * 5 + c);
          ++a.taps.n;
        }
    } else
    conv_taps(l, a.taps);
    a.isy = a.isx = l.stride; a.osy = a.osx = 1; a.ooy = a.oox = 0;
    a.OHt = out.H; a.OWt = out.W;
    *nlaunch = 1;
    if (count_only) return 0;
    static const bool no_xgroup = getenv("DFF_B200_NO_XGROUP") != nullptr;   // (A/B knob)
    static const int xgroup_sel = getenv("DFF_B200_XGROUP") ? atoi(getenv("DFF_B200_XGROUP")) : 7;   // (bit 0: stride-2 layers, bit 1: 1x3x3 8 -> 8, bit 2: 1x3x3 16 -> 16)
    if (wtc && use_fold && packed_base && l.has_xgroup && !no_xgroup && ((xgroup_sel >> (l.stride == 2 ? 0 : (l.cin == 8 ? 1 : 2))) & 1) && a.C1 == 0 &&
        a.C0 == l.cin && a.IW % l.xg_P == 0 && a.OW % (8 * l.xg_G) == 0 && !a.proj_w && !a.aux_add && !a.out_f32 && a.Cout == l.cout) {
      // x-grouped form (see Layer::has_xgroup)
      const int G = l.xg_G, P = l.xg_P, nq = l.xg_nq, Ng = (int)align_up(G * l.cout, 16);
      ConvArgs f = a;
      f.C0 = P * a.C0;
      f.IW = a.IW / P;
      f.taps.n = 0;
      f.use_kmask = 1;
      for (int ka = 0; ka < l.kd; ++ka)
        for (int kb = 0; kb < 3; ++kb)
          for (int q = 0; q < nq; ++q) {
            const int t = f.taps.n;
            f.taps.dz[t] = (int8_t)(ka - (l.kd - 1) / 2);
            f.taps.dy[t] = (int8_t)(kb - 1);
            f.taps.dx[t] = (int8_t)(q - 1);
            f.taps.widx[t] = (uint8_t)((ka * 3 + kb) * nq + q);
            // group -1 contributes its last pixel only (a group is 32 channels: K step 1), group +1 its first (K step 0)
            f.tap_kmask[t] = (uint8_t)(q == 0 ? 2 : (q == 2 ? 1 : 0));
            ++f.taps.n;
          }
      f.isy = l.stride; f.isx = 1;
      f.OW = a.OW / G; f.OWt = f.OW;
      f.Cout = G * l.cout;
      f.scale = (const float*)(packed_base + l.pk_ssxg);
      f.shift = f.scale + G * l.cout;
      if (conv_slab_supported(f, nullptr, 1, Ng)) return launch_conv_slab(f, nullptr, 1, packed_base + l.pk_wxg, Ng, nsm, st);
    }
    static const bool no_xpair = getenv("DFF_B200_NO_XPAIR") != nullptr;   // (A/B knob)
    static const int xpair_maxc = getenv("DFF_B200_XPAIR_MAXC") ? atoi(getenv("DFF_B200_XPAIR_MAXC")) : 8;   // (16-channel sources: measured neutral — half the rows against +33 % MMAs)
    if (wtc && use_fold && packed_base && l.has_xpair && !no_xpair && a.C1 == 0 && a.C0 == l.cin && l.cin <= xpair_maxc && a.IW % 2 == 0 &&
        a.Cout == out.C) {
      // x-paired form of a stride-2 layer (see Layer::has_xpair): source read as (.., W/2, 2C), stride 1 in x, taps on pairs -1 and 0
      ConvArgs f = a;
      f.C0 = 2 * a.C0;
      f.IW = a.IW / 2;
      f.taps.n = 0;
      for (int ka = 0; ka < l.kd; ++ka)
        for (int kb = 0; kb < 3; ++kb)
          for (int q = 0; q < 2; ++q) {
            f.taps.dz[f.taps.n] = (int8_t)(ka - (l.kd - 1) / 2);
            f.taps.dy[f.taps.n] = (int8_t)(kb - 1);
            f.taps.dx[f.taps.n] = (int8_t)(q - 1);
            f.taps.widx[f.taps.n] = (uint8_t)((ka * 3 + kb) * 2 + q);
            ++f.taps.n;
          }
      f.isy = 2; f.isx = 1;
      if (conv_slab_supported(f, nullptr, 1, l.Ntc)) return launch_conv_slab(f, nullptr, 1, packed_base + l.pk_wxp, l.Ntc, nsm, st);
    }
    static const bool no_yfold = getenv("DFF_B200_NO_YFOLD") != nullptr;   // (A/B knob)
    if (wtc && use_fold && packed_base && l.has_foldy && !no_yfold && !a.proj_w && !a.aux_add && !a.out_f32 && a.Cout == l.cout &&
        a.C0 == 8 && (a.C1 == 0 || a.C1 == 8) && out.H % l.gfold == 0) {
      // y-folded form: G vertically adjacent output pixels are the G*Cout channels of one GEMM row — tile rows G input rows apart through
      // the descriptor's row-group stride (as in the row-folded first layer), 8-channel output groups one row apart.  Unlike the
      // x-fold it needs no x-stride views, so the 8-channel sources are staged as merged rows (RX * 16 contiguous bytes per TMA row
      // instead of 16: the TMA unit retires ~1 row per clock, which bounded these layers).
      const int G = l.gfold, kp = l.kh + G - 1;
      ConvArgs f = a;
      f.taps.n = 0;
      for (int ka = 0; ka < l.kd; ++ka)
        for (int pq = 0; pq < kp; ++pq)
          for (int kc = 0; kc < l.kw; ++kc) {
            f.taps.dz[f.taps.n] = (int8_t)(ka - (l.kd - 1) / 2);
            f.taps.dy[f.taps.n] = (int8_t)(pq - 1);
            f.taps.dx[f.taps.n] = (int8_t)(kc - 1);
            f.taps.widx[f.taps.n] = (uint8_t)((ka * kp + pq) * l.kw + kc);
            ++f.taps.n;
          }
      f.isy = f.isx = 1; f.row_step = G; f.grp_rows = 1;
      f.osy = G; f.osx = 1; f.ooy = f.oox = 0;
      f.OHt = out.H / G; f.OWt = out.W;
      f.Cout = G * l.cout;
      f.scale = (const float*)(packed_base + l.pk_ssfold);
      f.shift = f.scale + G * l.cout;
      if (conv_slab_supported(f, nullptr, 1, G * l.cout))
        return launch_conv_slab(f, nullptr, 1, packed_base + l.pk_wfoldy, G * l.cout, nsm, st);
    }
    if (wtc && use_fold && packed_base && l.gfold > 1 && !a.proj_w && a.Cout == l.cout && a.OW % (8 * l.gfold) == 0) {
      // x-folded form: G adjacent output pixels are the G*Cout channels of one GEMM row; the input is read with x-stride G
      // (one staged view per residue), the output is the same memory viewed as (.., W/G, G*Cout)
      const int G = l.gfold, kq = l.kw + G - 1;
      ConvArgs f = a;
      f.taps.n = 0;
      for (int ka = 0; ka < l.kd; ++ka)
        for (int kb = 0; kb < l.kh; ++kb)
          for (int q = 0; q < kq; ++q) {
            f.taps.dz[f.taps.n] = (int8_t)(ka - (l.kd - 1) / 2);
            f.taps.dy[f.taps.n] = (int8_t)(kb - (l.kh - 1) / 2);
            f.taps.dx[f.taps.n] = (int8_t)(q - 1);
            f.taps.widx[f.taps.n] = (uint8_t)((ka * l.kh + kb) * kq + q);
            ++f.taps.n;
          }
      f.isy = 1; f.isx = G;
      f.OW = a.OW / G; f.OWt = f.OW;
      f.Cout = G * l.cout;
      f.scale = (const float*)(packed_base + l.pk_ssfold);
      f.shift = f.scale + G * l.cout;
      if (conv_slab_supported(f, nullptr, 1, G * l.cout))
        return launch_conv_slab(f, nullptr, 1, packed_base + l.pk_wfold, G * l.cout, nsm, st);
    }
    if (wtc) {
      if (wslab && use_row && conv_row_supported(a, l.Ntc)) return launch_conv_row(a, wslab, l.Ntc, nsm, st);
      if (wslab && conv_slab_supported(a, nullptr, 1, l.Ntc)) return launch_conv_slab(a, nullptr, 1, wslab, l.Ntc, nsm, st);
      DFF_TRY(ensure_tc_pack(lazy, st));
      return launch_conv_tc(a, wtc, l.pair_x ? 45 : l.ntaps, l.Ntc, nsm, st);
    }
    return launch_conv_ffma(a, bf16, st);
  }
  if (wtc && wslab && use_fold && packed_base && l.gfold == 2 && a.Cout == l.cout) {
    // x-folded transposed conv: two row phases, each GEMM row = the two adjacent output pixels (2j, 2j+1) as 2*Cout channels
    TapTable pt[2];
    for (int py = 0; py < 2; ++py) {
      TapTable& t = pt[py];
      t.n = 0;
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh) {
          if ((kh == 1) != (py == 0)) continue;   // row phase 0 uses kh = 1, row phase 1 uses kh = 0 (offset +1) and 2 (offset 0)
          for (int dxq = 0; dxq < 2; ++dxq) {
            t.dz[t.n] = (int8_t)(1 - kd);
            t.dy[t.n] = (int8_t)(kh == 0 ? 1 : 0);
            t.dx[t.n] = (int8_t)dxq;
            t.widx[t.n] = (uint8_t)((kd * 3 + kh) * 2 + dxq);
            ++t.n;
          }
        }
    }
    ConvArgs f = a;
    f.taps = pt[1];
    f.isy = f.isx = 1; f.osy = 2; f.osx = 1; f.ooy = f.oox = 0;
    f.OHt = in.H; f.OWt = in.W;
    f.OW = in.W;                 // output viewed as (.., 2*IH, IW, 2*Cout)
    f.Cout = 2 * l.cout;
    f.scale = (const float*)(packed_base + l.pk_ssfold);
    f.shift = f.scale + 2 * l.cout;
    f.proj_c = a.proj_w ? l.cout : 0;
    f.no_wstream = a.proj_w ? 0 : 1;   // (measured: with streamed weights the fold only pays when it also halves the classifier epilogue)
    if (conv_slab_supported(f, pt, 2, 2 * l.cout)) {
      *nlaunch = 1;
      return count_only ? 0 : launch_conv_slab(f, pt, 2, packed_base + l.pk_wfold, 2 * l.cout, nsm, st);
    }
  }
  if (wtc && wslab) {  // all four output-parity phases in one launch: the input planes are staged once
    TapTable pt[4];
    for (int ph = 0; ph < 4; ++ph) deconv_taps(ph >> 1, ph & 1, pt[ph]);
    a.taps = pt[3];
    a.isy = a.isx = 1; a.osy = a.osx = 2; a.ooy = a.oox = 0;
    a.OHt = in.H; a.OWt = in.W;
    if (conv_slab_supported(a, pt, 4, l.Ntc)) {
      *nlaunch = 1;
      return count_only ? 0 : launch_conv_slab(a, pt, 4, wslab, l.Ntc, nsm, st);
    }
  }
  *nlaunch = 4;
  if (count_only) return 0;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      deconv_taps(py, px, a.taps);
      a.isy = a.isx = 1; a.osy = a.osx = 2; a.ooy = py; a.oox = px;
      a.OHt = in.H; a.OWt = in.W;
      if (wtc) {
        if (wslab && conv_slab_supported(a, nullptr, 1, l.Ntc)) DFF_TRY(launch_conv_slab(a, nullptr, 1, wslab, l.Ntc, nsm, st));
        else {
          DFF_TRY(ensure_tc_pack(lazy, st));
          DFF_TRY(launch_conv_tc(a, wtc, l.ntaps, l.Ntc, nsm, st));
        }
      } else {
        DFF_TRY(launch_conv_ffma(a, bf16, st));
      }
    }
  return 0;
}

// ---- optional per-operator profiling (CUDA events on the launching stream) --------------------------------------
struct OpRecord {
  std::string name;
  double flops = 0, bytes = 0;  // algorithmic: 2*MACs without padding/zero taps; compulsory activation+weight bytes
  int launches = 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};
struct Profile {
  std::vector<OpRecord> ops;
  bool events = false;  // record CUDA events (false: only count launches / flops / bytes)
};

// ---- the forward schedule -----------------------------------------------------------------------------------
struct Runner {
  const Net& net;
  const char* packed;
  char* ws;
  size_t ws_bytes, off = 0;
  bool dry, bf16;
  cudaStream_t st;
  int rc = 0;
  Profile* prof = nullptr;
  bool use_tc = false, use_slab = true;
  bool use_fold = getenv("DFF_B200_NO_FOLD") == nullptr;   // x-folded small-Cout layers (A/B switch for measurements)

  void op_begin(const std::string& name, double flops, double bytes, int launches) {
    if (!prof) return;
    OpRecord r;
    r.name = name; r.flops = flops; r.bytes = bytes; r.launches = launches;
    if (prof->events && !dry && !rc) {
      cudaEventCreate(&r.e0);
      cudaEventCreate(&r.e1);
      cudaEventRecord(r.e0, st);
    }
    prof->ops.push_back(r);
  }
  void op_end() {
    if (prof && prof->events && !dry && !rc && prof->ops.back().e1) cudaEventRecord(prof->ops.back().e1, st);
  }

  size_t esize(bool f32) const { return (bf16 && !f32) ? 2 : 4; }
  Ten alloc(int B, int S, int H, int W, int C, bool f32 = false) {
    Ten t;
    t.B = B; t.S = S; t.H = H; t.W = W; t.C = C; t.f32 = f32;
    const size_t bytes = align_up((size_t)B * S * H * W * C * esize(f32), 256);
    if (!dry) {
      if (off + bytes > ws_bytes) {
        if (!rc) rc = fail(DFF_E_WORKSPACE, "workspace too small");
      } else {
        t.p = ws + off;
      }
    }
    off += bytes;
    return t;
  }
  Ten conv(const std::string& name, const Ten& in, EpiOpt e = EpiOpt()) {
    const Layer& l = net.layers[net.index.at(name)];
    Ten out;
    if (l.transposed) out = alloc(in.B, in.S, in.H * 2, in.W * 2, e.out_f32 ? l.cout : l.CoutP, e.out_f32);
    else out = alloc(in.B, in.S, in.H / l.stride, (l.pair_x && use_tc ? in.W - 2 : in.W) / l.stride, e.out_f32 ? l.cout : l.CoutP, e.out_f32);
    Ten aux;
    if (e.aux_add) {
      aux = alloc(out.B, out.S, out.H, out.W, out.C);
      aux_last = aux;
      e.aux_out = &aux;
    }
    Ten pj;
    if (e.proj) {
      pj = alloc(out.B, out.S, out.H, out.W, 1, true);
      proj_last = pj;
      e.proj_out = &pj;
    }
    {
      // algorithmic work: every output voxel sees cin*cout*taps MACs for a conv; a transposed conv spends
      // cin*cout*27 MACs per INPUT voxel (= 27/4 per output voxel).  Bytes: in + out + residuals + weights, once.
      const double ovox = (double)out.B * out.S * out.H * out.W, ivox = (double)in.B * in.S * in.H * in.W;
      const double macs = (l.transposed ? ivox : ovox) * l.cin * l.cout * l.ntaps;
      double bytes = ivox * l.CinP * esize(false) + ovox * out.C * esize(out.f32) + (double)l.ntaps * l.cin * l.cout * 4;
      if (e.res_pre) bytes += ovox * out.C * esize(false);
      if (e.res_post) bytes += ovox * out.C * esize(false);
      if (e.aux_add) bytes += 2 * ovox * out.C * esize(false);
      double flops = 2.0 * macs;
      if (e.proj) { flops += 2.0 * ovox * e.proj->cin; bytes += ovox * 4; }
      if (e.skip_out) bytes -= ovox * out.C * esize(out.f32);
      op_begin(name, flops, bytes, l.transposed ? 4 : 1);
    }
    if (!rc) {
      // in a dry run only the launch count is planned (pointers are placeholders that are never dereferenced)
      const char* pk = dry ? reinterpret_cast<const char*>(0x1000) : packed;
      int nl = 0;
      rc = run_conv(l, (const float*)(pk + l.pk_w), (const float*)(pk + l.pk_scale), (const float*)(pk + l.pk_shift), in, e, out,
                    bf16, st, use_tc ? pk + l.pk_wtc : nullptr, (use_tc && use_slab) ? pk + l.pk_wslab : nullptr, &nl, dry, true, pk,
                    use_tc && use_slab && use_fold);
      if (prof && !prof->ops.empty()) prof->ops.back().launches = nl;
    }
    op_end();
    return out;
  }
  Ten aux_last, proj_last;
  void* alloc_bytes(size_t bytes) {
    bytes = align_up(bytes, 256);
    void* p = nullptr;
    if (!dry) {
      if (off + bytes > ws_bytes) {
        if (!rc) rc = fail(DFF_E_WORKSPACE, "workspace too small");
      } else {
        p = ws + off;
      }
    }
    off += bytes;
    return p;
  }
  Ten pool(const Ten& in, int k, bool is_max) {
    Ten out = alloc(in.B, in.S, in.H / k, in.W / k, in.C);
    const double ivox = (double)in.B * in.S * in.H * in.W;
    op_begin(is_max ? "maxpool" : "avgpool", 0, ivox * in.C * esize(false) * (1.0 + 1.0 / (k * k)), 1);
    if (!dry && !rc) rc = launch_pool(in.p, out.p, in.B * in.S, in.H, in.W, in.C, k, is_max, bf16, st);
    op_end();
    return out;
  }
  static EpiOpt relu() {
    EpiOpt e;
    e.relu = true;
    return e;
  }
  // SRD / Feature_Extraction (reference :394-407)
  Ten srd(const std::string& p, const Ten& x) {
    Ten t = conv(p + ".Focus_Measure.conv.0.0", x, relu());
    EpiOpt e = relu();
    e.res_pre = &x;
    Ten f = conv(p + ".Focus_Measure.conv.2.0", t, e);
    const Layer& l0 = net.layers[net.index.at(p + ".N_ch_attention.0")];
    const Layer& l1 = net.layers[net.index.at(p + ".N_ch_attention.2")];
    if (use_tc && (f.C == 8 || f.C == 16 || f.C == 32) && l0.CinP == f.C && l0.CoutP == f.C && (f.H * f.W) % 16 == 0) {
      // both attention convolutions, both ReLUs and the residual in one bandwidth pass (the intermediate stays in registers):
      // warp-level tensor-core kernel (attention.cu); DFF_B200_ATTN=ffma selects the scalar C = 8 kernel, =mma2 the two slab launches
      static const char* mode = getenv("DFF_B200_ATTN");
      const bool scalar = mode && !strcmp(mode, "ffma") && f.C == 8;
      if (!(mode && !strcmp(mode, "mma2"))) {
        Ten o = alloc(f.B, f.S, f.H, f.W, f.C);
        const double vox = (double)f.B * f.S * f.H * f.W;
        op_begin(p + ".N_ch_attention(fused)", 2.0 * vox * f.C * f.C * 4, 2.0 * vox * f.C * esize(false) + 4.0 * f.C * f.C * 4, 1);
        if (!dry && !rc) {
          const float* w0 = (const float*)(packed + l0.pk_w);
          const float* w1 = (const float*)(packed + l1.pk_w);
          rc = scalar ? launch_srd_attention(f.p, w0, w1, o.p, f.B, f.S, f.H, f.W, f.C, st)
                      : launch_srd_attention_mma(f.p, w0, w1, o.p, f.B, f.S, f.H, f.W, f.C, num_sms_of_current_device(), st);
        }
        op_end();
        return o;
      }
    }
    Ten a = conv(p + ".N_ch_attention.0", f, relu());
    EpiOpt e2 = relu();
    e2.res_post = &f;
    return conv(p + ".N_ch_attention.2", a, e2);
  }
  // EFD / res_stride_conv_3d (reference :383-392)
  Ten efd(const std::string& p, const Ten& x) {
    Ten a = conv(p + ".stride_conv.0", x);
    Ten mp = pool(x, 2, true);
    EpiOpt e = relu();
    e.res_pre = &a;
    return conv(p + ".max_pooling.1.0", mp, e);
  }
  Ten tower(const std::string& a, const std::string& b, const Ten& x) {
    Ten r = conv(a + ".2.0", conv(a + ".0.0", x, relu()), relu());
    EpiOpt e;
    e.res_pre = &r;  // conv -> BN -> + residual, no ReLU (reference :252,255,258)
    return conv(b + ".2.0", conv(b + ".0.0", r, relu()), e);
  }
  // hourglassup (reference :247-273)
  Ten pyramid(const Ten& v3) {
    const std::string sp = "SPP_module.";
    Ten x8, x16, x32;
    static const bool no_pyr = getenv("DFF_B200_NO_PYRAMID_POOL") != nullptr;
    if (bf16 && v3.C % 8 == 0 && !no_pyr) {   // the three pools in one pass over V3
      x8 = alloc(v3.B, v3.S, v3.H / 2, v3.W / 2, v3.C);
      x16 = alloc(v3.B, v3.S, v3.H / 4, v3.W / 4, v3.C);
      x32 = alloc(v3.B, v3.S, v3.H / 8, v3.W / 8, v3.C);
      const double ivox = (double)v3.B * v3.S * v3.H * v3.W;
      op_begin("avgpool_pyramid", 0, ivox * v3.C * 2.0 * (1.0 + 1.0 / 4 + 1.0 / 16 + 1.0 / 64), 1);
      if (!dry && !rc) rc = launch_avgpool_pyramid(v3.p, x8.p, x16.p, x32.p, v3.B * v3.S, v3.H, v3.W, v3.C, st);
      op_end();
    } else {
      x8 = pool(v3, 2, false); x16 = pool(v3, 4, false); x32 = pool(v3, 8, false);
    }
    x8 = tower(sp + "dres8_0", sp + "dres8_1", x8);
    x16 = tower(sp + "dres16_0", sp + "dres16_1", x16);
    x32 = tower(sp + "dres32_0", sp + "dres32_1", x32);
    Ten c1 = conv(sp + "conv1", x8);
    EpiOpt e = relu();
    e.in1 = &x16;
    c1 = conv(sp + "combine1.0.0", c1, e);
    Ten c2 = conv(sp + "conv2.0.0", c1, relu());
    Ten c3 = conv(sp + "conv3", c2);
    EpiOpt e3 = relu();
    e3.in1 = &x32;
    c3 = conv(sp + "combine2.0.0", c3, e3);
    Ten c4 = conv(sp + "conv4.0.0", c3, relu());
    Ten r2 = conv(sp + "redir2.0", c2);
    EpiOpt e8 = relu();
    e8.res_pre = &r2;
    Ten c8 = conv(sp + "conv8.0", c4, e8);
    Ten r1 = conv(sp + "redir1.0", x8);
    EpiOpt e9 = relu();
    e9.res_pre = &r1;
    return conv(sp + "conv9.0", c8, e9);
  }
  // hourglass (reference :302-321).  Returns `out`; pre_1 through *pre1; out_in = skip + out through *out_in.
  // `classif` (tensor-core path only): the 1x1x1 classifier applied to out_in is fused into conv6's epilogue; *cost gets it.
  Ten hourglass(const std::string& p, const Ten& x, const Ten& skip_feat, const Ten* presqu, const Ten* postsqu, Ten* pre1,
                Ten* out_in, bool need_out, const char* classif = nullptr, Ten* cost = nullptr) {
    EpiOpt e0 = relu();
    e0.in1 = &skip_feat;
    *pre1 = conv(p + ".conv0.0.0", x, e0);
    Ten o = conv(p + ".conv1.0.0", *pre1, relu());
    EpiOpt e2 = relu();
    e2.res_pre = postsqu;
    Ten pre = conv(p + ".conv2.0", o, e2);
    o = conv(p + ".conv3.0.0", pre, relu());
    o = conv(p + ".conv4.0.0", o, relu());
    EpiOpt e5 = relu();
    e5.res_pre = presqu ? presqu : &pre;
    o = conv(p + ".conv5.0", o, e5);
    EpiOpt e6;
    const bool fuse = classif && use_tc;
    if (fuse) e6.proj = &net.layers[net.index.at(classif)];
    if (need_out) {
      e6.aux_add = &x;
      e6.proj_aux = true;
      Ten out = conv(p + ".conv6.0", o, e6);
      *out_in = aux_last;
      if (fuse) *cost = proj_last;
      return out;
    }
    e6.res_post = &x;  // last stage: only out2 + out is needed (reference :115) — and with the classifier fused, not even that
    e6.skip_out = fuse;
    *out_in = conv(p + ".conv6.0", o, e6);
    if (fuse) *cost = proj_last;
    return *out_in;
  }
};

// The focal stacks of a call: the reference's fp32 (B,3,S,H,W) tensor, or the datasets' uint8 (B,S,H0,W0,3) stacks (H0 <= H, W0 <= W;
// normalisation and the -1 padding happen in the staging kernel).
struct FwdIn {
  const float* FS = nullptr;
  const unsigned char* u8 = nullptr;
  int H0 = 0, W0 = 0;
  FwdIn() = default;
  FwdIn(const float* f) : FS(f) {}
  FwdIn(const unsigned char* u, int h0, int w0) : u8(u), H0(h0), W0(w0) {}
};

static int forward_impl(const void* packed, const FwdIn& fin, const float* fd, const int64_t* fds, int B, int S, int H, int W,
                        float* const* out4, float* const* cost4, void* ws, size_t ws_bytes, int mode, cudaStream_t st,
                        bool dry, size_t* need, Profile* prof = nullptr) {
  if (B < 1 || S < 1 || H < 32 || W < 32 || H % 32 || W % 32)
    return fail(DFF_E_ARG, "dff_forward: need B,S >= 1 and H,W positive multiples of 32 (pad with -1 like the reference dataloaders)");
  if (mode & DFF_TRAIN) return fail(DFF_E_UNSUPPORTED, "dff_forward: DFF_TRAIN is not available in this build");
  if (fin.u8 && (fin.H0 < 1 || fin.W0 < 1 || fin.H0 > H || fin.W0 > W))
    return fail(DFF_E_ARG, "dff_forward_u8: need 1 <= H0 <= H and 1 <= W0 <= W (H, W = the padded extent)");
  g_pdl_call = (double)B * S * H * W <= 8.0 * 10 * 384 * 576;   // up to 8 DDFF stacks' worth of voxels per call
  Runner r{net_of(DFF_NET_DFF), (const char*)packed, (char*)ws, ws_bytes, 0, dry, (mode & DFF_BF16) != 0, st};
  r.prof = prof;
  r.use_tc = r.bf16 && !(mode & DFF_NO_TC);
  r.use_slab = !(mode & DFF_NO_SLAB);
  const double vox = (double)B * S * H * W;
  const int c_in = r.use_tc ? 8 : 4;  // stored channels of the converted focal stack (TMA needs 16-byte pixels)
  Ten x0 = r.alloc(B, S, H, r.use_tc ? W + 2 : W, c_in);   // (tensor-core path: pair-packed first-layer input, see to_cl_pair_kernel)
  if (fin.u8 || (dry && fin.H0)) {
    r.op_begin("to_channels_last(u8)", 0, 3.0 * B * S * fin.H0 * fin.W0 + vox * c_in * r.esize(false), 1);
    if (!dry && !r.rc)
      r.rc = r.use_tc ? launch_u8_to_cl_pair(fin.u8, B, S, fin.H0, fin.W0, H, W, x0.p, st)
                      : launch_u8_to_cl(fin.u8, B, S, fin.H0, fin.W0, H, W, x0.p, c_in, r.bf16, st);
  } else {
    r.op_begin("to_channels_last", 0, vox * (12 + c_in * r.esize(false)), 1);
    if (!dry && !r.rc) r.rc = r.use_tc ? launch_to_cl_pair(fin.FS, B, S, H, W, x0.p, st) : launch_to_cl(fin.FS, B, 3, S, H, W, x0.p, c_in, r.bf16, st);
  }
  r.op_end();
  Ten t = r.conv("FM_measure.Focus_extraction.0.0", x0, Runner::relu());
  Ten v1 = r.srd("FM_measure.Focus_extraction.2", t);
  Ten v2 = r.srd("FM_conv1.1", r.efd("FM_conv1.0", v1));
  Ten v3 = r.srd("FM_conv2.1", r.efd("FM_conv2.0", v2));
  Ten vol = r.pyramid(v3);

  EpiOpt ec;
  ec.out_f32 = true;
  Ten cm = r.conv("confidence.2", r.conv("confidence.0.0", vol, Runner::relu()), ec);

  Ten x = r.conv("dres0.2.0", r.conv("dres0.0.0", vol, Runner::relu()), Runner::relu());
  x = r.conv("deconv_1.0", x);
  Ten pre, out_in, pre2, pre3;
  Ten cost1, cost2, cost3;
  Ten out = r.hourglass("dres2", x, v3, nullptr, nullptr, &pre, &out_in, true, "classif1.0", &cost1);
  if (!r.use_tc) cost1 = r.conv("classif1.0", out_in, ec);
  Ten o2 = r.conv("deconv_2.0", out_in);
  Ten out_in2;
  Ten outb = r.hourglass("dres3", o2, v2, &pre, &out, &pre2, &out_in2, true, "classif2.0", &cost2);
  if (!r.use_tc) cost2 = r.conv("classif2.0", out_in2, ec);
  Ten o3 = r.conv("deconv_3.0", out_in2);
  Ten out_in3;
  r.hourglass("dres4", o3, v1, &pre2, &outb, &pre3, &out_in3, false, "classif3.0", &cost3);
  if (!r.use_tc) cost3 = r.conv("classif3.0", out_in3, ec);
  if (need) *need = r.off;
  const Ten* costs[4] = {&cm, &cost1, &cost2, &cost3};
  // the four depth heads in one launch (focus_dists is read once): 4 cost volumes + S*H*W focus distances in, 4 maps out
  double hbytes = 4.0 * vox + 4.0 * 4.0 * B * H * W;
  for (int i = 0; i < 4; ++i) hbytes += 4.0 * B * S * costs[i]->H * costs[i]->W;
  if (dry) {
    r.op_begin("depth_heads", 0, hbytes, 1);
    return 0;
  }
  if (r.rc) return r.rc;
  r.op_begin("depth_heads", 0, hbytes, 1);
  const float* cp[4];
  int ch[4], cw[4];
  for (int i = 0; i < 4; ++i) { cp[i] = (const float*)costs[i]->p; ch[i] = costs[i]->H; cw[i] = costs[i]->W; }
  DFF_TRY(launch_depth_head4(cp, ch, cw, fd, fds, B, S, H, W, out4, r.bf16, st));
  r.op_end();
  for (int i = 0; i < 4; ++i)
    if (cost4 && cost4[i])
      DFF_CUDA(cudaMemcpyAsync(cost4[i], costs[i]->p, (size_t)B * S * costs[i]->H * costs[i]->W * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
  return 0;
}

// FlowNetwork.forward (reference End_to_End/End_to_End.py:63-104): three feature scales, coarse-to-fine estimation of the per-slice
// similarity warp (alpha: scale correction x 0.001, x shift, y shift), final warp of the focal stack.  FS_out (B,3,S,H,W) fp32.
static int flow_forward_impl(const void* packed, const float* FS, const float* fov, int B, int S, int H, int W, float* FS_out,
                             float* alpha_out, void* ws, size_t ws_bytes, int mode, cudaStream_t st, bool dry, size_t* need) {
  if (B < 1 || S < 1 || H < 4 || W < 4 || H % 4 || W % 4)
    return fail(DFF_E_ARG, "dff_flow_forward: need B,S >= 1 and H,W positive multiples of 4");
  g_pdl_call = (double)B * S * H * W <= 8.0 * 10 * 384 * 576;
  Runner r{net_of(DFF_NET_FLOW), (const char*)packed, (char*)ws, ws_bytes, 0, dry, (mode & DFF_BF16) != 0, st};
  r.use_tc = r.bf16 && !(mode & DFF_NO_TC);
  r.use_slab = !(mode & DFF_NO_SLAB);
  const int c_in = r.use_tc ? 8 : 4;
  Ten x0 = r.alloc(B, S, H, W, c_in);
  if (!dry && !r.rc) r.rc = launch_to_cl(FS, B, 3, S, H, W, x0.p, c_in, r.bf16, st);
  auto block = [&](const std::string& p, const Ten& x) {
    Ten t = r.conv(p + ".conv.0.0", x, Runner::relu());
    Ten f = r.conv(p + ".feature", x);
    EpiOpt e = Runner::relu();
    e.res_pre = &f;
    return r.conv(p + ".conv.2.0", t, e);
  };
  Ten fe1 = block("OF_feature.1", block("OF_feature.0", x0));
  Ten fe2 = block("OF_feature1.1", block("OF_feature1.0", fe1));
  Ten fe3 = block("OF_feature2.1", block("OF_feature2.0", fe2));
  const float* alpha = nullptr;
  const char* heads[3] = {"conv1", "conv2", "conv3"};
  const Ten* feats[3] = {&fe3, &fe2, &fe1};
  for (int k = 0; k < 3; ++k) {
    const Ten& f = *feats[k];
    const std::string h = heads[k];
    // (measured and rejected: warp + pair volume as ONE kernel with a 16-byte output piece per thread — 267 us against 59 + 133 us at
    // 512x768: three warp geometries per pixel instead of two, and the warp is issue-bound; profiles/r2_changes_measured.txt)
    Ten warped = r.alloc(f.B, f.S, f.H, f.W, f.C);
    Ten vol = r.alloc(f.B, f.S, f.H, f.W, 2 * f.C + 16);
    if (!dry && !r.rc) r.rc = launch_fov_warp_cl(f.p, alpha, fov, B, f.C, S, f.H, f.W, warped.p, r.bf16, st);
    if (!dry && !r.rc) r.rc = launch_pair_volume(warped.p, alpha, fov, B, f.C, S, f.H, f.W, vol.p, r.bf16, st, 2 * f.C + 16);
    Ten t = r.conv(h + ".0.0", vol, Runner::relu());
    t = r.conv(h + ".2.0", t, Runner::relu());
    t = r.conv(h + ".4.0", t, Runner::relu());
    EpiOpt eo;
    eo.out_f32 = true;
    Ten o = r.conv(h + ".6", t, eo);                     // (B,S,h,w,3) fp32, bias in the epilogue
    float* na = (float*)r.alloc_bytes((size_t)B * 3 * S * sizeof(float));
    // AdaptiveAvgPool3d((S,1,1)) + the 0.001 factor on the scale term + the running sum (reference :78-79, 88-90, 99-101)
    void* msc = r.alloc_bytes(spatial_mean_scratch_bytes(B, S, f.H, f.W));
    if (!dry && !r.rc) r.rc = launch_spatial_mean_accum((const float*)o.p, 3, B, S, f.H, f.W, alpha, 0.001f, 1.f, 1.f, na, msc, st);
    alpha = na;
  }
  if (need) *need = r.off;
  if (dry) return 0;
  if (r.rc) return r.rc;
  DFF_TRY(launch_fov_warp(FS, alpha, fov, B, 3, S, H, W, FS_out, nullptr, st));
  if (alpha_out) DFF_CUDA(cudaMemcpyAsync(alpha_out, alpha, (size_t)B * 3 * S * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

struct DeviceGuard {
  int prev = -1;
  int rc = 0;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (dev != prev) rc = check_cuda(cudaSetDevice(dev), "cudaSetDevice");
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};


// every kernel layout of one layer's weights (`w`: the reference layout) at its offsets inside `pk`
static int pack_layer_weights(const Layer& l, const float* w, char* pk, cudaStream_t st) {
  DFF_TRY(launch_pack_weight(w, (float*)(pk + l.pk_w), l.cout, l.cin, l.ntaps, l.CinP, l.CoutP, l.transposed ? 1 : 0, st));
  if (l.pair_x) {   // tensor-core packs of the paired-tap form: an ordinary (Cout, 8, 1, 9, 5) convolution weight
    DFF_TRY(launch_pair_weight(w, (float*)(pk + l.pk_pair), l.cout, st));
    DFF_TRY(launch_pack_weight_tc((const float*)(pk + l.pk_pair), pk + l.pk_wtc, l.cout, 8, 8, 45, l.Ntc, 0, st));
    DFF_TRY(launch_pack_weight_slab((const float*)(pk + l.pk_pair), pk + l.pk_wslab, l.cout, 8, 8, 45, l.Ntc, 0, st));
    if (l.pk_wfold != l.pk_ssfold) DFF_TRY(launch_pack_weight_slab_rowfold((const float*)(pk + l.pk_pair), pk + l.pk_wfold, l.cout, st));
  } else {
    DFF_TRY(launch_pack_weight_tc(w, pk + l.pk_wtc, l.cout, l.cin, l.CinT, l.ntaps, l.Ntc, l.transposed ? 1 : 0, st));
    DFF_TRY(launch_pack_weight_slab(w, pk + l.pk_wslab, l.cout, l.cin, l.CinT, l.ntaps, l.Ntc, l.transposed ? 1 : 0, st));
  }
  if (l.cout == 1 && l.ntaps == 1) DFF_TRY(launch_pack_weight(w, (float*)(pk + l.pk_proj), 1, l.cin, 1, l.CinT, 1, 0, st));
  if (l.gfold > 1) {
    if (l.transposed) DFF_TRY(launch_pack_weight_slab_deconv_fold(w, pk + l.pk_wfold, l.cout, l.cin, l.CinT, st));
    else DFF_TRY(launch_pack_weight_slab_fold(w, pk + l.pk_wfold, l.cout, l.cin, l.CinT, l.kd, l.kh, l.kw, l.gfold, st));
    if (l.has_foldy) DFF_TRY(launch_pack_weight_slab_fold(w, pk + l.pk_wfoldy, l.cout, l.cin, l.CinT, l.kd, l.kh, l.kw, l.gfold, st, 1));
  }
  if (l.has_wz) DFF_TRY(launch_pack_weight_slab_zmerge(w, pk + l.pk_wz, l.cout, l.cin, l.CinT, l.Ntc, st));
  if (l.has_xgroup) {
    const int Ng = (int)align_up(l.xg_G * l.cout, 16);
    DFF_TRY(launch_xgroup_weight(w, (float*)(pk + l.pk_xgw), l.cout, l.cin, l.kd, l.stride, l.xg_G, l.xg_P, l.xg_nq, st));
    DFF_TRY(launch_pack_weight_slab((const float*)(pk + l.pk_xgw), pk + l.pk_wxg, l.xg_G * l.cout, l.xg_P * l.cin, l.xg_P * l.cin,
                                    l.kd * 3 * l.xg_nq, Ng, 0, st));
  }
  if (l.has_xpair) {
    DFF_TRY(launch_xpair_weight(w, (float*)(pk + l.pk_xpw), l.cout, l.cin, l.kd, st));
    DFF_TRY(launch_pack_weight_slab((const float*)(pk + l.pk_xpw), pk + l.pk_wxp, l.cout, 2 * l.cin, 2 * l.cin, l.kd * 6, l.Ntc, 0, st));
  }
  return 0;
}
// scale/shift of the folded forms (replicated per pixel of the GEMM row) from the layer's scale/shift in `pk`
static int pack_layer_folded_ss(const Layer& l, char* pk, cudaStream_t st) {
  const int G = (l.pair_x && l.pk_wfold != l.pk_ssfold) ? 4 : l.gfold;
  if (G > 1)
    DFF_TRY(launch_replicate_ss((const float*)(pk + l.pk_scale), (const float*)(pk + l.pk_shift), (float*)(pk + l.pk_ssfold), l.cout, G, st));
  if (l.has_xgroup)
    DFF_TRY(launch_replicate_ss((const float*)(pk + l.pk_scale), (const float*)(pk + l.pk_shift), (float*)(pk + l.pk_ssxg), l.cout, l.xg_G, st));
  return 0;
}

}  // namespace dff

using namespace dff;

// =============================================================================================================
// C-ABI
// =============================================================================================================
extern "C" {

int dff_abi_version(void) { return DFF_ABI_VERSION; }
const char* dff_last_error(void) { return g_err.c_str(); }

int dff_check_device(int device) {
  cudaDeviceProp p;
  DFF_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return fail(DFF_E_DEVICE, std::string("dff_b200 needs an sm_100 (B200) device, found sm_") +
                                                   std::to_string(p.major) + std::to_string(p.minor));
  return 0;
}

int dff_param_count(int net) { return (int)net_of(net).params.size(); }
const char* dff_param_name(int net, int i) {
  const Net& n = net_of(net);
  return (i < 0 || i >= (int)n.params.size()) ? nullptr : n.params[i].name.c_str();
}
int64_t dff_param_numel(int net, int i) {
  const Net& n = net_of(net);
  return (i < 0 || i >= (int)n.params.size()) ? -1 : n.params[i].numel;
}
int64_t dff_param_offset(int net, int i) {
  const Net& n = net_of(net);
  return (i < 0 || i >= (int)n.params.size()) ? -1 : n.params[i].offset;
}
int64_t dff_raw_numel(int net) { return net_of(net).raw_numel; }
size_t dff_packed_bytes(int net) { return net_of(net).packed_bytes; }

int dff_pack_weights(int net, const float* raw, void* packed, int device, void* stream) {
  if (!raw || !packed) return fail(DFF_E_ARG, "dff_pack_weights: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  cudaStream_t st = (cudaStream_t)stream;
  const Net& n = net_of(net);
  char* pk = (char*)packed;
  for (const Layer& l : n.layers) {
    DFF_TRY(pack_layer_weights(l, raw + l.raw_w, pk, st));
    const bool bn = l.raw_gamma >= 0;
    DFF_TRY(launch_bn_fold(bn ? raw + l.raw_gamma : nullptr, bn ? raw + l.raw_beta : nullptr, bn ? raw + l.raw_mean : nullptr,
                           bn ? raw + l.raw_var : nullptr, l.raw_bias >= 0 ? raw + l.raw_bias : nullptr,
                           (float*)(pk + l.pk_scale), (float*)(pk + l.pk_shift), l.cout, l.CoutP, st));
    DFF_TRY(pack_layer_folded_ss(l, pk, st));
  }
  return 0;
}

size_t dff_workspace_bytes(int B, int S, int H, int W, int mode) {
  size_t need = 0;
  if (forward_impl(nullptr, FwdIn(), nullptr, nullptr, B, S, H, W, nullptr, nullptr, nullptr, 0, mode, nullptr, true, &need))
    return 0;
  return need;
}

int dff_forward(const void* packed, const float* FS, const float* fd, const int64_t fd_strides[4], int B, int S, int H,
                int W, float* const out4[4], float* const cost4[4], void* workspace, size_t workspace_bytes, int mode,
                int device, void* stream) {
  if (!packed || !FS || !fd || !fd_strides || !out4 || !workspace) return fail(DFF_E_ARG, "dff_forward: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return forward_impl(packed, FwdIn(FS), fd, fd_strides, B, S, H, W, out4, cost4, workspace, workspace_bytes, mode,
                      (cudaStream_t)stream, false, nullptr);
}

size_t dff_flow_workspace_bytes(int B, int S, int H, int W, int mode) {
  size_t need = 0;
  if (flow_forward_impl(nullptr, nullptr, nullptr, B, S, H, W, nullptr, nullptr, nullptr, 0, mode, nullptr, true, &need)) return 0;
  return need;
}

int dff_flow_forward(const void* packed_flow, const float* FS, const float* fov, int B, int S, int H, int W, float* FS_out,
                     float* alpha_out, void* workspace, size_t workspace_bytes, int mode, int device, void* stream) {
  if (!packed_flow || !FS || !fov || !FS_out || !workspace) return fail(DFF_E_ARG, "dff_flow_forward: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return flow_forward_impl(packed_flow, FS, fov, B, S, H, W, FS_out, alpha_out, workspace, workspace_bytes, mode,
                           (cudaStream_t)stream, false, nullptr);
}

int dff_forward_u8(const void* packed, const uint8_t* FS_u8, int H0, int W0, const float* fd, const int64_t fd_strides[4], int B,
                   int S, int H, int W, float* const out4[4], float* const cost4[4], void* workspace, size_t workspace_bytes,
                   int mode, int device, void* stream) {
  if (!packed || !FS_u8 || !fd || !fd_strides || !out4 || !workspace) return fail(DFF_E_ARG, "dff_forward_u8: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return forward_impl(packed, FwdIn(FS_u8, H0, W0), fd, fd_strides, B, S, H, W, out4, cost4, workspace, workspace_bytes, mode,
                      (cudaStream_t)stream, false, nullptr);
}

int dff_stage_u8(const uint8_t* FS_u8, int H0, int W0, int B, int S, int H, int W, float* FS, int device, void* stream) {
  if (!FS_u8 || !FS) return fail(DFF_E_ARG, "dff_stage_u8: null pointer");
  if (H0 < 1 || W0 < 1 || H0 > H || W0 > W) return fail(DFF_E_ARG, "dff_stage_u8: need 1 <= H0 <= H and 1 <= W0 <= W");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_u8_to_planar(FS_u8, B, S, H0, W0, H, W, FS, (cudaStream_t)stream);
}

int dff_forward_profiled(const void* packed, const float* FS, const float* fd, const int64_t fd_strides[4], int B, int S,
                         int H, int W, float* const out4[4], void* workspace, size_t workspace_bytes, int mode, int device,
                         void* stream, int max_ops, float* op_ms_host, double* op_flops_host, double* op_bytes_host,
                         int* op_launches_host, char* op_names_host, int* n_ops) {
  if (!n_ops) return fail(DFF_E_ARG, "dff_forward_profiled: null pointer");
  Profile prof;
  const bool dry = packed == nullptr;  // plan only: operator list, flops, bytes, launches (no GPU work)
  prof.events = !dry;
  int rc = 0;
  if (dry) {
    rc = forward_impl(nullptr, FwdIn(), nullptr, nullptr, B, S, H, W, nullptr, nullptr, nullptr, 0, mode, nullptr, true,
                      nullptr, &prof);
  } else {
    if (!FS || !fd || !fd_strides || !out4 || !workspace) return fail(DFF_E_ARG, "dff_forward_profiled: null pointer");
    DeviceGuard g(device);
    if (g.rc) return g.rc;
    rc = forward_impl(packed, FwdIn(FS), fd, fd_strides, B, S, H, W, out4, nullptr, workspace, workspace_bytes, mode,
                      (cudaStream_t)stream, false, nullptr, &prof);
    if (!rc) rc = check_cuda(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
  }
  const int n = (int)prof.ops.size();
  for (int i = 0; i < n; ++i) {
    OpRecord& o = prof.ops[i];
    float ms = 0.f;
    if (!rc && o.e0 && o.e1) cudaEventElapsedTime(&ms, o.e0, o.e1);
    if (o.e0) cudaEventDestroy(o.e0);
    if (o.e1) cudaEventDestroy(o.e1);
    if (i < max_ops) {
      if (op_ms_host) op_ms_host[i] = ms;
      if (op_flops_host) op_flops_host[i] = o.flops;
      if (op_bytes_host) op_bytes_host[i] = o.bytes;
      if (op_launches_host) op_launches_host[i] = o.launches;
      if (op_names_host) {
        snprintf(op_names_host + (size_t)i * 64, 64, "%s", o.name.c_str());
      }
    }
  }
  *n_ops = n;
  return rc;
}

// ---- host-buffer entry point: a three-stage software pipeline over micro-batches ----------------------------------------------
// stage k holds one micro-batch: FS | focus_dists | 4 depth maps.  The host->device copies of the next two chunks and chunk i-1's
// device->host reads run on two private copy streams while chunk i computes on the caller's stream (PCIe is full duplex).  The
// copies of a DDFF stack take 0.64 ms, its kernels 0.68 ms + 1.3 ms per launch sequence: with three stages the copy stream never
// idles, so only the first chunk's copy is exposed.
constexpr int kHostStages = 3;
constexpr int kHostMaxStages = 8;   // (two-stream schedule of the uint8 path: the same device buffer cut into smaller stages)
// focus-distance elements one chunk of n stacks addresses through the strides
static size_t fd_span(const int64_t* fds, int n, int S, int H, int W) {
  const int dims[4] = {n, S, H, W};
  size_t e = 1;
  for (int i = 0; i < 4; ++i) e += (size_t)(dims[i] - 1) * (size_t)fds[i];
  return e;
}
// one stage: input stacks (in_stack_bytes each) | focus_dists (fd_elems) | 4 depth maps
static size_t host_stage_bytes(int mb, size_t in_stack_bytes, size_t fd_elems, int H, int W) {
  return align_up((size_t)mb * in_stack_bytes, 256) + align_up(fd_elems * 4, 256) + 4 * align_up((size_t)mb * H * W * 4, 256);
}
size_t dff_host_io_bytes(int micro_batch, int S, int H, int W) {
  return kHostStages * host_stage_bytes(micro_batch, (size_t)3 * S * H * W * 4, (size_t)micro_batch * S * H * W, H, W);
}
size_t dff_host_io_bytes_u8(int micro_batch, int S, int H0, int W0, int H, int W, const int64_t fd_strides[4]) {
  if (!fd_strides || micro_batch < 1) return 0;
  return kHostStages * host_stage_bytes(micro_batch, (size_t)3 * S * H0 * W0, fd_span(fd_strides, micro_batch, S, H, W), H, W);
}

namespace {
struct HostPipe {
  cudaStream_t h2d = nullptr, d2h = nullptr, c2 = nullptr;   // copy streams; second compute stream (uint8 path)
  cudaEvent_t in_ready[kHostMaxStages] = {}, computed[kHostMaxStages] = {}, out_done[kHostMaxStages] = {};
  cudaEvent_t entry = nullptr, c2_done = nullptr;
  cudaEvent_t ticket_done[2] = {};   // asynchronous calls: everything of call `ticket` is in host memory
  bool ok = false;
  int device = -1;
  HostPipe() = default;
  HostPipe(const HostPipe&) = delete;
  HostPipe& operator=(const HostPipe&) = delete;
  // the calling thread ends (nn.DataParallel's worker threads are short-lived): give the streams and events back.  At process
  // exit the runtime may already be gone; the calls then fail harmlessly.
  ~HostPipe() {
    if (device < 0) return;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }
    for (int k = 0; k < kHostMaxStages; ++k) {
      if (in_ready[k]) cudaEventDestroy(in_ready[k]);
      if (computed[k]) cudaEventDestroy(computed[k]);
      if (out_done[k]) cudaEventDestroy(out_done[k]);
    }
    if (entry) cudaEventDestroy(entry);
    if (c2_done) cudaEventDestroy(c2_done);
    for (int t = 0; t < 2; ++t)
      if (ticket_done[t]) cudaEventDestroy(ticket_done[t]);
    if (h2d) cudaStreamDestroy(h2d);
    if (d2h) cudaStreamDestroy(d2h);
    if (c2) cudaStreamDestroy(c2);
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
  }
};
// one pipe per (host thread, device): nn.DataParallel drives each GPU from its own thread
HostPipe* host_pipe(int device) {
  thread_local std::map<int, HostPipe> pipes;
  HostPipe& hp = pipes[device];
  if (!hp.ok) {
    hp.device = device;
    if (cudaStreamCreateWithFlags(&hp.h2d, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&hp.d2h, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&hp.c2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&hp.entry, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&hp.c2_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (int t = 0; t < 2; ++t)
      if (cudaEventCreateWithFlags(&hp.ticket_done[t], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (int k = 0; k < kHostMaxStages; ++k) {
      if (cudaEventCreateWithFlags(&hp.in_ready[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&hp.computed[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&hp.out_done[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    hp.ok = true;
  }
  return &hp;
}

// `in_host`: fp32 (B,3,S,H,W) stacks (u8 == false) or uint8 (B,S,H0,W0,3) stacks
int forward_host_impl(const void* packed, const void* in_host, bool u8, int H0, int W0, const float* fd_host, const int64_t* fd_strides,
                      int B, int micro_batch, int S, int H, int W, float* const* out4_host, void* dev_io, void* workspace,
                      size_t workspace_bytes, int mode, int device, cudaStream_t st, int ticket = -1) {
  if (B < 1 || micro_batch < 1) return fail(DFF_E_ARG, "dff_forward_host: B and micro_batch must be >= 1");
  if (u8 && (H0 < 1 || W0 < 1 || H0 > H || W0 > W)) return fail(DFF_E_ARG, "dff_forward_host_u8: need 1 <= H0 <= H and 1 <= W0 <= W");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  HostPipe* hp = host_pipe(device);
  if (!hp) return fail(DFF_E_CUDA, "dff_forward_host: cannot create the copy streams");
  const int mb = micro_batch < B ? micro_batch : B;
  const size_t in_stack = u8 ? (size_t)3 * S * H0 * W0 : (size_t)3 * S * H * W * 4, map_px = (size_t)H * W;
  if (fd_span(fd_strides, mb, S, H, W) > (size_t)mb * S * H * W) return fail(DFF_E_ARG, "dff_forward_host: focus_dists strides exceed (B,S,H,W)");
  const size_t fd_elems = u8 ? fd_span(fd_strides, mb, S, H, W) : (size_t)mb * S * H * W;
  const size_t io_bytes = kHostStages * host_stage_bytes(mb, in_stack, fd_elems, H, W);   // what the caller sized `dev_io` for
  int rc = 0;
  // ---- chunk schedule ----
  // Every chunk costs a fixed ~1.3 ms of launch prologues / pipeline drains on top of its per-stack time; the first chunk's
  // host->device copy and the last chunk's device->host reads cannot overlap anything.  fp32 stacks (35 MB each) are copy-bound:
  // mb/4, mb/4, then mb/2 throughout keeps the copy stream busy (one compute stream).  uint8 stacks (6 MB each) are kernel-bound:
  // consecutive chunks ALTERNATE BETWEEN TWO COMPUTE STREAMS (each with its own half of the workspace) — two forwards in flight fill
  // each other's prologues and drains, so the fixed cost of a chunk is hidden (measured: 2 x 32 stacks on two streams take the time
  // of 1 x 64) and the chunks at both ends can be tiny: sizes mb/32, mb/10, equal middle parts <= mb/2, mb/10, mb/32.
  std::vector<int> sizes;
  static const int host_streams = getenv("DFF_B200_HOST_STREAMS") ? atoi(getenv("DFF_B200_HOST_STREAMS")) : 2;   // (A/B knob)
  bool two = false;
  if (u8 && host_streams >= 2 && B >= 16 && mb >= 16) {
    // (asynchronous calls overlap their ends with the neighbouring calls: no small head / tail chunks, just halves of mb)
    const int e0 = ticket >= 0 ? 0 : std::max(1, mb / 32), e1 = ticket >= 0 ? 0 : std::max(1, mb / 10), cap = mb / 2;
    const int mid = B - 2 * (e0 + e1);
    if (mid >= 2) {
      int parts = std::max(2, (mid + cap - 1) / cap);
      parts += parts & 1;
      if (e0) { sizes.push_back(e0); sizes.push_back(e1); }
      for (int i = 0, left = mid; i < parts; ++i) { const int n = (left + (parts - i) - 1) / (parts - i); sizes.push_back(n); left -= n; }
      if (e0) { sizes.push_back(e1); sizes.push_back(e0); }
      if (const char* sch = getenv("DFF_B200_HOST_SCHED")) {   // (experiments: explicit chunk sizes "2,6,24,24,6,2"; must sum to B)
        std::vector<int> v;
        int sum = 0;
        for (const char* q = sch; *q;) { const int n = atoi(q); if (n > 0) { v.push_back(n); sum += n; } while (*q && *q != ',') ++q; if (*q) ++q; }
        if (sum == B) sizes = v;
      }
      int maxn = 0;
      for (int n : sizes) maxn = std::max(maxn, n);
      // both compute streams need a workspace, and the stages of up to two computing + one arriving + one leaving chunk
      two = 2 * align_up(dff_workspace_bytes(maxn, S, H, W, mode), 256) <= workspace_bytes &&
            4 * host_stage_bytes(maxn, in_stack, fd_span(fd_strides, maxn, S, H, W), H, W) <= io_bytes;
    }
  }
  if (!two) {
    sizes.clear();
    const int head = mb >= 8 ? mb / 8 : 1;
    for (int i0 = 0, nchunk = 0; i0 < B; ++nchunk) {
      int n;
      if (!u8) n = nchunk < 2 ? (mb >= 4 ? mb / 4 : mb) : (mb >= 2 ? mb / 2 : mb);
      else if (B <= 2 * head + 1) n = B;
      else if (nchunk == 0) n = head;
      else {
        const int left = B - i0;
        if (left <= head) n = left;
        else {   // the middle, in equal parts no larger than mb
          const int mid = left - head, parts = (mid + mb - 1) / mb;
          n = (mid + parts - 1) / parts;
        }
      }
      if (n > mb) n = mb;
      if (n > B - i0) n = B - i0;
      sizes.push_back(n);
      i0 += n;
    }
  }
  int smax = 0;
  for (int n : sizes) smax = std::max(smax, n);
  const int sb = two ? smax : mb;                                   // stacks a stage holds
  const size_t sfd = two ? fd_span(fd_strides, sb, S, H, W) : fd_elems;
  const size_t stage = host_stage_bytes(sb, in_stack, sfd, H, W);
  const int nstages = two ? (int)std::min<size_t>(kHostMaxStages, io_bytes / stage) : kHostStages;
  const size_t ws_half = two ? (workspace_bytes / 2) & ~(size_t)255 : workspace_bytes;
  if (two) {   // the second stream starts after whatever the caller queued on its stream before this call
    DFF_CUDA(cudaEventRecord(hp->entry, st));
    DFF_CUDA(cudaStreamWaitEvent(hp->c2, hp->entry, 0));
  }
  int nchunk = 0;
  for (int i0 = 0; nchunk < (int)sizes.size() && !rc; i0 += sizes[nchunk], ++nchunk) {
    const int n = sizes[nchunk];
    const int k = nchunk % nstages;
    cudaStream_t cs = (two && (nchunk & 1)) ? hp->c2 : st;
    char* wsp = (char*)workspace + ((two && (nchunk & 1)) ? ws_half : 0);
    char* io = (char*)dev_io + k * stage;
    float* dfd = (float*)(io + align_up((size_t)sb * in_stack, 256));
    char* o = (char*)dfd + align_up(sfd * 4, 256);
    float* dout[4];
    for (int j = 0; j < 4; ++j) dout[j] = (float*)(o + j * align_up((size_t)sb * map_px * 4, 256));
    // stage k is free again once the maps of the chunk that used it last have left it
    if (nchunk >= nstages) DFF_CUDA(cudaStreamWaitEvent(hp->h2d, hp->out_done[k], 0));
    DFF_CUDA(cudaMemcpyAsync(io, (const char*)in_host + (size_t)i0 * in_stack, (size_t)n * in_stack, cudaMemcpyHostToDevice, hp->h2d));
    DFF_CUDA(cudaMemcpyAsync(dfd, fd_host + (size_t)i0 * fd_strides[0], fd_span(fd_strides, n, S, H, W) * 4, cudaMemcpyHostToDevice, hp->h2d));
    DFF_CUDA(cudaEventRecord(hp->in_ready[k], hp->h2d));
    DFF_CUDA(cudaStreamWaitEvent(cs, hp->in_ready[k], 0));
    const FwdIn fin = u8 ? FwdIn((const unsigned char*)io, H0, W0) : FwdIn((const float*)io);
    rc = forward_impl(packed, fin, dfd, fd_strides, n, S, H, W, dout, nullptr, wsp, ws_half, mode, cs, false, nullptr);
    if (rc) break;
    DFF_CUDA(cudaEventRecord(hp->computed[k], cs));
    DFF_CUDA(cudaStreamWaitEvent(hp->d2h, hp->computed[k], 0));
    for (int j = 0; j < 4; ++j)
      if (out4_host[j])
        DFF_CUDA(cudaMemcpyAsync(out4_host[j] + (size_t)i0 * map_px, dout[j], (size_t)n * map_px * 4, cudaMemcpyDeviceToHost, hp->d2h));
    DFF_CUDA(cudaEventRecord(hp->out_done[k], hp->d2h));
  }
  if (two) {   // the caller's stream is "after" everything this call computed
    cudaEventRecord(hp->c2_done, hp->c2);
    cudaStreamWaitEvent(st, hp->c2_done, 0);
  }
  if (ticket >= 0 && !rc) {
    // asynchronous call: everything is queued; the maps are complete when the device->host stream gets here (its reads wait for
    // the kernels, the kernels for the uploads).  The next call queues behind this one on the same streams — its uploads run during
    // this call's kernels, its first kernels fill this call's drain.
    DFF_CUDA(cudaEventRecord(hp->ticket_done[ticket], hp->d2h));
    return 0;
  }
  // the call returns with every map in host memory (and nothing of it still queued on the private streams)
  cudaError_t e1 = cudaStreamSynchronize(hp->h2d), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(hp->d2h);
  if (two && e2 == cudaSuccess) e2 = cudaStreamSynchronize(hp->c2);
  if (rc) return rc;
  DFF_CUDA(e1);
  DFF_CUDA(e2);
  DFF_CUDA(e3);
  return 0;
}
}  // namespace

int dff_forward_host(const void* packed, const float* FS_host, const float* fd_host, const int64_t fd_strides[4], int B,
                     int micro_batch, int S, int H, int W, float* const out4_host[4], void* dev_io, void* workspace,
                     size_t workspace_bytes, int mode, int device, void* stream) {
  if (!packed || !FS_host || !fd_host || !fd_strides || !out4_host || !dev_io || !workspace)
    return fail(DFF_E_ARG, "dff_forward_host: null pointer");
  return forward_host_impl(packed, FS_host, false, 0, 0, fd_host, fd_strides, B, micro_batch, S, H, W, out4_host, dev_io, workspace,
                           workspace_bytes, mode, device, (cudaStream_t)stream);
}

int dff_forward_host_u8(const void* packed, const uint8_t* FS_u8_host, int H0, int W0, const float* fd_host, const int64_t fd_strides[4],
                        int B, int micro_batch, int S, int H, int W, float* const out4_host[4], void* dev_io, void* workspace,
                        size_t workspace_bytes, int mode, int device, void* stream) {
  if (!packed || !FS_u8_host || !fd_host || !fd_strides || !out4_host || !dev_io || !workspace)
    return fail(DFF_E_ARG, "dff_forward_host_u8: null pointer");
  return forward_host_impl(packed, FS_u8_host, true, H0, W0, fd_host, fd_strides, B, micro_batch, S, H, W, out4_host, dev_io, workspace,
                           workspace_bytes, mode, device, (cudaStream_t)stream);
}

int dff_forward_host_u8_async(const void* packed, const uint8_t* FS_u8_host, int H0, int W0, const float* fd_host,
                              const int64_t fd_strides[4], int B, int micro_batch, int S, int H, int W, float* const out4_host[4],
                              void* dev_io, void* workspace, size_t workspace_bytes, int mode, int device, void* stream, int ticket) {
  if (!packed || !FS_u8_host || !fd_host || !fd_strides || !out4_host || !dev_io || !workspace)
    return fail(DFF_E_ARG, "dff_forward_host_u8_async: null pointer");
  if (ticket != 0 && ticket != 1) return fail(DFF_E_ARG, "dff_forward_host_u8_async: ticket must be 0 or 1");
  return forward_host_impl(packed, FS_u8_host, true, H0, W0, fd_host, fd_strides, B, micro_batch, S, H, W, out4_host, dev_io, workspace,
                           workspace_bytes, mode, device, (cudaStream_t)stream, ticket);
}

int dff_forward_host_wait(int device, int ticket) {
  if (ticket != 0 && ticket != 1) return fail(DFF_E_ARG, "dff_forward_host_wait: ticket must be 0 or 1");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  HostPipe* hp = host_pipe(device);
  if (!hp) return fail(DFF_E_CUDA, "dff_forward_host_wait: no pipeline on this thread and device");
  DFF_CUDA(cudaEventSynchronize(hp->ticket_done[ticket]));
  return 0;
}


static Layer adhoc_layer(int Cin, int Cout, int kd, int kh, int kw, int stride_hw, int dil_hw, bool transposed, bool tc, size_t* bytes) {
  Layer l{"adhoc", "", Cin, Cout, kd, kh, kw, stride_hw, dil_hw, transposed, false};
  l.ntaps = kd * kh * kw;
  l.CinP = tc ? (int)align_up(Cin, 8) : (int)align_up(Cin, 4);   // (stored channels: the caller's tensors are already padded)
  l.CoutP = (int)align_up(Cout, 8);
  l.CinT = (int)align_up(Cin, 8);
  l.Ntc = (int)align_up(Cout, 16);
  l.raw_w = l.raw_gamma = l.raw_beta = l.raw_mean = l.raw_var = l.raw_bias = -1;
  size_t n = 0;
  layout_layer(l, n);
  if (bytes) *bytes = n;
  return l;
}

size_t dff_conv3d_scratch_bytes(int Cin, int Cout, int kd, int kh, int kw) {
  // every kernel layout of the layer (FFMA, per-tap TMA, slab, folded forms) + a fused classifier's weights
  size_t n = 0, m = 0;
  adhoc_layer(Cin, Cout, kd, kh, kw, 1, kh == 9 ? 2 : 1, false, true, &n);
  adhoc_layer(Cin, Cout, 3, 3, 3, 2, 1, true, true, &m);
  if (kh == 3 && kw == 3) {   // (a stride-2 layer: its x-paired pack)
    size_t q = 0;
    adhoc_layer(Cin, Cout, kd, kh, kw, 2, 1, false, true, &q);
    n = n > q ? n : q;
  }
  return (n > m ? n : m) + align_up((size_t)align_up(Cout, 16) * 4 * sizeof(float), 256) + 1024;
}

int dff_conv3d_ex(const void* in0, int C0, const void* in1, int C1, int B, int S, int IH, int IW, const float* weight, int Cout,
                  int kd, int kh, int kw, int stride_hw, int dil_hw, int transposed, const float* scale, const float* shift,
                  const void* res_pre, const void* res_post, int relu, void* out, int elem, int plan, int pair_input,
                  const void* aux_add, void* aux_out, const float* proj_w, float* proj_out, int proj_on_aux, int skip_out,
                  void* scratch, int device, void* stream) {
  if (!in0 || !weight || !out || !scratch) return fail(DFF_E_ARG, "dff_conv3d: null pointer");
  g_pdl_call = false;   // the weights are packed by kernels launched just before the convolution: its prologue must not run ahead of them
  const bool out_f32 = (elem & DFF_OUT_F32) != 0;   // fp32 output from bf16 operands (cost volumes)
  elem &= ~DFF_OUT_F32;
  const bool tc = plan != 0;
  if (tc && (elem != DFF_BF16 || (C0 % 8) || (C1 % 8)))
    return fail(DFF_E_UNSUPPORTED, "dff_conv3d: the tensor-core path needs bf16 tensors with channel counts that are multiples of 8");
  if (kd * kh * kw > kMaxTaps) return fail(DFF_E_ARG, "dff_conv3d: too many taps");
  if (tc && ((scale == nullptr) != (shift == nullptr)))
    return fail(DFF_E_ARG, "dff_conv3d: the tensor-core path takes scale and shift together (or neither)");
  if (transposed && !(kd == 3 && kh == 3 && kw == 3 && stride_hw == 2 && dil_hw == 1))
    return fail(DFF_E_ARG, "dff_conv3d: transposed conv must be k=3, stride (1,2,2)");
  if (C0 % 4 || C1 % 4) return fail(DFF_E_ARG, "dff_conv3d: stored channels must be multiples of 4");
  if ((aux_add == nullptr) != (aux_out == nullptr)) return fail(DFF_E_ARG, "dff_conv3d: aux_add and aux_out go together");
  if ((proj_w == nullptr) != (proj_out == nullptr)) return fail(DFF_E_ARG, "dff_conv3d: proj_w and proj_out go together");
  if ((aux_add || proj_w || pair_input) && plan != 4)
    return fail(DFF_E_UNSUPPORTED, "dff_conv3d: second output / fused classifier / pair-packed input exist only in the forward's plan (plan 4)");
  if (pair_input && !(C0 == 8 && C1 == 0 && kd == 1 && kh == 9 && kw == 9 && dil_hw == 2 && stride_hw == 1 && !transposed))
    return fail(DFF_E_ARG, "dff_conv3d: the pair-packed input belongs to the 1x9x9 dilation-2 first layer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int Cin = pair_input ? 3 : C0 + C1;
  size_t nbytes = 0;
  Layer l = adhoc_layer(Cin, Cout, kd, kh, kw, stride_hw, dil_hw, transposed != 0, tc, &nbytes);
  if (!pair_input) { l.CinP = Cin; l.CinT = Cin; l.pair_x = false; }
  char* pk = (char*)scratch;
  if (plan == 4) {
    DFF_TRY(pack_layer_weights(l, weight, pk, st));
    // per-channel scale/shift as the forward keeps them (padded to CoutP; identity when absent), then the replicated folded forms
    DFF_TRY(launch_bn_fold(nullptr, nullptr, nullptr, nullptr, nullptr, (float*)(pk + l.pk_scale), (float*)(pk + l.pk_shift), Cout, l.CoutP, st));
    if (scale) {
      DFF_CUDA(cudaMemcpyAsync(pk + l.pk_scale, scale, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
      DFF_CUDA(cudaMemcpyAsync(pk + l.pk_shift, shift, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    DFF_TRY(pack_layer_folded_ss(l, pk, st));
  } else if (tc) {
    if (plan == 2) DFF_TRY(launch_pack_weight_tc(weight, pk + l.pk_wtc, Cout, Cin, Cin, l.ntaps, l.Ntc, transposed ? 1 : 0, st));
    else DFF_TRY(launch_pack_weight_slab(weight, pk + l.pk_wslab, Cout, Cin, Cin, l.ntaps, l.Ntc, transposed ? 1 : 0, st));
  } else {
    DFF_TRY(launch_pack_weight(weight, (float*)(pk + l.pk_w), Cout, Cin, l.ntaps, l.CinP, l.CoutP, transposed ? 1 : 0, st));
  }
  Layer pl = l;   // the fused classifier: its weights sit behind the layer's own packs
  if (proj_w) {
    pl.cin = Cout; pl.cout = 1; pl.CinT = (int)align_up(Cout, 8);
    pl.pk_proj = align_up(nbytes, 256);
    DFF_CUDA(cudaMemsetAsync(pk + pl.pk_proj, 0, (size_t)align_up(Cout, 16) * sizeof(float), st));
    DFF_CUDA(cudaMemcpyAsync(pk + pl.pk_proj, proj_w, Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  Ten in;
  in.p = const_cast<void*>(in0); in.B = B; in.S = S; in.H = IH; in.W = IW; in.C = C0;
  Ten t1;
  t1.p = const_cast<void*>(in1); t1.C = C1;
  Ten o;
  o.p = out; o.B = B; o.S = S; o.C = Cout;
  const int IWl = pair_input ? IW - 2 : IW;   // (the pair-packed input has two margin columns)
  o.H = transposed ? IH * 2 : IH / stride_hw;
  o.W = transposed ? IWl * 2 : IWl / stride_hw;
  Ten rp, rq, ax, ao, po;
  rp.p = const_cast<void*>(res_pre);
  rq.p = const_cast<void*>(res_post);
  ax.p = const_cast<void*>(aux_add);
  ao.p = aux_out;
  po.p = proj_out;
  EpiOpt e;
  e.in1 = (in1 && C1) ? &t1 : nullptr;
  e.res_pre = res_pre ? &rp : nullptr;
  e.res_post = res_post ? &rq : nullptr;
  e.relu = relu != 0;
  e.out_f32 = out_f32;
  o.f32 = out_f32;
  if (aux_add) { e.aux_add = &ax; e.aux_out = &ao; }
  if (proj_w) { e.proj = &pl; e.proj_out = &po; e.proj_aux = proj_on_aux != 0; e.skip_out = skip_out != 0; }
  // plan: 0 = FFMA; 1 = best tensor-core kernel for the plain form (row kernel with A in TMEM -> slab kernel -> per-tap TMA kernel);
  // 2 = force the per-tap TMA kernel; 3 = slab kernel (never the row kernel); 4 = what dff_forward runs for this layer: the folded forms
  // (x-fold, deconv-fold, row-folded first layer), the second output and the fused classifier included
  if (plan == 4) {
    static const bool no_fold = getenv("DFF_B200_NO_FOLD") != nullptr;
    return run_conv(l, (const float*)(pk + l.pk_w), (const float*)(pk + l.pk_scale), (const float*)(pk + l.pk_shift), in, e, o, true, st,
                    pk + l.pk_wtc, pk + l.pk_wslab, nullptr, false, true, pk, !no_fold);
  }
  LazyTcPack lazy{weight, pk + l.pk_wtc, Cout, Cin, Cin, l.ntaps, l.Ntc, transposed ? 1 : 0, plan == 2 || !tc};
  return run_conv(l, (const float*)(pk + l.pk_w), scale, shift, in, e, o, elem == DFF_BF16, st, tc ? pk + l.pk_wtc : nullptr,
                  (tc && plan != 2) ? pk + l.pk_wslab : nullptr, nullptr, false, plan == 1, nullptr, false, &lazy);
}

int dff_conv3d(const void* in0, int C0, const void* in1, int C1, int B, int S, int IH, int IW, const float* weight, int Cout,
               int kd, int kh, int kw, int stride_hw, int dil_hw, int transposed, const float* scale, const float* shift,
               const void* res_pre, const void* res_post, int relu, void* out, int elem, int use_tensor_cores, void* scratch,
               int device, void* stream) {
  return dff_conv3d_ex(in0, C0, in1, C1, B, S, IH, IW, weight, Cout, kd, kh, kw, stride_hw, dil_hw, transposed, scale, shift, res_pre,
                       res_post, relu, out, elem, use_tensor_cores, 0, nullptr, nullptr, nullptr, nullptr, 0, 0, scratch, device, stream);
}

// SRD channel-attention branch as one operator (reference train_codes/Depth_Estimation_Network.py:399-407)
int dff_srd_attention(const void* F, int B, int S, int H, int W, int C, const float* w_a, const float* w_b, void* out, void* scratch,
                      int device, void* stream) {
  if (!F || !w_a || !w_b || !out || !scratch) return fail(DFF_E_ARG, "dff_srd_attention: null pointer");
  g_pdl_call = false;   // (weights packed just before the launch, see dff_conv3d)
  if (C != 8 && C != 16 && C != 32) return fail(DFF_E_UNSUPPORTED, "dff_srd_attention: C must be 8, 16 or 32");
  if (((size_t)H * W) % 16) return fail(DFF_E_ARG, "dff_srd_attention: H*W must be a multiple of 16");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* w0 = (float*)scratch;            // [dz][ci][co]
  float* w1 = w0 + (size_t)3 * C * C;     // [c][co]
  DFF_TRY(launch_pack_weight(w_a, w0, C, C, 3, C, C, 0, st));
  DFF_TRY(launch_pack_weight(w_b, w1, C, C, 1, C, C, 0, st));
  return launch_srd_attention_mma(F, w0, w1, out, B, S, H, W, C, num_sms_of_current_device(), st);
}

int dff_depth_head(const float* cost, int h, int w, const float* fd, const int64_t fd_strides[4], int B, int S, int H, int W,
                   float* depth, int device, void* stream) {
  if (!cost || !fd || !fd_strides || !depth) return fail(DFF_E_ARG, "dff_depth_head: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_depth_head(cost, h, w, fd, fd_strides, B, S, H, W, depth, (cudaStream_t)stream);
}

int dff_depth_heads4(const float* const cost4[4], const float* fd, const int64_t fd_strides[4], int B, int S, int H, int W,
                     float* const depth4[4], int fast, int device, void* stream) {
  if (!cost4 || !fd || !fd_strides || !depth4) return fail(DFF_E_ARG, "dff_depth_heads4: null pointer");
  if (H % 8 || W % 8) return fail(DFF_E_ARG, "dff_depth_heads4: H and W must be multiples of 8");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  const int h[4] = {H / 8, H / 4, H / 2, H}, w[4] = {W / 8, W / 4, W / 2, W};
  return launch_depth_head4(cost4, h, w, fd, fd_strides, B, S, H, W, depth4, fast != 0, (cudaStream_t)stream);
}

size_t dff_depth_metrics_scratch_bytes(int B) { return depth_metrics_scratch_bytes(B) + 1024; }

int dff_depth_metrics(const float* est, const float* gt, const uint8_t* mask, const float* conf, int B, int H, int W, int Hc, int Wc,
                      float* out12, void* scratch, int device, void* stream) {
  if (!est || !gt || !out12 || !scratch) return fail(DFF_E_ARG, "dff_depth_metrics: null pointer");
  if (B < 1 || Hc < 1 || Wc < 1 || Hc > H || Wc > W) return fail(DFF_E_ARG, "dff_depth_metrics: need 1 <= Hc <= H, 1 <= Wc <= W");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_depth_metrics(est, gt, mask, conf, B, H, W, Hc, Wc, out12, (double*)scratch, (cudaStream_t)stream);
}

int dff_depth_to_jet(const float* est, int B, int H, int W, int Hc, int Wc, float lo, float hi, uint8_t* rgb, void* scratch, int device,
                     void* stream) {
  if (!est || !rgb || !scratch) return fail(DFF_E_ARG, "dff_depth_to_jet: null pointer");
  if (B < 1 || Hc < 1 || Wc < 1 || Hc > H || Wc > W || !(hi > lo)) return fail(DFF_E_ARG, "dff_depth_to_jet: bad crop or range");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_depth_to_jet(est, B, H, W, Hc, Wc, lo, hi, (unsigned char*)scratch, rgb, (cudaStream_t)stream);
}

int dff_fov_warp(const float* x, const float* alpha, const float* fov, int B, int C, int S, int H, int W, float* out,
                 float* flow, int device, void* stream) {
  if (!x || !fov || !out) return fail(DFF_E_ARG, "dff_fov_warp: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_fov_warp(x, alpha, fov, B, C, S, H, W, out, flow, (cudaStream_t)stream);
}

int dff_to_channels_last(const float* src, int B, int C, int S, int H, int W, void* dst, int Cp, int elem, int device,
                         void* stream) {
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_to_cl(src, B, C, S, H, W, dst, Cp, elem == DFF_BF16, (cudaStream_t)stream);
}
int dff_to_pair_packed(const float* FS, int B, int S, int H, int W, void* dst, int device, void* stream) {
  if (!FS || !dst) return fail(DFF_E_ARG, "dff_to_pair_packed: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_to_cl_pair(FS, B, S, H, W, dst, (cudaStream_t)stream);
}
int dff_from_channels_last(const void* src, int B, int C, int S, int H, int W, int Cp, int elem, float* dst, int device,
                           void* stream) {
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_from_cl(src, B, C, S, H, W, Cp, elem == DFF_BF16, dst, (cudaStream_t)stream);
}


// ---- train-mode building blocks -------------------------------------------------------------------------------------
static void taps_of(int kd, int kh, int kw, int dil, TapTable& t, bool negate) {
  Layer l{"", "", 0, 0, kd, kh, kw, 1, dil, false, false};
  conv_taps(l, t);
  if (negate)
    for (int i = 0; i < t.n; ++i) { t.dz[i] = (int8_t)-t.dz[i]; t.dy[i] = (int8_t)-t.dy[i]; t.dx[i] = (int8_t)-t.dx[i]; }
}

size_t dff_conv3d_dgrad_scratch_bytes(int Cin, int Cout, int kd, int kh, int kw) {
  // FFMA pack | tensor-core path: the adjoint weight (reference layout) + the forward operator's own scratch
  const size_t adj = align_up((size_t)kd * kh * kw * align_up(Cout, 8) * align_up(Cin, 8) * 4, 256);
  return adj + dff_conv3d_scratch_bytes((int)align_up(Cout, 8), (int)align_up(Cin, 8), kd, kh, kw);
}

int dff_conv3d_dgrad(const void* dy, int CoS, int B, int S, int OH, int OW, const float* weight, int Cin, int Cout, int kd, int kh,
                     int kw, int stride_hw, int dil_hw, int transposed, int ci0, int nci, void* dx, int elem, void* scratch,
                     int device, void* stream) {
  if (!dy || !weight || !dx || !scratch) return fail(DFF_E_ARG, "dff_conv3d_dgrad: null pointer");
  if (CoS % 4 || CoS < Cout || nci % 4 || ci0 < 0 || ci0 + nci > (int)align_up(Cin, 4))
    return fail(DFF_E_ARG, "dff_conv3d_dgrad: channel counts must be padded to multiples of 4");
  if (kd * kh * kw > kMaxTaps) return fail(DFF_E_ARG, "dff_conv3d_dgrad: too many taps");
  if ((transposed || stride_hw == 2) && !(kd == 3 && kh == 3 && kw == 3 && dil_hw == 1))
    return fail(DFF_E_ARG, "dff_conv3d_dgrad: strided / transposed layers must be k=3");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int ntaps = kd * kh * kw, CaP = CoS, CbP = (int)align_up(nci, 8);
  static const bool dgrad_ffma = getenv("DFF_B200_NO_TC") && atoi(getenv("DFF_B200_NO_TC")) != 0;
  if (elem == DFF_BF16 && !dgrad_ffma && CoS % 8 == 0 && nci % 8 == 0) {
    // tensor-core data gradient: the adjoint convolution through the forward operator (tcgen05 kernels), its weight re-laid out by
    // one small kernel (flipped taps for stride 1, a transposed convolution for stride 2 and vice versa)
    const int mode = transposed ? 2 : (stride_hw == 2 ? 1 : 0);
    float* wa = (float*)scratch;
    char* rest = (char*)scratch + align_up((size_t)ntaps * CoS * nci * 4, 256);
    DFF_TRY(launch_adjoint_weight(weight, wa, Cout, Cin, ntaps, ci0, nci, CoS, mode, st));
    if (mode == 1)
      return dff_conv3d_ex(dy, CoS, nullptr, 0, B, S, OH, OW, wa, nci, 3, 3, 3, 2, 1, 1, nullptr, nullptr, nullptr, nullptr, 0, dx, DFF_BF16, 1,
                           0, nullptr, nullptr, nullptr, nullptr, 0, 0, rest, device, stream);
    return dff_conv3d_ex(dy, CoS, nullptr, 0, B, S, OH, OW, wa, nci, kd, kh, kw, mode == 2 ? 2 : 1, dil_hw, 0, nullptr, nullptr, nullptr, nullptr,
                         0, dx, DFF_BF16, 1, 0, nullptr, nullptr, nullptr, nullptr, 0, 0, rest, device, stream);
  }
  DFF_TRY(launch_pack_weight_dgrad(weight, (float*)scratch, Cout, Cin, ntaps, ci0, nci, CaP, CbP, transposed ? 1 : 0, st));
  ConvArgs a{};
  a.in0 = dy; a.C0 = CoS; a.in1 = nullptr; a.C1 = 0;
  a.B = B; a.S = S; a.IH = OH; a.IW = OW;
  a.w = (const float*)scratch; a.CinP = CaP; a.CoutP = CbP;
  a.out = dx; a.Cout = nci;
  const bool bf16 = elem == DFF_BF16;
  if (transposed) {  // adjoint of the transposed convolution: an ordinary stride-2 convolution of dy with the same weights
    taps_of(3, 3, 3, 1, a.taps, false);
    a.isy = a.isx = 2; a.osy = a.osx = 1;
    a.OH = a.OHt = OH / 2; a.OW = a.OWt = OW / 2;
    return launch_conv_ffma(a, bf16, st);
  }
  if (stride_hw == 1) {  // adjoint of a stride-1 convolution: the same taps with negated offsets
    taps_of(kd, kh, kw, dil_hw, a.taps, true);
    a.isy = a.isx = 1; a.osy = a.osx = 1;
    a.OH = a.OHt = OH; a.OW = a.OWt = OW;
    return launch_conv_ffma(a, bf16, st);
  }
  // adjoint of the stride-(1,2,2) convolution: four output-parity phases of a transposed convolution
  a.OH = OH * 2; a.OW = OW * 2; a.OHt = OH; a.OWt = OW;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      deconv_taps(py, px, a.taps);
      a.isy = a.isx = 1; a.osy = a.osx = 2; a.ooy = py; a.oox = px;
      DFF_TRY(launch_conv_ffma(a, bf16, st));
    }
  return 0;
}

static int conv3d_wgrad_impl(const void* in0, int C0, const void* in1, int C1, int B, int S, int IH, int IW, const void* dy, int CoS,
                             int Cin, int Cout, int kd, int kh, int kw, int stride_hw, int dil_hw, int transposed, float* dw, int elem,
                             int device, void* stream, bool accumulate) {
  if (!in0 || !dy || !dw) return fail(DFF_E_ARG, "dff_conv3d_wgrad: null pointer");
  if (C0 % 4 || C1 % 4) return fail(DFF_E_ARG, "dff_conv3d_wgrad: stored channels must be multiples of 4");
  if (kd * kh * kw > kMaxTaps) return fail(DFF_E_ARG, "dff_conv3d_wgrad: too many taps");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int ntaps = kd * kh * kw;
  if (!accumulate) DFF_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cin * Cout * ntaps * sizeof(float), st));
  ConvArgs a{};
  a.in0 = in0; a.C0 = C0; a.in1 = (in1 && C1) ? in1 : nullptr; a.C1 = a.in1 ? C1 : 0;
  a.B = B; a.S = S; a.IH = IH; a.IW = IW;
  const bool bf16 = elem == DFF_BF16;
  if (!transposed) {
    Layer l{"", "", Cin, Cout, kd, kh, kw, stride_hw, dil_hw, false, false};
    conv_taps(l, a.taps);
    a.isy = a.isx = stride_hw; a.osy = a.osx = 1;
    a.OH = a.OHt = IH / stride_hw; a.OW = a.OWt = IW / stride_hw;
    return launch_conv_wgrad(a, dy, CoS, Cout, Cin, 0, dw, ntaps, 0, bf16, st);
  }
  a.OH = IH * 2; a.OW = IW * 2; a.OHt = IH; a.OWt = IW;
  if (bf16) {   // the four output-parity phases in one launch (they share the staged input region)
    TapTable pt[4];
    for (int ph = 0; ph < 4; ++ph) deconv_taps(ph >> 1, ph & 1, pt[ph]);
    a.isy = a.isx = 1; a.osy = a.osx = 2; a.ooy = a.oox = 0;
    bool handled = false;
    DFF_TRY(launch_conv_wgrad_tma(a, pt, 4, dy, CoS, Cout, Cin, 0, dw, ntaps, 1, st, &handled));
    if (handled) return 0;
  }
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      deconv_taps(py, px, a.taps);
      a.isy = a.isx = 1; a.osy = a.osx = 2; a.ooy = py; a.oox = px;
      DFF_TRY(launch_conv_wgrad(a, dy, CoS, Cout, Cin, 0, dw, ntaps, 1, bf16, st));
    }
  return 0;
}

int dff_conv3d_wgrad(const void* in0, int C0, const void* in1, int C1, int B, int S, int IH, int IW, const void* dy, int CoS,
                     int Cin, int Cout, int kd, int kh, int kw, int stride_hw, int dil_hw, int transposed, float* dw, int elem,
                     int device, void* stream) {
  return conv3d_wgrad_impl(in0, C0, in1, C1, B, S, IH, IW, dy, CoS, Cin, Cout, kd, kh, kw, stride_hw, dil_hw, transposed, dw, elem, device,
                           stream, false);
}
int dff_conv3d_wgrad_acc(const void* in0, int C0, const void* in1, int C1, int B, int S, int IH, int IW, const void* dy, int CoS,
                         int Cin, int Cout, int kd, int kh, int kw, int stride_hw, int dil_hw, int transposed, float* dw, int elem,
                         int device, void* stream) {
  return conv3d_wgrad_impl(in0, C0, in1, C1, B, S, IH, IW, dy, CoS, Cin, Cout, kd, kh, kw, stride_hw, dil_hw, transposed, dw, elem, device,
                           stream, true);
}

size_t dff_bn_scratch_bytes(int C) { return bn_partial_bytes(C); }

int dff_bn_train_forward(const void* x, int64_t npix, int C, int elem, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float momentum, float eps, const void* res_pre, const void* res_post, int relu,
                         void* out, float* save_mean, float* save_invstd, float* scale_shift, void* scratch, int device,
                         void* stream) {
  if (!x || !out) return fail(DFF_E_ARG, "dff_bn_train_forward: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf16 = elem == DFF_BF16;
  if (gamma) {
    if (!save_mean || !save_invstd || !scale_shift || !scratch) return fail(DFF_E_ARG, "dff_bn_train_forward: null pointer");
    DFF_TRY(launch_bn_stats(x, (size_t)npix, C, bf16, gamma, beta, running_mean, running_var, momentum, eps, scale_shift,
                            scale_shift + C, save_mean, save_invstd, scratch, st));
    return launch_bn_apply(x, scale_shift, scale_shift + C, res_pre, res_post, relu, (size_t)npix, C, bf16, out, st);
  }
  return launch_bn_apply(x, nullptr, nullptr, res_pre, res_post, relu, (size_t)npix, C, bf16, out, st);
}

int dff_bn_train_backward(const void* dy, const void* y_relu, const void* x, const float* save_mean, const float* save_invstd,
                          const float* gamma, int64_t npix, int C, int elem, void* dx, void* dres, float* dgamma, float* dbeta,
                          void* scratch, int device, void* stream) {
  if (!dy) return fail(DFF_E_ARG, "dff_bn_train_backward: null pointer");
  if (save_mean && (!x || !save_invstd || !dgamma || !dbeta || !scratch)) return fail(DFF_E_ARG, "dff_bn_train_backward: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_bn_backward(dy, y_relu, x, save_mean, save_invstd, gamma, (size_t)npix, C, elem == DFF_BF16, dx, dres, dgamma, dbeta,
                            scratch, (cudaStream_t)stream);
}

int dff_bn_eval_forward(const void* x, int64_t npix, int C, int elem, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, const void* res_pre, const void* res_post, int relu, void* out,
                        float* save_mean, float* save_invstd, float* scale_shift, int device, void* stream) {
  if (!x || !out || !running_mean || !running_var || !save_mean || !save_invstd || !scale_shift)
    return fail(DFF_E_ARG, "dff_bn_eval_forward: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  cudaStream_t st = (cudaStream_t)stream;
  DFF_TRY(launch_bn_eval_stats(gamma, beta, running_mean, running_var, eps, C, scale_shift, scale_shift + C, save_mean, save_invstd, st));
  return launch_bn_apply(x, scale_shift, scale_shift + C, res_pre, res_post, relu, (size_t)npix, C, elem == DFF_BF16, out, st);
}

int dff_bn_eval_backward(const void* dy, const void* y_relu, const void* x, const float* save_mean, const float* save_invstd,
                         const float* gamma, int64_t npix, int C, int elem, void* dx, void* dres, float* dgamma, float* dbeta,
                         void* scratch, int device, void* stream) {
  if (!dy || !x || !save_mean || !save_invstd || !dgamma || !dbeta || !scratch) return fail(DFF_E_ARG, "dff_bn_eval_backward: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_bn_backward(dy, y_relu, x, save_mean, save_invstd, gamma, (size_t)npix, C, elem == DFF_BF16, dx, dres, dgamma, dbeta,
                            scratch, (cudaStream_t)stream, true);
}

int dff_masked_mse(const float* const pred4[4], const float* gt, const uint8_t* mask, int64_t n, const float weights[4],
                   float* const grad4[4], float* stats, void* scratch, int device, void* stream) {
  if (!pred4 || !gt || !mask || !weights || !grad4 || !stats || !scratch) return fail(DFF_E_ARG, "dff_masked_mse: null pointer");
  for (int k = 0; k < 4; ++k)
    if (!pred4[k] || !grad4[k]) return fail(DFF_E_ARG, "dff_masked_mse: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_masked_mse(pred4, gt, mask, (size_t)n, weights, grad4, stats, (double*)scratch, (cudaStream_t)stream);
}

int dff_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1, double beta2,
                  double eps, int step, const float* grad_scale, int device, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq) return fail(DFF_E_ARG, "dff_adam_flat: null pointer");
  if (step < 1) return fail(DFF_E_ARG, "dff_adam_flat: step counts from 1");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_adam_flat(param, grad, exp_avg, exp_avg_sq, (size_t)n, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
}

int dff_add(const void* a, const void* b, int64_t n, int elem, void* out, int device, void* stream) {
  if (!a || !b || !out) return fail(DFF_E_ARG, "dff_add: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_add(a, b, (size_t)n, elem == DFF_BF16, out, (cudaStream_t)stream);
}

int dff_pool3d(const void* x, int BS, int H, int W, int C, int k, int is_max, int elem, void* out, int device, void* stream) {
  if (!x || !out) return fail(DFF_E_ARG, "dff_pool3d: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_pool(x, out, BS, H, W, C, k, is_max != 0, elem == DFF_BF16, (cudaStream_t)stream);
}

int dff_avgpool_pyramid(const void* x, int BS, int H, int W, int C, void* out2, void* out4, void* out8, int device, void* stream) {
  if (!x || !out2 || !out4 || !out8) return fail(DFF_E_ARG, "dff_avgpool_pyramid: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_avgpool_pyramid(x, out2, out4, out8, BS, H, W, C, (cudaStream_t)stream);
}

int dff_pool3d_backward(const void* x, const void* dy, int BS, int H, int W, int C, int k, int is_max, int elem, void* dx,
                        int device, void* stream) {
  if (!dy || !dx || (is_max && !x)) return fail(DFF_E_ARG, "dff_pool3d_backward: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_pool_bwd(x, dy, dx, BS, H, W, C, k, is_max != 0, elem == DFF_BF16, (cudaStream_t)stream);
}

int dff_depth_head_backward(const float* cost, int h, int w, const float* fd, const int64_t fd_strides[4], int B, int S, int H,
                            int W, const float* ddepth, float* dcost, int device, void* stream) {
  if (!cost || !fd || !fd_strides || !ddepth || !dcost) return fail(DFF_E_ARG, "dff_depth_head_backward: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_depth_head_bwd(cost, h, w, fd, fd_strides, B, S, H, W, ddepth, dcost, (cudaStream_t)stream);
}


// ---- End-to-End alignment network building blocks (channels-last) ----------------------------------------------------------
int dff_fov_warp_cl(const void* x, const float* alpha, const float* fov, int B, int C, int S, int H, int W, void* out, int elem,
                    int device, void* stream) {
  if (!x || !fov || !out) return fail(DFF_E_ARG, "dff_fov_warp_cl: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_fov_warp_cl(x, alpha, fov, B, C, S, H, W, out, elem == DFF_BF16, (cudaStream_t)stream);
}
int dff_pair_volume(const void* feat, const float* alpha, const float* fov, int B, int C, int S, int H, int W, void* out, int elem,
                    int device, void* stream) {
  if (!feat || !fov || !out) return fail(DFF_E_ARG, "dff_pair_volume: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_pair_volume(feat, alpha, fov, B, C, S, H, W, out, elem == DFF_BF16, (cudaStream_t)stream, 0);
}
int dff_spatial_mean_accum(const float* x, int Cs, int B, int S, int H, int W, const float* alpha_in, float s0, float s1, float s2,
                           float* alpha_out, int device, void* stream) {
  if (!x || !alpha_out) return fail(DFF_E_ARG, "dff_spatial_mean_accum: null pointer");
  DeviceGuard g(device);
  if (g.rc) return g.rc;
  return launch_spatial_mean_accum(x, Cs, B, S, H, W, alpha_in, s0, s1, s2, alpha_out, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
