// Implicit-GEMM 3-D convolution on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands, fp32 accumulate.
// Replaces the cuDNN conv3d / conv_transpose3d kernels behind every `nn.Conv3d` / `nn.ConvTranspose3d` of the network
// (reference train_codes/Depth_Estimation_Network.py:352-355, 43-50, 278-301) in DFF_BF16 mode.
//
// GEMM view of one CTA tile:   D[128 output pixels][N = Cout] += A[128 pixels][K] * B[N][K]^T ,  K = taps x Cin
//   * A is never materialised (no im2col, no torch.cat): for every (tap, source, channel chunk) ONE TMA box load
//     (KC channels x TW x TH pixels) of the channels-last activation lands in shared memory already in the UMMA
//     K-major canonical layout (128B / 64B / 32B swizzle, or 8x16B core matrices for 8-channel tensors).  The tap
//     offset is a coordinate shift; zero padding is TMA out-of-bounds fill; a stride-2 convolution reads one of
//     four parity-subsampled views (tensor maps with doubled strides); a transposed convolution is four output-parity
//     phases (3/6/6/12 taps) with a strided store; the hourglass `conv0` concat is a second set of tensor maps.
//   * B (weights, [tap][Cout][Cin] bf16) is streamed by TMA next to A — it stays L2 resident.
//   * one elected thread issues tcgen05.mma (M=128, N=Cout, K=16) into a double-buffered TMEM accumulator;
//     tcgen05.commit releases shared-memory stages and publishes finished accumulators through mbarriers.
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b), apply BatchNorm scale/shift (or bias), residual adds, ReLU,
//     and store bf16 channels-last (or fp32 cost volumes) while the next tile's MMAs run.
// Persistent CTAs (one per SM) walk the tile list; warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
#include <cstring>

#include <cstdlib>
#include "tc_common.cuh"

namespace dff {

constexpr int kTcThreads = 192;
constexpr int kTcMaxLoads = 96;
constexpr int kTcMaxStages = 8;

struct TcLoad {
  int32_t wrow;      // row coordinate into the weight tensor map (tap * N)
  int16_t wcol;      // input-channel coordinate into the weight tensor map
  int16_t c;         // channel coordinate into the activation tensor map
  int8_t map;        // which activation tensor map (source * 4 + parity)
  int8_t dz, dy, dx; // coordinate offsets (already divided for parity views)
};

struct alignas(64) TcParams {
  CUtensorMap amap[8];
  CUtensorMap wmap;
  int nsteps, nloads;      // K steps per tile; TMA load pairs per step (2 only for 8-channel tensors)
  int KC, N;               // channels per load; MMA N (>= 16)
  int S, OHt, OWt, OH, OW, osy, osx, ooy, oox;
  int TW, TH, twshift, tilesX, tilesY, ntiles;
  int isx;                 // input coordinate step per output position inside a view (always 1: strides live in the maps)
  int a_sub, b_sub, stage_bytes, nstages;
  int layout, sbo, lbo_a, lbo_b, tmem_cols;
  const float* scale;
  const float* shift;
  const void* res_pre;
  const void* res_post;
  void* out;
  void* out_aux;
  const void* aux_add;
  int cstore, relu, out_f32;
  const float* proj_w;
  float* proj_out;
  int proj_src, skip_out;
  TcLoad loads[kTcMaxLoads];
};



__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kTcMaxStages + 4];
  __shared__ uint32_t tmem_base_s;
  // dynamic smem base aligned to 1024 B (128B-swizzle atoms and UMMA base_offset = 0 need it)
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kTcMaxStages]);
  const uint32_t tfull0 = smem_u32(&bars[2 * kTcMaxStages]), tempty0 = smem_u32(&bars[2 * kTcMaxStages + 2]);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 8; ++i) prefetch_tmap(&p.amap[i]);
    prefetch_tmap(&p.wmap);
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  pdl_wait();   // (barriers, TMEM and descriptor prefetch above do not depend on the previous layer)
  const uint32_t tmem_base = tmem_base_s;
  const int tiles_per_bs = p.tilesX * p.tilesY;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)p.nloads * (p.a_sub + p.b_sub);
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int bs = tile / tiles_per_bs, r = tile % tiles_per_bs;
        const int ty0 = (r / p.tilesX) * p.TH, tx0 = (r % p.tilesX) * p.TW;
        const int b = bs / p.S, s = bs % p.S;
        for (int ks = 0; ks < p.nsteps; ++ks) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t sa = smem0 + stage * p.stage_bytes, sb = sa + p.nloads * p.a_sub;
          const uint32_t bar = full0 + 8 * stage;
          mbar_expect_tx(bar, tx_bytes);
          for (int j = 0; j < p.nloads; ++j) {
            const TcLoad& L = p.loads[ks * p.nloads + j];
            tma_load_5d(sa + j * p.a_sub, &p.amap[L.map], bar, L.c, tx0 + L.dx, ty0 + L.dy, s + L.dz, b);
            tma_load_2d(sb + j * p.b_sub, &p.wmap, bar, L.wcol, L.wrow);
          }
          if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
      const int kmma = p.nloads == 2 ? 1 : p.KC / 16;  // MMAs (K = 16) per step
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(tempty0 + 8 * buf, ((it >> 1) & 1) ^ 1);
        fence_after();
        const uint32_t dacc = tmem_base + buf * p.N;
        for (int ks = 0; ks < p.nsteps; ++ks) {
          mbar_wait(full0 + 8 * stage, phase);
          fence_after();
          const uint32_t sa = smem0 + stage * p.stage_bytes, sb = sa + p.nloads * p.a_sub;
          const uint64_t ad = make_desc(sa, p.lbo_a, p.sbo, p.layout), bd = make_desc(sb, p.lbo_b, p.sbo, p.layout);
          for (int k = 0; k < kmma; ++k)  // advance 32 B (16 bf16) along K inside the swizzle atom: +2 in the address field
            umma(dacc, ad + 2 * k, bd + 2 * k, idesc, (ks | k) != 0);
          umma_commit(empty0 + 8 * stage);
          if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull0 + 8 * buf);
      }
    }
  } else {
    // =============================== epilogue (4 warps, one TMEM lane = one output pixel) ===============================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int ty = row >> p.twshift, tx = row & (p.TW - 1);
    EpiArgs ep;
    ep.scale = p.scale; ep.shift = p.shift; ep.res_pre = p.res_pre; ep.res_post = p.res_post; ep.out = p.out;
    ep.out_aux = p.out_aux; ep.aux_add = p.aux_add; ep.cstore = p.cstore; ep.relu = p.relu; ep.out_f32 = p.out_f32; ep.N = p.N;
    ep.proj_w = p.proj_w; ep.proj_out = p.proj_out; ep.proj_src = p.proj_src; ep.skip_out = p.skip_out; ep.proj_c = 0; ep.grp_stride = 0; ep.pix_c = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int bs = tile / tiles_per_bs, r = tile % tiles_per_bs;
      const int oy = (r / p.tilesX) * p.TH + ty, ox = (r % p.tilesX) * p.TW + tx;
      const bool valid = oy < p.OHt && ox < p.OWt;
      const size_t pix = ((size_t)bs * p.OH + (oy * p.osy + p.ooy)) * p.OW + (ox * p.osx + p.oox);
      mbar_wait(tfull0 + 8 * buf, (it >> 1) & 1);
      fence_after();
      tc_epilogue_tile(ep, tmem_base + ((uint32_t)(q * 32) << 16) + buf * p.N, valid, pix);
      fence_before();
      mbar_arrive_relaxed(tempty0 + 8 * buf);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  // (A/B knob DFF_B200_TMAP_L2 = 0 | 64 | 128 | 256: L2 promotion of the tensor maps; default 256)
  static const int l2p = getenv("DFF_B200_TMAP_L2") ? atoi(getenv("DFF_B200_TMAP_L2")) : 256;
  const CUtensorMapL2promotion prom = l2p == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : l2p == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                      : l2p == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return 0;
}

// swizzle_bytes: 0 (none), 32, 64 or 128 — the swizzle span, equal to the bytes of the box's innermost row
int encode_tmap_bf16(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                     const unsigned* box, int swizzle_bytes) {
  cuuint64_t d[5], sb[4];
  cuuint32_t bx[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) sb[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  return encode(m, base, rank, d, sb, bx, sw);
}

bool conv_tc_supported(const ConvArgs& a, int Ntc) {
  if (a.C0 % 8 || a.C1 % 8 || a.C0 < 8) return false;
  if (Ntc < 16 || Ntc > 128 || Ntc % 16) return false;
  if (a.isy != a.isx || (a.isy != 1 && a.isy != 2)) return false;
  if (a.isy == 2 && (a.IH % 2 || a.IW % 2)) return false;
  int kc = a.C0 < 64 ? a.C0 : 64;
  if (a.C1 && a.C1 < kc) kc = a.C1;
  if (kc != 8 && kc != 16 && kc != 32 && kc != 64) return false;
  if (a.C0 % kc || a.C1 % kc) return false;
  const int loads = a.taps.n * ((a.C0 + a.C1) / kc);
  return loads + 1 <= kTcMaxLoads;
}

// `wtc`: bf16 weights [ntaps_total + 1][Ntc][Cin] (last tap block all zero), Cin = C0 + C1.
int launch_conv_tc(const ConvArgs& a, const void* wtc, int ntaps_total, int Ntc, int num_sms, cudaStream_t st) {
  if (!conv_tc_supported(a, Ntc)) return fail(-5, "conv_tc: unsupported layer shape");
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int Cin = a.C0 + a.C1;
  int KC = a.C0 < 64 ? a.C0 : 64;
  if (a.C1 && a.C1 < KC) KC = a.C1;
  p.KC = KC;
  p.N = Ntc;
  p.nloads = KC == 8 ? 2 : 1;
  CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                           : KC == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  p.layout = KC == 64 ? 2 : KC == 32 ? 4 : KC == 16 ? 6 : 0;
  p.a_sub = 128 * KC * 2;
  p.b_sub = Ntc * KC * 2;
  if (KC == 8) {  // 8x16B core matrices: SBO between 8-row groups, LBO between the two K halves (= the two sub-tiles)
    p.sbo = 128; p.lbo_a = p.a_sub; p.lbo_b = p.b_sub;
  } else {        // swizzled K-major: SBO = 8 rows, LBO unused
    p.sbo = 8 * KC * 2; p.lbo_a = 16; p.lbo_b = 16;
  }
  p.stage_bytes = (int)align_up((size_t)p.nloads * (p.a_sub + p.b_sub), 1024);
  p.nstages = (200 * 1024) / p.stage_bytes;
  if (p.nstages > kTcMaxStages) p.nstages = kTcMaxStages;
  if (p.nstages < 2) return fail(-5, "conv_tc: stage too large");
  int cols = 2 * Ntc;
  p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : 256;

  // ---- output tiling: the 128-pixel box shape with the fewest (padded) tiles ----------------------------------------
  const int cand[6][2] = {{16, 8}, {8, 16}, {32, 4}, {64, 2}, {128, 1}, {4, 32}};
  long best = -1;
  for (auto& c : cand) {
    const long n = (long)cdiv(a.OWt, c[0]) * cdiv(a.OHt, c[1]);
    if (best < 0 || n < best) { best = n; p.TW = c[0]; p.TH = c[1]; }
  }
  p.twshift = 0;
  while ((1 << p.twshift) < p.TW) ++p.twshift;
  p.tilesX = cdiv(a.OWt, p.TW);
  p.tilesY = cdiv(a.OHt, p.TH);
  p.ntiles = p.tilesX * p.tilesY * a.B * a.S;
  p.S = a.S; p.OHt = a.OHt; p.OWt = a.OWt; p.OH = a.OH; p.OW = a.OW;
  p.osy = a.osy; p.osx = a.osx; p.ooy = a.ooy; p.oox = a.oox;
  p.scale = a.scale; p.shift = a.shift;
  p.res_pre = a.res_pre; p.res_post = a.res_post; p.out = a.out; p.out_aux = a.out_aux; p.aux_add = a.aux_add;
  p.cstore = a.Cout; p.relu = a.relu; p.out_f32 = a.out_f32;
  p.proj_w = a.proj_w; p.proj_out = a.proj_out; p.proj_src = a.proj_src; p.skip_out = a.skip_out;

  // ---- activation tensor maps: (C, W, H, S, B) channels-last; stride 2 = four parity-subsampled views -----------------
  const int st2 = a.isy;
  bool have[8] = {false, false, false, false, false, false, false, false};
  for (int src = 0; src < (a.C1 ? 2 : 1); ++src) {
    const char* base = (const char*)(src ? a.in1 : a.in0);
    const cuuint64_t C = src ? a.C1 : a.C0;
    for (int par = 0; par < (st2 == 2 ? 4 : 1); ++par) {
      const int py = par >> 1, px = par & 1;
      cuuint64_t dims[5] = {C, (cuuint64_t)a.IW / st2, (cuuint64_t)a.IH / st2, (cuuint64_t)a.S, (cuuint64_t)a.B};
      cuuint64_t strides[4] = {C * 2 * st2, (cuuint64_t)a.IW * C * 2 * st2, (cuuint64_t)a.IH * a.IW * C * 2,
                               (cuuint64_t)a.S * a.IH * a.IW * C * 2};
      cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1, 1};
      DFF_TRY(encode(&p.amap[src * 4 + par], base + ((size_t)py * a.IW + px) * C * 2, 5, dims, strides, box, swz));
      have[src * 4 + par] = true;
    }
  }
  for (int i = 0; i < 8; ++i)  // unused slots still get prefetched: point them at a valid map
    if (!have[i]) p.amap[i] = p.amap[0];
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)(ntaps_total + 1) * Ntc};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)Ntc};
    DFF_TRY(encode(&p.wmap, wtc, 2, dims, strides, box, swz));
  }

  // ---- K-step table: tap-major, then source, then channel chunk ------------------------------------------------------
  int n = 0;
  for (int t = 0; t < a.taps.n; ++t)
    for (int src = 0; src < (a.C1 ? 2 : 1); ++src) {
      const int C = src ? a.C1 : a.C0;
      for (int c = 0; c < C; c += KC) {
        TcLoad& L = p.loads[n++];
        int dy = a.taps.dy[t], dx = a.taps.dx[t], par = 0;
        if (st2 == 2) {  // input coordinate 2*o + d  ->  parity view (d & 1), view coordinate o + (d - (d & 1)) / 2
          const int py = dy & 1, px = dx & 1;
          par = py * 2 + px;
          dy = (dy - py) / 2;
          dx = (dx - px) / 2;
        }
        L.map = (int8_t)(src * 4 + par);
        L.c = (int16_t)c;
        L.dz = a.taps.dz[t]; L.dy = (int8_t)dy; L.dx = (int8_t)dx;
        L.wrow = a.taps.widx[t] * Ntc;
        L.wcol = (int16_t)((src ? a.C0 : 0) + c);
      }
    }
  if (p.nloads == 2 && (n & 1)) {  // odd number of 8-channel loads: pair the last one with a zero-weight duplicate
    p.loads[n] = p.loads[n - 1];
    p.loads[n].wrow = ntaps_total * Ntc;
    ++n;
  }
  p.nsteps = n / p.nloads;

  const size_t smem = (size_t)p.nstages * p.stage_bytes + 1024;
  DFF_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.ntiles < num_sms ? p.ntiles : num_sms;
  DFF_CUDA(launch_pdl(conv_tc_kernel, dim3(grid), dim3(kTcThreads), smem, st, p));
  DFF_LAUNCH_CHECK("conv_tc");
  return 0;
}

// bf16 weight pack for the tensor-core path: reference layout -> [tap][Ntc][Cin] (+ one all-zero tap block)
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin, int CinP,
                                      int ntaps, int Ntc, int transposed) {
  const int n = (ntaps + 1) * Ntc * CinP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ci = i % CinP, co = (i / CinP) % Ntc, t = i / (CinP * Ntc);
    float v = 0.f;
    if (t < ntaps && co < Cout && ci < Cin)
      v = transposed ? w[((size_t)ci * Cout + co) * ntaps + t] : w[((size_t)co * Cin + ci) * ntaps + t];
    dst[i] = __float2bfloat16_rn(v);
  }
}

int launch_pack_weight_tc(const float* w, void* dst, int Cout, int Cin, int CinP, int ntaps, int Ntc, int transposed,
                          cudaStream_t st) {
  const int n = (ntaps + 1) * Ntc * CinP;
  int g = cdiv(n, 256);
  if (g > 512) g = 512;
  pack_weight_tc_kernel<<<g, 256, 0, st>>>(w, (__nv_bfloat16*)dst, Cout, Cin, CinP, ntaps, Ntc, transposed);
  DFF_LAUNCH_CHECK("pack_weight_tc");
  return 0;
}

}  // namespace dff
