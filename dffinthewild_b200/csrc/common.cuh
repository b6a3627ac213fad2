// Shared device/host helpers for the dff_b200 library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>

namespace dff {

// ---- element access: activations are stored either as fp32 (parity mode) or bf16 (throughput mode) ---------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
    uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    float4 v;
    v.x = __uint_as_float(r.x << 16);
    v.y = __uint_as_float(r.x & 0xffff0000u);
    v.z = __uint_as_float(r.y << 16);
    v.w = __uint_as_float(r.y & 0xffff0000u);
    return v;
  }
  static __device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
  static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// ---- host-side error plumbing ------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);
#define DFF_CUDA(x)                                 \
  do {                                              \
    int _rc = ::dff::check_cuda((x), #x);           \
    if (_rc) return _rc;                            \
  } while (0)
#define DFF_LAUNCH_CHECK(what) DFF_CUDA((cudaGetLastError()))
#define DFF_TRY(x)          \
  do {                      \
    int _rc = (x);          \
    if (_rc) return _rc;    \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// A kernel launched through launch_pdl may become resident while its predecessor in the stream is still draining: it runs its
// prologue (weights -> shared memory, tables, TMEM allocation — nothing a predecessor writes) and then blocks in pdl_wait() until
// the predecessor grid has completed and its memory is visible.  Every such kernel calls pdl_trigger() when it starts its last
// work item so that ITS successor may do the same while the tail drains (triggering at kernel start made early-resident successors
// compete with the running kernel: -9 % at 2 stacks but +2..4 % at 16..64).  Kernels launched the ordinary way in between simply serialise as usual.  DFF_B200_NO_PDL=1 turns it off.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- generic convolution description (shared by the FFMA and tcgen05 paths) --------------------------------
constexpr int kMaxTaps = 81;
struct TapTable {
  int n;
  int8_t dz[kMaxTaps], dy[kMaxTaps], dx[kMaxTaps];  // input offset of the tap (input coordinates)
  uint8_t widx[kMaxTaps];                           // index of the tap inside the packed weight
};

// A (phase of a) convolution as "strided gather over a tap table":
//   out[b,s, oy*osy+ooy, ox*osx+oox, co] = epi( sum_t sum_ci in[b, s+dz_t, oy*isy+dy_t, ox*isx+dx_t, ci] * w[widx_t][ci][co] )
struct ConvArgs {
  const void* in0;
  const void* in1;
  int C0, C1;              // stored channels of the two sources (C1 = 0: single source); multiples of 4
  int B, S, IH, IW;        // input extent
  int OHt, OWt;            // output positions computed by this launch (phase grid)
  int OH, OW;              // output tensor extent
  int isy, isx, osy, osx, ooy, oox;
  int dzmin, dymin, dxmin; // min tap offsets
  int RZ, RY, RX, RXP;     // staged input region per tile (RXP: padded row pitch)
  int TY;                  // tile rows (tile cols = 32)
  const float* w;          // [ntaps_total][CinP][CoutP] fp32
  int CinP, CoutP;
  const float* scale;      // [CoutP] or null
  const float* shift;      // [CoutP] or null
  const void* res_pre;     // added before ReLU (same layout as out) or null
  const void* res_post;    // added after ReLU or null
  int relu;
  void* out;
  void* out_aux;           // optional second output: out_aux = out_value + aux_add
  const void* aux_add;
  int Cout;                // stored channels of out
  int out_f32;             // store `out` as fp32 even when activations are bf16 (cost volumes feeding the depth head)
  const float* proj_w;     // optional fused C -> 1 projection in the tensor-core epilogues (see EpiArgs)
  float* proj_out;
  int proj_src, skip_out;
  int no_wstream;          // slab kernel: do not fall back to streamed weights (the caller has a better form for this layer)
  int row_step;            // > 0: consecutive tile rows are `row_step` input rows apart and the y stride is taken by the MMA descriptor
                           // (no row-parity views): the row-folded first layer.  Output: 8-channel groups `grp_rows` rows apart.
  int grp_rows;
  int proj_c;              // channels per pixel of the projection when a GEMM row holds several pixels (x-folded layers), 0 = all
  const void* wz;          // slab kernel, optional: the layer's weights in the focal-merged streaming layout (pack_weight_slab_zmerge)
  int use_kmask;           // slab kernel, single-phase launches: tap_kmask[t] != 0 lists the 16-channel K steps of tap t whose weights are
  uint8_t tap_kmask[kMaxTaps];   // not all zero (bit j = channels [16 j, 16 j + 16)); the others are not issued (banded weights of the x-grouped forms)
  TapTable taps;
};

}  // namespace dff
