// tcgen05 / TMEM / TMA / mbarrier primitives (inline PTX) and the fused convolution epilogue shared by the
// tensor-core convolution kernels (conv_tc.cu: per-tap TMA streaming; conv_slab.cu: TMA-staged shared-memory slab, resident or streamed weights).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dff {
namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrival without release semantics: used by the epilogue to hand a TMEM buffer back after tcgen05.wait::ld — a releasing
// arrive would first wait for the epilogue's outstanding global stores to be acknowledged (~1.5k clk per slice).
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (a hung box is a lost lease) — trap after ~2 s.  The slow path is kept
// out of line: the waiting warps are single instruction streams and every inlined instruction costs them latency.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// UMMA shared-memory matrix descriptor (K-major).  layout: 0 none, 2 128B, 4 64B, 6 32B swizzle.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// accumulate variant with the predicate folded into the instruction (no setp in the issue stream)
__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.eq.u32 p, 0, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
// The same instruction from 32-bit descriptor halves.  Both descriptors change only in their low word from one MMA to the next (the
// matrix start address and, for A, the leading-dimension offset live in bits 0-29; stride, version and layout in the constant high
// word), and a single issuing warp is bound by its instruction count on the uniform datapath (~11 uniform instructions per MMA with
// 64-bit descriptor arithmetic: 57 clk per MMA where the tensor pipe needs 40): one 32-bit add per descriptor, the 64-bit operands
// packed inside the asm block so that the constant high words can stay in the odd registers of the pairs.
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                      uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.eq.u32 p, 0, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// one lane of a fully active warp (the rest of the warp keeps executing the same uniform control flow)
__device__ __forceinline__ bool elect_one() {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok));
  return ok != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ float4 ld_bf16x4(const __nv_bfloat16* p) { return Elem<__nv_bfloat16>::load4(p); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
}  // namespace tc

// host: bf16 tiled tensor map without swizzle (conv_tc.cu); out-of-bounds elements read as zero
int encode_tmap_bf16(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                     const unsigned* box, int swizzle_bytes = 0);

// ---- fused epilogue of one 128-pixel x N accumulator tile --------------------------------------------------------------
// y = acc*scale[c] + shift[c] (BatchNorm eval / bias; scale and shift are both given or both null, 16-byte aligned) ; += res_pre ; ReLU ; += res_post ; store bf16 channels-last (or fp32
// for cost volumes) ; optional second output out_aux = y + aux_add.  (epilogue classes E1-E9.)
struct EpiArgs {
  const float* scale;
  const float* shift;
  const void* res_pre;
  const void* res_post;
  void* out;
  void* out_aux;
  const void* aux_add;
  int cstore, relu, out_f32, N;
  // optional fused C -> 1 projection (the 1x1x1 classifiers, reference :53-57, 105, 111, 116): proj_out[pixel] = sum_c v[c] * proj_w[c]
  // with v = the stored value (proj_src 0) or the second output out_aux (proj_src 1); skip_out: `out` itself is not needed
  const float* proj_w;
  float* proj_out;
  int proj_src, skip_out;
  long long grp_stride;   // != 0: 8-channel group g of a GEMM row is stored at pixel base + g * grp_stride elements (row-folded first layer)
  int pix_c;              // channels per stored pixel when grp_stride != 0
  int proj_c;   // x-folded rows hold cstore / proj_c pixels: one projection per pixel, proj_out[pix * (cstore / proj_c) + g]; 0 = one pixel
};

// `tacc`: TMEM address of the accumulator (lane quadrant and buffer column included).  Warp-collective (tcgen05.ld).
__device__ __forceinline__ void tc_epilogue_tile(const EpiArgs& p, uint32_t tacc, bool valid, size_t pix) {
  using namespace tc;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
  const __nv_bfloat16* rpre = reinterpret_cast<const __nv_bfloat16*>(p.res_pre);
  const __nv_bfloat16* rpost = reinterpret_cast<const __nv_bfloat16*>(p.res_post);
  __nv_bfloat16* oaux = reinterpret_cast<__nv_bfloat16*>(p.out_aux);
  const __nv_bfloat16* aadd = reinterpret_cast<const __nv_bfloat16*>(p.aux_add);
  float pacc = 0.f;
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tacc + c0, v);
    if (!valid || c0 >= p.cstore) continue;
    float f[16];
    if (p.scale) {  // 16-byte loads (generic: the slab kernel points scale/shift at shared memory)
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 sc = *reinterpret_cast<const float4*>(p.scale + c0 + j);
        const float4 sh = *reinterpret_cast<const float4*>(p.shift + c0 + j);
        f[j] = fmaf(__uint_as_float(v[j]), sc.x, sh.x);
        f[j + 1] = fmaf(__uint_as_float(v[j + 1]), sc.y, sh.y);
        f[j + 2] = fmaf(__uint_as_float(v[j + 2]), sc.z, sh.z);
        f[j + 3] = fmaf(__uint_as_float(v[j + 3]), sc.w, sh.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
    }
    const size_t o = pix * p.cstore + c0;
    if (p.out_f32) {  // cost volumes (Cout = 1): fp32, scalar stores
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < p.cstore) reinterpret_cast<float*>(p.out)[o + j] = p.relu ? fmaxf(f[j], 0.f) : f[j];
      continue;
    }
    const int nv = min(16, p.cstore - c0);  // 8 or 16 channels (stored channel counts are multiples of 8)
    if (rpre) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        if (j < nv) {
          const float4 t = ld_bf16x4(rpre + o + j);
          f[j] += t.x; f[j + 1] += t.y; f[j + 2] += t.z; f[j + 3] += t.w;
        }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (rpost) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        if (j < nv) {
          const float4 t = ld_bf16x4(rpost + o + j);
          f[j] += t.x; f[j + 1] += t.y; f[j + 2] += t.z; f[j + 3] += t.w;
        }
    }
    if (!p.skip_out) {
#pragma unroll
      for (int j = 0; j < 16; j += 8)
        if (j < nv) {
          uint4 w;
          w.x = pack2(f[j], f[j + 1]); w.y = pack2(f[j + 2], f[j + 3]);
          w.z = pack2(f[j + 4], f[j + 5]); w.w = pack2(f[j + 6], f[j + 7]);
          *reinterpret_cast<uint4*>(out + o + j) = w;
        }
    }
    if (oaux) {
#pragma unroll
      for (int j = 0; j < 16; j += 8)
        if (j < nv) {
          const float4 t0 = ld_bf16x4(aadd + o + j), t1 = ld_bf16x4(aadd + o + j + 4);
          uint4 w;
          w.x = pack2(f[j] + t0.x, f[j + 1] + t0.y); w.y = pack2(f[j + 2] + t0.z, f[j + 3] + t0.w);
          w.z = pack2(f[j + 4] + t1.x, f[j + 5] + t1.y); w.w = pack2(f[j + 6] + t1.z, f[j + 7] + t1.w);
          *reinterpret_cast<uint4*>(oaux + o + j) = w;
          if (p.proj_w && p.proj_src) {   // the projection sees what the next layer will read: the bf16-rounded sums
            f[j] += t0.x; f[j + 1] += t0.y; f[j + 2] += t0.z; f[j + 3] += t0.w;
            f[j + 4] += t1.x; f[j + 5] += t1.y; f[j + 6] += t1.z; f[j + 7] += t1.w;
          }
        }
    }
    if (p.proj_w) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        if (j < nv) {
          const float4 pw = __ldg(reinterpret_cast<const float4*>(p.proj_w + c0 + j));
          pacc = fmaf(__bfloat162float(__float2bfloat16_rn(f[j])), pw.x, pacc);
          pacc = fmaf(__bfloat162float(__float2bfloat16_rn(f[j + 1])), pw.y, pacc);
          pacc = fmaf(__bfloat162float(__float2bfloat16_rn(f[j + 2])), pw.z, pacc);
          pacc = fmaf(__bfloat162float(__float2bfloat16_rn(f[j + 3])), pw.w, pacc);
        }
    }
  }
  if (p.proj_out && valid) p.proj_out[pix] = pacc;
}


// ---- specialised epilogue (slab / row kernels: scale/shift in shared memory) ---------------------------------------------------
// The epilogue warps are single instruction streams and, for the layers with residual operands, the bottleneck of the kernel:
// this variant is compiled per operand pattern (no per-group branches on absent operands), reads residuals / writes outputs as
// 16-byte vectors, and takes scale/shift through shared-state-space loads.
//   RES: 0 none, 1 added before the ReLU (res_pre), 2 added after it (res_post).  AUX: second output out_aux = value + aux_add.
//   PROJ: fused C -> 1 projection of the stored value (proj_src 0) or of the second output (proj_src 1).
__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
  f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 w;
  w.x = tc::pack2(f[0], f[1]); w.y = tc::pack2(f[2], f[3]); w.z = tc::pack2(f[4], f[5]); w.w = tc::pack2(f[6], f[7]);
  return w;
}
// Residual / second-output operand of one GEMM row, 32 columns (four 16-byte vectors) at a time: issued as ONE batch of loads — by the
// caller before it waits for the accumulator (first batch) or at the start of the next 32 columns — instead of one exposed round trip
// per 8-channel half (the operand is a full-size activation tensor: DRAM latency).
// 32-byte global accesses (sm_100: LDG/STG.256): a GEMM row's 16 bf16 channels in one instruction — full 32-byte sectors per request
__device__ __forceinline__ void ldg_nc_v8(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg_v8(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w)
               : "memory");
}
template <int RES, bool AUX>
__device__ __forceinline__ void tc_epilogue_preload(const EpiArgs& p, bool valid, size_t pix, int c32, uint4* rv) {
  if (!(RES || AUX) || !valid) return;
  const __nv_bfloat16* const src =
      reinterpret_cast<const __nv_bfloat16*>(AUX ? p.aux_add : (RES == 1 ? p.res_pre : p.res_post)) + (p.grp_stride ? pix * p.pix_c : pix * p.cstore);
  if (p.grp_stride) {   // row-/y-folded forms: 8-channel group g of the GEMM row is the pixel g * grp_stride elements further
#pragma unroll
    for (int h = 0; h < 4; ++h)
      if (c32 + 8 * h < p.cstore) rv[h] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)((c32 >> 3) + h) * p.grp_stride));
    return;
  }
#pragma unroll
  for (int h = 0; h < 4; h += 2) {
    if (c32 + 8 * h + 16 <= p.cstore) ldg_nc_v8(src + c32 + 8 * h, rv[h], rv[h + 1]);   // (rows of >= 16 channels are 32-byte aligned)
    else if (c32 + 8 * h < p.cstore) rv[h] = __ldg(reinterpret_cast<const uint4*>(src + c32 + 8 * h));
  }
}
template <bool RELU, int RES, bool AUX, bool PROJ>
__device__ __forceinline__ void tc_epilogue_fast(const EpiArgs& p, uint32_t ss_s, uint32_t tacc, bool valid, size_t pix, uint4* rv) {
  using namespace tc;
  const size_t o0 = p.grp_stride ? pix * p.pix_c : pix * p.cstore;
  __nv_bfloat16* const out = reinterpret_cast<__nv_bfloat16*>(p.out) + o0;
  __nv_bfloat16* const oaux = reinterpret_cast<__nv_bfloat16*>(p.out_aux) + o0;
  float pacc = 0.f, pacc1 = 0.f;   // (x-folded rows: pixel 0 / pixel 1 of the row)
  const int pc = (PROJ && p.proj_c) ? p.proj_c : p.cstore;
  for (int c32 = 0; c32 < p.N; c32 += 32) {
    if (c32 > 0) tc_epilogue_preload<RES, AUX>(p, valid, pix, c32, rv);   // (the first 32 columns come preloaded)
#pragma unroll
   for (int g16 = 0; g16 < 2; ++g16) {
    const int c0 = c32 + 16 * g16;
    if (c0 >= p.N) break;
    uint32_t v[16];
    tmem_ld16(tacc + c0, v);
    if (!valid || c0 >= p.cstore) continue;
    const int nh = p.cstore - c0 >= 16 ? 2 : 1;   // 8-channel halves stored from this 16-column group
    uint4 po[2], pa[2];                           // packed output / second output of the two halves
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h >= nh) break;
      const int c = c0 + 8 * h;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        float4 sc, sh;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc.x), "=f"(sc.y), "=f"(sc.z), "=f"(sc.w) : "r"(ss_s + 4 * (c + j)));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sh.x), "=f"(sh.y), "=f"(sh.z), "=f"(sh.w) : "r"(ss_s + 4 * (p.N + c + j)));
        f[j] = fmaf(__uint_as_float(v[8 * h + j]), sc.x, sh.x);
        f[j + 1] = fmaf(__uint_as_float(v[8 * h + j + 1]), sc.y, sh.y);
        f[j + 2] = fmaf(__uint_as_float(v[8 * h + j + 2]), sc.z, sh.z);
        f[j + 3] = fmaf(__uint_as_float(v[8 * h + j + 3]), sc.w, sh.w);
      }
      const uint4 rvec = rv[g16 * 2 + h];
      if (RES) {
        float r[8];
        unpack8(rvec, r);
        if (RES == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += r[j];
        }
        if (RELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (RES == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += r[j];
        }
      } else if (RELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      po[h] = pack8(f);
      if (p.grp_stride) *reinterpret_cast<uint4*>(out + (size_t)(c >> 3) * p.grp_stride) = po[h];   // (second output / classifier do not exist in this form)
      if (AUX) {
        float a[8];
        unpack8(rvec, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += f[j];
        pa[h] = pack8(a);
        if (PROJ && p.proj_src) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = a[j];
        }
      }
      if (PROJ) {   // the projection sees what the next layer will read: the bf16-rounded values
        const int g = c >= pc ? 1 : 0;
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
          float4 pw;   // classifier weights, replicated per pixel of the row, live behind scale/shift in shared memory
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pw.x), "=f"(pw.y), "=f"(pw.z), "=f"(pw.w) : "r"(ss_s + 4 * (2 * p.N + c + j)));
          t = fmaf(__bfloat162float(__float2bfloat16_rn(f[j])), pw.x, t);
          t = fmaf(__bfloat162float(__float2bfloat16_rn(f[j + 1])), pw.y, t);
          t = fmaf(__bfloat162float(__float2bfloat16_rn(f[j + 2])), pw.z, t);
          t = fmaf(__bfloat162float(__float2bfloat16_rn(f[j + 3])), pw.w, t);
        }
        if (g) pacc1 += t; else pacc += t;
      }
    }
    if (nh == 2) {   // 16 channels = 32 bytes per row and instruction
      if (!p.grp_stride && !p.skip_out) stg_v8(out + c0, po[0], po[1]);
      if (AUX) stg_v8(oaux + c0, pa[0], pa[1]);
    } else {
      if (!p.grp_stride && !p.skip_out) *reinterpret_cast<uint4*>(out + c0) = po[0];
      if (AUX) *reinterpret_cast<uint4*>(oaux + c0) = pa[0];
    }
   }
  }
  if (PROJ && valid) {
    if (p.proj_c) { p.proj_out[2 * pix] = pacc; p.proj_out[2 * pix + 1] = pacc1; }
    else p.proj_out[pix] = pacc;
  }
}

}  // namespace dff
