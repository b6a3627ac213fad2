// "Slab" implicit-GEMM convolution on tcgen05/TMEM fed by TMA — the main bf16 kernel of the network.
//
// Replaces the cuDNN conv3d / conv_transpose3d calls of the reference (train_codes/Depth_Estimation_Network.py:352-355,
// 43-50, 278-301) for every layer but the five with >= 128 input channels or stride 2 on 64 (those stay on conv_tc.cu).
//
// What bounds a focal-volume convolution on B200 is not the tensor pipe but how often the activation tile is re-read: a
// per-tap TMA implicit GEMM (conv_tc.cu) pulls the same 128 pixels 27 times from L2.  Here one CTA owns an 8 x 16 pixel
// column of the focal volume and walks the S slices:
//   * one producer thread stages each slice's halo'd tile ONCE in shared memory with tiled TMA loads — one box of
//     8 channels x RX pixels x RY rows per 8-channel chunk and input-stride residue ("view": stride-2 and x-folded layers),
//     out-of-bounds = the convolution's zero padding, second tensor map = torch.cat — which lands as [row][pixel] x 16 B: the UMMA
//     no-swizzle K-major layout in which 8 horizontally adjacent pixels form one 128-byte core matrix.  A ring of NP planes holds
//     slices s-1, s, s+1 (the 3-tap focal dimension) plus look-ahead.  (Per-element 16-byte cp.async by two producer warps, the
//     first implementation and still the DFF_B200_SLAB_TMA=0 fallback, sustained only ~8 B/clk per SM and bounded a third of the
//     layers: profiles/r1_slab_roles.txt.)
//   * a convolution tap is then nothing but a byte offset added to the matrix-descriptor start address (dy rows, dx pixels,
//     dz = which ring slot) — no data movement per tap; the K=16 of one tcgen05.mma is two 8-channel chunk planes (LBO = plane
//     stride) or, for 8-channel tensors, two neighbouring taps (LBO = their distance);
//   * the layer's weights are loaded once per CTA into shared memory, already ordered by MMA — or, when they do not fit next to
//     the plane ring (64-channel 3x3x3 layers, stride-2 layers), STREAMED: one more warp feeds a ring of 3-4 slots with
//     cp.async.bulk copies of the next block of MMAs' weights (template parameter WS);
//   * one warp issues the MMAs of a slice (taps x Cin/16) from uniform registers into a double-buffered TMEM accumulator; 4 epilogue
//     warps (8 for multi-phase layers at <= 2 CTAs/SM, template parameter E2) apply BatchNorm/bias, residuals, ReLU, the fused
//     classifier and store while the next slice is being multiplied; residual operands are requested before the accumulator wait.
//   * row-folded first layer (a.row_step): tile rows 8 input rows apart through the descriptor's row-group stride, two row-parity
//     phases over one staged plane, 32 = 4 rows x 8 output channels per GEMM row (net.cu::run_conv).
// HBM/L2 traffic per output pixel drops from taps x Cin to ~1.4 x Cin, and there is no per-tap barrier round trip.
// Kernels launch with programmatic dependent launch: the prologue (weights, barriers, TMEM) overlaps the previous layer's tail.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_common.cuh"

namespace dff {

constexpr int kSlabProducers = 32;                     // warp 0: producer (one thread drives the TMA; all 32 in the cp.async fallback)
constexpr int kSlabMmaWarp = kSlabProducers / 32;      // warp 1: MMA issuer (+TMEM alloc)
constexpr int kSlabThreads = kSlabProducers + 32 + 128;  // + warps 2-5: epilogue (one per TMEM lane quadrant)
// variants: E2 = a second group of four epilogue warps (warps 6-9; the groups take alternate output phases) for multi-phase layers
// whose shared-memory footprint leaves at most two CTAs per SM; WS = weight streaming: + one last warp, the weight producer
__host__ __device__ constexpr int slab_threads(bool ws, bool e2) { return kSlabThreads + (e2 ? 128 : 0) + (ws ? 32 : 0); }
constexpr int kSlabMaxWSlot = 40 * 1024;                 // largest slot of the streamed-weight ring (bytes)
constexpr int kSlabMaxWSlots = 8;
constexpr int kSlabMaxOps = 128;
constexpr int kSlabMaxPlanes = 8;
constexpr int kSlabMaxAcc = 8;    // accumulator buffers: 2 (double buffer) or up to 8 slots (focal-merged form)
constexpr int kSlabTW = 8, kSlabTH = 16;
constexpr int kSlabSmemBudget = 228 * 1024;   // per SM; every co-resident CTA also costs ~1 KB static + 1 KB reserved

// DFF_SLAB_TRACE (compile-time, debugging only): CTA 0 records clock64 timestamps of its pipeline events per slice
#ifdef DFF_SLAB_TRACE
#define DFF_TR(slot, idx) do { if (p.trace && blockIdx.x == 0 && (idx) < 64 && (threadIdx.x & 31) == 0) p.trace[(idx) * 8 + (slot)] = clock64(); } while (0)
#else
#define DFF_TR(slot, idx) do { } while (0)
#endif

struct alignas(64) SlabParams {
  CUtensorMap tmap[8];   // (tma) per source (x4) and view: the view's pixels as a (C, W/stx, H/sty, S*B) bf16 tensor
  // An 8-channel source without x-stride views is described as (8*W, 1, H/sty, S*B) instead: a pixel IS the channel run, so a box row
  // is RX * 16 contiguous bytes.  The TMA unit retires about one innermost box row per clock: with 16-byte rows a 55 KB plane of the
  // first layer took ~4000 clk to land (13.7 B/clk per SM) and bounded that layer (profiles/r2_changes_measured.txt).
  int tmerge[2];
  // Wide-row mode (sources of 16 / 32 / 64n channels): a plane block is PIXEL-major — [row][pixel][cb channels], rb = 2 * cb = 32 / 64 /
  // 128 bytes per pixel — written by TMA boxes (cb, RX, RY) in the matching 32B / 64B / 128B swizzle and read through swizzled K-major
  // descriptors: ONE TMA row per pixel instead of one per 8 channels (the TMA unit saturates at 0.6 sixteen-byte rows per clock and SM:
  // 9.6 B/clk, tools/ubench/tma_row_bench.cu).  A tap is still a start-address offset (any pixel, any 32-byte K step: the swizzle is a
  // function of the absolute shared-memory address for TMA and MMA alike — tools/ubench/swz_probe.cu, base offset 0 everywhere).
  int wr, rb;
  const void* in0;
  const void* in1;
  const void* wslab;  // bf16 [tap][chunk][N][8]
  int C0, C1, nchunk, nch0;
  int B, S, IH, IW;
  int tile_sy, sbo;   // input rows per tile step; A-descriptor stride between 8-row groups (bytes)
  int sty, stx, nviews, vpy[4], vpx[4];   // input stride per output step (y, x); one staged view per input-coordinate residue
  int oy, ox, RX, RY, VB, CPS, plane_bytes, NP, LA, hz;   // VB: bytes of one (chunk, view) block of a plane (RY*RX*16 rounded up to 128)
  int N, nops, nph;          // MMA N; table entries; output phases (1 = convolution, 4 = fused transposed convolution, 2 = x-folded one)
  int phy[4], phx[4];        // output offset of each phase
  int g[12], ge[12];         // MMA groups by (phase, focal offset): table range [g[ph*3+k], ge[ph*3+k])
  int gw[12], nwu;           // resident weights: first weight slot of each group / number of slots — groups with the same weight sequence
                             // (the two row phases of the row-folded first layer) share their slots in shared memory
  int tilesX, tilesY, nsplit, slen, nitems;
  int OHt, OWt, OH, OW, osy, osx, ooy, oox;
  int w_bytes, tmem_cols, nelem, elem_off, ss_off, planes_off;
  long long* trace;
  int exp;                   // timing experiments only (DFF_SLAB_EXPERIMENT bit mask; results are wrong): 1 no loads, 2 no stores, 4 no MMAs
  EpiArgs epi;
  // streamed weights (layers whose weights do not fit in shared memory next to the plane ring): the MMA table is cut into blocks of
  // consecutive MMAs whose weights are one contiguous range of `wslab`; warp 7 streams them through a ring of `nwslots` slots
  int wstream, nwslots, nblk, wslot_bytes;
  // focal-merged form ("zmerge", 3-tap focal dimension, one output phase, resident weights): the kernel is INPUT-stationary over the
  // focal axis.  Plane z is multiplied once per spatial tap into the accumulators of the three output slices it feeds (z-1, z, z+1) with
  // ONE tcgen05.mma of N' = 3N columns — the accumulators of consecutive slices are adjacent column blocks of a ring of `nslot` TMEM
  // slots, the weights of the three focal taps adjacent row blocks of the B operand.  An SS-form MMA costs 32 + N'/4 clocks (the
  // A-operand fetch dominates), so three taps cost 32 + 3N/4 instead of 3 * (32 + N/4): 2.6x fewer tensor-pipe clocks at N = 16,
  // 2.1x at N = 32, 1.8x at N = 64; every plane is needed exactly once, so the plane ring needs no focal halo.
  int zmerge, nslot, lgslot, zT;    // on/off; accumulator slots (power of two) and its log2; spatial MMAs per plane
  // two-phase layers (x-folded transposed convolutions) in the focal-merged form: phase ph has its own spatial MMAs (zTp[ph], table
  // group 3*ph + 1), its own weight block (zwo[ph], in 16-byte rows) and its own ring of accumulator slots — columns
  // (ph * nslot + slot) * N, so that the slots of consecutive slices stay adjacent for the merged N' = cnt * N instruction
  int zTp[2], zwo[2];
  int egroups;                      // epilogue warp groups (1 or 2)
  int tma;                          // planes staged by ONE tiled TMA load each (single-chunk, single-view layers: box 8 ch x RX x RY)
  int gb[12], gbe[12];              // block range per MMA group
  uint8_t bop[kSlabMaxOps], bn[kSlabMaxOps];   // first MMA / number of MMAs of each block
  uint8_t wop[kSlabMaxOps];   // MMA table entry whose `wsrc` fills weight slot u
  alignas(16) uint64_t tab[kSlabMaxOps + 8];  // (+2 quads: the issuer prefetches two quads ahead) per MMA, zero-extended to 64 bits (added to the descriptor): (A byte offset inside a ring slot >> 4) | (LBO >> 4) << 16
  int16_t wsrc[2 * kSlabMaxOps];  // per MMA and K half: 8-channel weight block (tap * nchunk + chunk) in `wslab`, -1 = zeros
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait(int n) {  // wait until at most n groups are pending
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// staged-element table entry (built once per CTA): where one 16-byte piece of a slice plane comes from / goes to
constexpr int kSlabElemBias = 64;
struct SlabElem {
  int32_t rel;       // element offset from the tile-origin pixel of the slice, in the source's own channel stride
  uint16_t dst16;    // (byte offset inside the ring slot) >> 4, bit 15 = second source
  uint8_t gy, gx;    // input row / column relative to the tile origin, biased by kSlabElemBias (bounds check)
};

// Epilogue role of the slab kernel (4 warps, one TMEM lane quadrant each).  FAST = 0: generic epilogue (tc_epilogue_tile).
template <int FAST, bool RELU, int RES, bool AUX, bool PROJ>
__device__ __forceinline__ void slab_epilogue(const SlabParams& p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0, float* ss,
                                              uint32_t ss_s, const int neg) {
  using namespace tc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;                                   // TMEM lane quadrant of this warp
  const int eg = (warp - (kSlabMmaWarp + 1)) >> 2;          // epilogue group: takes the output phases eg, eg + neg, ...
  const int row = q * 32 + lane;
  const int ty = row >> 3, tx = row & 7;
  EpiArgs ep = p.epi;
  ep.scale = ss;
  ep.shift = ss + p.N;
  const int esz = ep.out_f32 ? 4 : 2;
  const void* const rsrc = ep.res_pre ? ep.res_pre : (ep.res_post ? ep.res_post : ep.aux_add);   // (a layer has at most one residual operand)
  const bool z2 = p.zmerge && p.nph == 2;
  int sc = 0;
  for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
    if (item + (int)gridDim.x >= p.nitems) pdl_trigger();
    int r = item;
    const int isp = r % p.nsplit; r /= p.nsplit;
    const int ox = (r % p.tilesX) * kSlabTW + tx; r /= p.tilesX;
    const int oy = (r % p.tilesY) * kSlabTH + ty;
    const int b = r / p.tilesY;
    const int s_begin = isp * p.slen, s_end = min(p.S, s_begin + p.slen);
    const bool valid = oy < p.OHt && ox < p.OWt;
    for (int s = s_begin; s < s_end; ++s, ++sc) {
      const int buf = p.zmerge ? (sc & (p.nslot - 1)) : (sc & 1);
      const uint32_t par = p.zmerge ? ((uint32_t)(sc >> p.lgslot) & 1u) : ((uint32_t)(sc >> 1) & 1u);
      const size_t row0 = ((size_t)b * p.S + s) * p.OH;
      uint4 rv[4];
      {
        // the residual / second-output operand does not depend on the accumulator: its first 32 columns are requested before the wait
        const size_t pix = (row0 + (oy * p.osy + p.phy[eg])) * p.OW + (ox * p.osx + p.phx[eg]);
        if (FAST && eg < p.nph) tc_epilogue_preload<RES, AUX>(ep, valid, pix, 0, rv);
        else if (valid && rsrc) {
          for (int ph = eg; ph < p.nph; ph += neg) {
            const size_t px = (row0 + (oy * p.osy + p.phy[ph])) * p.OW + (ox * p.osx + p.phx[ph]);
            const size_t ob = px * ep.cstore * esz;
            for (int k = 0; k < ep.cstore * esz; k += 128) prefetch_l1(reinterpret_cast<const char*>(rsrc) + ob + k);
          }
        }
      }
      mbar_wait(tfull0 + 8 * buf, par);
      fence_after();
      if (q == 3 && eg == 0) DFF_TR(3, sc);
      for (int ph = eg; ph < p.nph; ph += neg) {
        const size_t pix = (row0 + (oy * p.osy + p.phy[ph])) * p.OW + (ox * p.osx + p.phx[ph]);
        if (p.exp & 2) continue;
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (z2 ? ph * p.nslot + buf : buf * p.nph + ph) * p.N;   // (two-phase focal-merged: buf = slot, phase-major rings)
        if (FAST) {
          if (ph != eg) tc_epilogue_preload<RES, AUX>(ep, valid, pix, 0, rv);
          tc_epilogue_fast<RELU, RES, AUX, PROJ>(ep, ss_s, tacc, valid, pix, rv);
        } else tc_epilogue_tile(ep, tacc, valid, pix);
      }
      fence_before();
      mbar_arrive_relaxed(tempty0 + 8 * buf);
      if (q == 3 && eg == 0) DFF_TR(4, sc);
    }
  }
}

// I2 (with E2, without WS; one CTA per SM): a SECOND issuing warp (the last warp).  A two-phase layer whose planes leave room for one CTA
// per SM — the row-folded first layer — is bound by its single issuing warp (3.4 k clk per slice for 60 MMAs the tensor pipe does in
// 2.4 k); the phases have separate accumulators, so each issuer takes one phase and both commit to the (two-arrival) barriers.
// Z2: the focal-merged schedule of a TWO-phase layer (its own instantiation: with the phase loop in the common code every
// single-phase focal-merged layer ran 4 % slower — 17 % at the training batch — measured).
template <bool WS, bool E2, bool I2 = false, bool Z2 = false>
// (192 threads x 4 CTAs per SM: 85 registers per thread)
__global__ void __launch_bounds__(slab_threads(WS, E2) + (I2 ? 32 : 0), (WS || I2) ? 1 : (E2 ? 2 : 4)) conv_slab_kernel(const __grid_constant__ SlabParams p) {
  using namespace tc;
  static_assert(!(WS && I2), "the second issuer and the weight producer share the last warp");
  constexpr int kThreads = slab_threads(WS, E2) + (I2 ? 32 : 0);
  constexpr int kEG = E2 ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kSlabMaxPlanes + 2 * kSlabMaxAcc + (WS ? 2 * kSlabMaxWSlots : 0)];
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* const smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t w_s = smem0 + 1024;               // weights, MMA order: [op][half][N][8] bf16
  const uint32_t planes_s = smem0 + p.planes_off;  // ring of NP slice planes
  const SlabElem* const elems = reinterpret_cast<const SlabElem*>(smem_gen + p.elem_off);
  float* const ss = reinterpret_cast<float*>(smem_gen + p.ss_off);  // scale[N], shift[N], classifier weights[N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) DFF_TR(0, 63);   // (trace build: kernel entry; row 63 = entry, tables done, dependency wait over, all roles done)
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kSlabMaxPlanes]);
  const uint32_t tfull0 = smem_u32(&bars[2 * kSlabMaxPlanes]), tempty0 = smem_u32(&bars[2 * kSlabMaxPlanes + kSlabMaxAcc]);
  const uint32_t wfull0 = smem_u32(&bars[2 * kSlabMaxPlanes + 2 * kSlabMaxAcc]), wempty0 = wfull0 + 8 * kSlabMaxWSlots;   // (WS only)

  if (threadIdx.x == 0) {
    if (WS)
      for (int i = 0; i < p.nwslots; ++i) {
        mbar_init(wfull0 + 8 * i, 1);
        mbar_init(wempty0 + 8 * i, 1);
      }
    for (int i = 0; i < p.NP; ++i) {
      mbar_init(full0 + 8 * i, p.tma ? 1 : kSlabProducers);
      mbar_init(empty0 + 8 * i, I2 ? 2 : 1);
    }
    for (int i = 0; i < kSlabMaxAcc; ++i) {
      mbar_init(tfull0 + 8 * i, I2 ? 2 : 1);
      mbar_init(tempty0 + 8 * i, 128 * kEG);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kSlabMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- one-time: MMA table, weights (MMA order), BatchNorm scale/shift and the staging table into shared memory --------
  if (!WS && p.zmerge) {
    // [spatial op][K half][focal block j = 0,1,2 (dz = +1, 0, -1 -> output slices z-1, z, z+1)][N][8]
    const int N3 = 3 * p.N;
    const uint4* wg = reinterpret_cast<const uint4*>(p.wslab);
    for (int ph = 0; ph < (Z2 ? 2 : 1); ++ph) {
      const int total = 2 * p.zTp[ph] * N3;
      for (int i = threadIdx.x; i < total; i += kThreads) {
        const int op = i / (2 * N3), rem = i - op * 2 * N3;
        const int h = rem / N3, jn = rem - h * N3, j = jn / p.N, r = jn - j * p.N;
        const int src = p.wsrc[2 * (p.g[3 * ph + 2 - j] + op) + h];
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (src >= 0) v = __ldg(wg + (size_t)src * p.N + r);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(w_s + 16 * (p.zwo[ph] + i)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
    }
  } else if (!WS) {
    const int total = 2 * p.nwu * p.N;  // 16-byte rows
    const uint4* wg = reinterpret_cast<const uint4*>(p.wslab);
#pragma unroll 4
    for (int i = threadIdx.x; i < total; i += kThreads) {
      const int blk = i / p.N, r = i - blk * p.N;
      const int src = p.wsrc[2 * p.wop[blk >> 1] + (blk & 1)];
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (src >= 0) v = __ldg(wg + (size_t)src * p.N + r);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(w_s + 16 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  }
  for (int i = threadIdx.x; i < p.N; i += kThreads) {
    ss[i] = p.epi.scale ? __ldg(p.epi.scale + i) : 1.f;
    ss[p.N + i] = p.epi.shift ? __ldg(p.epi.shift + i) : 0.f;
    const int pc = p.epi.proj_c ? p.epi.proj_c : p.epi.cstore;   // (x-folded rows hold cstore / proj_c pixels)
    ss[2 * p.N + i] = (p.epi.proj_w && i < p.epi.cstore) ? __ldg(p.epi.proj_w + i % pc) : 0.f;
  }
  // (table order = global address order: consecutive producer threads copy consecutive 16-byte pieces — the chunks of a pixel, then
  // the next pixel of the row — whatever view the pixel belongs to; the shared-memory destination is free-form anyway.  With
  // strided views (stride-2 and x-folded layers) a view-major order would have every thread touch its own 32-byte sector.)
  for (int e = threadIdx.x; e < p.nelem; e += kThreads) {
    const int c = e % p.nchunk, t = e / p.nchunk;
    const int gxi = t % (p.stx * p.RX), gyi = t / (p.stx * p.RX);
    const int rx = gxi / p.stx, vx = gxi % p.stx, ry = gyi / p.sty, vy = gyi % p.sty;
    const int pixo = (vy * p.stx + vx) * p.VB + (ry * p.RX + rx) * 16;   // byte offset inside the chunk's block
    const int gy = p.sty * p.oy + gyi, gx = p.stx * p.ox + gxi;
    const bool second = c >= p.nch0;
    const int C = second ? p.C1 : p.C0, cc = second ? c - p.nch0 : c;
    SlabElem el;
    el.rel = (gy * p.IW + gx) * C + cc * 8;
    el.dst16 = (uint16_t)(((c * p.CPS + pixo) >> 4) | (second ? 0x8000 : 0));
    el.gy = (uint8_t)(gy + kSlabElemBias);
    el.gx = (uint8_t)(gx + kSlabElemBias);
    const_cast<SlabElem*>(elems)[e] = el;
  }
  fence_proxy_async();
  fence_before();
  __syncthreads();
  fence_after();
  if (threadIdx.x == 0) DFF_TR(1, 63);
  pdl_wait();      // everything above touched only weights and tables; from here on the previous layer's output is read
  if (threadIdx.x == 0) DFF_TR(2, 63);
  const uint32_t tmem_base = tmem_base_s;
  const int hz = p.hz;

  if (warp < kSlabMmaWarp && p.tma) {
    // =============================== producer: one tiled TMA load per slice plane ===============================
    // box = 8 channels x RX pixels x RY rows lands in shared memory as [row][pixel][16 B] — the plane layout itself; out-of-bounds
    // rows / columns (the convolution's zero padding) are filled by the TMA unit.  One thread, no per-element work.
    if (threadIdx.x == 0) {
      for (int i = 0; i < (p.C1 ? 8 : 4); ++i)
        if ((i & 3) < p.nviews) prefetch_tmap(&p.tmap[i]);
      int slot = 0;
      uint32_t ephase = 1;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        if (item + (int)gridDim.x >= p.nitems) pdl_trigger();
        int r = item;
        const int isp = r % p.nsplit; r /= p.nsplit;
        const int tx0 = (r % p.tilesX) * kSlabTW * p.stx; r /= p.tilesX;
        const int ty0 = (r % p.tilesY) * p.tile_sy;
        const int b = r / p.tilesY;
        const int s_begin = isp * p.slen, s_end = min(p.S, s_begin + p.slen);
        const int zlo = max(0, s_begin - hz), zhi = min(p.S, s_end + hz);
        for (int z = zlo; z < zhi; ++z) {
          mbar_wait(empty0 + 8 * slot, ephase);
          const uint32_t total = (uint32_t)(p.nchunk * p.nviews * p.RY * p.RX * 16);
          mbar_expect_tx(full0 + 8 * slot, total);
          if (!(p.exp & 1)) {
            const uint32_t dst0 = planes_s + slot * p.plane_bytes;
            const int x0 = tx0 / p.stx + p.ox, y0 = ty0 / p.sty + p.oy, bz = b * p.S + z;   // view coordinates of the staged region
            if (p.wr) {
              const int cb = p.rb >> 1, nblk = (p.nchunk * 8) / cb;
              for (int blk = 0; blk < nblk; ++blk) {
                const int cabs = blk * cb, src = cabs >= p.C0 ? 1 : 0, cc = src ? cabs - p.C0 : cabs;
                for (int v = 0; v < p.nviews; ++v)
                  tma_load_4d(dst0 + blk * p.CPS + v * p.VB, &p.tmap[src * 4 + v], full0 + 8 * slot, cc, x0, y0, bz);
              }
            } else
            for (int c = 0; c < p.nchunk; ++c) {
              const int src = c >= p.nch0 ? 1 : 0, cc = src ? c - p.nch0 : c;
              for (int v = 0; v < p.nviews; ++v)
                if (p.tmerge[src]) tma_load_4d(dst0 + c * p.CPS + v * p.VB, &p.tmap[src * 4 + v], full0 + 8 * slot, x0 * 8, 0, y0, bz);
                else tma_load_4d(dst0 + c * p.CPS + v * p.VB, &p.tmap[src * 4 + v], full0 + 8 * slot, cc * 8, x0, y0, bz);
            }
          } else {
            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * slot), "r"(total) : "memory");
          }
          if (++slot == p.NP) { slot = 0; ephase ^= 1; }
        }
      }
    } else {
      pdl_trigger();
    }
  } else if (warp < kSlabMmaWarp) {
    // =============================== producers: stage slice planes with cp.async ===============================
    const int ptid = threadIdx.x;
    const char* const base0 = reinterpret_cast<const char*>(p.in0);
    const char* const base1 = reinterpret_cast<const char*>(p.in1);
    int slot = 0;             // ring slot of the next plane
    int np = 0;
    uint32_t ephase = 1;      // parity to wait for on its `empty` barrier (a fresh barrier passes parity 1)
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      if (item + (int)gridDim.x >= p.nitems) pdl_trigger();   // last item of this CTA: the next kernel may start filling drained SMs
      int r = item;
      const int isp = r % p.nsplit; r /= p.nsplit;
      const int tx0 = (r % p.tilesX) * kSlabTW * p.stx; r /= p.tilesX;   // tile origin in input coordinates
      const int ty0 = (r % p.tilesY) * p.tile_sy;
      const int b = r / p.tilesY;
      const int s_begin = isp * p.slen, s_end = min(p.S, s_begin + p.slen);
      const int zlo = max(0, s_begin - hz), zhi = min(p.S, s_end + hz);
      for (int z = zlo; z < zhi; ++z) {
        mbar_wait(empty0 + 8 * slot, ephase);
        if (warp == 0) DFF_TR(5, np);
        const uint32_t dst0 = planes_s + slot * p.plane_bytes;
        const size_t pix0 = (((size_t)b * p.S + z) * p.IH + ty0) * p.IW + tx0;
        const char* const o0 = base0 + pix0 * p.C0 * 2;
        const char* const o1 = base1 + pix0 * p.C1 * 2;
        if (!(p.exp & 1))
        for (int e = ptid; e < p.nelem; e += kSlabProducers) {
          const SlabElem el = elems[e];
          const int gy = ty0 + (int)el.gy - kSlabElemBias, gx = tx0 + (int)el.gx - kSlabElemBias;
          const bool ok = gy >= 0 && gy < p.IH && gx >= 0 && gx < p.IW;
          const char* src = ((el.dst16 & 0x8000) ? o1 : o0) + (ptrdiff_t)el.rel * 2;
          cp_async16(dst0 + ((uint32_t)(el.dst16 & 0x7fff) << 4), ok ? src : base0, ok ? 16u : 0u);
        }
        // asynchronous publish: the slot's `full` barrier receives this thread's arrival when its copies have landed, so the
        // producers never wait for data — they run ahead as far as the ring has free slots
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full0 + 8 * slot) : "memory");
        if (++slot == p.NP) { slot = 0; ephase ^= 1; }
        if (warp == 0) DFF_TR(6, np);
        ++np;
      }
    }
  } else if (warp == kSlabMmaWarp || (I2 && warp == kThreads / 32 - 1)) {
    const int issuer = (I2 && warp != kSlabMmaWarp) ? 1 : 0;   // (I2: this warp issues the MMAs of output phase `issuer` only)
    // =============================== MMA issuer ===============================
    // The whole warp walks the schedule with warp-uniform values (everything derives from blockIdx and kernel parameters, the
    // MMA table is read from the parameter bank with a uniform index), so descriptors live in uniform registers and the
    // tcgen05.mma stream is issued back to back by one elected lane — no per-instruction lane serialisation.  A single warp
    // pays the full latency of every dependent instruction, so the per-slice bookkeeping is kept to a rotating window of three
    // plane descriptors (slices s-1, s, s+1) and the per-MMA work to two 64-bit adds.
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
    // descriptor high words: SBO (8-row group stride) | version 1 | no swizzle
    const uint64_t a_hi = ((uint64_t)(((uint32_t)p.sbo >> 4) | (1u << 14)) << 32) |
                          ((uint64_t)(p.wr ? (p.rb == 32 ? 6u : (p.rb == 64 ? 4u : 2u)) : 0u) << 61);   // (wide rows: 32B / 64B / 128B swizzle)
    const uint64_t b_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint64_t b_step = (uint32_t)(p.N * 32) >> 4;
    const uint64_t bd_base = b_hi | (uint64_t)((w_s >> 4) | (((uint32_t)(p.N * 16) >> 4) << 16));
    const uint32_t planes16 = planes_s >> 4, pb16 = (uint32_t)p.plane_bytes >> 4;
    const int NP = p.NP, ngrp = 3 * p.nph;
    int wslot = 0, sc = 0;
    uint32_t wphase = 0;
    int wsl = 0;          // (WS) slot / parity of the streamed-weight ring
    uint32_t wsph = 0;
    uint64_t a_prev = 0, a_cur = 0, a_next = 0;   // descriptor bases of the planes of slices s-1, s, s+1
    uint32_t e_prev = 0, e_cur = 0, e_next = 0;   // their `empty` barriers
    if (p.zmerge) {
      // ---- focal-merged schedule: one pass over the planes of an item, every plane multiplied into the slices it feeds ----
      const uint32_t N = (uint32_t)p.N, smask = (uint32_t)p.nslot - 1u;
      const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
      const uint64_t bdz = b_hi | (uint64_t)((w_s >> 4) | (((uint32_t)(3 * p.N * 16) >> 4) << 16));   // LBO = one K half of 3N rows
      const uint64_t bz_step = (uint32_t)(3 * p.N * 32) >> 4, bz_blk = (uint32_t)(p.N * 16) >> 4;
      int q0 = 0;          // running index of the first slice of the current item (slot = index & smask, use count = index >> lgslot)
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        if (item + (int)gridDim.x >= p.nitems) pdl_trigger();
        const int isp = item % p.nsplit;
        const int s_begin = isp * p.slen, s_end = min(p.S, s_begin + p.slen);
        const int zlo = max(0, s_begin - 1), zhi = min(p.S, s_end + 1);
        for (int z = zlo; z < zhi; ++z) {
          mbar_wait(full0 + 8 * wslot, wphase);
          const uint64_t a_pl = a_hi | (uint64_t)(planes16 + (uint32_t)wslot * pb16);
          const uint32_t e_pl = empty0 + 8 * wslot;
          if (++wslot == NP) { wslot = 0; wphase ^= 1; }
          fence_proxy_async();
          const int lo = max(z - 1, s_begin), hi = min(z + 1, s_end - 1);        // output slices this plane feeds
          const int nlo = (z == zlo) ? lo : (hi == z + 1 ? hi : hi + 1);          // first slice whose accumulator starts at this plane
          for (int sn = nlo; sn <= hi; ++sn) {     // a fresh slot: the epilogue must have drained its previous slice
            const uint32_t qi = (uint32_t)(q0 + sn - s_begin);
            mbar_wait(tempty0 + 8 * (qi & smask), ((qi >> p.lgslot) & 1u) ^ 1u);
          }
          fence_after();
          // The plane's MMAs as at most two (accumulator, B sub-block, N') segments per spatial op — the slices it feeds are consecutive
          // accumulator slots unless the ring wraps — computed once per plane (warp-uniform), so that the issue loop below is a
          // descriptor add and one or two tcgen05.mma per op.  Spatial op 0 starts the accumulator of the newest slice(s) (accumulate = 0).
          // Two-phase layers repeat this per phase on the phase's own ring of slots, MMAs and weight block.
#pragma unroll 1
          for (int zph = 0; zph < (Z2 ? 2 : 1); ++zph) {
          const uint32_t tb_ph = tmem_base + (uint32_t)(zph * p.nslot) * N;
          uint32_t sd[2], si[2];
          uint64_t sb[2];
          int nseg = 0;
          {
            int r0 = lo;
            while (r0 <= hi) {
              const uint32_t slot = (uint32_t)(q0 + r0 - s_begin) & smask;
              int cnt = hi - r0 + 1;
              if ((int)slot + cnt > p.nslot) cnt = p.nslot - (int)slot;
              sd[nseg] = tb_ph + slot * N;
              si[nseg] = idesc_base | ((((uint32_t)cnt * N) >> 3) << 17);
              sb[nseg] = (uint64_t)(r0 - (z - 1)) * bz_blk;
              ++nseg;
              r0 += cnt;
            }
          }
          // the plane's spatial MMAs, block by block: ONE block when the weights are resident; with streamed weights the whole warp
          // waits for a block's ring slot, the elected lane issues its MMAs and hands the slot back when they have read it
          const int g1 = p.g[3 * zph + 1];
          const int nblk = WS ? p.nblk : 1;
#pragma unroll 1
          for (int bk = 0; bk < nblk; ++bk) {
            const int i_lo = WS ? (int)p.bop[bk] : 0, i_hi = WS ? i_lo + (int)p.bn[bk] : p.zTp[zph];
            uint64_t bd = bdz + (WS ? (uint64_t)((uint32_t)(wsl * p.wslot_bytes) >> 4) : (uint64_t)(uint32_t)p.zwo[zph]);   // weights of op i_lo
            if (WS) {
              mbar_wait(wfull0 + 8 * wsl, wsph);
              fence_after();
            }
            if (leader) {
              int i = i_lo;
              if (i == 0) {   // spatial op 0: runs [lo, nlo-1] (accumulating) and [nlo, hi] (fresh), each split where the ring wraps
                const uint64_t ad = a_pl + p.tab[g1];
                for (int part = 0; part < 2; ++part) {
                  int r0 = part ? nlo : lo;
                  const int r1 = part ? hi : nlo - 1;
                  while (r0 <= r1) {
                    const uint32_t slot = (uint32_t)(q0 + r0 - s_begin) & smask;
                    int cnt = r1 - r0 + 1;
                    if ((int)slot + cnt > p.nslot) cnt = p.nslot - (int)slot;
                    umma(tb_ph + slot * N, ad, bd + (uint64_t)(r0 - (z - 1)) * bz_blk, idesc_base | ((((uint32_t)cnt * N) >> 3) << 17),
                         part ? 0u : 1u);
                    r0 += cnt;
                  }
                }
                i = 1;
                bd += bz_step;
              }
              if (nseg == 1) {
                const uint32_t d0 = sd[0], i0 = si[0];
                const uint64_t b0 = sb[0];
#pragma unroll 1
                for (; i < i_hi; ++i, bd += bz_step) umma_acc(d0, a_pl + p.tab[g1 + i], bd + b0, i0);
              } else {
                const uint32_t d0 = sd[0], i0 = si[0], d1 = sd[1], i1 = si[1];
                const uint64_t b0 = sb[0], b1 = sb[1];
#pragma unroll 1
                for (; i < i_hi; ++i, bd += bz_step) {
                  const uint64_t ad = a_pl + p.tab[g1 + i];
                  umma_acc(d0, ad, bd + b0, i0);
                  umma_acc(d1, ad, bd + b1, i1);
                }
              }
              if (WS) umma_commit(wempty0 + 8 * wsl);
            }
            if (WS) {
              if (++wsl == p.nwslots) { wsl = 0; wsph ^= 1; }
            }
          }
          }
          if (leader) {
            umma_commit(e_pl);                                   // the plane is not needed again
            if (z - 1 >= s_begin) umma_commit(tfull0 + 8 * ((uint32_t)(q0 + z - 1 - s_begin) & smask));        // slice z-1 is complete
            if (z == zhi - 1 && z < s_end) umma_commit(tfull0 + 8 * ((uint32_t)(q0 + z - s_begin) & smask));   // last plane: slice z too
          }
          __syncwarp();
        }
        q0 += s_end - s_begin;
      }
    } else
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      if (item + (int)gridDim.x >= p.nitems) pdl_trigger();
      const int isp = item % p.nsplit;
      const int s_begin = isp * p.slen, s_end = min(p.S, s_begin + p.slen);
      const int zlo = max(0, s_begin - hz), zhi = min(p.S, s_end + hz);
      int zn = zlo;  // next plane of this item to take from the ring (planes arrive in order)
      for (int s = s_begin; s < s_end; ++s, ++sc) {
        const int zlast = min(s + hz, zhi - 1);
        while (zn <= zlast) {
          mbar_wait(full0 + 8 * wslot, wphase);
          a_prev = a_cur; a_cur = a_next; a_next = a_hi | (uint64_t)(planes16 + (uint32_t)wslot * pb16);
          e_prev = e_cur; e_cur = e_next; e_next = empty0 + 8 * wslot;
          if (++wslot == NP) { wslot = 0; wphase ^= 1; }
          ++zn;
        }
        DFF_TR(0, sc);
        fence_proxy_async();  // cp.async wrote the planes through the generic proxy; the MMAs read them through the async proxy
        const bool have_next = zn - 1 > s;
        if (!have_next) { a_prev = a_cur; a_cur = a_next; e_prev = e_cur; e_cur = e_next; }  // the newest plane is slice s itself
        const int buf = sc & 1;
        mbar_wait(tempty0 + 8 * buf, ((sc >> 1) & 1) ^ 1);
        fence_after();
        DFF_TR(1, sc);
        if (WS) {
          // streamed weights: the whole warp waits for a block's ring slot (warp-uniform control flow keeps the descriptors in uniform
          // registers), the elected lane issues the block's MMAs and hands the slot back when they have read it
          uint32_t dacc = tmem_base + buf * p.nph * p.N;
          uint32_t acc = 0;
          int k = 0;
#pragma unroll 1
          for (int gi = 0; gi < ngrp; ++gi) {
            const int b0 = p.gb[gi], b1 = p.gbe[gi];
            const int z = s + k - 1;
            if (b1 > b0 && z >= 0 && z < p.S && !(p.exp & 4)) {
              const uint64_t ad0 = k == 0 ? a_prev : (k == 1 ? a_cur : a_next);
#pragma unroll 1
              for (int b = b0; b < b1; ++b) {
#ifdef DFF_SLAB_TRACE
                const long long tw0 = clock64();
#endif
                mbar_wait(wfull0 + 8 * wsl, wsph);
#ifdef DFF_SLAB_TRACE
                if (p.trace && blockIdx.x == 0 && sc < 64 && lane == 0) p.trace[sc * 8 + 7] += clock64() - tw0;   // clocks spent waiting for weights
#endif
                fence_after();
                if (leader) {
                  const int o0 = p.bop[b];
                  int n = p.bn[b];
                  uint64_t bd = bd_base + (uint64_t)((uint32_t)(wsl * p.wslot_bytes) >> 4);
                  const ulonglong2* tq = reinterpret_cast<const ulonglong2*>(p.tab + o0);  // blocks start on quad boundaries
                  ulonglong2 t01 = tq[0], t23 = tq[1];
#pragma unroll 1
                  for (; n >= 4; n -= 4) {
                    tq += 2;
                    const ulonglong2 n01 = tq[0], n23 = tq[1];
                    umma(dacc, ad0 + t01.x, bd, idesc, acc);
                    umma_acc(dacc, ad0 + t01.y, bd + b_step, idesc);
                    umma_acc(dacc, ad0 + t23.x, bd + 2 * b_step, idesc);
                    umma_acc(dacc, ad0 + t23.y, bd + 3 * b_step, idesc);
                    acc = 1;
                    bd += 4 * b_step;
                    t01 = n01; t23 = n23;
                  }
                  if (n > 0) {
                    umma(dacc, ad0 + t01.x, bd, idesc, acc);
                    if (n > 1) umma_acc(dacc, ad0 + t01.y, bd + b_step, idesc);
                    if (n > 2) umma_acc(dacc, ad0 + t23.x, bd + 2 * b_step, idesc);
                  }
                  umma_commit(wempty0 + 8 * wsl);
                }
                acc = 1;
                if (++wsl == p.nwslots) { wsl = 0; wsph ^= 1; }
              }
            }
            if (++k == 3) { k = 0; dacc += p.N; acc = 0; }
          }
        }
        if (leader) {
          uint32_t dacc = tmem_base + buf * p.nph * p.N;
          uint32_t acc = 0;
          int k = 0;
          if (!WS)
#pragma unroll 1
          for (int gi = 0; gi < ngrp; ++gi) {
            const int i0 = p.g[gi], n0 = p.ge[gi] - i0;
            const int z = s + k - 1;
            if (n0 > 0 && z >= 0 && z < p.S && !(p.exp & 4) && (!I2 || gi / 3 == issuer)) {  // (focal-dimension zero padding: nothing to multiply)
              const uint64_t ad0 = k == 0 ? a_prev : (k == 1 ? a_cur : a_next);
              uint64_t bd = bd_base + (uint32_t)p.gw[gi] * b_step;
              const ulonglong2* tq = reinterpret_cast<const ulonglong2*>(p.tab + i0);  // groups start on quad boundaries
              ulonglong2 t01 = tq[0], t23 = tq[1];
              int n = n0;
#pragma unroll 1
              for (; n >= 4; n -= 4) {
                tq += 2;
                const ulonglong2 n01 = tq[0], n23 = tq[1];  // next quad (the table has spare quads at the end)
                umma(dacc, ad0 + t01.x, bd, idesc, acc);
                umma_acc(dacc, ad0 + t01.y, bd + b_step, idesc);
                umma_acc(dacc, ad0 + t23.x, bd + 2 * b_step, idesc);
                umma_acc(dacc, ad0 + t23.y, bd + 3 * b_step, idesc);
                acc = 1;
                bd += 4 * b_step;
                t01 = n01; t23 = n23;
              }
              if (n > 0) {
                umma(dacc, ad0 + t01.x, bd, idesc, acc);
                acc = 1;
                if (n > 1) umma_acc(dacc, ad0 + t01.y, bd + b_step, idesc);
                if (n > 2) umma_acc(dacc, ad0 + t23.x, bd + 2 * b_step, idesc);
              }
            }
            if (++k == 3) { k = 0; dacc += p.N; acc = 0; }  // next output phase: next accumulator
          }
          umma_commit(tfull0 + 8 * buf);
          // ring slots whose last reader was this slice
          if (hz) {
            if (s > zlo) umma_commit(e_prev);
            if (s == s_end - 1) {
              umma_commit(e_cur);
              if (have_next) umma_commit(e_next);
            }
          } else {
            umma_commit(e_cur);
          }
        }
        __syncwarp();
        DFF_TR(2, sc);
      }
    }
  } else if (WS && warp == kThreads / 32 - 1) {
    // =============================== weight producer (streaming variant) ===============================
    // walks exactly the issuer's schedule; one bulk copy (global -> shared, completion on the slot's `full` barrier) per block
    pdl_trigger();
    if (lane == 0 && p.zmerge) {
      const char* const wg = reinterpret_cast<const char*>(p.wslab);     // (the focal-merged layout: launch_conv_slab passes a.wz)
      int wsl = 0;
      uint32_t eph = 1;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const int isp = item % p.nsplit;
        const int s_begin = isp * p.slen, s_end = min(p.S, s_begin + p.slen);
        const int zlo = max(0, s_begin - 1), zhi = min(p.S, s_end + 1);
        for (int z = zlo; z < zhi; ++z)
          for (int b = 0; b < p.nblk; ++b) {
            mbar_wait(wempty0 + 8 * wsl, eph);
            const uint32_t bytes = (uint32_t)p.bn[b] * (uint32_t)p.N * 96u;
            mbar_expect_tx(wfull0 + 8 * wsl, bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             w_s + (uint32_t)(wsl * p.wslot_bytes)),
                         "l"(wg + (size_t)p.bop[b] * p.N * 96), "r"(bytes), "r"(wfull0 + 8 * wsl)
                         : "memory");
            if (++wsl == p.nwslots) { wsl = 0; eph ^= 1; }
          }
      }
    } else if (lane == 0) {
      const char* const wg = reinterpret_cast<const char*>(p.wslab);
      const int ngrp = 3 * p.nph;
      int wsl = 0;
      uint32_t eph = 1;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const int isp = item % p.nsplit;
        const int s_begin = isp * p.slen, s_end = min(p.S, s_begin + p.slen);
        for (int s = s_begin; s < s_end; ++s) {
          int k = 0;
          for (int gi = 0; gi < ngrp; ++gi) {
            const int b0 = p.gb[gi], b1 = p.gbe[gi];
            const int z = s + k - 1;
            if (b1 > b0 && z >= 0 && z < p.S && !(p.exp & 4)) {
              for (int b = b0; b < b1; ++b) {
                mbar_wait(wempty0 + 8 * wsl, eph);
                const uint32_t bytes = (uint32_t)p.bn[b] * (uint32_t)p.N * 32u;
                const char* src = wg + (size_t)p.wsrc[2 * p.bop[b]] * p.N * 16;
                mbar_expect_tx(wfull0 + 8 * wsl, bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 w_s + (uint32_t)(wsl * p.wslot_bytes)),
                             "l"(src), "r"(bytes), "r"(wfull0 + 8 * wsl)
                             : "memory");
                if (++wsl == p.nwslots) { wsl = 0; eph ^= 1; }
              }
            }
            if (++k == 3) k = 0;
          }
        }
      }
    }
  } else {
    // =============================== epilogue ===============================
    // one compiled loop per operand pattern of the network's layers (see tc_epilogue_fast); mode 0 = the generic epilogue
    const EpiArgs& e = p.epi;
    const int res = e.res_pre ? 1 : (e.res_post ? 2 : 0);
    const bool aux = e.out_aux != nullptr, proj = e.proj_w != nullptr;
    const uint32_t ss_s = smem_u32(ss);
    if (e.out_f32 || (e.res_pre && e.res_post) || (res == 1 && aux) || (res == 2 && (aux || e.relu)) || (proj && !aux && res != 2)) slab_epilogue<0, false, 0, false, false>(p, tmem_base, tfull0, tempty0, ss, ss_s, kEG);
    else if (aux) { if (proj) slab_epilogue<1, false, 0, true, true>(p, tmem_base, tfull0, tempty0, ss, ss_s, kEG); else slab_epilogue<1, false, 0, true, false>(p, tmem_base, tfull0, tempty0, ss, ss_s, kEG); }
    else if (res == 2) { if (proj) slab_epilogue<1, false, 2, false, true>(p, tmem_base, tfull0, tempty0, ss, ss_s, kEG); else slab_epilogue<1, false, 2, false, false>(p, tmem_base, tfull0, tempty0, ss, ss_s, kEG); }
    else if (res == 1) { if (e.relu) slab_epilogue<1, true, 1, false, false>(p, tmem_base, tfull0, tempty0, ss, ss_s, kEG); else slab_epilogue<1, false, 1, false, false>(p, tmem_base, tfull0, tempty0, ss, ss_s, kEG); }
    else { if (e.relu) slab_epilogue<1, true, 0, false, false>(p, tmem_base, tfull0, tempty0, ss, ss_s, kEG); else slab_epilogue<1, false, 0, false, false>(p, tmem_base, tfull0, tempty0, ss, ss_s, kEG); }
  }
  fence_before();
  __syncthreads();
  if (threadIdx.x == 0) DFF_TR(3, 63);
  if (warp == kSlabMmaWarp) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side: plan (geometry, MMA table, ring depth) + launch
// ------------------------------------------------------------------------------------------------------------------
// `ptaps`/`nph`: tap table per output phase (nph = 1: a.taps; nph = 4: the parity phases of a transposed convolution).
// DFF_B200_DEBUG_PLAN=1 (works in a dry run without a GPU): report which rule of the plan rejected a layer
#define return_false do { if (getenv("DFF_B200_DEBUG_PLAN")) fprintf(stderr, "slab_plan: rejected at line %d\n", __LINE__); return false; } while (0)
static bool slab_plan(const ConvArgs& a, const TapTable* ptaps, int nph, int Ntc, int num_sms, SlabParams& p, size_t* smem_out,
                      int* occ_out) {
  memset(&p, 0, sizeof(p));
  if (a.C0 % 8 || a.C1 % 8 || a.C0 < 8) return_false;
  if (Ntc < 16 || Ntc > 128 || Ntc % 16) return_false;
  if ((a.isy != 1 && a.isy != 2) || (a.isx != 1 && a.isx != 2 && a.isx != 4) || a.isy * a.isx > 4) return_false;
  const int nchunk = (a.C0 + a.C1) / 8;
  if (nchunk != 1 && (nchunk & 1)) return_false;
  p.in0 = a.in0; p.in1 = a.in1; p.C0 = a.C0; p.C1 = a.C1; p.nchunk = nchunk; p.nch0 = a.C0 / 8;
  p.B = a.B; p.S = a.S; p.IH = a.IH; p.IW = a.IW; p.sty = a.row_step > 0 ? 1 : a.isy; p.stx = a.isx;
  p.nviews = p.sty * p.stx;
  for (int v = 0; v < 4; ++v) { p.vpy[v] = v / p.stx; p.vpx[v] = v % p.stx; }
  // ---- taps in view coordinates -------------------------------------------------------------------------------------
  struct VT { int dz, view, vy, vx, widx, ph, kmask; };
  std::vector<VT> vt;
  int vymin = 1000, vymax = -1000, vxmin = 1000, vxmax = -1000, dzmin = 1000, dzmax = -1000;
  if (nph != 1 && nph != 2 && nph != 4) return_false;
  p.nph = nph;
  for (int ph = 0; ph < 4; ++ph) {
    p.phy[ph] = nph == 4 ? (ph >> 1) : (nph == 2 ? ph : a.ooy);   // nph == 2: the two row phases of an x-folded transposed conv
    p.phx[ph] = nph == 4 ? (ph & 1) : (nph == 2 ? 0 : a.oox);
  }
  for (int ph = 0; ph < nph; ++ph)
  for (int t = 0; t < ptaps[ph].n; ++t) {
    const TapTable& tt = ptaps[ph];
    VT x;
    x.ph = ph;
    x.dz = tt.dz[t]; x.widx = tt.widx[t];
    x.kmask = (a.use_kmask && nph == 1 && ptaps == &a.taps) ? a.tap_kmask[t] : 0;
    int dy = tt.dy[t], dx = tt.dx[t];
    {  // input coordinate sty*o + d  ->  view (d mod sty, d mod stx), view coordinate o + floor(d / st)
      const int py = ((dy % p.sty) + p.sty) % p.sty, px = ((dx % p.stx) + p.stx) % p.stx;
      x.view = py * p.stx + px; x.vy = (dy - py) / p.sty; x.vx = (dx - px) / p.stx;
    }
    vymin = std::min(vymin, x.vy); vymax = std::max(vymax, x.vy);
    vxmin = std::min(vxmin, x.vx); vxmax = std::max(vxmax, x.vx);
    dzmin = std::min(dzmin, x.dz); dzmax = std::max(dzmax, x.dz);
    vt.push_back(x);
  }
  if (dzmin < -1 || dzmax > 1) return_false;
  p.hz = (dzmin < 0 || dzmax > 0) ? 1 : 0;
  p.oy = vymin; p.ox = vxmin;
  p.RY = (a.row_step > 0 ? (kSlabTH - 1) * a.row_step + 1 : kSlabTH) + (vymax - vymin);
  p.RX = kSlabTW + (vxmax - vxmin);
  // DFF_SLAB_EXPERIMENT=aligned: timing experiment only (wrong results) — force 128-byte aligned core matrices
  static const bool exp_aligned = getenv("DFF_SLAB_EXPERIMENT") && !strcmp(getenv("DFF_SLAB_EXPERIMENT"), "aligned");
  p.exp = (getenv("DFF_SLAB_EXPERIMENT") && !exp_aligned) ? atoi(getenv("DFF_SLAB_EXPERIMENT")) : 0;
  if (exp_aligned) p.RX = (p.RX + 7) & ~7;
  p.VB = (p.RY * p.RX * 16 + 127) & ~127;   // (each (chunk, view) block is the destination of one TMA box: 128-byte aligned)
  p.CPS = p.nviews * p.VB;
  p.plane_bytes = nchunk * p.CPS;
  {
    static const bool no_wr = getenv("DFF_B200_NO_WR") != nullptr;            // (A/B knob)
    static const int tma_mode_wr = getenv("DFF_B200_SLAB_TMA") ? atoi(getenv("DFF_B200_SLAB_TMA")) : 2;
    const int cb = a.C0 >= 64 ? 64 : a.C0;
    const bool tma_ok = tma_mode_wr > 1 && p.RX <= 256 && p.RY <= 256 && a.IW % p.stx == 0 && a.IH % p.sty == 0;
    if (!no_wr && !exp_aligned && tma_ok && a.row_step == 0 && (cb == 16 || cb == 32 || cb == 64) && a.C0 % cb == 0 && a.C1 % cb == 0) {
      p.wr = 1; p.rb = 2 * cb;
      p.VB = (p.RY * p.RX * p.rb + 1023) & ~1023;
      p.CPS = p.nviews * p.VB;
      p.plane_bytes = ((a.C0 + a.C1) / cb) * p.CPS;
    }
  }
  if (p.sty * p.oy < -kSlabElemBias || p.stx * p.ox < -kSlabElemBias || p.sty * (p.oy + p.RY) > 255 - kSlabElemBias ||
      p.stx * (p.ox + p.RX) > 255 - kSlabElemBias)
    return_false;   // (staging-table coordinates are biased bytes)
  p.tile_sy = a.row_step > 0 ? kSlabTH * a.row_step : kSlabTH * p.sty;
  p.sbo = (a.row_step > 0 ? a.row_step : 1) * p.RX * (p.wr ? p.rb : 16);
  if ((p.CPS >> 4) >= (1 << 14) || p.sbo >= (1 << 18)) return_false;
  auto aoff = [&](const VT& x, int chunk) {
    int o = chunk * p.CPS + x.view * p.VB + ((x.vy - p.oy) * p.RX + (x.vx - p.ox)) * 16;
    if (exp_aligned) o &= ~127;
    return o;
  };
  // ---- MMA table, grouped by focal offset dz = -1, 0, +1 ---------------------------------------------------------------
  int nops = 0;
  for (int ph = 0; ph < nph; ++ph)
  for (int k = 0; k < 3; ++k) {
    while (nops & 3) {  // groups start on a 16-byte table boundary (the issuer reads four entries per load); pads never run
      if (nops >= kSlabMaxOps) return_false;
      p.tab[nops] = 0; p.wsrc[2 * nops] = -1; p.wsrc[2 * nops + 1] = -1;
      ++nops;
    }
    p.g[ph * 3 + k] = nops;
    std::vector<VT> grp;
    for (auto& x : vt)
      if (x.dz == k - 1 && x.ph == ph) grp.push_back(x);
    if (nchunk > 1) {
      for (auto& x : grp)
        for (int j = 0; j < nchunk; j += 2) {
          if (x.kmask && !((x.kmask >> (j >> 1)) & 1)) continue;   // (a K step of all-zero weights: banded x-grouped forms)
          if (nops >= kSlabMaxOps) return_false;
          if (p.wr) {   // 16 channels = one 32-byte K step of the pixel's row in block (8 j) / cb
            const int cb = p.rb >> 1, c = 8 * j, blk = c / cb;
            const int o = blk * p.CPS + x.view * p.VB + ((x.vy - p.oy) * p.RX + (x.vx - p.ox)) * p.rb + (c - blk * cb) * 2;
            p.tab[nops] = (uint32_t)(o >> 4) | (1u << 16);
          } else
          p.tab[nops] = (uint32_t)(aoff(x, j) >> 4) | ((uint32_t)(p.CPS >> 4) << 16);
          p.wsrc[2 * nops] = (int16_t)(x.widx * nchunk + j);
          p.wsrc[2 * nops + 1] = (int16_t)(x.widx * nchunk + j + 1);
          ++nops;
        }
    } else {  // 8-channel tensor: K = 16 is two taps; an odd tap is paired with zero weights
      std::sort(grp.begin(), grp.end(), [&](const VT& l, const VT& r) { return aoff(l, 0) < aoff(r, 0); });
      for (size_t i = 0; i < grp.size(); i += 2) {
        if (nops >= kSlabMaxOps) return_false;
        const bool pair = i + 1 < grp.size();
        const int o1 = aoff(grp[i], 0), o2 = pair ? aoff(grp[i + 1], 0) : o1;
        if (((o2 - o1) >> 4) >= (1 << 14)) return_false;
        p.tab[nops] = (uint32_t)(o1 >> 4) | ((uint32_t)((o2 - o1) >> 4) << 16);
        p.wsrc[2 * nops] = (int16_t)grp[i].widx;
        p.wsrc[2 * nops + 1] = pair ? (int16_t)grp[i + 1].widx : (int16_t)-1;
        ++nops;
      }
    }
    p.ge[ph * 3 + k] = nops;
  }
  p.nops = nops;
  if (nops == 0) return_false;
  p.N = Ntc;
  // weight slots: a group whose weight sequence repeats an earlier group's shares its slots
  {
    int nwu = 0;
    for (int gi = 0; gi < 3 * nph; ++gi) {
      const int n = p.ge[gi] - p.g[gi];
      int same = -1;
      for (int gj = 0; gj < gi && same < 0; ++gj) {
        if (p.ge[gj] - p.g[gj] != n || n == 0) continue;
        bool eq = true;
        for (int i = 0; i < 2 * n && eq; ++i) eq = p.wsrc[2 * p.g[gi] + i] == p.wsrc[2 * p.g[gj] + i];
        if (eq) same = gj;
      }
      if (same >= 0) { p.gw[gi] = p.gw[same]; continue; }
      // (groups start on quad boundaries of the MMA table; their weight slots simply follow each other)
      p.gw[gi] = nwu;
      for (int i = 0; i < n; ++i) p.wop[nwu++] = (uint8_t)(p.g[gi] + i);
    }
    p.nwu = nwu;
  }
  p.w_bytes = p.nwu * Ntc * 32;
  // ---- shared memory: table + weights + scale/shift + staging table + ring; as many co-resident CTAs as fit -----------------
  // planes staged by TMA: one box (8 channels x RX x RY) per chunk and view (descriptors encoded at launch); DFF_B200_SLAB_TMA=0 keeps
  // the cp.async producers, =1 restricts TMA to the row-folded first layer
  static const int tma_mode = getenv("DFF_B200_SLAB_TMA") ? atoi(getenv("DFF_B200_SLAB_TMA")) : 2;
  p.tma = (tma_mode > 0 && (a.row_step > 0 || tma_mode > 1) && p.RX <= 256 && p.RY <= 256 && a.IW % p.stx == 0 && a.IH % p.sty == 0) ? 1 : 0;
  p.nelem = p.tma ? 0 : p.nviews * p.RY * p.RX * nchunk;   // (staging table of the cp.async producers)
  // focal-merged form: the three focal groups must be the same spatial MMAs (true for every 3x3x3 / strided / x-folded layer)
  static const bool no_zmerge = getenv("DFF_B200_NO_ZMERGE") != nullptr;
  static const bool no_zm2 = getenv("DFF_B200_NO_ZM2") != nullptr;   // (A/B knob: two-phase layers keep the per-slice schedule)
  bool zm = !no_zmerge && (nph == 1 || (nph == 2 && !no_zm2)) && p.hz == 1 && 3 * Ntc <= 256;
  if (zm) {
    for (int ph = 0; ph < nph && zm; ++ph) {
      const int g0 = 3 * ph;
      const int T = p.ge[g0 + 1] - p.g[g0 + 1];
      zm = T > 0 && p.ge[g0] - p.g[g0] == T && p.ge[g0 + 2] - p.g[g0 + 2] == T;
      for (int i = 0; zm && i < T; ++i) zm = p.tab[p.g[g0] + i] == p.tab[p.g[g0 + 1] + i] && p.tab[p.g[g0 + 2] + i] == p.tab[p.g[g0 + 1] + i];
      p.zTp[ph] = T;
    }
    p.zT = p.zTp[0];
    p.zwo[0] = 0;
    p.zwo[1] = 2 * p.zTp[0] * 3 * Ntc;
    if (zm) p.w_bytes = (p.zTp[0] + (nph == 2 ? p.zTp[1] : 0)) * 96 * Ntc;   // [op][K half][3 focal blocks][N][8] per phase
  }
  static const int zm_np = getenv("DFF_ZM_NPMIN") ? atoi(getenv("DFF_ZM_NPMIN")) : 2;   // (A/B knob: minimal plane-ring depth of the focal-merged form)
  const int np_min = zm ? zm_np : 2 * p.hz + 2;
  int cols = 2 * nph * Ntc;
  if (cols > 512) return_false;
  p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  int fixed = 0;
  auto lay_out = [&](int wregion) {   // offsets for a weight region of `wregion` bytes; returns the fixed part
    int off = 1024 + ((wregion + 127) & ~127);
    p.ss_off = off;
    off += (3 * Ntc * 4 + 127) & ~127;   // scale | shift | fused-classifier weights
    p.elem_off = off;
    off += (p.nelem * 8 + 127) & ~127;
    p.planes_off = off;
    return off + 128;  // + slack for the 128-byte alignment of the dynamic window
  };
  int occ = 4, NP = 0;
  fixed = lay_out(p.w_bytes);
  for (; occ >= 1; --occ) {
    // (focal-merged: at least 4 accumulator slots per CTA)
    if (occ * (zm ? std::max(32, nph * 4 * Ntc) : p.tmem_cols) > 512) continue;
    const int budget = kSlabSmemBudget / occ - 2048 - 256;  // static shared memory, the per-CTA reservation, allocation granularity
    NP = (budget - fixed) / p.plane_bytes;
    if (NP >= np_min) break;
  }
  if (zm && occ >= 1) {
    // as many slots as the CTA's share of tensor memory holds (8 at most): the ring wraps — and an MMA has to be split — once per
    // `nslot` planes
    int ns = 8;
    while (ns > 4 && occ * std::max(32, nph * ns * Ntc) > 512) ns >>= 1;
    p.zmerge = 1; p.nslot = ns; p.lgslot = ns == 8 ? 3 : 2;
    cols = nph * ns * Ntc;
    p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  }
  if (occ < 1) {
    // Resident weights do not fit next to the plane ring: stream them (one CTA per SM, warp 7 feeds a ring of 3-4 slots with bulk
    // copies).  Blocks = runs of MMAs of one group whose weights are contiguous in `wslab` (consecutive chunk pairs / taps).
    static const bool no_ws = getenv("DFF_B200_NO_WSTREAM") != nullptr;
    if (no_ws || a.no_wstream || nchunk < 2) return_false;
    // one CTA per SM leaves nothing to hide partially filled tiles behind: very sparsely filled tile grids stay on the per-tap kernel (measured: 1/16 and 1/32 resolution still gain)
    // (a fused transposed convolution competes with four launches of this same kernel, not with the per-tap kernel)
    static const double min_fill = getenv("DFF_B200_WS_FILL") ? atof(getenv("DFF_B200_WS_FILL")) : 0.5;
    if (nph != 4 && (double)a.OHt * a.OWt < min_fill * (double)(cdiv(a.OHt, kSlabTH) * kSlabTH) * (cdiv(a.OWt, kSlabTW) * kSlabTW)) return_false;
    // Every block costs the issuing warp a barrier wait and a tcgen05.commit (~250 clk that the tensor pipe idles), so blocks are made
    // as large as the ring allows: a group of MMAs is cut into the fewest blocks whose slot (<= 40 KB) still leaves room for three
    // slots and the minimal plane ring.
    static const bool no_zmws = getenv("DFF_B200_NO_ZMWS") != nullptr;   // (A/B knob)
    if (zm && nph == 1 && a.wz && !no_zmws && 8 * Ntc <= 512) {
      // focal-merged AND streamed: the weights of one spatial MMA (three focal taps, 96*N bytes) are contiguous in `wz`, so the ring
      // is fed block by block with one bulk copy each; every plane replays the same sequence of blocks
      const int T = p.zT, opz = 96 * Ntc;
      const int budget = kSlabSmemBudget - 2048 - 256 - 256;
      int per_slot = 0, nw = 0;
      for (int parts = 1; parts <= 32 && !per_slot; ++parts) {
        const int ps = cdiv(T, parts);
        if (ps * opz > kSlabMaxWSlot) continue;
        for (nw = 4; nw >= 3; --nw) {
          fixed = lay_out(nw * ps * opz);
          NP = (budget - fixed) / p.plane_bytes;
          if (NP >= 3) { per_slot = ps; break; }
        }
      }
      if (per_slot) {
        int nb = 0;
        for (int o = 0; o < T; o += per_slot) { p.bop[nb] = (uint8_t)o; p.bn[nb] = (uint8_t)std::min(per_slot, T - o); ++nb; }
        p.nblk = nb;
        occ = 1;
        p.wslot_bytes = per_slot * opz;
        p.wstream = 1;
        p.nwslots = nw;
        p.w_bytes = nw * p.wslot_bytes;
        p.zmerge = 1; p.nslot = 8; p.lgslot = 3;
        p.tmem_cols = 8 * Ntc <= 32 ? 32 : 8 * Ntc <= 64 ? 64 : 8 * Ntc <= 128 ? 128 : 8 * Ntc <= 256 ? 256 : 512;
      }
    }
    if (!p.wstream) {
    const int np_min = 2 * p.hz + 2;   // (shadows the focal-merged minimum: streamed-weight layers run the per-slice schedule)
    const int opb = Ntc * 32;
    int gmax = 0;
    for (int gi = 0; gi < 3 * nph; ++gi) gmax = std::max(gmax, p.ge[gi] - p.g[gi]);
    const int budget = kSlabSmemBudget - 2048 - 256 - 256;
    int per_slot = 0, nw = 0;
    for (int parts = 1; parts <= 16 && !per_slot; ++parts) {
      const int ps = (cdiv(gmax, parts) + 3) & ~3;   // (blocks start on quad boundaries of the MMA table)
      if (ps * opb > kSlabMaxWSlot) continue;
      for (nw = 4; nw >= 3; --nw) {
        fixed = lay_out(nw * ps * opb);
        NP = (budget - fixed) / p.plane_bytes;
        if (NP >= np_min + (nw > 3 ? 1 : 0)) { per_slot = ps; break; }   // a fourth slot only if the plane ring keeps one plane of look-ahead
      }
    }
    if (!per_slot) return_false;
    int nb = 0;
    for (int gi = 0; gi < 3 * nph; ++gi) {
      p.gb[gi] = nb;
      int o = p.g[gi];
      while (o < p.ge[gi]) {
        int n = 1;
        while (o + n < p.ge[gi] && n < per_slot && p.wsrc[2 * (o + n)] == p.wsrc[2 * (o + n - 1)] + 2 && p.wsrc[2 * (o + n) + 1] == p.wsrc[2 * (o + n)] + 1) ++n;
        if (p.wsrc[2 * o] < 0 || p.wsrc[2 * o + 1] != p.wsrc[2 * o] + 1) return_false;
        if (n < per_slot && o + n < p.ge[gi] && (n & 3)) return_false;   // a discontinuity off a quad boundary: not expressible
        p.bop[nb] = (uint8_t)o; p.bn[nb] = (uint8_t)n;
        ++nb;
        o += n;
      }
      p.gbe[gi] = nb;
    }
    p.nblk = nb;
    occ = 1;
    p.wslot_bytes = per_slot * opb;
    p.wstream = 1;
    p.nwslots = nw;
    p.w_bytes = nw * p.wslot_bytes;
    }
  }
  if (NP > kSlabMaxPlanes) NP = kSlabMaxPlanes;
  p.NP = NP;
  // multi-phase layers (transposed convolutions) at one or two CTAs per SM are bound by their four epilogue warps: give them eight
  static const bool no_e2 = getenv("DFF_B200_NO_EPI2") != nullptr;
  p.egroups = (nph >= 2 && occ <= 2 && !no_e2) ? 2 : 1;
  p.LA = std::min(5, NP - 2 * p.hz - 1);
  *smem_out = (size_t)fixed + (size_t)NP * p.plane_bytes;
  *occ_out = occ;
  // ---- work items ----------------------------------------------------------------------------------------------------------
  p.OHt = a.OHt; p.OWt = a.OWt; p.OH = a.OH; p.OW = a.OW;
  p.osy = a.osy; p.osx = a.osx; p.ooy = a.ooy; p.oox = a.oox;
  p.tilesX = cdiv(a.OWt, kSlabTW);
  p.tilesY = cdiv(a.OHt, kSlabTH);
  const int cols_items = a.B * p.tilesX * p.tilesY;
  int nsplit = 1;
  if (cols_items < 2 * num_sms * occ) nsplit = std::min(a.S, cdiv(2 * num_sms * occ, cols_items));
  p.slen = cdiv(a.S, nsplit);
  p.nsplit = cdiv(a.S, p.slen);
  p.nitems = cols_items * p.nsplit;
  p.epi.scale = a.scale; p.epi.shift = a.shift; p.epi.res_pre = a.res_pre; p.epi.res_post = a.res_post;
  p.epi.out = a.out; p.epi.out_aux = a.out_aux; p.epi.aux_add = a.aux_add;
  p.epi.cstore = a.Cout; p.epi.relu = a.relu; p.epi.out_f32 = a.out_f32; p.epi.N = Ntc;
  p.epi.proj_w = a.proj_w; p.epi.proj_out = a.proj_out; p.epi.proj_src = a.proj_src; p.epi.skip_out = a.skip_out;
  p.epi.proj_c = a.proj_c;
  p.epi.grp_stride = a.grp_rows > 0 ? (long long)a.grp_rows * a.OW * 8 : 0;
  p.epi.pix_c = 8;
  return true;
}

#undef return_false
bool conv_slab_supported(const ConvArgs& a, const TapTable* ptaps, int nph, int Ntc) {
  SlabParams p;
  size_t smem;
  int occ;
  return slab_plan(a, ptaps ? ptaps : &a.taps, ptaps ? nph : 1, Ntc, 148, p, &smem, &occ);
}

// `wslab`: bf16 weights [ntaps][Cin/8][Ntc][8]
// `ptaps` == null: one convolution phase described by a.taps; otherwise `nph` phases (fused transposed convolution: a.osy =
// a.osx = 2 and phase ph writes output parity (ph >> 1, ph & 1)).
int launch_conv_slab(const ConvArgs& a, const TapTable* ptaps, int nph, const void* wslab, int Ntc, int num_sms, cudaStream_t st) {
  SlabParams p;
  size_t smem = 0;
  int occ = 1;
  if (!slab_plan(a, ptaps ? ptaps : &a.taps, ptaps ? nph : 1, Ntc, num_sms, p, &smem, &occ))
    return fail(-5, "conv_slab: unsupported layer shape");
  p.wslab = (p.zmerge && p.wstream) ? a.wz : wslab;   // (focal-merged streaming reads its own layout)
  if (p.tma) {
    for (int src = 0; src < (a.C1 ? 2 : 1); ++src) {
      const unsigned long long C = src ? a.C1 : a.C0;
      const char* base = (const char*)(src ? a.in1 : a.in0);
      static const bool no_merge = getenv("DFF_B200_NO_TMERGE") != nullptr;   // (A/B knob)
      p.tmerge[src] = (!no_merge && C == 8 && p.stx == 1 && 8 * p.RX <= 256) ? 1 : 0;
      if (p.wr) p.tmerge[src] = 0;
      for (int v = 0; v < p.nviews; ++v) {
        const int vy = v / p.stx, vx = v % p.stx;
        if (p.wr) {   // pixel-major rows of cb channels, swizzled
          const unsigned long long dims[4] = {C, (unsigned long long)(a.IW / p.stx), (unsigned long long)(a.IH / p.sty), (unsigned long long)a.S * a.B};
          const unsigned long long strides[3] = {C * 2 * p.stx, (unsigned long long)a.IW * C * 2 * p.sty, (unsigned long long)a.IH * a.IW * C * 2};
          const unsigned box[4] = {(unsigned)(p.rb >> 1), (unsigned)p.RX, (unsigned)p.RY, 1u};
          DFF_TRY(encode_tmap_bf16(&p.tmap[src * 4 + v], base + ((size_t)vy * a.IW + vx) * C * 2, 4, dims, strides, box, p.rb));
          continue;
        }
        if (p.tmerge[src]) {   // (8*W, 1, H/sty, S*B): box rows of RX * 16 contiguous bytes
          const unsigned long long dims[4] = {8ull * a.IW, 1ull, (unsigned long long)(a.IH / p.sty), (unsigned long long)a.S * a.B};
          const unsigned long long strides[3] = {(unsigned long long)a.IW * 16, (unsigned long long)a.IW * 16 * p.sty, (unsigned long long)a.IH * a.IW * 16};
          const unsigned box[4] = {8u * (unsigned)p.RX, 1u, (unsigned)p.RY, 1u};
          DFF_TRY(encode_tmap_bf16(&p.tmap[src * 4 + v], base + (size_t)vy * a.IW * 16, 4, dims, strides, box));
          continue;
        }
        const unsigned long long dims[4] = {C, (unsigned long long)(a.IW / p.stx), (unsigned long long)(a.IH / p.sty), (unsigned long long)a.S * a.B};
        const unsigned long long strides[3] = {C * 2 * p.stx, (unsigned long long)a.IW * C * 2 * p.sty, (unsigned long long)a.IH * a.IW * C * 2};
        const unsigned box[4] = {8u, (unsigned)p.RX, (unsigned)p.RY, 1u};
        DFF_TRY(encode_tmap_bf16(&p.tmap[src * 4 + v], base + ((size_t)vy * a.IW + vx) * C * 2, 4, dims, strides, box));
      }
    }
  }
#ifdef DFF_SLAB_TRACE
  if (getenv("DFF_SLAB_TRACE")) { cudaMalloc(&p.trace, 64 * 8 * 8); cudaMemset(p.trace, 0, 64 * 8 * 8); }
#endif
  const int grid = p.nitems < num_sms * occ ? p.nitems : num_sms * occ;
#define DFF_SLAB_LAUNCH(WS_, E2_)                                                                                              \
  do {                                                                                                                             \
    DFF_CUDA(cudaFuncSetAttribute(conv_slab_kernel<WS_, E2_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    DFF_CUDA(launch_pdl(conv_slab_kernel<WS_, E2_>, dim3(grid), dim3(slab_threads(WS_, E2_)), smem, st, p));                        \
  } while (0)
  static const bool no_i2 = getenv("DFF_B200_NO_I2") != nullptr;   // (A/B knob: one issuing warp everywhere)
  if (p.zmerge && p.nph == 2 && !p.wstream) {
    if (p.egroups == 2) {
      DFF_CUDA(cudaFuncSetAttribute(conv_slab_kernel<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      DFF_CUDA(launch_pdl(conv_slab_kernel<false, true, false, true>, dim3(grid), dim3(slab_threads(false, true)), smem, st, p));
    } else {
      DFF_CUDA(cudaFuncSetAttribute(conv_slab_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      DFF_CUDA(launch_pdl(conv_slab_kernel<false, false, false, true>, dim3(grid), dim3(slab_threads(false, false)), smem, st, p));
    }
  } else
  if (!no_i2 && !p.wstream && p.egroups == 2 && !p.zmerge && p.nph == 2 && p.hz == 0 && occ == 1) {
    DFF_CUDA(cudaFuncSetAttribute(conv_slab_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DFF_CUDA(launch_pdl(conv_slab_kernel<false, true, true>, dim3(grid), dim3(slab_threads(false, true) + 32), smem, st, p));
  } else
  if (p.wstream) { if (p.egroups == 2) DFF_SLAB_LAUNCH(true, true); else DFF_SLAB_LAUNCH(true, false); }
  else { if (p.egroups == 2) DFF_SLAB_LAUNCH(false, true); else DFF_SLAB_LAUNCH(false, false); }
#undef DFF_SLAB_LAUNCH
  DFF_LAUNCH_CHECK("conv_slab");
#ifdef DFF_SLAB_TRACE
  if (p.trace) {
    long long h[64 * 8];
    cudaDeviceSynchronize();
    cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    printf("trace C0=%d C1=%d N=%d nops=%d NP=%d occ=%d grid=%d wstream=%d nwslots=%d nblk=%d (clk since first event): full-wait tempty-wait issued | tfull-wait epi-done | empty-wait loads-issued\n", p.C0, p.C1, p.N, p.nops, p.NP, occ, grid, p.wstream, p.nwslots, p.nblk);
    long long t0 = h[5] ? h[5] : h[0];
    for (int i = 0; i < 40; ++i) { for (int j = 0; j < 7; ++j) printf("%8lld", h[i * 8 + j] ? h[i * 8 + j] - t0 : -1); printf(" | w-wait %6lld\n", h[i * 8 + 7]); }
    printf("CTA 0 (clk since kernel entry): tables done %lld, dependency wait over %lld, first plane ready %lld, all roles done %lld\n",
           h[63 * 8 + 1] - h[63 * 8], h[63 * 8 + 2] - h[63 * 8], h[0] - h[63 * 8], h[63 * 8 + 3] - h[63 * 8]);
  }
#endif
  return 0;
}

__global__ void pack_weight_slab_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin, int CinP,
                                        int ntaps, int Ntc, int transposed) {
  const int nchunk = CinP / 8;
  const int n = ntaps * nchunk * Ntc * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = i & 7, co = (i >> 3) % Ntc, c = (i / (8 * Ntc)) % nchunk, t = i / (8 * Ntc * nchunk);
    const int ci = c * 8 + j;
    float v = 0.f;
    if (co < Cout && ci < Cin) v = transposed ? w[((size_t)ci * Cout + co) * ntaps + t] : w[((size_t)co * Cin + ci) * ntaps + t];
    dst[i] = __float2bfloat16_rn(v);
  }
}

// Weights of the x-folded form of a stride-1 convolution (kernel kd x kh x kw, kw taps along x): G adjacent output pixels become
// G*Cout output channels of ONE implicit-GEMM row, fed by kw+G-1 taps along x.  dst: bf16 [kd*kh*(kw+G-1)][Cin/8][G*Cout][8] with
// W_G[g*Cout + co][ci][kd][kh][q] = W[co][ci][kd][kh][q - g] for 0 <= q - g < kw, else 0.
__global__ void pack_weight_slab_fold_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin, int CinP,
                                             int kd, int kh, int kw, int G, int foldy) {
  // foldy: the same along y — G vertically adjacent output pixels per GEMM row, kh+G-1 taps along y:
  //   dst [kd*(kh+G-1)*kw][Cin/8][G*Cout][8],  W_G[g*Cout + co][ci][kd][p][kx] = W[co][ci][kd][p - g][kx] for 0 <= p - g < kh
  const int nchunk = CinP / 8, N = G * Cout, kq = (foldy ? kh : kw) + G - 1;
  const int n = kd * (foldy ? kw : kh) * kq * nchunk * N * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = i & 7, nn = (i >> 3) % N, c = (i / (8 * N)) % nchunk, t = i / (8 * N * nchunk);
    const int g = nn / Cout, co = nn % Cout, ci = c * 8 + j;
    int a, b, kx;
    if (foldy) {   // tap order (kd, p, kx)
      kx = t % kw;
      const int pq = (t / kw) % kq;
      a = t / (kw * kq);
      b = pq - g;
    } else {       // tap order (kd, kh, q)
      const int q = t % kq;
      b = (t / kq) % kh;
      a = t / (kq * kh);
      kx = q - g;
    }
    float v = 0.f;
    if (ci < Cin && kx >= 0 && kx < kw && b >= 0 && b < kh) v = w[((((size_t)co * Cin + ci) * kd + a) * kh + b) * kw + kx];
    dst[i] = __float2bfloat16_rn(v);
  }
}
int launch_pack_weight_slab_fold(const float* w, void* dst, int Cout, int Cin, int CinP, int kd, int kh, int kw, int G, cudaStream_t st,
                                 int foldy) {
  const int n = kd * (foldy ? kw * (kh + G - 1) : kh * (kw + G - 1)) * CinP * G * Cout;
  int g = cdiv(n, 256);
  if (g > 512) g = 512;
  pack_weight_slab_fold_kernel<<<g, 256, 0, st>>>(w, (__nv_bfloat16*)dst, Cout, Cin, CinP, kd, kh, kw, G, foldy);
  DFF_LAUNCH_CHECK("pack_weight_slab_fold");
  return 0;
}

// x-folded transposed convolution (k = 3, stride (1,2,2)): the two column phases px = 0, 1 of an output row are adjacent pixels, so
// they become 2*Cout channels of one GEMM row fed by the input taps dx in {0, +1}:  dst bf16 [(kd*3+kh)*2 + dxq][Cin/8][2*Cout][8],
//   px = 0: dxq == 0 ? Wt[ci][co][kd][kh][1] : 0          px = 1: dxq == 1 ? Wt[ci][co][kd][kh][0] : Wt[ci][co][kd][kh][2]
__global__ void pack_weight_slab_deconv_fold_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin,
                                                    int CinP) {
  const int nchunk = CinP / 8, N = 2 * Cout;
  const int n = 18 * nchunk * N * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = i & 7, nn = (i >> 3) % N, c = (i / (8 * N)) % nchunk, t = i / (8 * N * nchunk);
    const int dxq = t & 1, kk = t >> 1;   // kk = kd*3 + kh
    const int px = nn / Cout, co = nn % Cout, ci = c * 8 + j;
    int kw = -1;
    if (px == 0) kw = dxq == 0 ? 1 : -1;
    else kw = dxq == 1 ? 0 : 2;
    float v = 0.f;
    if (ci < Cin && kw >= 0) v = w[(((size_t)ci * Cout + co) * 9 + kk) * 3 + kw];
    dst[i] = __float2bfloat16_rn(v);
  }
}
int launch_pack_weight_slab_deconv_fold(const float* w, void* dst, int Cout, int Cin, int CinP, cudaStream_t st) {
  const int n = 18 * CinP * 2 * Cout;
  int g = cdiv(n, 256);
  if (g > 512) g = 512;
  pack_weight_slab_deconv_fold_kernel<<<g, 256, 0, st>>>(w, (__nv_bfloat16*)dst, Cout, Cin, CinP);
  DFF_LAUNCH_CHECK("pack_weight_slab_deconv_fold");
  return 0;
}

// Row-folded first layer (1 x 9 x 9, dilation 2, on the pair-packed input): the G = 4 output rows y, y+2, y+4, y+6 are the 4 x Cout
// channels of one GEMM row fed by the 12 x 5 taps (q, c) at input rows y + 2q - 8:  dst bf16 [q*5 + c][1][G*Cout][8] with
//   W_G[g*Cout + co][ci][q][c] = Wpair[co][ci][q - g][c] for 0 <= q - g < 9, else 0.     Wpair: (Cout, 8, 1, 9, 5) fp32.
__global__ void pack_weight_slab_rowfold_kernel(const float* __restrict__ wp, __nv_bfloat16* __restrict__ dst, int Cout) {
  const int N = 4 * Cout, n = 60 * N * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ci = i & 7, nn = (i >> 3) % N, t = i / (8 * N);
    const int c = t % 5, q = t / 5, g = nn / Cout, co = nn % Cout, ky = q - g;
    float v = 0.f;
    if (ky >= 0 && ky < 9) v = wp[((co * 8 + ci) * 9 + ky) * 5 + c];
    dst[i] = __float2bfloat16_rn(v);
  }
}
int launch_pack_weight_slab_rowfold(const float* wpair, void* dst, int Cout, cudaStream_t st) {
  pack_weight_slab_rowfold_kernel<<<cdiv(60 * 4 * Cout * 8, 256), 256, 0, st>>>(wpair, (__nv_bfloat16*)dst, Cout);
  DFF_LAUNCH_CHECK("pack_weight_slab_rowfold");
  return 0;
}

// Focal-merged streaming layout of a 3x3x3 convolution: bf16 [spatial tap u = kh*3+kw][chunk pair][K half][focal block j][Ntc][8] with
// focal block j = 0, 1, 2 <-> dz = +1, 0, -1 <-> kd = 2, 1, 0 — the B operand of one merged MMA (all three focal taps of a spatial tap
// and 16 input channels) is 96*Ntc contiguous bytes, so a block of consecutive merged MMAs is one bulk copy.
__global__ void pack_weight_slab_zmerge_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin, int CinP,
                                               int Ntc) {
  const int ncp = CinP / 16;
  const int n = 9 * ncp * 2 * 3 * Ntc * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int r = i;
    const int e = r & 7; r >>= 3;
    const int co = r % Ntc; r /= Ntc;
    const int j = r % 3; r /= 3;
    const int h = r & 1; r >>= 1;
    const int cp = r % ncp, u = r / ncp;
    const int ci = (2 * cp + h) * 8 + e, t = (2 - j) * 9 + u;
    float v = 0.f;
    if (co < Cout && ci < Cin) v = w[((size_t)co * Cin + ci) * 27 + t];
    dst[i] = __float2bfloat16_rn(v);
  }
}
int launch_pack_weight_slab_zmerge(const float* w, void* dst, int Cout, int Cin, int CinP, int Ntc, cudaStream_t st) {
  const int n = 27 * CinP * Ntc;
  int g = cdiv(n, 256);
  if (g > 512) g = 512;
  pack_weight_slab_zmerge_kernel<<<g, 256, 0, st>>>(w, (__nv_bfloat16*)dst, Cout, Cin, CinP, Ntc);
  DFF_LAUNCH_CHECK("pack_weight_slab_zmerge");
  return 0;
}

int launch_pack_weight_slab(const float* w, void* dst, int Cout, int Cin, int CinP, int ntaps, int Ntc, int transposed,
                            cudaStream_t st) {
  const int n = ntaps * CinP * Ntc;
  int g = cdiv(n, 256);
  if (g > 512) g = 512;
  pack_weight_slab_kernel<<<g, 256, 0, st>>>(w, (__nv_bfloat16*)dst, Cout, Cin, CinP, ntaps, Ntc, transposed);
  DFF_LAUNCH_CHECK("pack_weight_slab");
  return 0;
}

}  // namespace dff
