// "Row" implicit-GEMM convolution with the activation operand in TENSOR MEMORY (tcgen05.mma TS form) — the kernel for the
// small-channel, high-resolution layers of the network (stride-1 3x3x3 and 1x3x3 convolutions with Cin <= 64, Cout <= 32 at full
// and half resolution: SRD blocks, hourglass conv0/conv2; reference train_codes/Depth_Estimation_Network.py:278-301, 361-367).
//
// Why: with both operands in shared memory an M=128, K=16 tcgen05.mma occupies the tensor pipe for ~39 clk whatever N <= 32 is,
// because it fetches 4 KB of A (profiles/r1_ubench.txt) — a 16-channel 3x3x3 layer then costs 27 x 39 clk per 128 pixels, 3x its
// HBM time.  With A in TMEM the same MMA takes N/2 clk (9-16).  A TMEM operand cannot be shifted along M, so the schedule is
// INPUT-STATIONARY: M = 128 consecutive pixels of one image row; one input row segment (slice z, row r), shifted by dx = -1,0,+1,
// is written to TMEM once and multiplied into the NINE accumulators of the outputs it touches (slices z-1..z+1 x rows r-1..r+1).
// The accumulators of rows r-1, r, r+1 of one slice are adjacent TMEM columns, so the three dy taps are ONE instruction with
// N' = 3N and the weight blocks of dy = +1, 0, -1 stacked along N: 3 x Cin/16 x 3 MMAs per input row instead of 27 x Cin/16 (the
// issuing warp is latency-bound per instruction, profiles/r1_ubench.txt).  Accumulators (3 slices x TH rows) rotate through
// TMEM and are always accumulated into: the epilogue re-zeroes a buffer (tcgen05.st) after draining it.
//
// Pipeline of one CTA (owns a 128-pixel x TH-row band and walks the S slices):
//   warp  0    loader: one elected thread issues one TMA box load per 8-channel plane of the row (130 pixels x 8 channels; the
//              halo pixels outside the image are the TMA's out-of-bounds zero fill, the second source of a torch.cat is a
//              second tensor map) into a shared-memory FIFO (mbarrier expect_tx);
//   warps 2-5  transfer: lane = pixel; conflict-free LDS.128 of the dx-shifted 8-channel pieces -> tcgen05.st into the A ring
//              (each global byte is loaded once, each shared byte read 3x, and none of it by the MMA);
//   warp  6    issues tcgen05.mma [D], [A_tmem], B_smem (warp-uniform loop, elected lane); ONE tcgen05.commit per input row
//              (a commit costs the issuing thread ~150 clk): its barrier both frees the A slot and publishes the outputs that
//              row completed;
//   warps 7-10 epilogue: tcgen05.ld, BatchNorm/bias, residuals, ReLU, bf16 store — a 128-pixel row segment is one contiguous
//              run of memory in the channels-last layout, so every store instruction writes whole lines.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace dff {

constexpr int kRowXferWarp0 = 2, kRowMmaWarp = 6, kRowEpiWarp0 = 7;
constexpr int kRowThreads = 11 * 32;
constexpr int kRowMaxNR = 16, kRowMaxRA = 4, kRowMaxAcc = 12, kRowDone = 16;
constexpr int kRowPx = 128 + 2;                 // pixels of a staged row segment (one halo pixel each side)
constexpr int kRowPlane = 17 * 128;             // bytes reserved per 8-channel plane of a staged row (130 x 16 B, 128-byte aligned)

// DFF_SLAB_TRACE (compile-time, debugging only): CTA 0 records clock64 timestamps of its pipeline events per input row
#ifdef DFF_SLAB_TRACE
#define DFF_RT(slot, idx) do { if (p.trace && blockIdx.x == 0 && (idx) < 64 && (threadIdx.x & 31) == 0) p.trace[(idx) * 8 + (slot)] = clock64(); } while (0)
#else
#define DFF_RT(slot, idx) do { } while (0)
#endif

struct alignas(64) RowParams {
  CUtensorMap map0, map1;   // (C, W, H, S, B) channels-last views of the two sources, box (8, 130, 1, 1, 1)
  const void* wslab;        // bf16 [tap][chunk8][N][8]
  int nchunk8, nch0, c8shift;
  int B, S, H, W;
  int KD, N, NA, TH, RA, NR, NACC;
  int tilesX, bandsY, nitems;
  int row_bytes, w_blocks, ss_off, rows_off, tmem_cols, acc_col0;
  EpiArgs epi;
  // per (band class, row of the band): what an input row touches.  x = idesc (N' of the merged-dy MMA), y = first accumulator
  // column relative to the slice plane, z = first weight row, w = slots whose occupancy starts (bits 0-11) / completes (12-23)
  uint4 tab[4][6];
  long long* trace;
  int exp;   // timing experiments only (DFF_ROW_EXPERIMENT bit mask; results are wrong): 1 no loads, 2 no epilogue math/stores, 4 no MMAs, 8 no transfer
};

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.eq.u32 p, 0, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc)
      : "memory");
}
// Predicated forms: the whole issuing warp runs the (warp-uniform) schedule, only the elected lane's instruction takes effect —
// no divergent region, so the compiler keeps every descriptor and counter in uniform registers.
__device__ __forceinline__ void umma_ts_if(uint32_t pred, uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.u32 q, 0, 0;\n"
      "@p tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t pred, uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %1, 0;\n"
      "@p tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z)
               : "memory");
}

__global__ void __launch_bounds__(kRowThreads, 2) conv_row_kernel(const __grid_constant__ RowParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kRowMaxNR + kRowMaxRA + kRowDone + kRowMaxAcc];
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t w_s = smem0;                     // weights in MMA order
  const uint32_t rows_s = smem0 + p.rows_off;     // FIFO of NR staged rows
  float* const ss = reinterpret_cast<float*>(smem_gen + p.ss_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rfull0 = smem_u32(&bars[0]), rempty0 = smem_u32(&bars[kRowMaxNR]);
  const uint32_t afull0 = smem_u32(&bars[2 * kRowMaxNR]);
  const uint32_t rdone0 = smem_u32(&bars[2 * kRowMaxNR + kRowMaxRA]);
  const uint32_t tempty0 = smem_u32(&bars[2 * kRowMaxNR + kRowMaxRA + kRowDone]);

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.map0);
    prefetch_tmap(&p.map1);
    for (int i = 0; i < p.NR; ++i) { mbar_init(rfull0 + 8 * i, 1); mbar_init(rempty0 + 8 * i, 128); }
    for (int i = 0; i < p.RA; ++i) mbar_init(afull0 + 8 * i, 128);
    for (int i = 0; i < kRowDone; ++i) mbar_init(rdone0 + 8 * i, 1);
    for (int i = 0; i < p.NACC; ++i) mbar_init(tempty0 + 8 * i, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kRowMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // weights in MMA order: block (dz, a, K half) = [3 dy (reversed) x N rows] x 16 B ; BatchNorm scale/shift
    const int rows_per_blk = 3 * p.N;
    const int total = p.KD * p.NA * 2 * rows_per_blk;  // 16-byte rows
    const uint4* wg = reinterpret_cast<const uint4*>(p.wslab);
    for (int i = threadIdx.x; i < total; i += kRowThreads) {
      const int n = i % p.N;
      int t = i / p.N;
      const int dyrev = t % 3; t /= 3;
      const int h = t & 1; t >>= 1;
      const int a = t % p.NA, dzi = t / p.NA;
      const int j = 2 * a + h, dxi = j >> p.c8shift, c8 = j & (p.nchunk8 - 1);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (dxi < 3) {
        const int tap = (dzi * 3 + (2 - dyrev)) * 3 + dxi;
        v = __ldg(wg + ((size_t)tap * p.nchunk8 + c8) * p.N + n);
      }
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(w_s + 16 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    for (int i = threadIdx.x; i < p.N; i += kRowThreads) {
      ss[i] = p.epi.scale ? __ldg(p.epi.scale + i) : 1.f;
      ss[p.N + i] = p.epi.shift ? __ldg(p.epi.shift + i) : 0.f;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int hz = p.KD == 3 ? 1 : 0;
  const int bands_x = p.tilesX * p.bandsY;

  if (warp == 0) {
    // =============================== loader: TMA -> shared row FIFO ===============================
    if (lane == 0) {
      int rs = 0, nrow = 0;
      uint32_t ephase = 1;
      const uint32_t tx_bytes = (uint32_t)p.nchunk8 * kRowPx * 16;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const int b = item / bands_x, rem = item % bands_x;
        const int x0 = (rem % p.tilesX) * 128, y0 = (rem / p.tilesX) * p.TH;
        const int rlo = max(y0 - 1, 0), rhi = min(y0 + p.TH, p.H - 1);
        for (int z = 0; z < p.S; ++z)
          for (int r = rlo; r <= rhi; ++r) {
            mbar_wait(rempty0 + 8 * rs, ephase);
            DFF_RT(0, nrow);
            const uint32_t dst0 = rows_s + rs * p.row_bytes, bar = rfull0 + 8 * rs;
            if (!(p.exp & 1)) {
              mbar_expect_tx(bar, tx_bytes);
              for (int c8 = 0; c8 < p.nchunk8; ++c8) {
                const bool second = c8 >= p.nch0;
                tma_load_5d(dst0 + c8 * kRowPlane, second ? &p.map1 : &p.map0, bar, (second ? c8 - p.nch0 : c8) * 8, x0 - 1, r, z, b);
              }
            } else {
              mbar_arrive(bar);
            }
            if (++rs == p.NR) { rs = 0; ephase ^= 1; }
            DFF_RT(1, nrow);
            ++nrow;
          }
      }
    }
  } else if (warp >= kRowXferWarp0 && warp < kRowMmaWarp) {
    // =============================== transfer: shared row -> TMEM A ring (lane = pixel) ===============================
    const int q = warp & 3, pxl = q * 32 + lane;
    int rs = 0, as = 0, nrow = 0;
    uint32_t rphase = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int band = (item % bands_x) / p.tilesX;
      const int nrows = p.TH + 2 - (band == 0 ? 1 : 0) - (band == p.bandsY - 1 ? 1 : 0);
      for (int n = 0; n < p.S * nrows; ++n) {
        mbar_wait(rfull0 + 8 * rs, rphase);
        if (nrow >= p.RA) {   // the MMAs that read this A slot RA rows ago must have completed
          const int m = nrow - p.RA;
          mbar_wait(rdone0 + 8 * (m & (kRowDone - 1)), (uint32_t)(m >> 4) & 1u);
          fence_after();
        }
        if (q == 2) DFF_RT(2, nrow);
        const uint32_t src0 = rows_s + rs * p.row_bytes + pxl * 16;
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + as * (p.NA * 8);
        if (!(p.exp & 8))
        for (int a = 0; a < p.NA; ++a) {
          uint32_t v[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = 2 * a + h, dxi = j >> p.c8shift, c8 = j & (p.nchunk8 - 1);
            if (dxi < 3) {
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(v[4 * h]), "=r"(v[4 * h + 1]), "=r"(v[4 * h + 2]), "=r"(v[4 * h + 3])
                           : "r"(src0 + c8 * kRowPlane + dxi * 16));
            } else {
              v[4 * h] = v[4 * h + 1] = v[4 * h + 2] = v[4 * h + 3] = 0u;
            }
          }
          tmem_st8(t0 + a * 8, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        fence_before();
        mbar_arrive(afull0 + 8 * as);
        mbar_arrive_relaxed(rempty0 + 8 * rs);
        if (q == 2) DFF_RT(3, nrow);
        ++nrow;
        if (++rs == p.NR) { rs = 0; rphase ^= 1; }
        if (++as == p.RA) as = 0;
      }
    }
  } else if (warp == kRowMmaWarp) {
    // =============================== MMA issuer (TS form: A in TMEM) ===============================
    // Single instruction stream, latency-bound: per input row one table entry (parameter bank, uniform load) says which
    // accumulator columns / weight rows / N' the merged-dy MMA uses and which accumulators start; per slice the three (dz)
    // planes rotate.
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t blk2 = 6u * (uint32_t)p.N;      // (dz, a) weight tile = two K halves of 3N rows, in 16-byte descriptor units
    const uint64_t bd_base = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | (uint64_t)((w_s >> 4) | ((3u * p.N) << 16));
    const uint32_t planeN = (uint32_t)(p.TH * p.N), acc0 = tmem_base + p.acc_col0;
    const uint32_t dzw = (uint32_t)p.NA * blk2;    // weight stride between dz groups
    int as = 0, nrow = 0;
    uint32_t aphase = 0, te_bits = 0u;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int band = (item % bands_x) / p.tilesX;
      const int cls = (band == 0 ? 1 : 0) | (band == p.bandsY - 1 ? 2 : 0);
      const int nrows = p.TH + 2 - (cls & 1) - (cls >> 1);
      int zm3 = 0;   // plane of output slice s = z
      for (int z = 0; z < p.S; ++z) {
        // output slices touched by input slice z: s = z+1 (dzi 0), z (dzi 1), z-1 (dzi 2); KD == 1: s = z only (as dzi 0)
        const int p_up = zm3 == 2 ? 0 : zm3 + 1, p_dn = zm3 == 0 ? 2 : zm3 - 1;
        const int pl0 = hz ? p_up : zm3, pl1 = zm3, pl2 = p_dn;
        const bool ok0 = hz ? (z + 1 < p.S) : true, ok1 = hz != 0, ok2 = hz && z > 0;
        const bool fst0 = true, fst1 = z == 0;   // s = z+1 always starts at z; s = z starts here only for z == 0 (dzi 2 never starts)
        for (int ri = 0; ri < nrows; ++ri) {
          const uint4 e = p.tab[cls][ri];
          mbar_wait(afull0 + 8 * as, aphase);
          DFF_RT(4, nrow);
          // accumulators that start with this row must have been drained (and re-zeroed) by the epilogue
          if (e.w & 0xfffu) {
            if (ok0 && fst0)
              for (uint32_t w = e.w & 0xfffu; w; w >>= 4) {
                const int sl = pl0 * p.TH + (int)(w & 0xfu) - 1;
                mbar_wait(tempty0 + 8 * sl, (te_bits >> sl) & 1u);
                te_bits ^= 1u << sl;
              }
            if (ok1 && fst1)
              for (uint32_t w = e.w & 0xfffu; w; w >>= 4) {
                const int sl = pl1 * p.TH + (int)(w & 0xfu) - 1;
                mbar_wait(tempty0 + 8 * sl, (te_bits >> sl) & 1u);
                te_bits ^= 1u << sl;
              }
          }
          fence_after();
          const uint32_t a_t = tmem_base + as * (p.NA * 8);
          const uint64_t bd0 = bd_base + (uint64_t)e.z;
          if (!(p.exp & 4)) {
            if (ok0) {
              const uint32_t dacc = acc0 + pl0 * planeN + e.y;
              uint64_t bd = bd0;
              for (int a = 0; a < p.NA; ++a, bd += blk2) umma_ts_if(leader, dacc, a_t + a * 8, bd, e.x);
            }
            if (ok1) {
              const uint32_t dacc = acc0 + pl1 * planeN + e.y;
              uint64_t bd = bd0 + dzw;
              for (int a = 0; a < p.NA; ++a, bd += blk2) umma_ts_if(leader, dacc, a_t + a * 8, bd, e.x);
            }
            if (ok2) {
              const uint32_t dacc = acc0 + pl2 * planeN + e.y;
              uint64_t bd = bd0 + 2 * dzw;
              for (int a = 0; a < p.NA; ++a, bd += blk2) umma_ts_if(leader, dacc, a_t + a * 8, bd, e.x);
            }
          }
          umma_commit_if(leader, rdone0 + 8 * (nrow & (kRowDone - 1)));   // frees the A slot AND publishes the outputs this row completed
          DFF_RT(5, nrow);
          ++nrow;
          if (++as == p.RA) { as = 0; aphase ^= 1; }
        }
        if (++zm3 == 3) zm3 = 0;
      }
    }
  } else if (warp >= kRowEpiWarp0) {
    // =============================== epilogue ===============================
    const int q = warp & 3, pxl = q * 32 + lane;
    EpiArgs ep = p.epi;
    ep.scale = ss;
    ep.shift = ss + p.N;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + p.acc_col0;
    // every MMA accumulates: hand all accumulators over zeroed
    for (int c = 0; c < p.NACC * p.N; c += 16) tmem_zero16(tq + c);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    fence_before();
    for (int sl = 0; sl < p.NACC; ++sl) mbar_arrive(tempty0 + 8 * sl);
    int nout = 0, nrow = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int b = item / bands_x, rem = item % bands_x;
      const int band = rem / p.tilesX;
      const int x0 = (rem % p.tilesX) * 128, y0 = band * p.TH;
      const int cls = (band == 0 ? 1 : 0) | (band == p.bandsY - 1 ? 2 : 0);
      const int nrows = p.TH + 2 - (cls & 1) - (cls >> 1);
      const bool valid = x0 + pxl < p.W;
      int zm3 = 0;
      for (int z = 0; z < p.S; ++z) {
        for (int ri = 0; ri < nrows; ++ri, ++nrow) {
          const uint32_t cm = p.tab[cls][ri].w >> 12;
          if (!cm) continue;
          bool waited = false;
          for (int dzi = 0; dzi < p.KD; ++dzi) {
            const int dz = hz ? dzi - 1 : 0, s = z - dz;
            if (s < 0 || s >= p.S) continue;
            if (hz && z != min(s + 1, p.S - 1)) continue;   // only the last slice of s completes rows
            int sm3 = zm3 - dz;
            sm3 = sm3 < 0 ? sm3 + 3 : (sm3 > 2 ? sm3 - 3 : sm3);
            if (!waited) {   // all MMAs up to and including this input row have completed
              mbar_wait(rdone0 + 8 * (nrow & (kRowDone - 1)), (uint32_t)(nrow >> 4) & 1u);
              fence_after();
              waited = true;
            }
            for (uint32_t w = cm; w; w >>= 4) {
              const int c = (int)(w & 0xfu) - 1, slot = sm3 * p.TH + c;
              const size_t pix = (((size_t)b * p.S + s) * p.H + y0 + c) * p.W + x0 + pxl;
              if (!(p.exp & 2)) tc_epilogue_tile(ep, tq + slot * p.N, valid, pix);
              for (int c0 = 0; c0 < p.N; c0 += 16) tmem_zero16(tq + slot * p.N + c0);
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              fence_before();
              mbar_arrive(tempty0 + 8 * slot);
              if (q == 3) { DFF_RT(6, nout); ++nout; }
            }
          }
        }
        if (++zm3 == 3) zm3 = 0;
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == kRowMmaWarp) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*RowEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static RowEncodeFn row_get_encode() {
  static RowEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (RowEncodeFn)f;
  }();
  return fn;
}
static int row_encode(CUtensorMap* m, const void* base, int C, int B, int S, int H, int W) {
  RowEncodeFn fn = row_get_encode();
  if (!fn) return fail(-3, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)S * H * W * C * 2};
  cuuint32_t box[5] = {8, (cuuint32_t)kRowPx, 1, 1, 1}, es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return 0;
}

static bool row_plan(const ConvArgs& a, int Ntc, RowParams& p, size_t* smem_out, int* occ_out) {
  memset(&p, 0, sizeof(p));
  // Opt-in (DFF_B200_ROW=1): correct on every supported shape (tests/test_gpu_ops.py) but its single-warp issue loop still costs
  // ~300 instructions per input row, which leaves it 0-25 % behind the slab kernel (profiles/r1_row_kernel.txt).
  static const bool enabled = getenv("DFF_B200_ROW") && atoi(getenv("DFF_B200_ROW")) != 0;
  if (!enabled) return false;
  if (a.C0 % 8 || a.C1 % 8 || a.C0 < 8) return false;
  const int nchunk8 = (a.C0 + a.C1) / 8;
  if (nchunk8 != 1 && nchunk8 != 2 && nchunk8 != 4) return false;   // Cin in {8, 16, 32}
  if (Ntc != 16 && Ntc != 32) return false;
  if (a.isy != 1 || a.isx != 1 || a.osy != 1 || a.osx != 1 || a.ooy || a.oox) return false;
  if (a.OH != a.IH || a.OW != a.IW || a.OHt != a.IH || a.OWt != a.IW) return false;
  if (a.IW < 256 || (a.IW % 128 != 0 && a.IW % 128 < 32)) return false;   // M = 128 pixels of a row: keep the padding waste small
  // taps must be exactly the KD x 3 x 3 box in canonical order (dilation 1)
  const int KD = a.taps.n == 27 ? 3 : (a.taps.n == 9 ? 1 : 0);
  if (!KD) return false;
  for (int t = 0; t < a.taps.n; ++t) {
    const int kd = t / 9, kh = (t / 3) % 3, kw = t % 3;
    if (a.taps.dz[t] != (KD == 3 ? kd - 1 : 0) || a.taps.dy[t] != kh - 1 || a.taps.dx[t] != kw - 1 || a.taps.widx[t] != t) return false;
  }
  p.nchunk8 = nchunk8; p.nch0 = a.C0 / 8;
  p.c8shift = nchunk8 == 1 ? 0 : nchunk8 == 2 ? 1 : 2;
  p.B = a.B; p.S = a.S; p.H = a.IH; p.W = a.IW;
  p.KD = KD; p.N = Ntc;
  p.NA = (3 * nchunk8 + 1) / 2;
  p.w_blocks = KD * p.NA * 2;   // (dz, a, K half) blocks of 3N rows
  const int w_bytes = (p.w_blocks * 3 * Ntc * 16 + 127) & ~127;
  p.ss_off = w_bytes;
  p.rows_off = (w_bytes + 2 * Ntc * 4 + 1023) & ~1023;
  p.row_bytes = nchunk8 * kRowPlane;
  // TMEM: A ring + 3 x TH accumulators; prefer two CTAs per SM (256 columns each)
  int best_occ = 0;
  for (int occ = 2; occ >= 1 && !best_occ; --occ)
    for (int TH = 4; TH >= 2 && !best_occ; --TH)
      for (int RA = 3; RA >= 2 && !best_occ; --RA) {
        const int cols = RA * p.NA * 8 + 3 * TH * Ntc;
        if (cols > 512 / occ) continue;
        if (p.H % TH) continue;   // bands are whole (the row table is per band class)
        if (3 * TH > kRowMaxAcc) continue;
        const int budget = 220 * 1024 / occ - 3072;
        const int NR = std::min(kRowMaxNR, (budget - p.rows_off) / p.row_bytes);
        if (NR < 4) continue;
        p.TH = TH; p.RA = RA; p.NR = NR; best_occ = occ;
        p.acc_col0 = RA * p.NA * 8;
        p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
      }
  if (!best_occ) return false;
  p.NACC = 3 * p.TH;
  for (int cls = 0; cls < 4; ++cls) {
    const bool top = cls & 1, bottom = (cls & 2) != 0;
    int ri = 0;
    for (int rr = top ? 0 : -1; rr <= (bottom ? p.TH - 1 : p.TH); ++rr, ++ri) {
      const int ylo = std::max(rr - 1, 0), yhi = std::min(rr + 1, p.TH - 1);
      uint32_t wait = 0, commit = 0;
      int nw = 0, nc = 0;
      if (yhi == rr + 1) wait |= (uint32_t)(yhi + 1) << (4 * nw++);   // the row below starts here
      if (top && rr == 0) wait |= (uint32_t)(0 + 1) << (4 * nw++);    // image row 0 has no row above: it starts here too
      if (ylo == rr - 1) commit |= (uint32_t)(ylo + 1) << (4 * nc++); // the row above is complete
      if (bottom && rr == p.TH - 1) commit |= (uint32_t)(rr + 1) << (4 * nc++);   // the last image row completes with itself
      const uint32_t nsel = (uint32_t)(yhi - ylo + 1) * Ntc;
      p.tab[cls][ri].x = (1u << 4) | (1u << 7) | (1u << 10) | ((nsel >> 3) << 17) | ((128u >> 4) << 24);
      p.tab[cls][ri].y = (uint32_t)ylo * Ntc;
      p.tab[cls][ri].z = (uint32_t)(ylo - (rr - 1)) * Ntc;
      p.tab[cls][ri].w = wait | (commit << 12);
    }
  }
  p.tilesX = cdiv(p.W, 128);
  p.bandsY = p.H / p.TH;
  p.nitems = p.B * p.tilesX * p.bandsY;
  p.epi.scale = a.scale; p.epi.shift = a.shift; p.epi.res_pre = a.res_pre; p.epi.res_post = a.res_post;
  p.epi.out = a.out; p.epi.out_aux = a.out_aux; p.epi.aux_add = a.aux_add;
  p.epi.cstore = a.Cout; p.epi.relu = a.relu; p.epi.out_f32 = a.out_f32; p.epi.N = Ntc;
  p.epi.proj_w = a.proj_w; p.epi.proj_out = a.proj_out; p.epi.proj_src = a.proj_src; p.epi.skip_out = a.skip_out;
  p.epi.proj_c = 0; p.epi.grp_stride = 0; p.epi.pix_c = 0;
  *smem_out = (size_t)p.rows_off + (size_t)p.NR * p.row_bytes + 1024;
  *occ_out = best_occ;
  return true;
}

bool conv_row_supported(const ConvArgs& a, int Ntc) {
  RowParams p;
  size_t smem;
  int occ;
  return row_plan(a, Ntc, p, &smem, &occ);
}

// `wslab`: bf16 weights [ntaps][Cin/8][Ntc][8] (the slab kernel's pack)
int launch_conv_row(const ConvArgs& a, const void* wslab, int Ntc, int num_sms, cudaStream_t st) {
  RowParams p;
  size_t smem = 0;
  int occ = 1;
  if (!row_plan(a, Ntc, p, &smem, &occ)) return fail(-5, "conv_row: unsupported layer shape");
  p.wslab = wslab;
  p.exp = getenv("DFF_ROW_EXPERIMENT") ? atoi(getenv("DFF_ROW_EXPERIMENT")) : 0;
  DFF_TRY(row_encode(&p.map0, a.in0, a.C0, a.B, a.S, a.IH, a.IW));
  if (a.C1) DFF_TRY(row_encode(&p.map1, a.in1, a.C1, a.B, a.S, a.IH, a.IW));
  else p.map1 = p.map0;
  DFF_CUDA(cudaFuncSetAttribute(conv_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.nitems < num_sms * occ ? p.nitems : num_sms * occ;
#ifdef DFF_SLAB_TRACE
  if (getenv("DFF_SLAB_TRACE")) { cudaMalloc(&p.trace, 64 * 8 * 8); cudaMemset(p.trace, 0, 64 * 8 * 8); }
#endif
  conv_row_kernel<<<grid, kRowThreads, smem, st>>>(p);
  DFF_LAUNCH_CHECK("conv_row");
#ifdef DFF_SLAB_TRACE
  if (p.trace) {
    long long h[64 * 8];
    cudaDeviceSynchronize();
    cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    printf("row trace nchunk8=%d N=%d KD=%d NA=%d TH=%d RA=%d NR=%d occ=%d grid=%d (clk): loader slot-free, issued | transfer ready, done | mma A-ready, issued | epilogue k-th output done\n",
           p.nchunk8, p.N, p.KD, p.NA, p.TH, p.RA, p.NR, occ, grid);
    long long t0 = h[0];
    for (int i = 0; i < 40; ++i) { for (int j = 0; j < 7; ++j) printf("%8lld", h[i * 8 + j] ? h[i * 8 + j] - t0 : -1); printf("\n"); }
  }
#endif
  return 0;
}

}  // namespace dff
