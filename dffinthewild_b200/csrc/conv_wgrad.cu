// Weight gradient of the generic 3-D convolution (train path; autograd of `nn.Conv3d` / `nn.ConvTranspose3d`, invoked by
// `Total.backward()` at reference train_codes/train_code_Defocus.py:167 — ATen's convolution_backward, 74 % of the reference's
// CPU training time).
//
//   dw[widx_t][ci][co] = sum over output positions o of  in[o*is + d_t][ci] * dy[o*os + oo][co]
//
// in the same "strided gather over a tap table" form as the forward kernels (conv_ffma.cu), so ordinary, strided and
// (per output-parity phase) transposed convolutions and the two-source concat all go through one kernel.  One CTA owns a
// 32 x TY tile of output positions of one (batch, slice): it stages the input region (halo included) channel-planar for CK
// input channels and the dy tile for all output channels in shared memory; a thread owns one (tap, ci) and a block of COB output
// channels, sums over all positions of the tile in registers (one activation load feeds COB FMAs, the dy row is a shared-memory
// broadcast), and adds its partial sums to the fp32 gradient in the reference's weight layout with red.global.add.f32.
// fp32 accumulate; the order of the global adds is not fixed, which moves results by ~1e-7 relative (gate: gradient cosine).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

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

template <typename T, int CK, int COB>
__global__ void __launch_bounds__(kWgThreads, 2) conv_wgrad_kernel(const __grid_constant__ WgradArgs g) {
  extern __shared__ __align__(16) float smem[];
  const ConvArgs& a = g.a;
  const int REG = a.RZ * a.RY * a.RXP;
  const int NP = 32 * a.TY;              // positions per tile
  const int CoP = (g.Cout + 3) & ~3;
  float* in_s = smem;                     // [CK][REG]
  float* dy_s = smem + CK * REG;          // [NP][CoP]
  const int tid = threadIdx.x;
  const int bs = blockIdx.z, b = bs / a.S, s = bs % a.S;
  const int ty0 = blockIdx.y * a.TY, tx0 = blockIdx.x * 32;
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

  const int Ctot = a.C0 + a.C1;
  const int npos = a.RZ * a.RY * a.RX;
  constexpr int NQ = CK / 4;
  const int cosplit = CoP / COB;                       // output-channel blocks per (tap, ci)
  const int nitems = a.taps.n * CK * cosplit;          // work items of one channel chunk: (tap, ci, co block)

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
    // a thread owns (tap, ci) x COB output channels and sums over all positions of the tile: one activation load feeds COB FMAs,
    // the dy row is a shared-memory broadcast for the warp
    for (int it = tid; it < nitems; it += kWgThreads) {
      const int cs = it % cosplit, ci = (it / cosplit) % CK, t = it / (cosplit * CK);
      const int cig = g.ci_base + c0 + ci;   // channel inside the layer's Cin (stored channels beyond Cin are padding)
      if (cig >= g.Cin) continue;
      const float* ip = in_s + ci * REG + ((a.taps.dz[t] - a.dzmin) * a.RY + (a.taps.dy[t] - a.dymin)) * a.RXP +
                        (a.taps.dx[t] - a.dxmin);
      const float* dp = dy_s + cs * COB;
      float acc[COB];
#pragma unroll
      for (int j = 0; j < COB; ++j) acc[j] = 0.f;
      for (int py = 0; py < a.TY; ++py) {
        const float* irow = ip + py * a.isy * a.RXP;
        const float* drow = dp + (py * 32) * CoP;
#pragma unroll 4
        for (int px = 0; px < 32; ++px) {
          const float av = irow[px * a.isx];
#pragma unroll
          for (int j4 = 0; j4 < COB / 4; ++j4) {
            const float4 d4 = *reinterpret_cast<const float4*>(drow + px * CoP + 4 * j4);
            acc[4 * j4] = fmaf(av, d4.x, acc[4 * j4]);
            acc[4 * j4 + 1] = fmaf(av, d4.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(av, d4.z, acc[4 * j4 + 2]);
            acc[4 * j4 + 3] = fmaf(av, d4.w, acc[4 * j4 + 3]);
          }
        }
      }
      const int wi = a.taps.widx[t];
#pragma unroll
      for (int j = 0; j < COB; ++j) {
        const int co = cs * COB + j;
        if (co >= g.Cout) continue;
        const size_t o = g.wt_transposed ? ((size_t)cig * g.Cout + co) * g.ntaps_total + wi
                                         : ((size_t)co * g.Cin + cig) * g.ntaps_total + wi;
        atomicAdd(g.dw + o, acc[j]);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// bf16 weight gradient on the tensor cores (warp-level mma.sync m16n8k16, fp32 accumulate).
//
//   dW_t[ci][co] = sum over the tile's positions p of  X[p*is + d_t][ci] * dY[p][co]        = (X_t^T) (dY):  M = ci, N = co, K = positions
//
// Both operands are channels-last, i.e. K-rows of contiguous M / N elements — exactly what ldmatrix.trans turns into the A (row) and
// B (col) fragments: a tap is a different base address of the same staged input region, a stride-2 layer a different row pitch, so
// ordinary, strided and (per output-parity phase) transposed convolutions and the two-source concat share the kernel.  One CTA owns
// 32 x TY output positions of one (batch, slice); it stages the halo'd input region for CK channels and the dy tile (all output
// channels) in shared memory as bf16, its 8 warps split the (tap, 16-channel block, 8-output-channel block) accumulators — 8-channel
// tensors pair two taps in one m16 block — and each accumulator runs over the tile's K = 32*TY positions in registers before it is
// added to the fp32 gradient (reference weight layout) with red.global.add.f32.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4_t(unsigned addr, unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(unsigned addr, unsigned& r0, unsigned& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* d, unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 16-byte global -> shared copy without a register round trip; ok == false writes zeros (the convolution's padding)
__device__ __forceinline__ void wg_cp16(unsigned dst, const void* src, bool ok) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
}

struct WgPiece {          // one staged 16-byte piece
  int rel;                // source element offset relative to the tile / region origin
  unsigned dst;           // byte offset inside the staging buffer
  unsigned char x, y, z, pad;   // tile-relative coordinates (z bit 7: second source)
};

struct WgradMmaArgs {
  WgradArgs w;
  int CK, TY;      // channels staged per pass (8, 16, 32 or 64); tile rows
  int nsplit;      // CTAs per tile: low-resolution layers have few tiles but thousands of accumulators — the accumulator list is dealt
                   // round-robin to `nsplit` CTAs that stage the same (L2-resident) region (blockIdx.x = tile_x * nsplit + split)
};

// Persistent over tiles: CTA (split, c) owns the accumulators `split, split + nsplit, ...` of the layer's list and walks the tiles
// c, c + ncta, ... keeping them in registers (at most kWgMaxAcc m16 x n16 accumulators per warp), so the fp32 atomics into the
// gradient happen once per CTA instead of once per tile.
constexpr int kWgMaxAcc = 4;
__global__ void __launch_bounds__(kWgThreads, 2) conv_wgrad_mma_kernel(const __grid_constant__ WgradMmaArgs ga) {
  extern __shared__ __align__(16) unsigned char smem_b[];
  const WgradArgs& g = ga.w;
  const ConvArgs& a = g.a;
  const int CK = ga.CK, TY = ga.TY;
  const int REGPOS = a.RZ * a.RY * a.RX;
  const int NP = 32 * TY;
  const int CoP = (g.Cout + 7) & ~7;
  __nv_bfloat16* in_s = reinterpret_cast<__nv_bfloat16*>(smem_b);            // [REGPOS][CK]
  __nv_bfloat16* dy_s = in_s + (size_t)REGPOS * CK;                           // [NP][CoP]
  WgPiece* in_tab = reinterpret_cast<WgPiece*>(dy_s + (size_t)NP * CoP);      // [REGPOS * CK/8]
  WgPiece* dy_tab = in_tab + (size_t)REGPOS * (CK / 8);                       // [NP * CoP/8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x % ga.nsplit, cta = blockIdx.x / ga.nsplit, ncta = gridDim.x / ga.nsplit;
  const int tilesX = (a.OWt + 31) / 32, tilesY = (a.OHt + TY - 1) / TY;
  const int ntile = tilesX * tilesY * a.B * a.S;
  const int Ctot = a.C0 + a.C1;
  const int s8 = CK / 8;                         // 8-channel slots per tap
  const int nslots = a.taps.n * s8;
  const int nmt = (nslots + 1) / 2;              // m16 blocks: two consecutive slots each
  const int nnt = CoP / 8;
  const int ngrp = (nnt + 1) / 2;                // n8 blocks are taken two at a time
  const int nitems = nmt * ngrp;
  const unsigned in_u = (unsigned)__cvta_generic_to_shared(in_s), dy_u = (unsigned)__cvta_generic_to_shared(dy_s);
  // ldmatrix row this lane supplies: matrix j = lane / 8 -> (m half j & 1, k half j >> 1), row r = lane % 8 -> position k = 8*(j>>1) + r
  const int lj = lane >> 3, lr = lane & 7, lk = 8 * (lj >> 1) + lr;
  const int gq = lane >> 2, t4 = lane & 3;
  const int q8 = CoP / 8;
  const int wstep = ga.nsplit * (kWgThreads / 32);   // stride of this warp's accumulator list

  const __nv_bfloat16* const src0 = reinterpret_cast<const __nv_bfloat16*>(a.in0);
  const __nv_bfloat16* const src1 = reinterpret_cast<const __nv_bfloat16*>(a.in1);
  const __nv_bfloat16* const dyg = reinterpret_cast<const __nv_bfloat16*>(g.dy);
  const int rowlen = a.RX * s8, nin = a.RZ * a.RY * rowlen;      // 16-byte pieces of the staged input region
  const int ndy = NP * q8;
  for (int c0 = 0; c0 < Ctot; c0 += CK) {
    float acc[kWgMaxAcc][8];
#pragma unroll
    for (int u = 0; u < kWgMaxAcc; ++u)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[u][e] = 0.f;
    // ---- tables of the 16-byte pieces a tile stages (the same for every tile of this channel pass): destination in shared memory,
    // source offset relative to the tile origin, coordinates for the bounds check — so that the per-tile staging loops are a table
    // read, a few compares and a cp.async instead of five integer divisions per piece ----
    __syncthreads();
    for (int e = tid; e < nin; e += kWgThreads) {
      const int row = e / rowlen, rem = e - row * rowlen;
      const int x = rem / s8, q = rem - x * s8;
      const int z = row / a.RY, y = row - z * a.RY;
      const int cabs = c0 + 8 * q;
      const bool second = cabs >= a.C0;
      const int Csrc = second ? a.C1 : a.C0, cs = second ? cabs - a.C0 : cabs;
      WgPiece pc;
      pc.rel = (int)((((long long)z * a.IH + y) * a.IW + x) * Csrc + cs);
      pc.dst = 2u * (unsigned)((row * a.RX + x) * CK + 8 * q);
      pc.x = (unsigned char)x; pc.y = (unsigned char)y; pc.z = (unsigned char)(z | (second ? 0x80 : 0)); pc.pad = 0;
      in_tab[e] = pc;
    }
    if (c0 == 0)
      for (int e = tid; e < ndy; e += kWgThreads) {
        const int q = e % q8, pos = e / q8;
        const int px = pos & 31, py = pos >> 5;
        WgPiece pc;
        pc.rel = (int)((((long long)py * a.osy) * a.OW + (long long)px * a.osx) * g.CoS + 8 * q);
        pc.dst = 2u * (unsigned)(pos * CoP + 8 * q);
        pc.x = (unsigned char)px; pc.y = (unsigned char)py; pc.z = (unsigned char)(8 * q < g.CoS ? 0 : 1); pc.pad = 0;
        dy_tab[e] = pc;
      }
    __syncthreads();
    for (int tile = cta; tile < ntile; tile += ncta) {
      int r = tile;
      const int tx0 = (r % tilesX) * 32; r /= tilesX;
      const int ty0 = (r % tilesY) * TY; r /= tilesY;
      const int s = r % a.S, b = r / a.S;
      const int gy0 = ty0 * a.isy + a.dymin, gx0 = tx0 * a.isx + a.dxmin, gz0 = s + a.dzmin;
      __syncthreads();   // the previous tile's MMAs have read the staging buffers
      // ---- stage the dy tile and the halo'd input region with fire-and-forget 16-byte cp.async copies (zero fill outside the
      // tensors / the phase grid): every thread has all of its copies in flight before anyone waits — the first version loaded
      // and stored piece by piece and spent its time in exposed global-memory latency ----
      {
        // dy: the tile-relative pixel offset of every piece comes from the table built once per CTA
        const size_t dpix0 = (((size_t)b * a.S + s) * a.OH + ((size_t)ty0 * a.osy + a.ooy)) * a.OW + ((size_t)tx0 * a.osx + a.oox);
        const __nv_bfloat16* dbase = dyg + dpix0 * g.CoS;
        for (int e = tid; e < ndy; e += kWgThreads) {
          const WgPiece pc = dy_tab[e];
          const bool ok = tx0 + (int)pc.x < a.OWt && ty0 + (int)pc.y < a.OHt && pc.z == 0;
          wg_cp16(dy_u + pc.dst, ok ? (const void*)(dbase + pc.rel) : (const void*)dyg, ok);
        }
        // input region: rel is relative to the region origin (gz0, gy0, gx0) in the piece's own source (channel stride Csrc)
        const long long org = (((long long)b * a.S + gz0) * a.IH + gy0) * (long long)a.IW + gx0;
        const __nv_bfloat16* b0 = src0 + org * a.C0;
        const __nv_bfloat16* b1 = src1 ? src1 + org * a.C1 : src0;
        for (int e = tid; e < nin; e += kWgThreads) {
          const WgPiece pc = in_tab[e];
          const int gz = gz0 + (int)(pc.z & 0x7f), gy = gy0 + (int)pc.y, gx = gx0 + (int)pc.x;
          const bool ok = gz >= 0 && gz < a.S && gy >= 0 && gy < a.IH && gx >= 0 && gx < a.IW;
          const __nv_bfloat16* sp = (pc.z & 0x80) ? b1 : b0;
          wg_cp16(in_u + pc.dst, ok ? (const void*)(sp + pc.rel) : (const void*)src0, ok);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
#pragma unroll
      for (int u = 0; u < kWgMaxAcc; ++u) {
        const int it = split * (kWgThreads / 32) + warp + u * wstep;
        if (it >= nitems) break;
        const int mt = it / ngrp, ng = it - mt * ngrp;
        const int slot0 = 2 * mt, slot1 = min(2 * mt + 1, nslots - 1);
        const int myslot = (lj & 1) ? slot1 : slot0;
        const int tap = myslot / s8, c8 = myslot - tap * s8;
        // element offset of this lane's row at k-step 0: tap origin + its position's pixel + its channel group
        const int rowbase = ((((int)a.taps.dz[tap] - a.dzmin) * a.RY + ((int)a.taps.dy[tap] - a.dymin)) * a.RX + ((int)a.taps.dx[tap] - a.dxmin) +
                             lk * a.isx) * CK + 8 * c8;
        const int n0 = 2 * ng, n1 = min(2 * ng + 1, nnt - 1);
        const unsigned brow = dy_u + 2u * (unsigned)((lane & 15) * CoP);
        for (int ks = 0; ks < 2 * TY; ++ks) {
          // 16 consecutive positions of one tile row: row ks / 2, columns 16 * (ks & 1) ...
          const int koff = ((ks >> 1) * a.isy * a.RX + (ks & 1) * 16 * a.isx) * CK;
          unsigned a0, a1, a2, a3, b0, b1;
          ldsm_x4_t(in_u + 2u * (unsigned)(rowbase + koff), a0, a1, a2, a3);
          const unsigned bk = brow + 2u * (unsigned)(ks * 16 * CoP);
          ldsm_x2_t(bk + 16u * (unsigned)n0, b0, b1);
          mma_bf16_16816(acc[u], a0, a1, a2, a3, b0, b1);
          if (n1 != n0) {
            ldsm_x2_t(bk + 16u * (unsigned)n1, b0, b1);
            mma_bf16_16816(acc[u] + 4, a0, a1, a2, a3, b0, b1);
          }
        }
      }
    }
    // ---- flush this pass's accumulators: d0,d1: (m = gq, n = 2*t4, 2*t4+1) ; d2,d3: (m = gq + 8, ...): m < 8 -> slot0, m >= 8 -> slot1 ----
#pragma unroll
    for (int u = 0; u < kWgMaxAcc; ++u) {
      const int it = split * (kWgThreads / 32) + warp + u * wstep;
      if (it >= nitems) break;
      const int mt = it / ngrp, ng = it - mt * ngrp;
      const int n0 = 2 * ng, n1 = min(2 * ng + 1, nnt - 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int slot = 2 * mt + half;
        if (slot >= nslots) continue;
        const int tp = slot / s8, cc = slot - tp * s8;
        const int cig = g.ci_base + c0 + 8 * cc + gq;
        if (cig >= g.Cin) continue;
        const int wi = a.taps.widx[tp];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          if (nb == 1 && n1 == n0) continue;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int co = 8 * (nb ? n1 : n0) + 2 * t4 + e;
            if (co >= g.Cout) continue;
            const size_t o = g.wt_transposed ? ((size_t)cig * g.Cout + co) * g.ntaps_total + wi
                                             : ((size_t)co * g.Cin + cig) * g.ntaps_total + wi;
            atomicAdd(g.dw + o, acc[u][4 * nb + 2 * half + e]);
          }
        }
      }
    }
  }
}

// plan: input region extent for a tile of TY rows; returns the shared-memory bytes of the mma kernel for (CK, TY)
static size_t wg_mma_plan(ConvArgs& a, int CK, int TY, int Cout) {
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
  a.TY = TY;
  a.RZ = dzmax - a.dzmin + 1;
  a.RY = (TY - 1) * a.isy + (dymax - a.dymin) + 1;
  a.RX = 31 * a.isx + (dxmax - a.dxmin) + 1;
  a.RXP = a.RX;
  const int CoP = (Cout + 7) & ~7;
  const size_t pieces = (size_t)a.RZ * a.RY * a.RX * (CK / 8) + (size_t)32 * TY * (CoP / 8);
  return ((size_t)a.RZ * a.RY * a.RX * CK + (size_t)32 * TY * CoP) * 2 + pieces * sizeof(WgPiece);
}

static int launch_conv_wgrad_mma(ConvArgs a, const void* dy, int CoS, int Cout, int Cin, int ci_base, float* dw, int ntaps_total,
                                 int wt_transposed, cudaStream_t st, bool* handled) {
  *handled = false;
  if (a.C0 % 8 || a.C1 % 8 || CoS % 8 || a.C0 < 8) return 0;
  // channels per pass: the largest of 64, 32, 16 that divides the (virtually concatenated) input; a pass may straddle the two sources
  // (8-channel tensors: 8, two taps per m16 block)
  int CK = 8;
  for (int c = 64; c >= 16; c >>= 1)
    if ((a.C0 + a.C1) % c == 0) { CK = c; break; }
  const size_t budget = 110 * 1024;   // two CTAs per SM
  int TY = 0;
  size_t smem = 0;
  for (;; CK >>= 1) {
    for (int ty = 16; ty >= 2; ty >>= 1) {
      if (ty > 2 && (ty >> 1) >= a.OHt) continue;      // no taller than the phase grid needs
      ConvArgs t = a;
      const size_t sz = wg_mma_plan(t, CK, ty, Cout);
      if (sz <= budget) { TY = ty; smem = sz; break; }
    }
    if (TY || CK == 8) break;
  }
  if (!TY) return 0;
  WgradMmaArgs ga{};
  wg_mma_plan(a, CK, TY, Cout);
  ga.w.a = a; ga.w.dy = dy; ga.w.CoS = CoS; ga.w.Cout = Cout; ga.w.Cin = Cin; ga.w.dw = dw; ga.w.ntaps_total = ntaps_total;
  ga.w.wt_transposed = wt_transposed; ga.w.ci_base = ci_base;
  ga.CK = CK; ga.TY = TY;
  // the accumulator list is dealt to `nsplit` CTA groups so that a warp holds at most kWgMaxAcc of them in registers; each group gets
  // an equal share of ~2 CTAs per SM, and every CTA of a group walks the tiles with that stride
  const int ntile = cdiv(a.OWt, 32) * cdiv(a.OHt, TY) * a.B * a.S;
  const int s8 = CK / 8, nmt = (a.taps.n * s8 + 1) / 2, ngrp = (((Cout + 7) / 8) + 1) / 2;
  const int nitems = nmt * ngrp;
  const int nsplit = cdiv(nitems, kWgMaxAcc * (kWgThreads / 32));
  int ncta = std::max(1, std::min(ntile, (2 * 148) / nsplit));
  if (nsplit * ncta < 148 && ntile > ncta) ncta = std::min(ntile, cdiv(148, nsplit));
  ga.nsplit = nsplit;
  dim3 grid(nsplit * ncta, 1, 1);
  static const bool log = getenv("DFF_B200_WGRAD_LOG") != nullptr;   // (profiling aid: one line per launch, joins with an ncu launch list)
  if (log)
    fprintf(stderr, "wgrad_mma Cin=%d+%d Cout=%d taps=%d B=%d S=%d OHt=%d OWt=%d is=%d os=%d CK=%d TY=%d nsplit=%d ncta=%d ntile=%d smem=%zu\n",
            a.C0, a.C1, Cout, a.taps.n, a.B, a.S, a.OHt, a.OWt, a.isx, a.osx, CK, TY, nsplit, ncta, ntile, smem);
  DFF_CUDA(cudaFuncSetAttribute(conv_wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_wgrad_mma_kernel<<<grid, kWgThreads, smem, st>>>(ga);
  DFF_LAUNCH_CHECK("conv_wgrad_mma");
  *handled = true;
  return 0;
}


// ------------------------------------------------------------------------------------------------------------------
// TMA-staged, double-buffered form of the tensor-core weight gradient (the default for bf16 tensors).
//
// Same contraction as conv_wgrad_mma_kernel (M = (tap, input channel), N = output channel, K = output positions, mma.sync m16n8k16),
// re-organised around what bounded that kernel (profiles/r2_train_wgrad.txt): per-thread staging with index arithmetic, 2..8-way
// shared-memory bank conflicts of the [position][CK] layout, one ldmatrix pair per MMA, no overlap of staging and math.
//  * A tile (TX x TY positions of one (batch, slice), TX = 8 / 16 / 32 by the layer's width) is staged by ONE thread with tiled TMA
//    loads — a 5-D box (8 channels, RX, RY, RZ, 1) per 8-channel chunk of the input region (zero padding = out-of-bounds fill, the
//    second tensor map = the concat source) and a 4-D box (8, TX, TY, 1) per 8 output channels of dy (a transposed convolution's
//    output-parity phase = a strided tensor map) — into two stages: tile i+1 lands while the warps work on tile i.
//  * Layout [8-channel chunk][position][8]: the 8 rows of every ldmatrix are 128 contiguous bytes (conflict-free), and a tap is
//    still only a different base address.
//  * A warp owns a block of up to MW m16-tiles x NW n8-tiles (64 fp32 accumulators at most): the dy fragments of a k-step are loaded
//    once for the whole block, each input fragment feeds NW MMAs.  Blocks are dealt to the 8 warps of `nsub` CTA groups; the input
//    channels are split into `npass` groups of CK so that a CTA stages only the channels it multiplies.
//  * The four phases of a transposed convolution run in ONE launch: they read the same input region, every tap belongs to exactly
//    one phase, so a phase is just the dy set (one of four staged sub-tiles) a tap's m16-tile multiplies with.
// Accumulators stay in registers over all tiles a CTA walks; one red.global.add.f32 per element and CTA at the end.
// ------------------------------------------------------------------------------------------------------------------
constexpr size_t kWg2Smem = 110 * 1024;   // two CTAs per SM
struct Wg2Args {
  CUtensorMap tx[2];
  CUtensorMap tdy[4];
  TapTable taps;                 // the taps of all dy sets, concatenated
  uint8_t tap_set[kMaxTaps];     // dy set (phase) of every tap
  int nset;
  int C0, S;
  int tilesX, tilesY, ntile;
  int isx, isy, dzmin, dymin, dxmin, RX, RY, REGP;
  int TX, TY, NP;
  int CK, s8, nsub, ncta;
  int ksplit;                   // warps sharing one accumulator block, each taking every ksplit-th k-row group of a tile (few-tap layers
                                 // have fewer blocks than warps: one warp walking all 64 k-steps of a tile alone was instruction-latency-bound)
  int nst, cs;                  // pipeline stages (2..4); CTAs per cluster (1, 2, 4, 8: the accumulators of a cluster are summed
                                 // through distributed shared memory before they go to the gradient)
  int MT, NT, NB, MB, mper, nslots;
  int Cout, Cin, ci_base, ntaps_total, wt_transposed;
  float* dw;
  unsigned x_bytes, dy_bytes, tx_bytes;   // per stage: input part, one dy set; bytes the TMA loads of one tile deliver
};

__device__ __forceinline__ void wg2_issue(const Wg2Args& g, int tile, int c0, uint32_t dst, uint32_t bar) {
  using namespace tc;
  int r = tile;
  const int tx0 = (r % g.tilesX) * g.TX; r /= g.tilesX;
  const int ty0 = (r % g.tilesY) * g.TY; r /= g.tilesY;
  const int s = r % g.S, b = r / g.S;
  const int gx0 = tx0 * g.isx + g.dxmin, gy0 = ty0 * g.isy + g.dymin, gz0 = s + g.dzmin;
  mbar_expect_tx(bar, g.tx_bytes);
  for (int c = 0; c < g.s8; ++c) {
    const int cabs = c0 + 8 * c;
    const int src = cabs >= g.C0 ? 1 : 0;
    tma_load_5d(dst + (uint32_t)(c * g.REGP) * 16u, &g.tx[src], bar, src ? cabs - g.C0 : cabs, gx0, gy0, gz0, b);
  }
  for (int set = 0; set < g.nset; ++set)
    for (int nt = 0; nt < g.NT; ++nt)
      tma_load_4d(dst + g.x_bytes + (uint32_t)set * g.dy_bytes + (uint32_t)(nt * g.NP) * 16u, &g.tdy[set], bar, nt * 8, tx0, ty0, b * g.S + s);
}

template <int MW, int NW>
__global__ void __launch_bounds__(kWgThreads, 2) conv_wgrad_tma_kernel(const __grid_constant__ Wg2Args g) {
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem_w[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = smem_u32(smem_w);
  const uint32_t stage_bytes = g.x_bytes + (uint32_t)g.nset * g.dy_bytes;
  const uint32_t bar0 = s0 + (uint32_t)g.nst * stage_bytes;
  int r = blockIdx.x;
  const int cta = r % g.ncta; r /= g.ncta;
  const int sub = r % g.nsub, pass = r / g.nsub;
  const int c0 = pass * g.CK;
  if (tid == 0) {
    for (int i = 0; i < g.nst; ++i) mbar_init(bar0 + 8u * (uint32_t)i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&g.tx[0]);
    prefetch_tmap(&g.tdy[0]);
  }
  __syncthreads();
  if (tid == 0)   // prologue: the first nst - 1 tiles of this CTA are in flight before anyone waits
    for (int i = 0; i < g.nst - 1; ++i) {
      const int t = cta + i * g.ncta;
      if (t < g.ntile) wg2_issue(g, t, c0, s0 + (uint32_t)i * stage_bytes, bar0 + 8u * (uint32_t)i);
    }

  // ---- this warp's block of accumulators: m16-tiles [mt0, mt0 + mcount) x n8-tiles [nt0, nt0 + NW) ----
  const int nblk = g.MB * g.NB, ks = g.ksplit;
  const int item = ks > 1 ? warp % nblk : sub * (kWgThreads / 32) + warp;
  const int kpart = ks > 1 ? warp / nblk : 0;
  const bool active = ks > 1 ? kpart < ks : item < nblk;
  const int mb = active ? item / g.NB : 0, nb = active ? item - mb * g.NB : 0;
  const int mt0 = mb * g.mper;
  const int mcount = active ? max(0, min(g.mper, g.MT - mt0)) : 0;
  const int nt0 = nb * NW;
  // ldmatrix row of this lane: matrix j = lane / 8 -> (m half j & 1, k half j >> 1), row r -> position k = 8 * (j >> 1) + r of the k-step
  const int lj = lane >> 3, lr = lane & 7, lk = 8 * (lj >> 1) + lr;
  const int gq = lane >> 2, t4 = lane & 3;
  const int lanepos = g.TX == 8 ? ((lk >> 3) * g.isy * g.RX + (lk & 7) * g.isx) : lk * g.isx;
  uint32_t abase[MW];
  uint32_t sets = 0;
#pragma unroll
  for (int i = 0; i < MW; ++i) {
    const int mt = min(mt0 + i, g.MT - 1);
    const int slot = min(2 * mt + (lj & 1), g.nslots - 1);
    const int tap = slot / g.s8, c8 = slot - tap * g.s8;
    const int tapoff = (((int)g.taps.dz[tap] - g.dzmin) * g.RY + ((int)g.taps.dy[tap] - g.dymin)) * g.RX + ((int)g.taps.dx[tap] - g.dxmin);
    abase[i] = 16u * (uint32_t)(c8 * g.REGP + tapoff + lanepos);
    sets |= (uint32_t)g.tap_set[(2 * mt) / g.s8] << (2 * i);
  }
  const uint32_t boff = g.x_bytes + 16u * (uint32_t)(nt0 * g.NP + (lane & 15) + (NW >= 2 ? (lane >> 4) * g.NP : 0));
  const int nh = g.TX == 32 ? 2 : 1;
  const uint32_t halfstep = 256u * (uint32_t)g.isx;
  const uint32_t rowstep = 16u * (uint32_t)((g.TX == 8 ? 2 : 1) * g.isy * g.RX);
  const int nrow = g.NP / (16 * nh);
  float acc[MW][NW][4];
#pragma unroll
  for (int i = 0; i < MW; ++i)
#pragma unroll
    for (int n = 0; n < NW; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][n][e] = 0.f;

  int stage = 0;
  uint32_t par = 0;   // bit s: parity the next wait on stage s uses
  for (int tile = cta; tile < g.ntile; tile += g.ncta) {
    {   // the stage consumed in the previous iteration takes the tile nst - 1 ahead
      const int nxt = tile + (g.nst - 1) * g.ncta;
      const int ns = stage == 0 ? g.nst - 1 : stage - 1;
      if (tid == 0 && nxt < g.ntile) wg2_issue(g, nxt, c0, s0 + (uint32_t)ns * stage_bytes, bar0 + 8u * (uint32_t)ns);
    }
    mbar_wait(bar0 + 8u * (uint32_t)stage, (par >> stage) & 1u);
    par ^= 1u << stage;
    const uint32_t sb = s0 + (uint32_t)stage * stage_bytes;
    if (mcount > 0) {
#pragma unroll 1
      for (int rr = kpart; rr < nrow; rr += ks) {
#pragma unroll 1
        for (int h = 0; h < nh; ++h) {
          const uint32_t ka = sb + (uint32_t)rr * rowstep + (uint32_t)h * halfstep;
          const uint32_t kb = sb + boff + 256u * (uint32_t)(rr * nh + h);
          int cur = -1;
          unsigned b[NW][2];
#pragma unroll
          for (int i = 0; i < MW; ++i) {
            if (i < mcount) {
              const int set = (int)((sets >> (2 * i)) & 3u);
              if (set != cur) {
                cur = set;
                const uint32_t kbs = kb + (uint32_t)set * g.dy_bytes;
                if (NW == 1) ldsm_x2_t(kbs, b[0][0], b[0][1]);
                else {
#pragma unroll
                  for (int n = 0; n < NW; n += 2) ldsm_x4_t(kbs + 16u * (uint32_t)(n * g.NP), b[n][0], b[n][1], b[n + 1 < NW ? n + 1 : n][0], b[n + 1 < NW ? n + 1 : n][1]);
                }
              }
              unsigned a0, a1, a2, a3;
              ldsm_x4_t(ka + abase[i], a0, a1, a2, a3);
#pragma unroll
              for (int n = 0; n < NW; ++n) mma_bf16_16816(acc[i][n], a0, a1, a2, a3, b[n][0], b[n][1]);
            }
          }
        }
      }
    }
    __syncthreads();   // every warp is done with this stage: the next iteration's TMA may overwrite it
    if (++stage == g.nst) stage = 0;
  }
  // ---- flush.  Fragment: d0,d1 = (m = gq, n = 2*t4, 2*t4+1), d2,d3 = (m = gq + 8, ...): m < 8 -> slot 2*mt, m >= 8 -> slot 2*mt + 1.
  // Every CTA of a (pass, sub) group holds the same accumulator blocks for different tiles; the gradient takes one fp32 atomic per
  // element and *cluster*: the CTAs of a cluster park their fragments in their own shared memory, CTA `rank` sums warp-slots
  // rank, rank + cs, ... of all of them through distributed shared memory and issues the atomics (the atomics — every CTA adding
  // its whole block to a few KB of gradient — were what bounded the low-channel layers) ----
  if (g.cs > 1) {
    float* dump = reinterpret_cast<float*>(smem_w);
#pragma unroll
    for (int i = 0; i < MW; ++i)
#pragma unroll
      for (int n = 0; n < NW; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) dump[((((warp * MW + i) * NW + n) * 4 + e) << 5) + lane] = acc[i][n][e];
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int ws = (int)rank; ws < kWgThreads / 32; ws += g.cs) {
      const int it2 = ks > 1 ? ws % nblk : sub * (kWgThreads / 32) + ws;   // (k-split: every warp slot of a block is one more addend)
      if (ks > 1 ? ws >= nblk * ks : it2 >= nblk) continue;
      const int mb2 = it2 / g.NB, nb2 = it2 - mb2 * g.NB;
      const int mt02 = mb2 * g.mper, mc2 = max(0, min(g.mper, g.MT - mt02)), nt02 = nb2 * NW;
      for (int pi = warp; pi < mc2 * NW; pi += kWgThreads / 32) {
        const int i = pi / NW, n = pi - i * NW;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t la = s0 + 4u * (uint32_t)(((((ws * MW + i) * NW + n) * 4) << 5) + lane);
        for (int q = 0; q < g.cs; ++q) {
          uint32_t ra;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(q));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x;
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(x) : "r"(ra + 128u * (uint32_t)e));
            v[e] += x;
          }
        }
        const int mt = mt02 + i;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int slot = 2 * mt + (e >> 1);
          if (slot >= g.nslots) continue;
          const int tp = slot / g.s8, cc = slot - tp * g.s8;
          const int cig = g.ci_base + c0 + 8 * cc + gq;
          const int co = 8 * (nt02 + n) + 2 * t4 + (e & 1);
          if (cig >= g.Cin || co >= g.Cout) continue;
          const int wi = g.taps.widx[tp];
          const size_t o = g.wt_transposed ? ((size_t)cig * g.Cout + co) * g.ntaps_total + wi
                                           : ((size_t)co * g.Cin + cig) * g.ntaps_total + wi;
          atomicAdd(g.dw + o, v[e]);
        }
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers may still be reading this CTA's fragments
    return;
  }
#pragma unroll
  for (int i = 0; i < MW; ++i) {
    if (i >= mcount) continue;
    const int mt = mt0 + i;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int slot = 2 * mt + half;
      if (slot >= g.nslots) continue;
      const int tp = slot / g.s8, cc = slot - tp * g.s8;
      const int cig = g.ci_base + c0 + 8 * cc + gq;
      if (cig >= g.Cin) continue;
      const int wi = g.taps.widx[tp];
#pragma unroll
      for (int n = 0; n < NW; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int co = 8 * (nt0 + n) + 2 * t4 + e;
          if (co >= g.Cout) continue;
          const size_t o = g.wt_transposed ? ((size_t)cig * g.Cout + co) * g.ntaps_total + wi
                                           : ((size_t)co * g.Cin + cig) * g.ntaps_total + wi;
          atomicAdd(g.dw + o, acc[i][n][2 * half + e]);
        }
    }
  }
}

int encode_tmap_bf16(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                     const unsigned* box, int swizzle_bytes);

// `a`: geometry / sources / strides as for the other kernels; `pt[nph]`: the tap tables of the phases (nph = 1: an ordinary launch
// with a.ooy/a.oox; nph = 4: the four output-parity phases (py, px) = (ph >> 1, ph & 1) of a stride-2 transposed convolution).
int launch_conv_wgrad_tma(const ConvArgs& a, const TapTable* pt, int nph, const void* dy, int CoS, int Cout, int Cin, int ci_base,
                          float* dw, int ntaps_total, int wt_transposed, cudaStream_t st, bool* handled) {
  *handled = false;
  static const bool no_tma = getenv("DFF_B200_WGRAD_NO_TMA") != nullptr;   // A/B switch: the per-thread-staged mma kernel
  if (no_tma) return 0;
  const int Ctot = a.C0 + a.C1;
  if (a.C0 % 8 || a.C1 % 8 || CoS % 8 || a.C0 < 8 || nph < 1 || nph > 4) return 0;
  const int NT = CoS / 8;
  const int NW = NT >= 4 ? 4 : NT;
  if (NW == 3 || NT % NW) return 0;
  const int MW = NW == 4 ? 4 : 8;
  Wg2Args ga{};
  ga.taps.n = 0;
  int dzmax = -100, dymax = -100, dxmax = -100;
  ga.dzmin = ga.dymin = ga.dxmin = 100;
  for (int ph = 0; ph < nph; ++ph)
    for (int t = 0; t < pt[ph].n; ++t) {
      const int k = ga.taps.n++;
      if (k >= kMaxTaps) return 0;
      ga.taps.dz[k] = pt[ph].dz[t]; ga.taps.dy[k] = pt[ph].dy[t]; ga.taps.dx[k] = pt[ph].dx[t]; ga.taps.widx[k] = pt[ph].widx[t];
      ga.tap_set[k] = (uint8_t)ph;
      ga.dzmin = std::min(ga.dzmin, (int)pt[ph].dz[t]); dzmax = std::max(dzmax, (int)pt[ph].dz[t]);
      ga.dymin = std::min(ga.dymin, (int)pt[ph].dy[t]); dymax = std::max(dymax, (int)pt[ph].dy[t]);
      ga.dxmin = std::min(ga.dxmin, (int)pt[ph].dx[t]); dxmax = std::max(dxmax, (int)pt[ph].dx[t]);
    }
  const int ntap = ga.taps.n;
  if (!ntap) return 0;
  // channels per pass: the smallest group that still gives every warp of a CTA a block (a CTA then stages only what it multiplies);
  // merged phases need both halves of an m16-tile in one tap (CK >= 16)
  const int NB = NT / NW;
  int CK = 0;
  for (int c = (nph > 1 ? 16 : 8); c <= 64; c <<= 1) {
    if (Ctot % c) continue;
    CK = c;
    const int MT = (ntap * (c / 8) + 1) / 2;
    if (cdiv(MT, MW) * NB >= kWgThreads / 32) break;
  }
  if (!CK) return 0;
  const int s8 = CK / 8, nslots = ntap * s8, MT = (nslots + 1) / 2;
  int MB = cdiv(MT, MW);
  if (MB * NB < kWgThreads / 32) MB = std::min(MT, std::max(1, (kWgThreads / 32) / NB));
  const int mper = cdiv(MT, MB);
  MB = cdiv(MT, mper);
  const int nsub = cdiv(MB * NB, kWgThreads / 32);
  // tile: TX by the layer's width; TY as tall as three pipeline stages allow (up to 1024 positions: every tile costs a block-wide
  // barrier round and a TMA round trip) but no taller than leaves every CTA a few tiles; at least two stages must fit
  const int npass = Ctot / CK, G = npass * nsub;
  const int TX = a.OWt > 16 ? 32 : (a.OWt > 8 ? 16 : 8);
  int RX = 0, RY = 0, RZ = 0, REGP = 0, TY = 0;
  unsigned xb = 0, db = 0;
  for (int want = 3; want >= 2 && !TY; --want)
    for (int ty = 1024 / TX; TX * ty >= 16 && (TX != 8 || ty % 2 == 0); ty /= 2) {
      if (ty / 2 >= a.OHt && TX * (ty / 2) >= 16 && (TX != 8 || (ty / 2) % 2 == 0)) continue;   // no taller than the phase grid needs
      if (TX * ty > 256 && (long long)cdiv(a.OWt, TX) * cdiv(a.OHt, ty) * a.B * a.S < 4LL * std::max(1, 296 / G)) continue;
      RZ = dzmax - ga.dzmin + 1;
      RY = (ty - 1) * a.isy + (dymax - ga.dymin) + 1;
      RX = (TX - 1) * a.isx + (dxmax - ga.dxmin) + 1;
      REGP = (RZ * RY * RX + 7) & ~7;
      xb = (unsigned)(s8 * REGP) * 16u;
      db = (unsigned)(NT * TX * ty) * 16u;
      if (RX <= 256 && RY <= 256 && RZ <= 256 && want * (size_t)(xb + nph * db) + 64 <= kWg2Smem) { TY = ty; break; }
    }
  if (!TY) return 0;
  ga.nset = nph;
  ga.C0 = a.C0; ga.S = a.S;
  ga.tilesX = cdiv(a.OWt, TX); ga.tilesY = cdiv(a.OHt, TY);
  ga.ntile = ga.tilesX * ga.tilesY * a.B * a.S;
  if (ga.ntile <= 0) { *handled = true; return 0; }
  ga.isx = a.isx; ga.isy = a.isy; ga.RX = RX; ga.RY = RY; ga.REGP = REGP;
  ga.TX = TX; ga.TY = TY; ga.NP = TX * TY;
  ga.CK = CK; ga.s8 = s8; ga.nsub = nsub;
  static const bool no_ksplit = getenv("DFF_B200_WGRAD_NO_KSPLIT") != nullptr;   // (A/B knob)
  ga.ksplit = (!no_ksplit && nsub == 1 && MB * NB <= (kWgThreads / 32) / 2) ? (kWgThreads / 32) / (MB * NB) : 1;
  ga.MT = MT; ga.NT = NT; ga.NB = NB; ga.MB = MB; ga.mper = mper; ga.nslots = nslots;
  ga.Cout = Cout; ga.Cin = Cin; ga.ci_base = ci_base; ga.ntaps_total = ntaps_total; ga.wt_transposed = wt_transposed;
  ga.dw = dw;
  ga.x_bytes = xb; ga.dy_bytes = db;
  ga.tx_bytes = (unsigned)(s8 * RZ * RY * RX) * 16u + (unsigned)nph * db;
  // clusters of up to 8 CTAs of one group reduce their accumulators on chip.  The grid is sized to ONE wave of co-resident clusters
  // (cudaOccupancyMaxActiveClusters: clusters are placed per GPC, 2 x 148 CTAs of 8-CTA clusters do not all fit at once — the
  // left-over clusters ran as a second wave and doubled the time of the full-resolution layers)
  static const int cs_max = getenv("DFF_B200_WGRAD_CLUSTER") ? atoi(getenv("DFF_B200_WGRAD_CLUSTER")) : 8;   // (A/B knob; 1 = off)
  const size_t stage = (size_t)xb + (size_t)nph * db;
  int want_cta = std::max(1, std::min(ga.ntile, (2 * 148) / G));
  int cs = 1;
  while (cs * 2 <= cs_max && cs * 2 <= want_cta && cs * 2 <= 8) cs *= 2;
  const size_t dump = cs > 1 ? (size_t)(kWgThreads / 32) * MW * NW * 4 * 32 * sizeof(float) : 0;
  int nst = (int)std::min<size_t>(4, (kWg2Smem - 64) / stage);
  nst = std::max(2, std::min(nst, cdiv(ga.ntile, want_cta) + 1));
  const size_t smem = std::max((size_t)nst * stage + 64, dump);
  const void* kfn = NW == 4 ? (const void*)conv_wgrad_tma_kernel<4, 4> : NW == 2 ? (const void*)conv_wgrad_tma_kernel<8, 2> : (const void*)conv_wgrad_tma_kernel<8, 1>;
  DFF_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 1) {
    cudaLaunchConfig_t qc = {};
    qc.gridDim = dim3(cs * 64); qc.blockDim = dim3(kWgThreads); qc.dynamicSmemBytes = smem;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = (unsigned)cs; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    qc.attrs = qa; qc.numAttrs = 1;
    static thread_local unsigned long long ckey[16];   // tiny per-thread cache of the driver query: (variant, smem, cluster) -> clusters
    static thread_local int cval[16], cn = 0;
    const unsigned long long key = ((unsigned long long)NW << 40) | ((unsigned long long)smem << 8) | (unsigned)cs;
    int ncl = 0;
    for (int i = 0; i < cn; ++i)
      if (ckey[i] == key) ncl = cval[i];
    if (!ncl) {
      if (cudaOccupancyMaxActiveClusters(&ncl, kfn, &qc) != cudaSuccess) { (void)cudaGetLastError(); ncl = 0; }
      if (ncl > 0 && cn < 16) { ckey[cn] = key; cval[cn] = ncl; ++cn; }
    }
    if (ncl > 0) want_cta = std::min(want_cta, std::max(cs, ncl * cs / G));
  }
  ga.ncta = std::max(cs, want_cta / cs * cs);
  ga.cs = cs;
  ga.nst = nst;
  for (int src = 0; src < (a.C1 ? 2 : 1); ++src) {
    const unsigned long long C = src ? a.C1 : a.C0;
    const unsigned long long dims[5] = {C, (unsigned long long)a.IW, (unsigned long long)a.IH, (unsigned long long)a.S, (unsigned long long)a.B};
    const unsigned long long strides[4] = {C * 2, (unsigned long long)a.IW * C * 2, (unsigned long long)a.IH * a.IW * C * 2,
                                           (unsigned long long)a.S * a.IH * a.IW * C * 2};
    const unsigned box[5] = {8u, (unsigned)RX, (unsigned)RY, (unsigned)RZ, 1u};
    DFF_TRY(encode_tmap_bf16(&ga.tx[src], src ? a.in1 : a.in0, 5, dims, strides, box));
  }
  for (int ph = 0; ph < nph; ++ph) {
    const int ooy = nph > 1 ? (ph >> 1) : a.ooy, oox = nph > 1 ? (ph & 1) : a.oox;
    const unsigned long long dims[4] = {(unsigned long long)CoS, (unsigned long long)a.OWt, (unsigned long long)a.OHt, (unsigned long long)a.S * a.B};
    const unsigned long long strides[3] = {(unsigned long long)CoS * 2 * a.osx, (unsigned long long)a.OW * CoS * 2 * a.osy,
                                           (unsigned long long)a.OH * a.OW * CoS * 2};
    const unsigned box[4] = {8u, (unsigned)TX, (unsigned)TY, 1u};
    DFF_TRY(encode_tmap_bf16(&ga.tdy[ph], (const char*)dy + ((size_t)ooy * a.OW + oox) * CoS * 2, 4, dims, strides, box));
  }
  static const bool log = getenv("DFF_B200_WGRAD_LOG") != nullptr;   // (profiling aid: one line per launch, joins with an ncu launch list)
  if (log)
    fprintf(stderr, "wgrad_tma Cin=%d+%d Cout=%d taps=%d nph=%d B=%d S=%d OHt=%d OWt=%d is=%d os=%d CK=%d TX=%d TY=%d MW=%d NW=%d MB=%d NB=%d mper=%d nsub=%d npass=%d ncta=%d ntile=%d nst=%d cs=%d smem=%zu\n",
            a.C0, a.C1, Cout, ntap, nph, a.B, a.S, a.OHt, a.OWt, a.isx, a.osx, CK, TX, TY, MW, NW, MB, NB, mper, nsub, npass, ga.ncta, ga.ntile, nst, cs, smem);
  dim3 grid(G * ga.ncta, 1, 1);
#define DFF_WG2(MW_, NW_)                                                                                                     \
  do {                                                                                                                        \
    cudaLaunchConfig_t cfg = {};                                                                                              \
    cfg.gridDim = grid; cfg.blockDim = dim3(kWgThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;                        \
    cudaLaunchAttribute at[1];                                                                                                \
    at[0].id = cudaLaunchAttributeClusterDimension;                                                                           \
    at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                            \
    cfg.attrs = at; cfg.numAttrs = 1;                                                                                         \
    DFF_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad_tma_kernel<MW_, NW_>, ga));                                                   \
  } while (0)
  if (NW == 4) DFF_WG2(4, 4);
  else if (NW == 2) DFF_WG2(8, 2);
  else DFF_WG2(8, 1);
#undef DFF_WG2
  DFF_LAUNCH_CHECK("conv_wgrad_tma");
  *handled = true;
  return 0;
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
  const int CoP = (Cout + 3) & ~3;
  a.TY = CoP <= 32 ? kWgTY : (CoP <= 64 ? kWgTY / 2 : kWgTY / 4);   // keep the dy tile <= 32 KB
  a.RZ = dzmax - a.dzmin + 1;
  a.RY = (a.TY - 1) * a.isy + (dymax - a.dymin) + 1;
  a.RX = 31 * a.isx + (dxmax - a.dxmin) + 1;
  a.RXP = a.RX | 1;
  return ((size_t)CK * a.RZ * a.RY * a.RXP + (size_t)32 * a.TY * CoP) * sizeof(float);
}

// `a`: geometry/taps/in0/in1 as for the forward launch of the same (phase of the) layer.
int launch_conv_wgrad(ConvArgs a, const void* dy, int CoS, int Cout, int Cin, int ci_base, float* dw, int ntaps_total,
                      int wt_transposed, bool bf16, cudaStream_t st) {
  if (a.C0 % 4 || a.C1 % 4) return fail(-1, "conv_wgrad: stored input channels must be multiples of 4");
  static const bool no_mma = getenv("DFF_B200_WGRAD_FFMA") != nullptr;   // A/B switch: the fp32-FMA kernel also for bf16 tensors
  if (bf16 && !no_mma) {
    bool handled = false;
    DFF_TRY(launch_conv_wgrad_tma(a, &a.taps, 1, dy, CoS, Cout, Cin, ci_base, dw, ntaps_total, wt_transposed, st, &handled));
    if (handled) return 0;
  }
  if (bf16 && !no_mma) {
    bool handled = false;
    DFF_TRY(launch_conv_wgrad_mma(a, dy, CoS, Cout, Cin, ci_base, dw, ntaps_total, wt_transposed, st, &handled));
    if (handled) return 0;
  }
  WgradArgs g{};
  int CK = (a.C0 % 8 == 0 && a.C1 % 8 == 0) ? 8 : 4;
  size_t smem = wg_plan(a, CK, Cout);
  if (smem > 100 * 1024 && CK == 8) { CK = 4; smem = wg_plan(a, CK, Cout); }
  if (smem > 227 * 1024) return fail(-5, "conv_wgrad: tile does not fit shared memory");
  g.a = a; g.dy = dy; g.CoS = CoS; g.Cout = Cout; g.Cin = Cin; g.dw = dw; g.ntaps_total = ntaps_total;
  g.wt_transposed = wt_transposed; g.ci_base = ci_base;
  dim3 grid(cdiv(a.OWt, 32), cdiv(a.OHt, a.TY), a.B * a.S);
  const int CoP = (Cout + 3) & ~3;
  const int COB = CoP >= 32 ? 32 : (CoP >= 16 ? 16 : (CoP >= 8 ? 8 : 4));
  if (CoP % COB) return fail(-5, "conv_wgrad: unsupported output channel count");
#define DFF_WG(T, CK_, COB_)                                                                                        \
  do {                                                                                                              \
    auto k = conv_wgrad_kernel<T, CK_, COB_>;                                                                       \
    DFF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
    k<<<grid, kWgThreads, smem, st>>>(g);                                                                           \
  } while (0)
#define DFF_WG_T(T)                                                               \
  do {                                                                            \
    if (CK == 8) {                                                                \
      if (COB == 32) DFF_WG(T, 8, 32); else if (COB == 16) DFF_WG(T, 8, 16);      \
      else if (COB == 8) DFF_WG(T, 8, 8); else DFF_WG(T, 8, 4);                   \
    } else {                                                                      \
      if (COB == 32) DFF_WG(T, 4, 32); else if (COB == 16) DFF_WG(T, 4, 16);      \
      else if (COB == 8) DFF_WG(T, 4, 8); else DFF_WG(T, 4, 4);                   \
    }                                                                             \
  } while (0)
  if (bf16) DFF_WG_T(__nv_bfloat16);
  else DFF_WG_T(float);
#undef DFF_WG_T
#undef DFF_WG
  DFF_LAUNCH_CHECK("conv_wgrad");
  return 0;
}

}  // namespace dff
