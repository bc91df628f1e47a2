// tcgen05 implicit-GEMM weight gradient of the 3x3 convolution for sm_100a.
//
//   dW[co][ci][dy][dx] += sum_{n,h,w} dz[n,h,w,co] * view[n,h+dy-1,w+dx-1,ci]
//
// Replaces cuDNN's Conv2d wgrad reached through `loss.backward()` (reference train.py:95) for the
// Conv2DBlock convolutions (reference model.py:8,13); `view` is the layer's logical input with the producer's
// BatchNorm/ReLU/MaxPool/Upsample/cat fused into the gather exactly as in the forward kernel (conv.cu).
//
// GEMM per tap: M = 128 output channels, N = NT <= 128 input channels, K = pixels; both operands are MN-major
// views of planar tiles [plane = 8 channels][pixel][16 B]. A tcgen05.mma in SS mode reads (M + N) x 32 bytes of
// shared memory per K step, so small N starves the tensor pipe on shared-memory bandwidth (N = 32 caps at ~40 %):
// a CTA therefore owns ONE filter row dy (3 taps x NT <= 384 TMEM columns) with NT up to 128, streams 4x16-pixel
// K tiles of its pixel range through a 3-stage ring, and red.global.add's its 3 taps into the OIHW gradient.
// dz arrives pre-split (bf16 hi/lo, written by the BatchNorm backward kernel), so its fill is a pure 32-byte copy;
// only the view operand is transformed (BN affine, ReLU, pool/upsample, bf16 split) on the way into shared memory.
//
// Warp roles (512 threads, registers re-balanced with setmaxnreg): warp 0 issues tcgen05.mma; warps 4-15
// (384 threads) are producers and issue every global load of a K tile before consuming the first; warps 4-7 also
// run the epilogue.
#include "igemm.cuh"
#include "prof.cuh"
#include <type_traits>
#include <algorithm>
#include <cstdlib>

namespace tnb {

static constexpr int kThreads = 512;
static constexpr int kFillThreads = 384;
static constexpr int kHdrBytes = 256;
static constexpr int kTileH = 4, kTileW = 16;    // 64 pixels per K tile
static constexpr int kHaloW = kTileW + 2;        // 18
static constexpr int kStages = 3;

struct WgradArgs {
  ViewDesc view;
  const uint8_t* dz;  // pre-split bf16 [N,H,W][2 (hi, lo)][Cout]
  float* dw;          // [Cout][CinReal][3][3], accumulated with atomics (must be zeroed by the caller)
  int Cout, CinReal, NT, nterms, variant;
  int ndy;  // filter rows per CTA: 3 when all 9 taps fit in TMEM (9 * NT <= 512), else 1 (grid.x carries dy)
  int tiles_h, tiles_w, ktiles, ktiles_per_cta, ncot, ncit;
  int ci_tile_base;  // first input-channel tile of this launch (concat views with two gather modes use two launches)
  float* ws;         // optional split-K slabs [splits][9][Cin_view][Cout] (see wgrad_scatter_kernel); nullptr: atomics into dw
  long long slab;    // floats per slab = 9 * Cin_view * Cout
  int fmt;           // 16-bit format of BOTH operands: 0 = fp16 hi/lo, 1 = bf16 hi/lo (mixing is an illegal instruction)
  const float* dz_mul;  // optional device scalar: the power of two dz was stored multiplied by (results are divided by it)
};

template <int MODE> TNB_DEVINL int view_off_t(const SrcDesc& s, int n, int h, int w) {
  if (MODE == SRC_AFFINE_RELU_POOL) return (n * s.Hs + 2 * h) * s.Ws + 2 * w;
  if (MODE == SRC_AFFINE_RELU_UP) return (n * s.Hs + (h >> 1)) * s.Ws + (w >> 1);
  return (n * s.Hs + h) * s.Ws + w;
}

// VMODE: gather mode of the view operand, compile-time so that every instantiation carries exactly one gather path
// (the producers are register-limited; a run-time switch over all modes costs spills in the hot loop)
template <int VMODE>
__global__ void __launch_bounds__(kThreads, 1) wgrad3x3_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const ViewDesc& V = a.view;
  const int NT = a.NT, NPL = NT / 8;
  const int TP = a.nterms > 1 ? 2 : 1;
  const int DZPL = pad_px(kTileH * kTileW) * 16;  // 66 * 16
  const int kViewPx = (kTileH + a.ndy - 1) * kHaloW;  // 72 (one filter row) or 108 (all three)
  const int VPL = pad_px(kViewPx) * 16;
  const int DZ_BYTES = TP * 16 * DZPL;
  const int STAGE = DZ_BYTES + TP * NPL * VPL;
  constexpr int S = kStages;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + S;
  uint64_t* tmem_full = full + 2 * S;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full + 2 * S + 1);
  uint8_t* st_base = smem + kHdrBytes;

  int bx = blockIdx.x;
  const int ndy = a.ndy, ngrp = 3 / ndy;
  const int dy0 = (bx % ngrp) * ndy; bx /= ngrp;  // first filter row of this CTA
  const int co0 = (bx % a.ncot) * 128;
  const int ci0 = (bx / a.ncot + a.ci_tile_base) * NT;
  const int cvalid = min(128, a.Cout - co0);
  const int npld = cvalid / 8;  // dz planes actually filled (8 or 16)
  const int kt0 = blockIdx.y * a.ktiles_per_cta;
  const int kt1 = min(a.ktiles, kt0 + a.ktiles_per_cta);
  const int tmem_cols = 512;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < S; ++i) { mbar_init(&full[i], kFillThreads); mbar_init(&empty[i], 1); }
      mbar_init(tmem_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, tmem_cols);
  }
  // dz planes that are never filled (64 output channels): zero them once so the unused accumulator rows stay finite
  if (npld < 16) {
    for (int i = tid; i < S * TP * (16 - npld) * (DZPL / 16); i += kThreads) {
      int r = i;
      const int px = r % (DZPL / 16); r /= (DZPL / 16);
      const int pl = npld + r % (16 - npld); r /= (16 - npld);
      const int term = r % TP; r /= TP;
      *reinterpret_cast<uint4*>(st_base + r * STAGE + (term * 16 + pl) * DZPL + px * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // the prologue above touched shared memory and TMEM only: it overlaps the previous kernel's tail

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // whole warp runs the uniform loops (descriptor math in uniform registers); one lane issues
      const bool lead = elect_one();
      const uint32_t idesc = make_idesc(128, NT, a.fmt /* A and B formats must match (mixed = illegal instruction) */, 1, 1);
      // MN-major planar tiles: SBO = plane stride (next 8 channels), LBO = 128 B (next 8 pixels).
      uint32_t a_lbo = 128, a_sbo = DZPL, b_lbo = 128, b_sbo = VPL;
      if (a.variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
      if (a.variant & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      const uint64_t a_desc0 = make_smem_desc(smem_u32(st_base), a_lbo, a_sbo);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(st_base) + DZ_BYTES, b_lbo, b_sbo);
      const uint32_t stage16 = STAGE >> 4, a_lo16 = (16 * DZPL) >> 4, b_lo16 = (NPL * VPL) >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t a_st = a_desc0 + (uint64_t)(s * stage16);
        const uint64_t b_st = b_desc0 + (uint64_t)(s * stage16);
        // Issue order: for each tap (= one TMEM accumulator) chain all K steps of this tile (4 rows x nterms) back
        // to back; the tensor pipe pays a drain whenever the accumulator changes, so chains must be long.
        for (int dyl = 0; dyl < ndy; ++dyl) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint32_t d_tmem = tmem_base + (dyl * 3 + dx) * NT;
#pragma unroll
            for (int r = 0; r < kTileH; ++r) {
              const uint64_t a_hi = a_st + (uint64_t)(r * kTileW);
              const uint64_t b_hi = b_st + (uint64_t)((r + dyl) * kHaloW + dx);
              const uint32_t acc = (kt != kt0 || r != 0);
              if (lead && !(a.variant & 4)) {
                umma_f16(d_tmem, a_hi, b_hi, idesc, acc);
                if (a.nterms > 1) {
                  umma_f16(d_tmem, a_hi + a_lo16, b_hi, idesc, 1);
                  umma_f16(d_tmem, a_hi, b_hi + b_lo16, idesc, 1);
                }
              }
            }
          }
        }
        if (lead) umma_commit(&empty[s]);
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (lead) umma_commit(tmem_full);
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int ftid = tid - 128;
    // dz items: plane dpl, pixels dpx0 + u*DG (u < 3);  view items: plane vpl, view pixels vpx0 + u*VG (u < 3)
    const int dpl = ftid % npld, dpx0 = ftid / npld, DG = kFillThreads / npld;  // DG = 48 or 24
    const int VG = kFillThreads / NPL;  // NPL 4..16 -> 96..24
    const bool vactive = ftid < VG * NPL;
    const int vpl = ftid % NPL, vpx0 = ftid / NPL;
    const int vch = ci0 + vpl * 8;
    const bool vsecond = vch >= V.C0;
    const SrcDesc& VS = vsecond ? V.s[1] : V.s[0];
    const int vcc = vsecond ? vch - V.C0 : vch;
    float sc[8], sh[8];
    if (VMODE != SRC_IDENTITY && VMODE != SRC_PRESPLIT && vactive) { ld8(VS.scale + vcc, sc); ld8(VS.shift + vcc, sh); }
    const uint8_t* dz_base = a.dz + (size_t)(co0 + dpl * 8) * 2;  // dz: [pixel][2 (hi, lo)][Cout] bf16
    const size_t dz_pix_stride = (size_t)a.Cout * 4;
    const int dz_lo = a.Cout * 2;
    const int per_img = a.tiles_h * a.tiles_w;

    // tile-independent part of this thread's (up to) 3 + 3 copies of a K tile (pre-split operands only): shared-memory
    // offset inside the stage (-1 = no copy), pixel offset (dh, dw) from the tile origin for the bounds test, and byte
    // offset from the tile origin's address. A half-resolution source (SRC_PRESPLIT_UP) is addressed at (h >> 1, w >> 1);
    // tile origins are even, so (h0 + dh) >> 1 = (h0 >> 1) + (dh >> 1) with an arithmetic shift.
    struct CopyItem { int soff, dh, dw, goff; };
    CopyItem cdz[3], cvw[3];
    const int cvup = VS.mode == SRC_PRESPLIT_UP ? 1 : 0;
    const int cvHs = VS.Hs, cvWs = VS.Ws;
    const size_t cvstride = (size_t)VS.C * 4;  // [pixel][2][C] 16-bit
    const int cv_lo = VS.C * 2;
    const uint8_t* cvbase = reinterpret_cast<const uint8_t*>(VS.ptr);
    if (VMODE == SRC_PRESPLIT) {
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int px = dpx0 + u * DG;
        cdz[u].soff = px < kTileH * kTileW ? dpl * DZPL + px * 16 : -1;
        cdz[u].dh = px >> 4; cdz[u].dw = px & 15;
        cdz[u].goff = (cdz[u].dh * V.W + cdz[u].dw) * (int)dz_pix_stride + (co0 + dpl * 8) * 2;
        const int p = vpx0 + u * VG;
        const int hr = p / kHaloW, hc = p - hr * kHaloW;
        cvw[u].soff = (vactive && p < kViewPx) ? DZ_BYTES + vpl * VPL + p * 16 : -1;
        cvw[u].dh = hr + dy0 - 1; cvw[u].dw = hc - 1;
        cvw[u].goff = ((cvw[u].dh >> cvup) * cvWs + (cvw[u].dw >> cvup)) * (int)cvstride + vcc * 2;
      }
    }
    int s = 0;
    uint32_t ph = 0;
    // (n, h0, w0) of the K tile advance incrementally: no divisions in the loop
    int n_i = kt0 / per_img, th_i = (kt0 - n_i * per_img) / a.tiles_w, tw_i = kt0 - n_i * per_img - th_i * a.tiles_w;
    for (int kt = kt0; kt < kt1; ++kt) {
      const int n = n_i, h0 = th_i * kTileH, w0 = tw_i * kTileW;
      if (++tw_i == a.tiles_w) { tw_i = 0; if (++th_i == a.tiles_h) { th_i = 0; ++n_i; } }
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* stage = st_base + s * STAGE;
      uint8_t* dzp = stage + dpl * DZPL;
      uint8_t* vwp = stage + DZ_BYTES + vpl * VPL;

      constexpr bool kCopy = (VMODE == SRC_PRESPLIT);  // both operands are plain copies
      if (kCopy && (a.variant & 8)) {  // ablation: barrier traffic only
        cp_async_mbar_arrive_noinc(&full[s]);
        if (++s == S) { s = 0; ph ^= 1; }
        continue;
      }
      if (kCopy) {
        // Both operands are already (hi, lo) 16-bit pairs in HBM: the fill is 16-byte cp.async copies straight into
        // the planar tiles. The thread never waits for its own loads (the stage's mbarrier is armed with a
        // cp.async-completion arrival), so up to kStages K tiles of HBM latency are in flight per thread.
        // The producers are bound by their own instruction stream (12 warps, ~30 instructions per copy when every
        // address is rebuilt from (n, h, w)), so everything that does not depend on the tile is hoisted into
        // CopyItem: per K tile a copy costs one bounds test, one 64-bit add and the LDGSTS pair.
        const uint8_t* dzt = a.dz + ((size_t)(n * V.H + h0) * V.W + w0) * dz_pix_stride;  // warp-uniform tile origins
        const uint8_t* vwt = cvbase + ((ptrdiff_t)(n * cvHs + (h0 >> cvup)) * cvWs + (w0 >> cvup)) * (ptrdiff_t)cvstride;
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          if (cdz[u].soff >= 0) {
            const bool ok = h0 + cdz[u].dh < V.H && w0 + cdz[u].dw < V.W;
            const uint8_t* q = ok ? dzt + cdz[u].goff : a.dz;
            cp_async16(stage + cdz[u].soff, q, ok ? 16u : 0u, true);
            if (a.nterms > 1) cp_async16(stage + cdz[u].soff + 16 * DZPL, q + dz_lo, ok ? 16u : 0u, true);
          }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          if (cvw[u].soff >= 0) {
            const bool ok = (unsigned)(h0 + cvw[u].dh) < (unsigned)V.H && (unsigned)(w0 + cvw[u].dw) < (unsigned)V.W;
            const uint8_t* q = ok ? vwt + cvw[u].goff : cvbase;
            cp_async16(stage + cvw[u].soff, q, ok ? 16u : 0u, true);
            if (a.nterms > 1) cp_async16(stage + cvw[u].soff + NPL * VPL, q + cv_lo, ok ? 16u : 0u, true);
          }
        }
        cp_async_mbar_arrive_noinc(&full[s]);
        if (++s == S) { s = 0; ph ^= 1; }
        continue;
      }
      // ---- issue every load of this K tile: dz copies first, then the view gathers ----
      Raw8 draw[3];
      bool dok[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int px = dpx0 + u * DG;
        const int h = h0 + (px >> 4), w = w0 + (px & 15);
        dok[u] = px < kTileH * kTileW && h < V.H && w < V.W;
        if (dok[u]) {
          const uint8_t* q = dz_base + ((size_t)(n * V.H + h) * V.W + w) * dz_pix_stride;
          const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(q)), x1 = __ldg(reinterpret_cast<const uint4*>(q + dz_lo));
          draw[u].a = *reinterpret_cast<const float4*>(&x0);
          draw[u].b = *reinterpret_cast<const float4*>(&x1);
        }
      }
      auto view_run = [&](auto mode_tag, auto batch_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        constexpr int U = decltype(batch_tag)::value;
        for (int p0 = vpx0; p0 < kViewPx; p0 += VG * U) {
          Raw8 raw[U][RawCount<MODE>::value];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = p0 + u * VG;
            const int hr = p / kHaloW, hc = p - hr * kHaloW;
            const int h = h0 + hr + dy0 - 1, w = w0 - 1 + hc;
            ok[u] = p < kViewPx && h >= 0 && h < V.H && w >= 0 && w < V.W;
            if (ok[u]) view_issue<MODE>(VS, view_off_t<MODE>(VS, n, h, w), vcc, raw[u]);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = p0 + u * VG;
            if (p < kViewPx) {
              uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
              if (ok[u]) {
                if (MODE == SRC_PRESPLIT) {  // materialised bf16 view (tnb_view_presplit): pure copy
                  hi = *reinterpret_cast<const uint4*>(&raw[u][0].a);
                  lo = *reinterpret_cast<const uint4*>(&raw[u][0].b);
                } else {
                  float v[8];
                  view_finish<MODE>(raw[u], sc, sh, 1.f, v);
                  if (a.fmt == 0) split8<0>(v, hi, lo); else split8<1>(v, hi, lo);
                }
              }
              *reinterpret_cast<uint4*>(vwp + p * 16) = hi;
              if (a.nterms > 1) *reinterpret_cast<uint4*>(vwp + p * 16 + NPL * VPL) = lo;
            }
          }
        }
      };
      if (vactive)
        view_run(std::integral_constant<int, VMODE>{},
                 std::integral_constant<int, (VMODE == SRC_AFFINE_RELU_POOL) ? 1 : 3>{});
      // ---- dz: already (hi, lo) bf16 -> two 16-byte stores, no arithmetic ----
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int px = dpx0 + u * DG;
        if (px < kTileH * kTileW) {
          uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
          if (dok[u]) {
            hi = *reinterpret_cast<const uint4*>(&draw[u].a);
            lo = *reinterpret_cast<const uint4*>(&draw[u].b);
          }
          *reinterpret_cast<uint4*>(dzp + px * 16) = hi;
          if (a.nterms > 1) *reinterpret_cast<uint4*>(dzp + px * 16 + 16 * DZPL) = lo;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full[s]);
      if (++s == S) { s = 0; ph ^= 1; }
    }

    if (warp < 8) {
      const int q = warp & 3;
      const float out_mul = a.dz_mul != nullptr ? 1.f / *a.dz_mul : 1.f;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int row = 32 * q + lane;
      for (int t = 0; t < 3 * ndy; ++t) {
        for (int col0 = 0; col0 < NT; col0 += 16) {
          uint32_t rg[16];
          tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * NT + col0), rg);
          tmem_ld_wait();
          if (row < cvalid) {
            // With a slab buffer every split-K CTA owns its slab: plain stores (lanes = consecutive output channels, one
            // 128-byte store per warp instruction), summed in split order by wgrad_scatter_kernel - bit-identical from
            // run to run, like the reference's cudnn.deterministic = True (train.py:205). Without one: atomics into dw.
            float* slab = a.ws != nullptr ? a.ws + (size_t)blockIdx.y * a.slab : nullptr;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ci = ci0 + col0 + j;
              const float v = kt1 > kt0 ? __uint_as_float(rg[j]) * out_mul : 0.f;
              if (slab != nullptr)
                slab[((size_t)(dy0 * 3 + t) * V.C + ci) * a.Cout + co0 + row] = v;
              else if (ci < a.CinReal && kt1 > kt0)
                atomicAdd(a.dw + ((size_t)(co0 + row) * a.CinReal + ci) * 9 + dy0 * 3 + t, v);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// =============================================================================================
// CTA-pair variant (tcgen05 cta_group::2) for Cout % 256 == 0 with a 128-channel input tile and pre-split operands.
//
// The single-CTA kernel above is bound by what it has to move, not by the tensor pipe: an M = 128, N = 128 MMA reads
// (128 + 128) x 32 B of shared memory per K step - 64 clocks at 128 B/clk, exactly its math time - and every K tile
// costs 68 KB of L2 -> SM traffic (the fill alone runs at ~6.7 TB/s chip-wide). Two CTAs of a cluster (the two SMs
// of a TPC) computing 256 output channels x the SAME 128 input channels share the view operand: each CTA loads its
// own 128 dz rows and HALF of the view tile (64 channels), the pair's tensor cores read both halves through
// distributed shared memory. Per CTA: 50 KB per K tile instead of 68, (128 + 64) x 32 B per MMA instead of 256 x 32
// (48 clocks < 64 of math), and the smaller stage buys a fourth pipeline stage.
//
// Protocol: rank 0 issues every tcgen05.mma.cta_group::2 (M = 256: TMEM lanes 0-127 of each CTA hold that CTA's 128
// output channels) and multicasts the commits to both CTAs' empty / tmem_full barriers. Rank 1's producers arrive on
// rank 1's own full barrier; its warp 0 relays each completed phase to rank 0's full barrier (expected count
// kFillThreads + 1) with a cluster-scope release arrive.
// =============================================================================================
static constexpr int kPStages = 4;
// plane stride in pixels: mode 0 = 1 (mod 8) pixels, the fastest fill measured (256->256: 0.221 ms; 2 (mod 8) = pad_px
// 0.226, 4 (mod 8) 0.248, 128-byte aligned planes 0.338: the cp.async stores of neighbouring planes then collide).
// Modes 1-3 (variant bits 7-8) keep the other strides reachable for tools/ablate_pair.py.
__host__ __device__ inline int pad_sel(int px, int mode) {
  const int r = px & 7, want = mode == 0 ? 1 : mode == 1 ? 0 : mode == 2 ? 2 : 4;
  return px + ((want - r) & 7);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    wgrad3x3_pair_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const ViewDesc& V = a.view;
  const uint32_t rank = cluster_ctarank();
  constexpr int NPLC = 8;  // view planes per CTA: 64 of the pair's 128 input channels
  const int TP = a.nterms > 1 ? 2 : 1;
  const int padm = (a.variant >> 7) & 3;  // experiment: plane stride padding (0 = pad_px)
  const int DZPL = pad_sel(kTileH * kTileW, padm) * 16;
  const int kViewPx = kTileH * kHaloW;  // one filter row per CTA pair: 4 x 18 halo pixels
  const int VPL = pad_sel(kViewPx, padm) * 16;
  const int DZ_BYTES = TP * 16 * DZPL;
  const int STAGE = DZ_BYTES + TP * NPLC * VPL;
  constexpr int S = kPStages;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + S;
  uint64_t* tmem_full = full + 2 * S;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full + 2 * S + 1);
  uint8_t* st_base = smem + kHdrBytes;

  int bx = blockIdx.x >> 1;  // the two CTAs of a pair are neighbours in x
  const int dy0 = bx % 3; bx /= 3;
  const int npair = a.ncot >> 1;
  const int co0 = ((bx % npair) * 2 + (int)rank) * 128;
  const int ci0 = (bx / npair) * 128;
  const int kt0 = blockIdx.y * a.ktiles_per_cta;
  const int kt1 = min(a.ktiles, kt0 + a.ktiles_per_cta);
  const int tmem_cols = 512;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < S; ++i) { mbar_init(&full[i], kFillThreads + (rank == 0 ? 1 : 0)); mbar_init(&empty[i], 1); }
      mbar_init(tmem_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_ptr, tmem_cols);
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && rank == 0) {
      const bool lead = elect_one();
      const uint32_t idesc = make_idesc(256, 128, a.fmt, 1, 1);  // both operands MN-major and of one 16-bit format, M = 2 x 128
      const uint64_t a_desc0 = make_smem_desc(smem_u32(st_base), 128, DZPL);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(st_base) + DZ_BYTES, 128, VPL);
      const uint32_t stage16 = STAGE >> 4, a_lo16 = (16 * DZPL) >> 4, b_lo16 = (NPLC * VPL) >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait_cluster(&full[s], ph);
        tc_fence_after();
        const uint64_t a_st = a_desc0 + (uint64_t)(s * stage16);
        const uint64_t b_st = b_desc0 + (uint64_t)(s * stage16);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const uint32_t d_tmem = tmem_base + dx * 128;
#pragma unroll
          for (int r = 0; r < kTileH; ++r) {
            const uint64_t a_hi = a_st + (uint64_t)(r * kTileW);
            const uint64_t b_hi = b_st + (uint64_t)(r * kHaloW + dx);
            const uint32_t acc = (kt != kt0 || r != 0);
            if (lead && !(a.variant & 4)) {
              umma_f16_pair(d_tmem, a_hi, b_hi, idesc, acc);
              if (a.nterms > 1) {
                umma_f16_pair(d_tmem, a_hi + a_lo16, b_hi, idesc, 1);
                umma_f16_pair(d_tmem, a_hi, b_hi + b_lo16, idesc, 1);
              }
            }
          }
        }
        if (lead) umma_commit_pair(&empty[s]);
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (lead) umma_commit_pair(tmem_full);
      __syncwarp();
    } else if (warp == 0) {
      // rank 1: forward every completed fill phase of this CTA to the MMA issuer's barrier
      int s = 0;
      uint32_t ph = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(&full[s], ph);
        fence_proxy_async_smem();
        if (lane == 0) mbar_arrive_remote(&full[s], 0);
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int ftid = tid - 128;
    const int dpl = ftid & 15, dpx0 = ftid >> 4, DG = kFillThreads / 16;      // dz: 16 planes x 64 pixels
    const int vpl = ftid & (NPLC - 1), vpx0 = ftid / NPLC, VG = kFillThreads / NPLC;  // view: 8 planes x 72 pixels
    const bool vsecond = ci0 >= V.C0;  // a 128-channel tile never straddles the two sources (pick_nt)
    const SrcDesc& VS = vsecond ? V.s[1] : V.s[0];
    const int vcc = (vsecond ? ci0 - V.C0 : ci0) + (int)rank * 64 + vpl * 8;
    const size_t dz_pix_stride = (size_t)a.Cout * 4;  // dz: [pixel][2 (hi, lo)][Cout] bf16
    const int dz_lo = a.Cout * 2;
    const int per_img = a.tiles_h * a.tiles_w;
    struct CopyItem { int soff, dh, dw, goff; };  // see wgrad3x3_kernel
    CopyItem cdz[3], cvw[3];
    const int cvup = VS.mode == SRC_PRESPLIT_UP ? 1 : 0;
    const int cvHs = VS.Hs, cvWs = VS.Ws;
    const size_t cvstride = (size_t)VS.C * 4;
    const int cv_lo = VS.C * 2;
    const uint8_t* cvbase = reinterpret_cast<const uint8_t*>(VS.ptr);
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int px = dpx0 + u * DG;
      cdz[u].soff = px < kTileH * kTileW ? dpl * DZPL + px * 16 : -1;
      cdz[u].dh = px >> 4; cdz[u].dw = px & 15;
      cdz[u].goff = (cdz[u].dh * V.W + cdz[u].dw) * (int)dz_pix_stride + (co0 + dpl * 8) * 2;
      const int p = vpx0 + u * VG;
      const int hr = p / kHaloW, hc = p - hr * kHaloW;
      cvw[u].soff = p < kViewPx ? DZ_BYTES + vpl * VPL + p * 16 : -1;
      cvw[u].dh = hr + dy0 - 1; cvw[u].dw = hc - 1;
      cvw[u].goff = ((cvw[u].dh >> cvup) * cvWs + (cvw[u].dw >> cvup)) * (int)cvstride + vcc * 2;
    }
    int s = 0;
    uint32_t ph = 0;
    int n_i = kt0 / per_img, th_i = (kt0 - n_i * per_img) / a.tiles_w, tw_i = kt0 - n_i * per_img - th_i * a.tiles_w;
    for (int kt = kt0; kt < kt1; ++kt) {
      const int n = n_i, h0 = th_i * kTileH, w0 = tw_i * kTileW;
      if (++tw_i == a.tiles_w) { tw_i = 0; if (++th_i == a.tiles_h) { th_i = 0; ++n_i; } }
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* stage = st_base + s * STAGE;
      const uint8_t* dzt = a.dz + ((size_t)(n * V.H + h0) * V.W + w0) * dz_pix_stride;
      const uint8_t* vwt = cvbase + ((ptrdiff_t)(n * cvHs + (h0 >> cvup)) * cvWs + (w0 >> cvup)) * (ptrdiff_t)cvstride;
      if (!(a.variant & 8)) {  // (ablation bit 8: barrier traffic only)
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        if (cdz[u].soff >= 0) {
          const bool ok = h0 + cdz[u].dh < V.H && w0 + cdz[u].dw < V.W;
          const uint8_t* q = ok ? dzt + cdz[u].goff : a.dz;
          cp_async16(stage + cdz[u].soff, q, ok ? 16u : 0u, true);
          if (a.nterms > 1) cp_async16(stage + cdz[u].soff + 16 * DZPL, q + dz_lo, ok ? 16u : 0u, true);
        }
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        if (cvw[u].soff >= 0) {
          const bool ok = (unsigned)(h0 + cvw[u].dh) < (unsigned)V.H && (unsigned)(w0 + cvw[u].dw) < (unsigned)V.W;
          const uint8_t* q = ok ? vwt + cvw[u].goff : cvbase;
          cp_async16(stage + cvw[u].soff, q, ok ? 16u : 0u, true);
          if (a.nterms > 1) cp_async16(stage + cvw[u].soff + NPLC * VPL, q + cv_lo, ok ? 16u : 0u, true);
        }
      }
      }
      cp_async_mbar_arrive_noinc(&full[s]);
      if (++s == S) { s = 0; ph ^= 1; }
    }

    if (warp < 8) {
      const int q = warp & 3;
      const float out_mul = a.dz_mul != nullptr ? 1.f / *a.dz_mul : 1.f;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int row = 32 * q + lane;  // TMEM lane = output channel co0 + row of THIS CTA
      for (int t = 0; t < 3; ++t) {
        for (int col0 = 0; col0 < 128; col0 += 16) {
          uint32_t rg[16];
          tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * 128 + col0), rg);
          tmem_ld_wait();
          {
            float* slab = a.ws != nullptr ? a.ws + (size_t)blockIdx.y * a.slab : nullptr;  // see wgrad3x3_kernel
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ci = ci0 + col0 + j;
              const float v = kt1 > kt0 ? __uint_as_float(rg[j]) * out_mul : 0.f;
              if (slab != nullptr)
                slab[((size_t)(dy0 * 3 + t) * V.C + ci) * a.Cout + co0 + row] = v;
              else if (ci < a.CinReal && kt1 > kt0)
                atomicAdd(a.dw + ((size_t)(co0 + row) * a.CinReal + ci) * 9 + dy0 * 3 + t, v);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's tensor core reads this CTA's shared memory until the last commit has landed
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
}

// =============================================================================================
// Tap-stacked variant for 64 output channels (the full-resolution layers, where the generic kernel wastes half of
// every MMA's M = 128 rows on zero planes and runs three CTAs - one per filter row - over the same pixels).
//
// Roles are swapped: M = input channels, N = 64 output channels, and the input halo tile is stored
// [halo row][plane = 8 channels][halo column][16 B]. With SBO = one plane (RP bytes) the 16 eight-channel row groups
// of an M = 128 operand walk 8 planes of halo row r and then - with no gap - the 8 planes of halo row r + 1: ONE
// tcgen05.mma computes the taps (dy, dx) and (dy + 1, dx) of a 64-channel tile (or dy = 0..3 of a 32-channel tile)
// from the same shared-memory bytes. A CTA therefore owns all 9 taps: 6 MMAs per K step instead of 9 (3 for the
// 27(32)-channel first layer), the dz tile and the input tile are fetched once instead of three times.
// Accumulators (TMEM columns): [dx] = rows {dy 0 | dy 1}, [3 + dx] = rows {dy 2 | unused}; 32-channel tiles: [dx] =
// rows {dy 0 | 1 | 2 | unused}. Both operands must be pre-split bf16 (tnb_view_presplit / BatchNorm backward).
// =============================================================================================
static constexpr int kSStages = 4;
static constexpr int kSRP = 19 * 16;        // bytes per (halo row, plane): 18 columns + 1 pad (odd -> spread banks)
static constexpr int kSRows = kTileH + 3;   // 6 halo rows filled + 1 slack row read by the unused half of dy = 2

struct WgradSArgs {
  const uint8_t* view;  // pre-split bf16 [N,H,W][2][C] (first source: channels [0, C0))
  const uint8_t* view1; // second source of a concat view, channels [C0, C)
  int C0, Cs0, Cs1;     // channels of the first source in the view; channel strides of the two source tensors
  int up0, Hs0, Ws0;    // first source is at half resolution (SRC_PRESPLIT_UP) with these dimensions
  const uint8_t* dz;    // pre-split bf16 [N,H,W][2][Cout]
  float* dw;
  int N, H, W, C, Cout, CinReal, P /*planes per ci tile: 8 or 4*/, nterms, ncit, ncot;
  int tiles_h, tiles_w, ktiles, ktiles_per_cta;
  float* ws;  // optional split-K slabs [splits][9][Cout][C] (see wgrad_scatter_kernel); nullptr: atomics into dw
  long long slab;  // floats per slab = 9 * Cout * C
  int fmt;         // see WgradArgs
  const float* dz_mul;
};

__global__ void __launch_bounds__(kThreads, 1) wgrad3x3_stacked_kernel(const __grid_constant__ WgradSArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int P = a.P, CI = P * 8;
  const int TP = a.nterms > 1 ? 2 : 1;
  const int DZPL = pad_px(kTileH * kTileW) * 16;
  const int A_TERM = kSRows * P * kSRP;  // one term of the input tile
  const int A_BYTES = TP * A_TERM;
  const int B_TERM = 8 * DZPL;
  const int STAGE = A_BYTES + TP * B_TERM;
  constexpr int S = kSStages;
  const int nacc = P == 8 ? 6 : 3;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + S;
  uint64_t* tmem_full = full + 2 * S;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full + 2 * S + 1);
  uint8_t* st_base = smem + kHdrBytes;

  const int ci0 = (blockIdx.x % a.ncit) * CI;
  const int co0 = (blockIdx.x / a.ncit) * 64;
  const int kt0 = blockIdx.y * a.ktiles_per_cta;
  const int kt1 = min(a.ktiles, kt0 + a.ktiles_per_cta);
  const int tmem_cols = 512;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < S; ++i) { mbar_init(&full[i], kFillThreads); mbar_init(&empty[i], 1); }
      mbar_init(tmem_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, tmem_cols);
  }
  // the slack halo row is never written by the fill: clear it once (its products land in accumulator rows nobody reads,
  // but they should not be signalling garbage)
  for (int i = tid; i < S * TP * P * (kSRP / 16); i += kThreads) {
    int r = i;
    const int c = r % (kSRP / 16); r /= (kSRP / 16);
    const int pl = r % P; r /= P;
    const int term = r % TP; r /= TP;
    *reinterpret_cast<uint4*>(st_base + r * STAGE + term * A_TERM + ((kSRows - 1) * P + pl) * kSRP + c * 16) =
        make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      const bool lead = elect_one();
      const uint32_t idesc = make_idesc(128, 64, a.fmt, 1, 1);  // both operands MN-major and of one 16-bit format
      const uint64_t a_desc0 = make_smem_desc(smem_u32(st_base), 128, kSRP);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(st_base) + A_BYTES, 128, DZPL);
      const uint32_t stage16 = STAGE >> 4, a_lo16 = A_TERM >> 4, b_lo16 = B_TERM >> 4;
      const uint32_t row16 = (P * kSRP) >> 4;  // one halo row of the input tile
      int s = 0;
      uint32_t ph = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t a_st = a_desc0 + (uint64_t)(s * stage16);
        const uint64_t b_st = b_desc0 + (uint64_t)(s * stage16);
        for (int acc_i = 0; acc_i < nacc; ++acc_i) {  // one accumulator = one chain of 4 rows x nterms MMAs
          const int dx = acc_i % 3, dyb = (acc_i / 3) * 2;
          const uint32_t d_tmem = tmem_base + acc_i * 64;
#pragma unroll
          for (int r = 0; r < kTileH; ++r) {
            const uint64_t a_hi = a_st + (uint64_t)((r + dyb) * row16 + dx);
            const uint64_t b_hi = b_st + (uint64_t)(r * kTileW);
            const uint32_t acc = (kt != kt0 || r != 0);
            if (lead) {
              umma_f16(d_tmem, a_hi, b_hi, idesc, acc);
              if (a.nterms > 1) {
                umma_f16(d_tmem, a_hi + a_lo16, b_hi, idesc, 1);
                umma_f16(d_tmem, a_hi, b_hi + b_lo16, idesc, 1);
              }
            }
          }
        }
        if (lead) umma_commit(&empty[s]);
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (lead) umma_commit(tmem_full);
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int ftid = tid - 128;
    // The copy list of a K tile is the same for every tile: item -> (shared offset, pixel offset, channel offset).
    // dz items: 64 pixels x 8 planes; input items: 6 x 18 halo pixels x P planes. 8 (or P) neighbouring threads copy
    // the 128 (64) contiguous bytes of one pixel.
    // The producers are bound by their own instruction stream, so everything tile-independent is hoisted: per item the
    // shared-memory offsets of both terms, the pixel offset from the tile origin (bounds test) and the byte offset from
    // the tile origin's address; per K tile the two (warp-uniform) tile-origin addresses.
    constexpr int MAXI = 4;
    const int ndz = kTileH * kTileW * 8, nin = (kTileH + 2) * kHaloW * P;
    // a CTA's 64 (32) input channels come from ONE source of the view (launcher: C0 is a multiple of the tile)
    const bool second = ci0 >= a.C0;
    const uint8_t* vsrc = second ? a.view1 : a.view;
    const int vC = second ? a.Cs1 : a.Cs0;
    const int vshift = (!second && a.up0) ? 1 : 0;  // half-resolution source: pixel (h, w) of the view is its (h/2, w/2)
    const int vH = vshift ? a.Hs0 : a.H, vW = vshift ? a.Ws0 : a.W;
    const size_t dz_stride = (size_t)a.Cout * 4, v_stride = (size_t)vC * 4;
    const int dz_lo = a.Cout * 2, v_lo = vC * 2;
    const int per_img = a.tiles_h * a.tiles_w;
    int soff[MAXI], slo[MAXI], dh[MAXI], dwv[MAXI], goff[MAXI], glo[MAXI];  // soff < 0: no item
    bool isdz[MAXI];
#pragma unroll
    for (int u = 0; u < MAXI; ++u) {
      const int it = ftid + u * kFillThreads;
      isdz[u] = it < ndz;
      if (isdz[u]) {
        const int pl = it & 7, px = it >> 3;
        soff[u] = A_BYTES + pl * DZPL + px * 16; slo[u] = B_TERM; glo[u] = dz_lo;
        dh[u] = px >> 4; dwv[u] = px & 15;
        goff[u] = (dh[u] * a.W + dwv[u]) * (int)dz_stride + (co0 + pl * 8) * 2;
      } else {
        const int j = it - ndz;
        const int pl = j % P, hp = j / P;
        const int hr = hp / kHaloW, hc = hp - hr * kHaloW;
        soff[u] = it < ndz + nin ? (hr * P + pl) * kSRP + hc * 16 : -1; slo[u] = A_TERM; glo[u] = v_lo;
        dh[u] = hr - 1; dwv[u] = hc - 1;
        // tile origins are even, so (h0 + dh) >> 1 = (h0 >> 1) + (dh >> 1) with an arithmetic shift
        goff[u] = ((dh[u] >> vshift) * vW + (dwv[u] >> vshift)) * (int)v_stride + (ci0 - (second ? a.C0 : 0) + pl * 8) * 2;
      }
    }

    int s = 0;
    uint32_t ph = 0;
    int n_i = kt0 / per_img, th_i = (kt0 - n_i * per_img) / a.tiles_w, tw_i = kt0 - n_i * per_img - th_i * a.tiles_w;
    for (int kt = kt0; kt < kt1; ++kt) {
      const int n = n_i, h0 = th_i * kTileH, w0 = tw_i * kTileW;
      if (++tw_i == a.tiles_w) { tw_i = 0; if (++th_i == a.tiles_h) { th_i = 0; ++n_i; } }
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* stage = st_base + s * STAGE;
      const uint8_t* dzt = a.dz + ((size_t)(n * a.H + h0) * a.W + w0) * dz_stride;
      const uint8_t* vwt = vsrc + ((size_t)(n * vH + (h0 >> vshift)) * vW + (w0 >> vshift)) * v_stride;
#pragma unroll
      for (int u = 0; u < MAXI; ++u) {
        if (soff[u] >= 0) {
          const bool ok = (unsigned)(h0 + dh[u]) < (unsigned)a.H && (unsigned)(w0 + dwv[u]) < (unsigned)a.W;
          const uint8_t* q = ok ? (isdz[u] ? dzt : vwt) + goff[u] : a.dz;
          cp_async16(stage + soff[u], q, ok ? 16u : 0u, true);
          if (a.nterms > 1) cp_async16(stage + soff[u] + slo[u], q + glo[u], ok ? 16u : 0u, true);
        }
      }
      cp_async_mbar_arrive_noinc(&full[s]);
      if (++s == S) { s = 0; ph ^= 1; }
    }

    if (warp < 8) {
      const int q = warp & 3;
      const float out_mul = a.dz_mul != nullptr ? 1.f / *a.dz_mul : 1.f;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int m = 32 * q + lane;  // accumulator row = (tap slot, input channel)
      const int slot = m / CI, ci = ci0 + m % CI;
      for (int acc_i = 0; acc_i < nacc; ++acc_i) {
        const int dx = acc_i % 3;
        const int dy = P == 8 ? (acc_i / 3) * 2 + slot : slot;
        const bool live = (P == 8 ? (acc_i < 3 || slot == 0) : slot < 3) && ci < a.CinReal;
        float* slab = a.ws != nullptr ? a.ws + (size_t)blockIdx.y * a.slab : nullptr;  // see wgrad3x3_kernel
        for (int col0 = 0; col0 < 64; col0 += 16) {
          uint32_t rg[16];
          tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc_i * 64 + col0), rg);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float v = kt1 > kt0 ? __uint_as_float(rg[j]) * out_mul : 0.f;
              if (slab != nullptr)  // lanes = consecutive input channels: one 128-byte store per warp instruction
                slab[((size_t)(dy * 3 + dx) * a.Cout + co0 + col0 + j) * a.C + ci] = v;
              else if (kt1 > kt0)
                atomicAdd(a.dw + ((size_t)(co0 + col0 + j) * a.CinReal + ci) * 9 + dy * 3 + dx, v);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

static bool stacked_applicable(const ViewDesc& view, int Cout) {
  if (Cout != 64 || !(view.C == 32 || view.C % 64 == 0)) return false;
  const bool two = view.C0 < view.C;
  if (two && view.C0 % 64 != 0) return false;
  const SrcDesc& s0 = view.s[0];
  const bool ok0 = (s0.mode == SRC_PRESPLIT && s0.Hs == view.H && s0.Ws == view.W) ||
                   (s0.mode == SRC_PRESPLIT_UP && s0.Hs * 2 == view.H && s0.Ws * 2 == view.W);
  const bool ok1 = !two || (view.s[1].mode == SRC_PRESPLIT && view.s[1].Hs == view.H && view.s[1].Ws == view.W);
  return ok0 && ok1;
}

// dw[co][ci][tap] = sum over the split-K slabs, in split order (fixed order: the result is bit-identical from run to run),
// written to the reference's OIHW layout with plain stores.
// layout 0: slab[tap][ci][co] (generic / pair kernel), layout 1: slab[tap][co][ci] (stacked kernel); ci over the padded
// view channels.
__global__ void __launch_bounds__(256) wgrad_scatter_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cout,
                                                            int Cin, int CinReal, int layout, int splits) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)9 * Cout * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i / ((long long)Cout * Cin));
    const int r = (int)(i - (long long)tap * Cout * Cin);
    const int co = layout == 0 ? r % Cout : r / Cin, ci = layout == 0 ? r / Cout : r % Cin;
    if (ci < CinReal) {
      float acc = 0.f;
      for (int sp = 0; sp < splits; ++sp) acc += ws[(size_t)sp * total + i];
      dw[((size_t)co * CinReal + ci) * 9 + tap] = acc;
    }
  }
}
static int launch_wgrad_scatter(const float* ws, float* dw, int Cout, int Cin, int CinReal, int layout, int splits,
                                cudaStream_t st) {
  const long long total = (long long)9 * Cout * Cin;
  return launch_pdl(wgrad_scatter_kernel, dim3((unsigned)std::min<long long>((total + 255) / 256, 148 * 8)), dim3(256), 0, st,
                    ws, dw, Cout, Cin, CinReal, layout, splits);
}

// input-channel tile: divides Cin and (for concat views) the first source, so that a CTA's channels come from ONE
// source (one gather mode per CTA: no divergence between the up-sampled and the skip half)
static int pick_nt(int cin, int c0) {
  const int cand[] = {128, 96, 64, 48, 32};
  for (int c : cand)
    if (cin % c == 0 && (c0 == cin || c0 % c == 0)) return c;
  return 0;
}

// Which kernel, which tiling, how many split-K CTAs: shared by the launcher and the scratch-size query.
struct WgradPlan {
  int kind;  // 0 single-CTA generic, 1 CTA pair, 2 tap-stacked
  int NT, ndy, ncot, ncit, P, gx, splits, ktiles, ktiles_per_cta, tiles_h, tiles_w, mode0, mode1;
};
static int plan_wgrad(const ViewDesc& view, int Cout, int variant, WgradPlan* p) {
  TNB_REQUIRE(view.C % 32 == 0 && Cout % 64 == 0, "wgrad3x3: unsupported channels Cin=%d Cout=%d", view.C, Cout);
  p->tiles_h = (view.H + kTileH - 1) / kTileH;
  p->tiles_w = (view.W + kTileW - 1) / kTileW;
  p->ktiles = view.N * p->tiles_h * p->tiles_w;
  auto canon = [](int m) { return m == SRC_PRESPLIT_UP ? (int)SRC_PRESPLIT : m; };
  p->mode0 = canon(view.s[0].mode);
  p->mode1 = canon((view.C0 < view.C) ? view.s[1].mode : view.s[0].mode);
  p->NT = 0; p->ndy = 1; p->P = 0;
  if (!(variant & 32) && stacked_applicable(view, Cout)) {  // variant bit 32: force the generic kernel (tests, ablation)
    p->kind = 2;
    p->P = view.C == 32 ? 4 : 8;
    p->ncit = view.C / (p->P * 8); p->ncot = Cout / 64;
    p->gx = p->ncit * p->ncot;
  } else {
    p->NT = pick_nt(view.C, view.C0);
    TNB_REQUIRE(p->NT > 0, "wgrad3x3: no input-channel tile for Cin=%d (first source %d)", view.C, view.C0);
    p->ncot = (Cout + 127) / 128;
    p->ncit = view.C / p->NT;
    p->ndy = (9 * p->NT <= 512) ? 3 : 1;
    p->gx = p->ncot * p->ncit * (3 / p->ndy);
    // CTA pairs (cta_group::2) share the view operand: 256 output channels x one 128-channel input tile per pair.
    // variant bit 64 / TNB_WGRAD_PAIR=0 force the single-CTA kernel (tests, ablation).
    static const int pair_env = [] { const char* e = getenv("TNB_WGRAD_PAIR"); return e ? atoi(e) : 1; }();
    p->kind = (pair_env && !(variant & 64) && Cout % 256 == 0 && p->NT == 128 && p->ndy == 1 && p->mode0 == SRC_PRESPLIT &&
               p->mode1 == SRC_PRESPLIT) ? 1 : 0;
  }
  // split the pixel (K) range so that the grid is ~1 wave of 148 SMs (every CTA ends with its epilogue stores, so
  // fewer, longer CTAs are better), each CTA owning >= 8 K tiles; never more CTAs than SMs: a 149th would cost a second wave
  int splits = 148 / p->gx;
  if (splits > (p->ktiles + 7) / 8) splits = (p->ktiles + 7) / 8;
  if (splits < 1) splits = 1;
  p->ktiles_per_cta = (p->ktiles + splits - 1) / splits;
  p->splits = (p->ktiles + p->ktiles_per_cta - 1) / p->ktiles_per_cta;  // every split owns at least one K tile
  return 0;
}
size_t wgrad3x3_ws_floats(const ViewDesc& view, int Cout) {
  WgradPlan p;
  if (plan_wgrad(view, Cout, 0, &p)) return 0;
  return (size_t)p.splits * 9 * Cout * view.C;
}

static int launch_wgrad3x3_stacked(const ViewDesc& view, const void* dz_presplit, float* dw, int Cout, int CinReal,
                                   int nterms, const WgradPlan& p, cudaStream_t st, float* ws, int fmt, const float* dz_mul) {
  WgradSArgs a;
  a.ws = ws; a.slab = (long long)9 * Cout * view.C; a.fmt = fmt; a.dz_mul = dz_mul;
  a.view = reinterpret_cast<const uint8_t*>(view.s[0].ptr); a.dz = (const uint8_t*)dz_presplit; a.dw = dw;
  a.view1 = reinterpret_cast<const uint8_t*>(view.C0 < view.C ? view.s[1].ptr : view.s[0].ptr);
  a.C0 = view.C0; a.Cs0 = view.s[0].C; a.Cs1 = view.C0 < view.C ? view.s[1].C : view.s[0].C;
  a.up0 = view.s[0].mode == SRC_PRESPLIT_UP ? 1 : 0; a.Hs0 = view.s[0].Hs; a.Ws0 = view.s[0].Ws;
  a.N = view.N; a.H = view.H; a.W = view.W; a.C = view.C; a.Cout = Cout; a.CinReal = CinReal; a.nterms = nterms;
  a.P = p.P; a.ncit = p.ncit; a.ncot = p.ncot;
  a.tiles_h = p.tiles_h; a.tiles_w = p.tiles_w; a.ktiles = p.ktiles; a.ktiles_per_cta = p.ktiles_per_cta;
  const int TP = nterms > 1 ? 2 : 1;
  const size_t smem = kHdrBytes + kSStages * (size_t)(TP * kSRows * a.P * kSRP + TP * 8 * pad_px(kTileH * kTileW) * 16);
  TNB_REQUIRE(smem <= 232448, "wgrad3x3 (stacked): shared memory plan too large (%zu)", smem);
  TNB_CHECK_CUDA(cudaFuncSetAttribute(wgrad3x3_stacked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(PROF_WGRAD, st, view.N, view.H, view.W, view.C, Cout);
  if (int rc = launch_pdl(wgrad3x3_stacked_kernel, dim3(p.gx, p.splits), dim3(kThreads), smem, st, a)) return rc;
  if (ws != nullptr) return launch_wgrad_scatter(ws, dw, Cout, view.C, CinReal, 1, p.splits, st);
  return 0;
}

int launch_wgrad3x3(const ViewDesc& view, const void* dz_presplit, float* dw, int Cout, int CinReal, int nterms,
                    int variant, cudaStream_t st, float* ws, int fmt, const float* dz_mul) {
  TNB_REQUIRE(fmt == 0 || fmt == 1, "wgrad3x3: operand format %d (0 = fp16, 1 = bf16)", fmt);
  WgradPlan p;
  if (int rc = plan_wgrad(view, Cout, variant, &p)) return rc;
  if (p.kind == 2) return launch_wgrad3x3_stacked(view, dz_presplit, dw, Cout, CinReal, nterms, p, st, ws, fmt, dz_mul);
  WgradArgs a;
  a.ws = ws; a.slab = (long long)9 * Cout * view.C; a.fmt = fmt; a.dz_mul = dz_mul;
  a.view = view; a.dz = (const uint8_t*)dz_presplit; a.dw = dw; a.Cout = Cout; a.CinReal = CinReal;
  a.nterms = nterms; a.variant = variant; a.ci_tile_base = 0;
  a.NT = p.NT; a.tiles_h = p.tiles_h; a.tiles_w = p.tiles_w; a.ktiles = p.ktiles; a.ncot = p.ncot; a.ncit = p.ncit;
  a.ndy = p.ndy; a.ktiles_per_cta = p.ktiles_per_cta;
  const int gx = p.gx, splits = p.splits;
  const int TP = nterms > 1 ? 2 : 1;
  const int mode0 = p.mode0, mode1 = p.mode1;
  ProfScope prof(PROF_WGRAD, st, view.N, view.H, view.W, view.C, Cout);
  if (p.kind == 1) {
    const int padm = (variant >> 7) & 3;
    const size_t psmem = kHdrBytes + kPStages * (size_t)(TP * 16 * pad_sel(kTileH * kTileW, padm) * 16 +
                                                         TP * 8 * pad_sel(kTileH * kHaloW, padm) * 16);
    TNB_REQUIRE(psmem <= 232448, "wgrad3x3 (pair): shared memory plan too large (%zu)", psmem);
    TNB_CHECK_CUDA(cudaFuncSetAttribute(wgrad3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    // gx = ncot * ncit * 3 with ncot even
    if (int rc = launch_pdl(wgrad3x3_pair_kernel, dim3(gx, splits), dim3(kThreads), psmem, st, a)) return rc;
    if (ws != nullptr) return launch_wgrad_scatter(ws, dw, Cout, view.C, CinReal, 0, splits, st);
    return 0;
  }
  const size_t smem = kHdrBytes + kStages * (size_t)(TP * 16 * pad_px(kTileH * kTileW) * 16 +
                                                     TP * (a.NT / 8) * pad_px((kTileH + a.ndy - 1) * kHaloW) * 16);
  TNB_REQUIRE(smem <= 232448, "wgrad3x3: shared memory plan too large (%zu)", smem);
  // one gather mode per launch: a concat view whose two halves use different modes is split by the channel tiling
  // (pick_nt keeps every CTA inside one source), but the kernel is instantiated per mode.
  // a half-resolution pre-split source differs from a full-resolution one by an address shift only: same instantiation
  auto launch = [&](auto tag, const WgradArgs& args, int gx_count) -> int {
    constexpr int M = decltype(tag)::value;
    TNB_CHECK_CUDA(cudaFuncSetAttribute(wgrad3x3_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return launch_pdl(wgrad3x3_kernel<M>, dim3(gx_count, splits), dim3(kThreads), smem, st, args);
  };
  auto dispatch = [&](int mode, const WgradArgs& args, int gx_count) -> int {
    switch (mode) {
      case SRC_IDENTITY: return launch(std::integral_constant<int, SRC_IDENTITY>{}, args, gx_count);
      case SRC_AFFINE_RELU: return launch(std::integral_constant<int, SRC_AFFINE_RELU>{}, args, gx_count);
      case SRC_AFFINE_RELU_POOL: return launch(std::integral_constant<int, SRC_AFFINE_RELU_POOL>{}, args, gx_count);
      case SRC_AFFINE_RELU_UP: return launch(std::integral_constant<int, SRC_AFFINE_RELU_UP>{}, args, gx_count);
      case SRC_PRESPLIT: return launch(std::integral_constant<int, SRC_PRESPLIT>{}, args, gx_count);
      default: tnb::set_last_error("wgrad3x3: bad view mode %d", mode); return -2;
    }
  };
  if (mode0 == mode1) {
    if (int rc = dispatch(mode0, a, gx)) return rc;
  } else {
    // two launches, one per source: ci tiles [0, C0/NT) use mode0, the rest mode1 (ci_base shifts blockIdx.x)
    // (each launch under-fills the GPU: the K split was chosen for the combined grid. Only views whose two halves need
    // different on-the-fly gathers come here; the network's backward pass materialises both halves pre-split.)
    WgradArgs a0 = a, a1 = a;
    const int t0 = view.C0 / a.NT;
    a0.ncit = t0; a0.ci_tile_base = 0;
    a1.ncit = a.ncit - t0; a1.ci_tile_base = t0;
    if (int rc = dispatch(mode0, a0, a.ncot * a0.ncit * (3 / a.ndy))) return rc;
    if (int rc = dispatch(mode1, a1, a.ncot * a1.ncit * (3 / a.ndy))) return rc;
  }
  if (ws != nullptr) return launch_wgrad_scatter(ws, dw, Cout, view.C, CinReal, 0, splits, st);
  return 0;
}

}  // namespace tnb
