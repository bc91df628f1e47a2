// CTA-pair variant of the persistent tcgen05 implicit-GEMM 3x3 convolution (forward and dgrad) - EXPERIMENTAL, off by
// default (TNB_CONV_PAIR=1), not yet run on a GPU. conv.cu holds the validated single-CTA kernel; this file is the same
// kernel body with the pair protocol of wgrad3x3_pair_kernel (wgrad.cu) added, kept in its own translation unit so that
// the shipped kernel's code stays byte-identical until the pair version has been measured:
//
//   * clusters of two CTAs (the two SMs of a TPC) work on two M tiles of the same output-channel tile; each CTA
//     gathers its own halo tile and loads HALF of every weight stage (output-channel rows [rank * BN/2, +BN/2) of each
//     plane; tap images packed [rank][term][plane][BN/2 rows][8], launch_pack_weights layout 2);
//   * rank 0 issues every MMA as tcgen05.mma.cta_group::2 with M = 256 (TMEM lanes 0-127 of each CTA = that CTA's 128
//     pixels) and multicasts the commits to both CTAs' empty_A / empty_B / tmem_full barriers;
//   * rank 1's warp 0 relays its full_A / full_B phases to rank 0's barriers (expected counts + 1), rank 1's epilogue
//     warps arrive on rank 0's tmem_empty (count 8);
//   * an odd number of M tiles leaves rank 1 of the last pair a duplicate of the last tile with all stores off.
// Why: the weights are the larger part of the L2 -> SM traffic of these kernels (590 KB per 256-pixel tile on a
// 128 -> 128 layer against 166 KB of activations), an N = 128 MMA reads 64 clocks of operands for 64 of math, and the
// 64-wide layers are bound by per-stage handshakes (profiles/r1_final.md section 10): a pair halves all three per SM.
#include "igemm.cuh"
#include <cstdlib>
#include "prof.cuh"
#include <type_traits>

namespace tnb {

static constexpr int kThreads = 384;
static constexpr int kFillThreads = 192;
static constexpr int kEpiThreads = 128;
static constexpr int kMaxSmem = 232448;  // 227 KB opt-in limit per CTA on sm_100
static constexpr int kHdrBytes = 512;

struct ConvPairArgs {
  ViewDesc view;
  const uint16_t* wpack;
  float* out;        // [N,H,W,Cout]
  float* stat_part;  // [ntiles][2][Cout] per-tile (sum, sumsq) partials, or nullptr
  // dgrad fused with the BatchNorm-backward reduction of the layer that produced the view this gradient belongs to:
  // with bz != nullptr the partials are (sum g, sum g * xhat), g = out * [relu'(bsc * z + bsh)], xhat = (z - bmu) * bis
  const float *bz, *bsc, *bsh, *bmu, *bis;
  int Cout, BN, MT, SA, SB, G, nbuf, nterms, variant, tmem_cols;  // G: filter taps per weight stage (1 or 3)
  int tiles_h, tiles_w, ntiles, nwork;
  int ntiles_p, nwork_p;  // CTA-pair kernel: pairs of M tiles per n-tile, pair work items (see conv3x3_kernel<..., PAIR>)
  int merged;  // weights packed [plane][hi | lo][BN rows]: x_hi * [w_hi | w_lo] is ONE MMA of width 2 * BN (see conv3x3_merged)
  int tall;  // tile orientation: 0 = 16 rows x 8*MT columns (halo tile row-major), 1 = 8*MT rows x 16 columns
             // (halo tile column-major: the 8-pixel core-matrix groups then run down the image)
};

// 31-shuffle transpose-reduce: on return lane j holds the sum over the 32 lanes of v[j].
TNB_DEVINL float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// M0 / M1: gather modes of the (up to two) concatenated view sources, compile-time so that every instantiation carries
// only the gather paths it needs (the producers are register-limited; a run-time switch over all modes costs spills)
//
// PAIR (experimental, TNB_CONV_PAIR=1; launched as clusters of two CTAs): the two CTAs of a pair work on two M tiles of
// the same output-channel tile and share the weight operand through tcgen05 cta_group::2 - each CTA gathers its own
// halo tile and loads HALF of every weight stage (rows [rank * BN/2, +BN/2) of each plane), rank 0 issues every MMA
// with M = 256 and multicasts the commits; rank 1's warp 0 relays its full_A / full_B phases to rank 0's barriers and
// rank 1's epilogue warps arrive on rank 0's tmem_empty. Same protocol as wgrad3x3_pair_kernel (wgrad.cu).
template <int FMT, int M0, int M1>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_pair_kernel(const __grid_constant__ ConvPairArgs a) {
  constexpr bool PAIR = true, BWD = false;  // the body is conv3x3_kernel's (conv.cu) with the pair protocol switched on
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  // persistent loop over work items: (M tile, n-tile), or for PAIR (pair of M tiles, n-tile) per pair of CTAs
  const int wstart = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int wstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int wcount = PAIR ? a.nwork_p : a.nwork;
  // work item -> (n-tile, M tile of THIS CTA, dummy): an odd tile count leaves rank 1 of the last pair without a tile;
  // it then repeats the last one (the pair must stay in lock step) with every global store switched off
  auto decode_work = [&](int work, int& nt, int& tile, bool& dummy) {
    if (PAIR) {
      nt = work / a.ntiles_p;
      tile = 2 * (work - nt * a.ntiles_p) + rank;
      dummy = tile >= a.ntiles;
      if (dummy) tile = a.ntiles - 1;
    } else {
      nt = work / a.ntiles;
      tile = work - nt * a.ntiles;
      dummy = false;
    }
  };

  const ViewDesc& V = a.view;
  const int MT = a.MT, BN = a.BN;
  const int PITCH = 8 * MT + 2;
  const int HALO_PX = 18 * PITCH;
  const int PLANE = pad_px(HALO_PX) * 16;  // bytes
  const int TP = a.nterms > 1 ? 2 : 1;     // operand term planes stored (hi[,lo])
  const int A_STAGE = TP * 4 * PLANE;
  const bool MG = a.merged != 0;
  const int BROWS = PAIR ? BN / 2 : BN;    // weight rows (output channels) held by THIS CTA
  const int B_TAP = (MG ? 2 : TP) * 64 * BROWS;  // bytes of one tap: [term][4 planes][rows][16B], merged: [4 planes][term][rows][16B]
  const int B_STAGE = a.G * B_TAP;         // a stage holds G consecutive taps of one 32-channel chunk
  const int nchunks = V.C / 32;
  const int ACCW = (MG && a.nterms > 1) ? 2 * BN : BN;  // TMEM columns per M tile (merged: [x*w_hi + x_lo*w_hi | x_hi*w_lo])
  const int BUFCOLS = MT * ACCW;

  uint64_t* full_A = reinterpret_cast<uint64_t*>(smem);       // [4]
  uint64_t* empty_A = full_A + 4;                             // [4]
  uint64_t* full_B = full_A + 8;                              // [8]
  uint64_t* empty_B = full_A + 16;                            // [8]
  uint64_t* tmem_full = full_A + 24;                          // [2]
  uint64_t* tmem_empty = full_A + 26;                         // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full_A + 28);
  int2* table = reinterpret_cast<int2*>(smem + kHdrBytes);
  const int table_bytes = (HALO_PX * 8 + 127) & ~127;
  float* sstat = reinterpret_cast<float*>(smem + kHdrBytes + table_bytes);  // [4 warps][2][BN]
  float4* btab = reinterpret_cast<float4*>(sstat + 4 * 2 * BN);  // [BN] (scale, shift, mean, invstd), fused BN-bwd only
  uint8_t* a_base = smem + kHdrBytes + table_bytes + 4 * 2 * BN * 4 + (BWD ? BN * 16 : 0);
  uint8_t* b_base = a_base + a.SA * A_STAGE;

  // ---- one-time setup ----
  if (warp == 0) {
    if (elect_one()) {
      // PAIR, rank 0: one extra arrival per phase from rank 1's relay (full_A, full_B) / 4 more from its epilogue warps
      const int extra = (PAIR && rank == 0) ? 1 : 0;
      for (int i = 0; i < a.SA; ++i) { mbar_init(&full_A[i], kFillThreads + extra); mbar_init(&empty_A[i], 1); }
      for (int i = 0; i < a.SB; ++i) { mbar_init(&full_B[i], 1 + extra); mbar_init(&empty_B[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4 + 4 * extra); }
      fence_mbar_init();
    }
    __syncwarp();
    if (PAIR) tmem_alloc_pair(tmem_ptr, a.tmem_cols); else tmem_alloc(tmem_ptr, a.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // dgrad: dz is multiplied by a power of two on the way in (so that its fp16 hi/lo split keeps ~22 bits) and
  // the accumulator by the inverse on the way out; forward views use BN scale/shift instead (in_mul = 1).
  const float in_mul = (V.s[0].mode == SRC_IDENTITY) ? pow2_scale_for(V.s[0].scale) : 1.f;
  const float out_mul = 1.f / in_mul;

  if (PAIR && warp == 0 && rank != 0) {
    // =========================== rank 1 of a pair: relay ===========================
    // forward every completed fill phase of this CTA (halo tile, weight half) to the MMA issuer's barriers, in the
    // order in which the issuer waits for them
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    for (int work = wstart; work < wcount; work += wstep) {
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full_A[sa], pha);
        fence_proxy_async_smem();
        if (lane == 0) mbar_arrive_remote(&full_A[sa], 0);
        __syncwarp();
        for (int t0 = 0; t0 < 9; t0 += a.G) {
          mbar_wait(&full_B[sb], phb);
          if (lane == 0) mbar_arrive_remote(&full_B[sb], 0);
          __syncwarp();
          if (++sb == a.SB) { sb = 0; phb ^= 1; }
        }
        if (++sa == a.SA) { sa = 0; pha ^= 1; }
      }
    }
  } else if (warp == 0) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the (warp-uniform) loops so that the descriptor arithmetic stays in uniform
    // registers; only the tcgen05 instructions themselves are predicated on one elected lane.
    {
      const bool lead = elect_one();
      const uint32_t idesc = make_idesc(PAIR ? 256 : 128, BN, FMT, 0, 0);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
        if (PAIR) umma_f16_pair(d, da, db, id, acc); else umma_f16(d, da, db, id, acc);
      };
      auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
      auto wait = [&](uint64_t* bar, uint32_t parity) { if (PAIR) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity); };
      // A: K-major planar halo tile. LBO = plane stride (next 8 channels), SBO = halo row pitch (next 8 output
      // pixels = next image row of the 16x8 tile). B: K-major packed weights. variant bits swap them (probe).
      uint32_t a_lbo = PLANE, a_sbo = PITCH * 16;
      uint32_t b_lbo = (MG ? 2 : 1) * BROWS * 16, b_sbo = 128;
      const uint32_t idesc2 = make_idesc(128, 2 * BN, FMT, 0, 0);  // merged: B = [w_hi | w_lo], 2 * BN rows per plane
      if (a.variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
      if (a.variant & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      // descriptors are base + (byte offset >> 4): the start-address field never carries into the next field
      const uint64_t a_desc0 = make_smem_desc(smem_u32(a_base), a_lbo, a_sbo);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(b_base), b_lbo, b_sbo);
      const uint32_t a_stage16 = A_STAGE >> 4, b_stage16 = B_STAGE >> 4, b_tap16 = B_TAP >> 4;
      const uint32_t a_k16 = (2 * PLANE) >> 4, a_lo16 = (4 * PLANE) >> 4;
      const uint32_t b_k16 = (2 * b_lbo) >> 4, b_lo16 = MG ? (uint32_t)(BROWS * 16) >> 4 : (uint32_t)(4 * BROWS * 16) >> 4;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int k = 0;
      for (int work = wstart; work < wcount; work += wstep, ++k) {
        const int buf = k % a.nbuf;
        const uint32_t use = (uint32_t)(k / a.nbuf);
        wait(&tmem_empty[buf], (use & 1) ^ 1);  // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_buf = tmem_base + buf * BUFCOLS;
        for (int c = 0; c < nchunks; ++c) {
          wait(&full_A[sa], pha);
          tc_fence_after();
          const uint64_t a_st = a_desc0 + (uint64_t)(sa * a_stage16);
          for (int t0 = 0; t0 < 9; t0 += a.G) {
            wait(&full_B[sb], phb);
            tc_fence_after();
            const uint64_t b_st = b_desc0 + (uint64_t)(sb * b_stage16);
            // Issue order: ALL MMAs of this tap group that accumulate into one TMEM tile are issued back to back
            // (G taps x 2 K-steps x nterms), then the next tile. Measured: the tensor pipe retires a chain into one
            // accumulator at full rate but pays a drain when the accumulator changes, so short chains starve it.
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t d_tmem = d_buf + mt * ACCW;
              for (int tg = 0; tg < a.G; ++tg) {
                const int t = t0 + tg;
                const int dy = t / 3, dx = t - dy * 3;
                const uint64_t a_tap = a_st + (uint64_t)((a.tall ? dx * PITCH + dy : dy * PITCH + dx) + 8 * mt);
                const uint64_t b_tap = b_st + (uint64_t)(tg * b_tap16);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  const uint64_t a_hi = a_tap + (uint64_t)(kk * a_k16);
                  const uint64_t b_hi = b_tap + (uint64_t)(kk * b_k16);
                  const uint32_t acc = (c | t | kk) != 0;
                  if (lead && !(a.variant & 4)) {
                    if (MG && a.nterms > 1) {
                      // x_hi * [w_hi | w_lo] in one MMA of width 2 * BN (the A tile is read once for both products),
                      // then x_lo * w_hi into the first half; the epilogue adds the two halves
                      mma(d_tmem, a_hi, b_hi, idesc2, acc);
                      mma(d_tmem, a_hi + a_lo16, b_hi, idesc, 1);
                    } else {
                      mma(d_tmem, a_hi, b_hi, idesc, acc);
                      if (a.nterms > 1) {
                        mma(d_tmem, a_hi + a_lo16, b_hi, idesc, 1);
                        mma(d_tmem, a_hi, b_hi + b_lo16, idesc, 1);
                      }
                    }
                  }
                }
              }
            }
            if (lead) commit(&empty_B[sb]);
            if (++sb == a.SB) { sb = 0; phb ^= 1; }
          }
          if (lead) commit(&empty_A[sa]);
          if (++sa == a.SA) { sa = 0; pha ^= 1; }
        }
        if (lead) commit(&tmem_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== weight loader (bulk TMA) ===========================
    if (elect_one()) {
      int sb = 0;
      uint32_t phb = 0;
      for (int work = wstart; work < wcount; work += wstep) {
        int nt, tile_unused; bool dummy_unused;
        decode_work(work, nt, tile_unused, dummy_unused);
        // PAIR: a tap image is [rank][term][4 planes][BN/2 rows][16 B]; this CTA fetches its own half
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)nt * nchunks * 9 * (size_t)(128 * BN) +
                              (PAIR ? (size_t)rank * (size_t)(64 * BN) : 0);
        for (int i = 0; i < nchunks * 9; i += a.G) {
          mbar_wait(&empty_B[sb], phb ^ 1);
          mbar_arrive_expect_tx(&full_B[sb], (uint32_t)B_STAGE);
          for (int g = 0; g < a.G; ++g)  // each tap image is [hi | lo]; with one term only the hi half is fetched
            bulk_g2s(b_base + sb * B_STAGE + g * B_TAP, wsrc + (size_t)(i + g) * (128 * BN), (uint32_t)B_TAP, &full_B[sb]);
          if (++sb == a.SB) { sb = 0; phb ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // =========================== epilogue (warps 2..5) ===========================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = 32 * q + lane;
    const int r = row >> 3, cc = row & 7;
    const int et = tid - 64;  // 0..127
    int k = 0;
    int tab_nt = -1;
    for (int work = wstart; work < wcount; work += wstep, ++k) {
      int nt, tile; bool dummy;
      decode_work(work, nt, tile, dummy);
      if (BWD && nt != tab_nt) {  // per-channel BatchNorm constants of this output-channel tile
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int j = et; j < BN; j += kEpiThreads) {
          const int c = nt * BN + j;
          btab[j] = make_float4(a.bsc[c], a.bsh[c], a.bmu[c], a.bis[c]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        tab_nt = nt;
      }
      const int tile_id = tile;
      const int tw = tile % a.tiles_w; tile /= a.tiles_w;
      const int th = tile % a.tiles_h;
      const int n = tile / a.tiles_h;
      const int h0 = a.tall ? th * 8 * MT : th * 16, w0 = a.tall ? tw * 16 : tw * 8 * MT, n0 = nt * BN;
      const int buf = k % a.nbuf;
      const uint32_t use = (uint32_t)(k / a.nbuf);
      if (BWD) {  // pull this tile's slice of the producer's z towards L2 while the MMAs of the tile are still running
        for (int mt = 0; mt < MT; ++mt) {
          const int h = a.tall ? h0 + 8 * mt + cc : h0 + r, w = a.tall ? w0 + r : w0 + 8 * mt + cc;
          if (h < V.H && w < V.W) {
            const float* zp = a.bz + ((size_t)(n * V.H + h) * V.W + w) * a.Cout + n0;
            for (int col0 = 0; col0 < BN; col0 += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(zp + col0));
          }
        }
      }
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      for (int col0 = 0; col0 < BN; col0 += 32) {
        float csum = 0.f, csq = 0.f;
        for (int mt = 0; mt < MT; ++mt) {
          const int h = a.tall ? h0 + 8 * mt + cc : h0 + r, w = a.tall ? w0 + r : w0 + 8 * mt + cc;
          const bool valid = (h < V.H) && (w < V.W) && !(a.variant & 16) && !dummy;
          float zz[BWD ? 32 : 1];
          if (BWD) {  // the producer's z for these 32 channels: issued before the TMEM load so the latencies overlap
            const float4* zp =
                reinterpret_cast<const float4*>(a.bz + ((size_t)(n * V.H + h) * V.W + w) * a.Cout + n0 + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 t = valid ? __ldg(zp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
              zz[4 * i] = t.x; zz[4 * i + 1] = t.y; zz[4 * i + 2] = t.z; zz[4 * i + 3] = t.w;
            }
          }
          uint32_t rg[32];
          tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * BUFCOLS + mt * ACCW + col0), rg);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rg[i]);
          if (ACCW != BN) {  // merged weights: the x_hi * w_lo products sit BN columns further
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * BUFCOLS + mt * ACCW + BN + col0), rg);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(rg[i]);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= out_mul;
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(a.out + ((size_t)(n * V.H + h) * V.W + w) * a.Cout + n0 + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (a.stat_part != nullptr) {
            if (!valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            float s[32];
            if (BWD) {
              // g = dL/da masked by the producer's ReLU (same expression as the forward gather / bn_bwd_kernel);
              // column sums of g and g * (z - mean)
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float4 t = btab[col0 + i];
                const float g = fmaf(zz[i], t.x, t.y) > 0.f ? v[i] : 0.f;
                v[i] = g;
                s[i] = g * (zz[i] - t.z);
              }
              csum += warp_transpose_sum(v, lane);
              csq += warp_transpose_sum(s, lane);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) s[i] = v[i];
              csum += warp_transpose_sum(s, lane);
#pragma unroll
              for (int i = 0; i < 32; ++i) s[i] = v[i] * v[i];
              csq += warp_transpose_sum(s, lane);
            }
          }
        }
        if (a.stat_part != nullptr) {
          if (BWD) csq *= btab[col0 + lane].w;  // sum g * (z - mean) -> sum g * xhat
          sstat[(q * 2 + 0) * BN + col0 + lane] = csum;
          sstat[(q * 2 + 1) * BN + col0 + lane] = csq;
        }
      }
      // all TMEM reads of this buffer are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // PAIR: the MMA issuer (rank 0) waits for the epilogue warps of both CTAs
        if (PAIR && rank != 0) mbar_arrive_remote(&tmem_empty[buf], 0); else mbar_arrive(&tmem_empty[buf]);
      }
      if (a.stat_part != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int j = et; j < 2 * BN; j += kEpiThreads) {
          const int which = j / BN, col = j - which * BN;
          const float s = sstat[(0 * 2 + which) * BN + col] + sstat[(1 * 2 + which) * BN + col] +
                          sstat[(2 * 2 + which) * BN + col] + sstat[(3 * 2 + which) * BN + col];
          if (!dummy) a.stat_part[((size_t)tile_id * 2 + which) * a.Cout + n0 + col] = s;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // sstat is reused by the next tile
      }
    }
  } else {
    // =========================== A producers: gather + BN/ReLU/pool/upsample + split ===========
    const int ftid = tid - 192;
    const int j = ftid & 3;       // plane (8 channels) this thread fills: fixed, kFillThreads % 4 == 0
    const int pbase = ftid >> 2;  // first halo pixel; stride kFillThreads/4 pixels
    int sa = 0;
    uint32_t pha = 0;
    for (int work = wstart; work < wcount; work += wstep) {
      int nt, tile; bool dummy;
      decode_work(work, nt, tile, dummy);
      const int tw = tile % a.tiles_w; tile /= a.tiles_w;
      const int th = tile % a.tiles_h;
      const int n = tile / a.tiles_h;
      const int h0 = a.tall ? th * 8 * MT : th * 16, w0 = a.tall ? tw * 16 : tw * 8 * MT;
      asm volatile("bar.sync 2, 192;" ::: "memory");  // previous tile's table is no longer read
      for (int p = ftid; p < HALO_PX; p += kFillThreads) {
        const int major = p / PITCH, minor = p - major * PITCH;  // halo pixel index = major * PITCH + minor
        const int hr = a.tall ? minor : major, hc = a.tall ? major : minor;
        const int h = h0 - 1 + hr, w = w0 - 1 + hc;
        int2 e = make_int2(-1, -1);
        if (h >= 0 && h < V.H && w >= 0 && w < V.W) {
          e.x = view_pix_off(V.s[0], n, h, w);
          if (V.C0 < V.C) e.y = view_pix_off(V.s[1], n, h, w);
        }
        table[p] = e;
      }
      asm volatile("bar.sync 2, 192;" ::: "memory");
      for (int c = 0; c < nchunks; ++c) {
        const int cch = c * 32 + j * 8;
        const bool second = cch >= V.C0;
        const SrcDesc& S = second ? V.s[1] : V.s[0];
        const int cc = second ? cch - V.C0 : cch;
        float sc[8], sh[8];
        if (S.mode != SRC_IDENTITY && S.mode != SRC_PRESPLIT) { ld8(S.scale + cc, sc); ld8(S.shift + cc, sh); }
        mbar_wait(&empty_A[sa], pha ^ 1);
        uint8_t* stage = a_base + sa * A_STAGE + j * PLANE;
        auto run = [&](auto mode_tag, auto batch_tag) {
          constexpr int MODE = decltype(mode_tag)::value;
          constexpr int U = decltype(batch_tag)::value;
          if (a.variant & 8) return;  // ablation: barrier traffic only
          for (int p0 = pbase; p0 < HALO_PX; p0 += (kFillThreads / 4) * U) {
            Raw8 raw[U][RawCount<MODE>::value];
            int off[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int p = p0 + (kFillThreads / 4) * u;
              off[u] = -1;
              if (p < HALO_PX) {
                const int2 e = table[p];
                off[u] = second ? e.y : e.x;
                if (off[u] >= 0) view_issue<MODE>(S, off[u], cc, raw[u]);
              }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int p = p0 + (kFillThreads / 4) * u;
              if (p < HALO_PX) {
                uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
                if (off[u] >= 0) {
                  if (MODE == SRC_PRESPLIT) {  // already (hi, lo): pure copy, no arithmetic
                    hi = *reinterpret_cast<const uint4*>(&raw[u][0].a);
                    lo = *reinterpret_cast<const uint4*>(&raw[u][0].b);
                  } else {
                    float v[8];
                    view_finish<MODE>(raw[u], sc, sh, in_mul, v);
                    split8<FMT>(v, hi, lo);
                  }
                }
                uint8_t* dst = stage + p * 16;
                *reinterpret_cast<uint4*>(dst) = hi;
                if (a.nterms > 1) *reinterpret_cast<uint4*>(dst + 4 * PLANE) = lo;
              }
            }
          }
        };
        if (M0 == SRC_PRESPLIT && M1 == SRC_PRESPLIT) {
          // operands already (hi, lo) pairs in HBM (dgrad: dz written pre-split by the BatchNorm backward kernel):
          // 16-byte cp.async copies straight into the planar tile; the stage's mbarrier gets a cp.async-completion
          // arrival, so the thread never waits for its own loads and SA stages of HBM latency stay in flight
          if (!(a.variant & 8)) {
            const uint8_t* sbase = reinterpret_cast<const uint8_t*>(S.ptr) + (size_t)cc * 2;  // [pixel][2][C] 16-bit
            const size_t sstride = (size_t)S.C * 4;
            const int lo_off = S.C * 2;
            const bool ca = (a.variant & 256) != 0;
            for (int p = pbase; p < HALO_PX; p += kFillThreads / 4) {
              const int2 e = table[p];
              const int off = second ? e.y : e.x;
              const uint8_t* q = off >= 0 ? sbase + (size_t)off * sstride : sbase;
              cp_async16(stage + p * 16, q, off >= 0 ? 16u : 0u, ca);
              if (a.nterms > 1) cp_async16(stage + p * 16 + 4 * PLANE, q + lo_off, off >= 0 ? 16u : 0u, ca);
            }
          }
          cp_async_mbar_arrive_noinc(&full_A[sa]);
          if (++sa == a.SA) { sa = 0; pha ^= 1; }
          continue;
        }
        constexpr int U0 = (M0 == SRC_AFFINE_RELU_POOL) ? 2 : 4, U1 = (M1 == SRC_AFFINE_RELU_POOL) ? 2 : 4;
        if (second) run(std::integral_constant<int, M1>{}, std::integral_constant<int, U1>{});
        else        run(std::integral_constant<int, M0>{}, std::integral_constant<int, U0>{});
        fence_proxy_async_smem();
        mbar_arrive(&full_A[sa]);
        if (++sa == a.SA) { sa = 0; pha ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's tensor core reads this CTA's weight half until the last commit has landed
  if (warp == 0) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, a.tmem_cols); else tmem_dealloc(tmem_base, a.tmem_cols);
  }
}


int launch_conv3x3_pair(const ViewDesc& view, const uint16_t* wpack, float* out, float* stat_part, int Cout, int nterms,
                        int fmt, int variant, const ConvPlan& p, cudaStream_t st) {
  const int m0 = view.s[0].mode, m1 = (view.C0 < view.C) ? view.s[1].mode : view.s[0].mode;
  ConvPairArgs a;
  a.view = view; a.wpack = wpack; a.out = out; a.stat_part = stat_part;
  a.bz = a.bsc = a.bsh = a.bmu = a.bis = nullptr;
  a.Cout = Cout; a.BN = p.BN; a.MT = p.MT; a.SA = p.SA; a.SB = p.SB; a.G = p.G; a.nbuf = p.nbuf; a.nterms = nterms;
  a.variant = variant | (cp_async_ca_env() ? 256 : 0); a.tmem_cols = p.tmem_cols; a.tiles_h = p.tiles_h; a.tiles_w = p.tiles_w;
  a.tall = p.tall; a.merged = 0;
  a.ntiles = view.N * p.tiles_h * p.tiles_w;
  a.nwork = a.ntiles * (Cout / p.BN);
  a.ntiles_p = (a.ntiles + 1) / 2;
  a.nwork_p = a.ntiles_p * (Cout / p.BN);
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int pairs = a.nwork_p < sms / 2 ? a.nwork_p : sms / 2;
  ProfScope prof(view.s[0].mode == SRC_PRESPLIT ? PROF_CONV_DGRAD : PROF_CONV_FWD, st, view.N, view.H, view.W, view.C, Cout);
  auto go = [&](auto kern) -> int {
    TNB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    TNB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return 0;
  };
  int rc2 = -2;
#define TNB_CONVP_CASE(F, A, B) if (fmt == F && m0 == A && m1 == B) rc2 = go(conv3x3_pair_kernel<F, A, B>); else
  TNB_CONVP_CASE(0, SRC_IDENTITY, SRC_IDENTITY)
  TNB_CONVP_CASE(0, SRC_AFFINE_RELU, SRC_AFFINE_RELU)
  TNB_CONVP_CASE(0, SRC_AFFINE_RELU_POOL, SRC_AFFINE_RELU_POOL)
  TNB_CONVP_CASE(0, SRC_AFFINE_RELU_UP, SRC_AFFINE_RELU)
  TNB_CONVP_CASE(1, SRC_PRESPLIT, SRC_PRESPLIT)
  { tnb::set_last_error("conv3x3 (pair): unsupported (fmt %d, source modes %d/%d) combination", fmt, m0, m1); return -2; }
#undef TNB_CONVP_CASE
  if (rc2) return rc2;
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tnb
