// tcgen05 implicit-GEMM 3x3 convolution kernels for sm_100a.
//
// Replaces, for the TrackNet hot path, what the reference delegates to cuDNN:
//   * nn.Conv2d(3x3, padding='same', bias=False) forward            (reference model.py:8, 13)
//   * its autograd dgrad and wgrad                                   (reference train.py:95 loss.backward())
// and fuses into the operand load what the reference runs as separate ATen kernels:
//   BatchNorm2d apply + ReLU (model.py:9-10,14-15), MaxPool2d (model.py:59,61,63),
//   Upsample x2 + torch.cat (model.py:65,67,69).
//
// Design:
//   * All tensors in HBM are fp32 NHWC. Producer warps gather the layer's logical input ("view"),
//     apply BN/ReLU/pool/upsample/concat in registers, split every fp32 value into a 16-bit (hi, lo)
//     pair and store both into a planar shared-memory tile  [plane = 8 channels][pixel][16 bytes].
//   * The same planar tile is a valid SWIZZLE_NONE K-major operand (rows = pixels, K = channels: the
//     forward / dgrad GEMM) and MN-major operand (rows = channels, K = pixels: the wgrad GEMM), and
//     any 3x3 tap is just a different 16-byte-aligned start address inside ONE halo tile - no im2col,
//     no per-tap reload.
//   * One elected thread issues tcgen05.mma (kind::f16, fp32 accumulate in TMEM). With nterms = 3 the
//     product is hi*hi + lo*hi + hi*lo, i.e. fp32-faithful (error ~2^-22) at 3x the MMA work; nterms = 1
//     is the TF32-class single pass.
//   * Weight tiles are pre-packed into the exact smem image and staged with 1-D bulk TMA
//     (cp.async.bulk + mbarrier complete_tx).
#include "igemm.cuh"
#include "prof.cuh"
#include <stdio.h>
#include <type_traits>

namespace tnb {

static constexpr int kThreads = 320;      // wgrad: warp0 MMA issue + TMEM alloc, warp1 idle, warps 2..9 fill
static constexpr int kFillThreads = 256;  // warps 2..5 double as the epilogue warps
static constexpr int kMaxSmem = 232448;   // 227 KB opt-in limit per CTA on sm_100
static constexpr int kHdrBytes = 256;

// =============================================================================================
// weight packing: OIHW fp32 -> per (n-tile, k-chunk, tap) smem images [term][plane(4)][BN][8] (uint16)
// mode 0 (forward):  B[n = co][k = ci] = W[co][ci][dy][dx]
// mode 1 (dgrad):    B[n = ci][k = co] = W[co][ci][2-dy][2-dx]      (180-degree rotated, transposed)
// =============================================================================================
template <int FMT>
__global__ void pack_weights_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Co, int Ci,
                                    int Nside, int Kpad, int BN, int mode) {
  const int nchunks = Kpad / 32;
  const long long total = (long long)(Nside / BN) * nchunks * 9 * 4 * BN;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int nl = (int)(r % BN); r /= BN;
    const int plane = (int)(r % 4); r /= 4;
    const int tap = (int)(r % 9); r /= 9;
    const int chunk = (int)(r % nchunks); r /= nchunks;
    const int ntile = (int)r;
    const int n = ntile * BN + nl;
    const int dy = tap / 3, dx = tap % 3;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = chunk * 32 + plane * 8 + e;
      float x = 0.f;
      if (mode == 0) {
        if (n < Co && k < Ci) x = w[(((size_t)n * Ci + k) * 3 + dy) * 3 + dx];
      } else {
        if (n < Ci && k < Co) x = w[(((size_t)k * Ci + n) * 3 + (2 - dy)) * 3 + (2 - dx)];
      }
      v[e] = x;
    }
    uint4 hi, lo;
    split8<FMT>(v, hi, lo);
    // stage base (uint16 elements): ((ntile*nchunks + chunk)*9 + tap) * (2*4*BN*8)
    const size_t stage = (((size_t)ntile * nchunks + chunk) * 9 + tap) * (size_t)(64 * BN);
    uint4* dst_hi = reinterpret_cast<uint4*>(out + stage + ((size_t)plane * BN + nl) * 8);
    uint4* dst_lo = reinterpret_cast<uint4*>(out + stage + (size_t)32 * BN + ((size_t)plane * BN + nl) * 8);
    *dst_hi = hi;
    *dst_lo = lo;
  }
}

size_t conv3x3_wpack_elems(int Kside, int Nside) {
  const int Kpad = (Kside + 31) / 32 * 32;
  return (size_t)Nside * Kpad * 9 * 2;
}

int launch_pack_weights(const float* w, uint16_t* out, int Co, int Ci, int mode, int fmt, int BN,
                        cudaStream_t st) {
  const int Nside = mode == 0 ? Co : Ci;
  const int Kside = mode == 0 ? Ci : Co;
  const int Kpad = (Kside + 31) / 32 * 32;
  TNB_REQUIRE(BN > 0 && Nside % BN == 0, "pack_weights: N side %d not divisible by BN %d", Nside, BN);
  const long long total = (long long)(Nside / BN) * (Kpad / 32) * 9 * 4 * BN;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads);
  if (fmt == 0)
    pack_weights_kernel<0><<<blocks, threads, 0, st>>>(w, out, Co, Ci, Nside, Kpad, BN, mode);
  else
    pack_weights_kernel<1><<<blocks, threads, 0, st>>>(w, out, Co, Ci, Nside, Kpad, BN, mode);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// conv3x3 wgrad:  dW[co][ci][dy][dx] += sum_{n,h,w} dz[n,h,w,co] * view[n,h+dy-1,w+dx-1,ci]
// GEMM per tap: M = co (128 rows), N = ci tile (32 or 48), K = pixels. Both operands are MN-major views
// of planar tiles; 9 taps x NT columns of fp32 accumulators live in TMEM for the whole CTA lifetime.
// =============================================================================================
struct WgradArgs {
  ViewDesc view;
  const float* dz;       // [N,H,W,Cout]
  const float* dz_amax;  // optional: max|dz| (device scalar) -> power-of-two pre-scaling for the fp16 split
  float* dw;             // [Cout][CinReal][3][3], accumulated with atomics (must be zeroed by the caller)
  int Cout, CinReal, NT, nterms, variant;
  int tiles_h, tiles_w, ktiles, ktiles_per_cta, ncot;
};

static constexpr int kWgTileH = 8, kWgTileW = 16;   // pixels per K tile = 128
static constexpr int kWgHaloW = kWgTileW + 2;       // 18
static constexpr int kWgHaloPx = (kWgTileH + 2) * kWgHaloW;  // 180

template <int FMT>
__global__ void __launch_bounds__(kThreads, 1) wgrad3x3_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const ViewDesc& V = a.view;
  const int NT = a.NT, NPL = NT / 8;
  const int TP = a.nterms > 1 ? 2 : 1;
  const int DZPL = pad_px(128) * 16;        // 2080
  const int VPL = pad_px(kWgHaloPx) * 16;   // 2976
  const int DZ_BYTES = TP * 16 * DZPL;
  const int STAGE = DZ_BYTES + TP * NPL * VPL;
  constexpr int S = 2;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + S;
  uint64_t* tmem_full = full + 2 * S;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full + 2 * S + 1);
  uint8_t* st_base = smem + kHdrBytes;

  const int co0 = (blockIdx.x % a.ncot) * 128;
  const int ci0 = (blockIdx.x / a.ncot) * NT;
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
  // rows of dz planes that are never filled (cvalid == 64) must still be finite-free garbage-tolerant:
  // zero them once so the unused accumulator rows stay harmless.
  if (npld < 16) {
    for (int i = tid; i < S * TP * (16 - npld) * (DZPL / 16); i += kThreads) {
      int r = i;
      const int px = r % (DZPL / 16); r /= (DZPL / 16);
      const int pl = npld + r % (16 - npld); r /= (16 - npld);
      const int term = r % TP; r /= TP;
      const int s = r;
      *reinterpret_cast<uint4*>(st_base + s * STAGE + (term * 16 + pl) * DZPL + px * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const float dz_mul = pow2_scale_for(a.dz_amax);
  const float out_mul = 1.f / dz_mul;

  if (warp == 0) {
    // whole warp runs the uniform loops (descriptor math in uniform registers); one lane issues
    {
      const bool lead = elect_one();
      const uint32_t idesc = make_idesc(128, NT, FMT, 1, 1);
      // MN-major planar tiles: SBO = plane stride (next 8 channels), LBO = 128 B (next 8 pixels).
      uint32_t a_lbo = 128, a_sbo = DZPL, b_lbo = 128, b_sbo = VPL;
      if (a.variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
      if (a.variant & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      const uint64_t a_desc0 = make_smem_desc(smem_u32(st_base), a_lbo, a_sbo);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(st_base) + DZ_BYTES, b_lbo, b_sbo);
      const uint32_t stage16 = STAGE >> 4, a_lo16 = (16 * DZPL) >> 4, b_lo16 = (NPL * VPL) >> 4;
      int it = 0;
      for (int kt = kt0; kt < kt1; ++kt, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t a_st = a_desc0 + (uint64_t)(s * stage16);
        const uint64_t b_st = b_desc0 + (uint64_t)(s * stage16);
        for (int r = 0; r < kWgTileH; ++r) {
          const uint64_t a_hi = a_st + (uint64_t)(r * 16);
          const uint32_t acc = (it | r) != 0;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int dy = t / 3, dx = t % 3;
            const uint64_t b_hi = b_st + (uint64_t)((r + dy) * kWgHaloW + dx);
            const uint32_t d_tmem = tmem_base + t * NT;
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
      }
      if (lead) umma_commit(tmem_full);
    }
    __syncwarp();
  } else if (warp >= 2) {
    const int ftid = tid - 64;
    const int tiles_per_img = a.tiles_h * a.tiles_w;
    // dz items: thread owns plane dpl, pixels dpx0 + k*DG;  view items: plane vpl, halo pixels vpx0 + k*VG
    const int dpl = ftid % npld, dpx0 = ftid / npld, DG = kFillThreads / npld;
    const int VG = kFillThreads / NPL;
    const bool vactive = ftid < VG * NPL;
    const int vpl = ftid % NPL, vpx0 = ftid / NPL;
    const int vch = ci0 + vpl * 8;
    const bool vsecond = vch >= V.C0;
    const SrcDesc& VS = vsecond ? V.s[1] : V.s[0];
    const int vcc = vsecond ? vch - V.C0 : vch;
    float sc[8], sh[8];
    if (VS.mode != SRC_IDENTITY && vactive) { ld8(VS.scale + vcc, sc); ld8(VS.shift + vcc, sh); }
    int it = 0;
    for (int kt = kt0; kt < kt1; ++kt, ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      const int n = kt / tiles_per_img;
      const int trem = kt - n * tiles_per_img;
      const int th = trem / a.tiles_w, tw = trem - th * a.tiles_w;
      const int h0 = th * kWgTileH, w0 = tw * kWgTileW;
      uint8_t* stage = st_base + s * STAGE;
      // ---- dz tile: 128 pixels x npld planes, 4 pixels per batch ----
      {
        uint8_t* dstp = stage + dpl * DZPL;
        for (int px0 = dpx0; px0 < 128; px0 += DG * 4) {
          Raw8 raw[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int px = px0 + u * DG;
            const int h = h0 + (px >> 4), w = w0 + (px & 15);
            ok[u] = px < 128 && h < V.H && w < V.W;
            if (ok[u]) raw[u] = ld_raw8(a.dz + ((size_t)(n * V.H + h) * V.W + w) * a.Cout + co0 + dpl * 8);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int px = px0 + u * DG;
            if (px < 128) {
              uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
              if (ok[u]) {
                float v[8];
                raw_to_arr(raw[u], v);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= dz_mul;
                split8<FMT>(v, hi, lo);
              }
              *reinterpret_cast<uint4*>(dstp + px * 16) = hi;
              if (a.nterms > 1) *reinterpret_cast<uint4*>(dstp + px * 16 + 16 * DZPL) = lo;
            }
          }
        }
      }
      // ---- view halo tile: 180 pixels x NPL planes ----
      if (vactive) {
        uint8_t* dstp = stage + DZ_BYTES + vpl * VPL;
        auto run = [&](auto mode_tag, auto batch_tag) {
          constexpr int MODE = decltype(mode_tag)::value;
          constexpr int U = decltype(batch_tag)::value;
          for (int p0 = vpx0; p0 < kWgHaloPx; p0 += VG * U) {
            Raw8 raw[U][RawCount<MODE>::value];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int p = p0 + u * VG;
              const int hr = p / kWgHaloW, hc = p - hr * kWgHaloW;
              const int h = h0 - 1 + hr, w = w0 - 1 + hc;
              ok[u] = p < kWgHaloPx && h >= 0 && h < V.H && w >= 0 && w < V.W;
              if (ok[u]) view_issue<MODE>(VS, view_pix_off(VS, n, h, w), vcc, raw[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int p = p0 + u * VG;
              if (p < kWgHaloPx) {
                uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
                if (ok[u]) {
                  float v[8];
                  view_finish<MODE>(raw[u], sc, sh, 1.f, v);
                  split8<FMT>(v, hi, lo);
                }
                *reinterpret_cast<uint4*>(dstp + p * 16) = hi;
                if (a.nterms > 1) *reinterpret_cast<uint4*>(dstp + p * 16 + NPL * VPL) = lo;
              }
            }
          }
        };
        switch (VS.mode) {
          case SRC_IDENTITY: run(std::integral_constant<int, SRC_IDENTITY>{}, std::integral_constant<int, 2>{}); break;
          case SRC_AFFINE_RELU_POOL: run(std::integral_constant<int, SRC_AFFINE_RELU_POOL>{}, std::integral_constant<int, 1>{}); break;
          case SRC_AFFINE_RELU_UP: run(std::integral_constant<int, SRC_AFFINE_RELU_UP>{}, std::integral_constant<int, 2>{}); break;
          default: run(std::integral_constant<int, SRC_AFFINE_RELU>{}, std::integral_constant<int, 2>{}); break;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full[s]);
    }
    if (warp < 6) {
      const int q = warp & 3;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int row = 32 * q + lane;
      for (int t = 0; t < 9; ++t) {
        for (int col0 = 0; col0 < NT; col0 += 16) {
          uint32_t rg[16];
          tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * NT + col0), rg);
          tmem_ld_wait();
          if (row < cvalid && kt1 > kt0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ci = ci0 + col0 + j;
              if (ci < a.CinReal)
                atomicAdd(a.dw + ((size_t)(co0 + row) * a.CinReal + ci) * 9 + t, __uint_as_float(rg[j]) * out_mul);
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

int launch_wgrad3x3(const ViewDesc& view, const float* dz, const float* dz_amax, float* dw, int Cout, int CinReal,
                    int nterms, int fmt, int variant, cudaStream_t st) {
  TNB_REQUIRE(view.C % 32 == 0 && Cout % 64 == 0, "wgrad3x3: unsupported channels Cin=%d Cout=%d", view.C, Cout);
  WgradArgs a;
  a.view = view; a.dz = dz; a.dz_amax = dz_amax; a.dw = dw; a.Cout = Cout; a.CinReal = CinReal; a.nterms = nterms; a.variant = variant;
  a.NT = (view.C % 48 == 0) ? 48 : 32;
  a.tiles_h = (view.H + kWgTileH - 1) / kWgTileH;
  a.tiles_w = (view.W + kWgTileW - 1) / kWgTileW;
  a.ktiles = view.N * a.tiles_h * a.tiles_w;
  a.ncot = (Cout + 127) / 128;
  const int gx = a.ncot * (view.C / a.NT);
  // split the pixel (K) range so that the grid is ~3 waves of 148 SMs, each CTA owning >= 4 K tiles
  int splits = (3 * 148 + gx - 1) / gx;
  if (splits > (a.ktiles + 3) / 4) splits = (a.ktiles + 3) / 4;
  if (splits < 1) splits = 1;
  a.ktiles_per_cta = (a.ktiles + splits - 1) / splits;
  splits = (a.ktiles + a.ktiles_per_cta - 1) / a.ktiles_per_cta;
  const int TP = nterms > 1 ? 2 : 1;
  const size_t smem = kHdrBytes + 2 * (size_t)(TP * 16 * pad_px(128) * 16 + TP * (a.NT / 8) * pad_px(kWgHaloPx) * 16);
  auto kern = fmt == 0 ? wgrad3x3_kernel<0> : wgrad3x3_kernel<1>;
  TNB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(PROF_WGRAD, st, view.N, view.H, view.W, view.C, Cout);
  kern<<<dim3(gx, splits), kThreads, smem, st>>>(a);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tnb
