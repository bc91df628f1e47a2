// tcgen05 implicit-GEMM weight gradient of the 3x3 convolution for sm_100a.
//
//   dW[co][ci][dy][dx] += sum_{n,h,w} dz[n,h,w,co] * view[n,h+dy-1,w+dx-1,ci]
//
// Replaces cuDNN's Conv2d wgrad reached through `loss.backward()` (reference train.py:95) for the
// Conv2DBlock convolutions (reference model.py:8,13); `view` is the layer's logical input with the producer's
// BatchNorm/ReLU/MaxPool/Upsample/cat fused into the gather exactly as in the forward kernel (conv.cu).
//
// GEMM per tap: M = 128 output channels, N = 32 or 48 input channels, K = pixels. Both operands are MN-major
// views of planar tiles [plane = 8 channels][pixel][16 B]; the 9 taps are 9 start addresses inside one halo tile
// of the view, and their 9 x N fp32 accumulator columns stay resident in TMEM while the CTA streams 8x16-pixel
// K tiles of its pixel range (split-K over CTAs; epilogue red.global.add.f32 into the OIHW gradient).
//
// Warp roles (512 threads, registers re-balanced with setmaxnreg): warp 0 issues tcgen05.mma; warps 4-15
// (384 threads) gather dz and the view: all global loads of a K tile are issued before the first is consumed
// and the NEXT tile's lines are prefetched into L2, so HBM latency is paid once per tile; warps 4-7 run the epilogue.
#include "igemm.cuh"
#include "prof.cuh"
#include <type_traits>

namespace tnb {

static constexpr int kThreads = 512;
static constexpr int kFillThreads = 384;
static constexpr int kHdrBytes = 256;
static constexpr int kWgTileH = 8, kWgTileW = 16;            // pixels per K tile = 128
static constexpr int kWgHaloW = kWgTileW + 2;                // 18
static constexpr int kWgHaloPx = (kWgTileH + 2) * kWgHaloW;  // 180
static constexpr int kStages = 2;

struct WgradArgs {
  ViewDesc view;
  const float* dz;       // [N,H,W,Cout]
  const float* dz_amax;  // optional: max|dz| (device scalar) -> power-of-two pre-scaling for the fp16 split
  float* dw;             // [Cout][CinReal][3][3], accumulated with atomics (must be zeroed by the caller)
  int Cout, CinReal, NT, nterms, variant;
  int tiles_h, tiles_w, ktiles, ktiles_per_cta, ncot;
};

struct TileCoord { int n, h0, w0; };
TNB_DEVINL TileCoord tile_coord(const WgradArgs& a, int kt) {
  const int per_img = a.tiles_h * a.tiles_w;
  TileCoord t;
  t.n = kt / per_img;
  const int rem = kt - t.n * per_img;
  const int th = rem / a.tiles_w;
  t.h0 = th * kWgTileH;
  t.w0 = (rem - th * a.tiles_w) * kWgTileW;
  return t;
}

template <int FMT>
__global__ void __launch_bounds__(kThreads, 1) wgrad3x3_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const ViewDesc& V = a.view;
  const int NT = a.NT, NPL = NT / 8;
  const int TP = a.nterms > 1 ? 2 : 1;
  const int DZPL = pad_px(128) * 16;       // 2080
  const int VPL = pad_px(kWgHaloPx) * 16;  // 2976
  const int DZ_BYTES = TP * 16 * DZPL;
  const int STAGE = DZ_BYTES + TP * NPL * VPL;
  constexpr int S = kStages;

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
  const float dz_mul = pow2_scale_for(a.dz_amax);
  const float out_mul = 1.f / dz_mul;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // whole warp runs the uniform loops (descriptor math in uniform registers); one lane issues
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
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int ftid = tid - 128;
    // dz items: thread owns plane dpl, pixels dpx0 + u*DG;  view items: plane vpl, halo pixels vpx0 + u*VG
    const int dpl = ftid % npld, dpx0 = ftid / npld, DG = kFillThreads / npld;  // DG = 48 or 24
    const int VG = kFillThreads / NPL;                                           // 96 or 64
    const int vpl = ftid % NPL, vpx0 = ftid / NPL;
    const int vch = ci0 + vpl * 8;
    const bool vsecond = vch >= V.C0;
    const SrcDesc& VS = vsecond ? V.s[1] : V.s[0];
    const int vcc = vsecond ? vch - V.C0 : vch;
    float sc[8], sh[8];
    if (VS.mode != SRC_IDENTITY) { ld8(VS.scale + vcc, sc); ld8(VS.shift + vcc, sh); }
    const float* dz_base = a.dz + co0 + dpl * 8;

    int it = 0;
    for (int kt = kt0; kt < kt1; ++kt, ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      const TileCoord tc = tile_coord(a, kt);
      // ---- L2 prefetch of the next K tile (one 128-byte line covers 4 planes) ----
      if (kt + 1 < kt1) {
        const TileCoord nx = tile_coord(a, kt + 1);
        if ((dpl & 3) == 0) {
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int px = dpx0 + u * DG;
            const int h = nx.h0 + (px >> 4), w = nx.w0 + (px & 15);
            if (px < 128 && h < V.H && w < V.W)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(dz_base + ((size_t)(nx.n * V.H + h) * V.W + w) * a.Cout));
          }
        }
        if ((vpl & 3) == 0) {
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const int p = vpx0 + u * VG;
            const int hr = p / kWgHaloW, hc = p - hr * kWgHaloW;
            const int h = nx.h0 - 1 + hr, w = nx.w0 - 1 + hc;
            if (p < kWgHaloPx && h >= 0 && h < V.H && w >= 0 && w < V.W) {
              const float* q = VS.ptr + (size_t)view_pix_off(VS, nx.n, h, w) * VS.C + vcc;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
              if (VS.mode == SRC_AFFINE_RELU_POOL) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q + VS.C));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q + (size_t)VS.Ws * VS.C));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q + (size_t)VS.Ws * VS.C + VS.C));
              }
            }
          }
        }
      }
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* stage = st_base + s * STAGE;
      uint8_t* dzp = stage + dpl * DZPL;
      uint8_t* vwp = stage + DZ_BYTES + vpl * VPL;

      auto dz_issue = [&](int u0, Raw8 (&raw)[3], bool (&ok)[3]) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int px = dpx0 + (u0 + u) * DG;
          const int h = tc.h0 + (px >> 4), w = tc.w0 + (px & 15);
          ok[u] = px < 128 && h < V.H && w < V.W;
          if (ok[u]) raw[u] = ld_raw8(dz_base + ((size_t)(tc.n * V.H + h) * V.W + w) * a.Cout);
        }
      };
      auto dz_finish = [&](int u0, const Raw8 (&raw)[3], const bool (&ok)[3]) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int px = dpx0 + (u0 + u) * DG;
          if (px < 128) {
            uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
            if (ok[u]) {
              float v[8];
              raw_to_arr(raw[u], v);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= dz_mul;
              split8<FMT>(v, hi, lo);
            }
            *reinterpret_cast<uint4*>(dzp + px * 16) = hi;
            if (a.nterms > 1) *reinterpret_cast<uint4*>(dzp + px * 16 + 16 * DZPL) = lo;
          }
        }
      };
      auto view_run = [&](auto mode_tag, auto batch_tag, auto with_dz_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        constexpr int U = decltype(batch_tag)::value;
        constexpr bool WITH_DZ = decltype(with_dz_tag)::value;
        Raw8 draw[3];
        bool dok[3];
        if (WITH_DZ) dz_issue(0, draw, dok);  // dz and view loads of this K tile are all in flight together
        for (int p0 = vpx0; p0 < kWgHaloPx; p0 += VG * U) {
          Raw8 raw[U][RawCount<MODE>::value];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = p0 + u * VG;
            const int hr = p / kWgHaloW, hc = p - hr * kWgHaloW;
            const int h = tc.h0 - 1 + hr, w = tc.w0 - 1 + hc;
            ok[u] = p < kWgHaloPx && h >= 0 && h < V.H && w >= 0 && w < V.W;
            if (ok[u]) view_issue<MODE>(VS, view_pix_off(VS, tc.n, h, w), vcc, raw[u]);
          }
          if (WITH_DZ && p0 == vpx0) dz_finish(0, draw, dok);
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
              *reinterpret_cast<uint4*>(vwp + p * 16) = hi;
              if (a.nterms > 1) *reinterpret_cast<uint4*>(vwp + p * 16 + NPL * VPL) = lo;
            }
          }
        }
      };
      using T = std::true_type;
      using F = std::false_type;
      switch (VS.mode) {
        case SRC_IDENTITY: view_run(std::integral_constant<int, SRC_IDENTITY>{}, std::integral_constant<int, 3>{}, T{}); break;
        case SRC_AFFINE_RELU_UP: view_run(std::integral_constant<int, SRC_AFFINE_RELU_UP>{}, std::integral_constant<int, 3>{}, T{}); break;
        case SRC_AFFINE_RELU_POOL: {
          Raw8 draw[3];
          bool dok[3];
          dz_issue(0, draw, dok);
          dz_finish(0, draw, dok);
          view_run(std::integral_constant<int, SRC_AFFINE_RELU_POOL>{}, std::integral_constant<int, 1>{}, F{});
          break;
        }
        default: view_run(std::integral_constant<int, SRC_AFFINE_RELU>{}, std::integral_constant<int, 3>{}, T{}); break;
      }
      if (npld == 16) {  // second half of the dz pixels (DG = 24: six pixels per thread)
        Raw8 draw[3];
        bool dok[3];
        dz_issue(3, draw, dok);
        dz_finish(3, draw, dok);
      }
      fence_proxy_async_smem();
      mbar_arrive(&full[s]);
    }

    if (warp < 8) {
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
  a.view = view; a.dz = dz; a.dz_amax = dz_amax; a.dw = dw; a.Cout = Cout; a.CinReal = CinReal; a.nterms = nterms;
  a.variant = variant;
  a.NT = (view.C % 48 == 0) ? 48 : 32;
  a.tiles_h = (view.H + kWgTileH - 1) / kWgTileH;
  a.tiles_w = (view.W + kWgTileW - 1) / kWgTileW;
  a.ktiles = view.N * a.tiles_h * a.tiles_w;
  a.ncot = (Cout + 127) / 128;
  const int gx = a.ncot * (view.C / a.NT);
  // split the pixel (K) range so that the grid is ~2 waves of 148 SMs, each CTA owning >= 4 K tiles
  int splits = (2 * 148 + gx - 1) / gx;
  if (splits > (a.ktiles + 3) / 4) splits = (a.ktiles + 3) / 4;
  if (splits < 1) splits = 1;
  a.ktiles_per_cta = (a.ktiles + splits - 1) / splits;
  splits = (a.ktiles + a.ktiles_per_cta - 1) / a.ktiles_per_cta;
  const int TP = nterms > 1 ? 2 : 1;
  const size_t smem =
      kHdrBytes + kStages * (size_t)(TP * 16 * pad_px(128) * 16 + TP * (a.NT / 8) * pad_px(kWgHaloPx) * 16);
  auto kern = fmt == 0 ? wgrad3x3_kernel<0> : wgrad3x3_kernel<1>;
  TNB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(PROF_WGRAD, st, view.N, view.H, view.W, view.C, Cout);
  kern<<<dim3(gx, splits), kThreads, smem, st>>>(a);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tnb
