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

static constexpr int kThreads = 320;      // warp0: MMA issue + TMEM alloc, warp1: weight TMA, warps 2..9: fill
static constexpr int kFillThreads = 256;  // warps 2..5 double as the epilogue warps
static constexpr int kMaxSmem = 232448;   // 227 KB opt-in limit per CTA on sm_100
static constexpr int kHdrBytes = 256;

__host__ __device__ inline int pad_px(int px) {  // plane stride ≡ 2 (mod 8) pixels -> conflict-free fill stores
  int r = px & 7;
  return px + ((2 - r) & 7);
}

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

static int pick_bn(int nside) {
  if (nside % 256 == 0) return 256;
  if (nside % 192 == 0) return 192;
  if (nside % 128 == 0) return 128;
  if (nside % 64 == 0) return 64;
  if (nside % 32 == 0) return 32;
  return 0;
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
// conv3x3 forward / dgrad
// =============================================================================================
struct ConvArgs {
  ViewDesc view;
  const uint16_t* wpack;
  float* out;        // [N,H,W,Cout]
  float* stat_part;  // [gridDim.x][2][Cout] per-tile (sum, sumsq) partials, or nullptr
  int Cout, BN, MT, SA, SB, nterms, variant, tmem_cols;
  int tiles_h, tiles_w;
};

// 31-shuffle transpose-reduce: on return lane j holds sum over the 32 lanes of v[j].
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

template <int FMT>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_kernel(const __grid_constant__ ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  const ViewDesc& V = a.view;
  const int MT = a.MT, BN = a.BN;
  const int PITCH = 8 * MT + 2;
  const int HALO_PX = 18 * PITCH;
  const int PLANE = pad_px(HALO_PX) * 16;  // bytes
  const int TP = a.nterms > 1 ? 2 : 1;     // operand term planes stored (hi[,lo])
  const int A_STAGE = TP * 4 * PLANE;
  const int B_STAGE = TP * 64 * BN;        // bytes: [term][4 planes][BN][16B]
  const int nchunks = V.C / 32;

  uint64_t* full_A = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_A = full_A + 2;
  uint64_t* full_B = full_A + 4;
  uint64_t* empty_B = full_A + 12;
  uint64_t* tmem_full = full_A + 20;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(full_A + 21);
  int2* table = reinterpret_cast<int2*>(smem + kHdrBytes);
  const int table_bytes = (HALO_PX * 8 + 127) & ~127;
  uint8_t* a_base = smem + kHdrBytes + table_bytes;
  uint8_t* b_base = a_base + a.SA * A_STAGE;

  // ---- tile coordinates ----
  int tile = blockIdx.x;
  const int tw = tile % a.tiles_w; tile /= a.tiles_w;
  const int th = tile % a.tiles_h;
  const int n = tile / a.tiles_h;
  const int h0 = th * 16, w0 = tw * 8 * MT;
  const int n0 = blockIdx.y * BN;

  // ---- one-time setup ----
  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < a.SA; ++i) { mbar_init(&full_A[i], kFillThreads); mbar_init(&empty_A[i], 1); }
      for (int i = 0; i < a.SB; ++i) { mbar_init(&full_B[i], 1); mbar_init(&empty_B[i], 1); }
      mbar_init(tmem_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, a.tmem_cols);
  }
  for (int p = tid; p < HALO_PX; p += kThreads) {
    const int hr = p / PITCH, hc = p - hr * PITCH;
    const int h = h0 - 1 + hr, w = w0 - 1 + hc;
    int2 e = make_int2(-1, -1);
    if (h >= 0 && h < V.H && w >= 0 && w < V.W) {
      e.x = view_pix_off(V.s[0], n, h, w);
      if (V.C0 < V.C) e.y = view_pix_off(V.s[1], n, h, w);
    }
    table[p] = e;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // dgrad: dz is multiplied by a power of two on the way in (so that its fp16 hi/lo split keeps ~22 bits) and
  // the accumulator by the inverse on the way out; forward views use BN scale/shift instead (in_mul = 1).
  const float in_mul = (V.s[0].mode == SRC_IDENTITY) ? pow2_scale_for(V.s[0].scale) : 1.f;
  const float out_mul = 1.f / in_mul;

  if (warp == 0) {
    // =========================== MMA issuer (single elected thread) ===========================
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128, BN, FMT, 0, 0);
      // A: K-major planar halo tile. LBO = plane stride (next 8 channels), SBO = halo row pitch (next
      // 8 output pixels = next image row of the 16x8 tile). variant bit0 swaps the two (bring-up probe).
      uint32_t a_lbo = PLANE, a_sbo = PITCH * 16;
      uint32_t b_lbo = BN * 16, b_sbo = 128;
      if (a.variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
      if (a.variant & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      int sb = 0; uint32_t phb = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int sa = c % a.SA;
        const uint32_t pha = (c / a.SA) & 1;
        mbar_wait(&full_A[sa], pha);
        tc_fence_after();
        const uint32_t a_stage = smem_u32(a_base + sa * A_STAGE);
        for (int t = 0; t < 9; ++t) {
          mbar_wait(&full_B[sb], phb);
          tc_fence_after();
          const uint32_t b_stage = smem_u32(b_base + sb * B_STAGE);
          const int dy = t / 3, dx = t - dy * 3;
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t d_tmem = tmem_base + mt * BN;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint32_t a_addr = a_stage + (2 * kk) * PLANE + (dy * PITCH + dx + 8 * mt) * 16;
              const uint32_t b_addr = b_stage + (2 * kk) * BN * 16;
              const uint64_t a_hi = make_smem_desc(a_addr, a_lbo, a_sbo);
              const uint64_t b_hi = make_smem_desc(b_addr, b_lbo, b_sbo);
              const uint32_t acc = (c | t | kk) != 0;
              umma_f16(d_tmem, a_hi, b_hi, idesc, acc);
              if (a.nterms > 1) {
                const uint64_t a_lo = make_smem_desc(a_addr + 4 * PLANE, a_lbo, a_sbo);
                const uint64_t b_lo = make_smem_desc(b_addr + 4 * BN * 16, b_lbo, b_sbo);
                umma_f16(d_tmem, a_lo, b_hi, idesc, 1);
                umma_f16(d_tmem, a_hi, b_lo, idesc, 1);
              }
            }
          }
          umma_commit(&empty_B[sb]);
          if (++sb == a.SB) { sb = 0; phb ^= 1; }
        }
        umma_commit(&empty_A[sa]);
      }
      umma_commit(tmem_full);
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== weight loader (bulk TMA) ===========================
    if (elect_one()) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpack) +
                            (size_t)blockIdx.y * nchunks * 9 * (size_t)(128 * BN);
      int sb = 0; uint32_t phb = 0;
      for (int i = 0; i < nchunks * 9; ++i) {
        mbar_wait(&empty_B[sb], phb ^ 1);
        mbar_arrive_expect_tx(&full_B[sb], (uint32_t)B_STAGE);
        bulk_g2s(b_base + sb * B_STAGE, wsrc + (size_t)i * (128 * BN), (uint32_t)B_STAGE, &full_B[sb]);
        if (++sb == a.SB) { sb = 0; phb ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =========================== A producers: gather + BN/ReLU/pool/upsample + split ===========
    const int ftid = tid - 64;
    const int j = ftid & 3;            // plane (8 channels) this thread fills: fixed, 256 % 4 == 0
    const int pbase = ftid >> 2;       // first halo pixel; stride 64 pixels
    for (int c = 0; c < nchunks; ++c) {
      const int sa = c % a.SA;
      const uint32_t pha = (c / a.SA) & 1;
      const int cch = c * 32 + j * 8;
      const bool second = cch >= V.C0;
      const SrcDesc& S = second ? V.s[1] : V.s[0];
      const int cc = second ? cch - V.C0 : cch;
      float sc[8], sh[8];
      if (S.mode != SRC_IDENTITY) { ld8(S.scale + cc, sc); ld8(S.shift + cc, sh); }
      mbar_wait(&empty_A[sa], pha ^ 1);
      uint8_t* stage = a_base + sa * A_STAGE + j * PLANE;
      auto run = [&](auto mode_tag, auto batch_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        constexpr int U = decltype(batch_tag)::value;
        for (int p0 = pbase; p0 < HALO_PX; p0 += 64 * U) {
          Raw8 raw[U][RawCount<MODE>::value];
          int off[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = p0 + 64 * u;
            off[u] = -1;
            if (p < HALO_PX) {
              const int2 e = table[p];
              off[u] = second ? e.y : e.x;
              if (off[u] >= 0) view_issue<MODE>(S, off[u], cc, raw[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = p0 + 64 * u;
            if (p < HALO_PX) {
              uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
              if (off[u] >= 0) {
                float v[8];
                view_finish<MODE>(raw[u], sc, sh, in_mul, v);
                split8<FMT>(v, hi, lo);
              }
              uint8_t* dst = stage + p * 16;
              *reinterpret_cast<uint4*>(dst) = hi;
              if (a.nterms > 1) *reinterpret_cast<uint4*>(dst + 4 * PLANE) = lo;
            }
          }
        }
      };
      switch (S.mode) {
        case SRC_IDENTITY: run(std::integral_constant<int, SRC_IDENTITY>{}, std::integral_constant<int, 4>{}); break;
        case SRC_AFFINE_RELU_POOL: run(std::integral_constant<int, SRC_AFFINE_RELU_POOL>{}, std::integral_constant<int, 2>{}); break;
        case SRC_AFFINE_RELU_UP: run(std::integral_constant<int, SRC_AFFINE_RELU_UP>{}, std::integral_constant<int, 4>{}); break;
        default: run(std::integral_constant<int, SRC_AFFINE_RELU>{}, std::integral_constant<int, 4>{}); break;
      }
      fence_proxy_async_smem();
      mbar_arrive(&full_A[sa]);
    }

    // =========================== epilogue (warps 2..5) ===========================
    if (warp < 6) {
      const int q = warp & 3;  // TMEM lane quarter this warp may access
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      float* sstat = reinterpret_cast<float*>(a_base);  // [4 warps][2][BN], A stages are dead by now
      const int row = 32 * q + lane;
      const int r = row >> 3, cc = row & 7;
      for (int col0 = 0; col0 < BN; col0 += 32) {
        float csum = 0.f, csq = 0.f;
        for (int mt = 0; mt < MT; ++mt) {
          const int h = h0 + r, w = w0 + 8 * mt + cc;
          const bool valid = (h < V.H) && (w < V.W);
          uint32_t rg[32];
          tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(mt * BN + col0), rg);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rg[i]) * out_mul;
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(a.out + ((size_t)(n * V.H + h) * V.W + w) * a.Cout + n0 + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (a.stat_part != nullptr) {
            float s1[32], s2[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = valid ? v[i] : 0.f;
              s1[i] = x;
              s2[i] = x * x;
            }
            csum += warp_transpose_sum(s1, lane);
            csq += warp_transpose_sum(s2, lane);
          }
        }
        if (a.stat_part != nullptr) {
          sstat[(q * 2 + 0) * BN + col0 + lane] = csum;
          sstat[(q * 2 + 1) * BN + col0 + lane] = csq;
        }
      }
      if (a.stat_part != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps only
        const int et = tid - 64;                        // 0..127
        for (int j = et; j < 2 * BN; j += 128) {
          const int which = j / BN, col = j - which * BN;
          const float s = sstat[(0 * 2 + which) * BN + col] + sstat[(1 * 2 + which) * BN + col] +
                          sstat[(2 * 2 + which) * BN + col] + sstat[(3 * 2 + which) * BN + col];
          a.stat_part[((size_t)blockIdx.x * 2 + which) * a.Cout + n0 + col] = s;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

static int pow2_cols(int c) {
  int p = 32;
  while (p < c) p <<= 1;
  return p;
}

int conv3x3_plan(int N, int H, int W, int Cin, int Cout, int nterms, ConvPlan* plan) {
  TNB_REQUIRE(Cin % 32 == 0, "conv3x3: view channels %d must be a multiple of 32", Cin);
  const int BN = pick_bn(Cout);
  TNB_REQUIRE(BN >= 32 && BN % 16 == 0, "conv3x3: unsupported output channel count %d", Cout);
  const int TP = nterms > 1 ? 2 : 1;
  int MT = 512 / BN;
  if (MT > 4) MT = 4;
  while (MT > 1 && 8 * (MT - 1) >= W) --MT;  // do not tile wider than the image
  int SA = 2, SB = 0;
  size_t smem = 0;
  for (;; --MT) {
    const int pitch = 8 * MT + 2, halo = 18 * pitch;
    const size_t a_stage = (size_t)TP * 4 * pad_px(halo) * 16;
    const size_t b_stage = (size_t)TP * 64 * BN;
    const size_t fixed = kHdrBytes + ((halo * 8 + 127) & ~127) + SA * a_stage;
    if (fixed + 2 * b_stage <= (size_t)kMaxSmem) {
      SB = (int)((kMaxSmem - fixed) / b_stage);
      if (SB > 8) SB = 8;
      smem = fixed + SB * b_stage;
      break;
    }
    TNB_REQUIRE(MT > 1, "conv3x3: no shared-memory plan for Cin=%d Cout=%d", Cin, Cout);
  }
  plan->BN = BN; plan->MT = MT; plan->SA = SA; plan->SB = SB;
  plan->tmem_cols = pow2_cols(MT * BN);
  plan->smem_bytes = smem;
  plan->tiles_h = (H + 15) / 16;
  plan->tiles_w = (W + 8 * MT - 1) / (8 * MT);
  (void)N;
  return 0;
}

int conv3x3_num_stat_rows(int N, int H, int W, int Cin, int Cout, int nterms) {
  ConvPlan p;
  if (conv3x3_plan(N, H, W, Cin, Cout, nterms, &p)) return -1;
  return N * p.tiles_h * p.tiles_w;
}

int launch_conv3x3(const ViewDesc& view, const uint16_t* wpack, float* out, float* stat_part, int Cout,
                   int nterms, int fmt, int variant, cudaStream_t st) {
  ConvPlan p;
  int rc = conv3x3_plan(view.N, view.H, view.W, view.C, Cout, nterms, &p);
  if (rc) return rc;
  ConvArgs a;
  a.view = view; a.wpack = wpack; a.out = out; a.stat_part = stat_part;
  a.Cout = Cout; a.BN = p.BN; a.MT = p.MT; a.SA = p.SA; a.SB = p.SB; a.nterms = nterms; a.variant = variant;
  a.tmem_cols = p.tmem_cols; a.tiles_h = p.tiles_h; a.tiles_w = p.tiles_w;
  dim3 grid(view.N * p.tiles_h * p.tiles_w, Cout / p.BN);
  auto kern = fmt == 0 ? conv3x3_kernel<0> : conv3x3_kernel<1>;
  TNB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
  ProfScope prof((view.s[0].mode == SRC_IDENTITY && view.s[0].scale != nullptr) ? PROF_CONV_DGRAD : PROF_CONV_FWD, st, view.N, view.H, view.W, view.C, Cout);
  kern<<<grid, kThreads, p.smem_bytes, st>>>(a);
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
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128, NT, FMT, 1, 1);
      // MN-major planar tiles: SBO = plane stride (next 8 channels), LBO = 128 B (next 8 pixels).
      uint32_t a_lbo = 128, a_sbo = DZPL, b_lbo = 128, b_sbo = VPL;
      if (a.variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
      if (a.variant & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      int it = 0;
      for (int kt = kt0; kt < kt1; ++kt, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t dz_s = smem_u32(st_base + s * STAGE);
        const uint32_t v_s = dz_s + DZ_BYTES;
        for (int r = 0; r < kWgTileH; ++r) {
          const uint32_t a_addr = dz_s + (r * 16) * 16;
          const uint64_t a_hi = make_smem_desc(a_addr, a_lbo, a_sbo);
          const uint64_t a_lo = make_smem_desc(a_addr + 16 * DZPL, a_lbo, a_sbo);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int dy = t / 3, dx = t % 3;
            const uint32_t b_addr = v_s + ((r + dy) * kWgHaloW + dx) * 16;
            const uint64_t b_hi = make_smem_desc(b_addr, b_lbo, b_sbo);
            const uint32_t d_tmem = tmem_base + t * NT;
            const uint32_t acc = (it | r) != 0;
            umma_f16(d_tmem, a_hi, b_hi, idesc, acc);
            if (a.nterms > 1) {
              const uint64_t b_lo = make_smem_desc(b_addr + NPL * VPL, b_lbo, b_sbo);
              umma_f16(d_tmem, a_lo, b_hi, idesc, 1);
              umma_f16(d_tmem, a_hi, b_lo, idesc, 1);
            }
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
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
          case SRC_IDENTITY: run(std::integral_constant<int, SRC_IDENTITY>{}, std::integral_constant<int, 4>{}); break;
          case SRC_AFFINE_RELU_POOL: run(std::integral_constant<int, SRC_AFFINE_RELU_POOL>{}, std::integral_constant<int, 2>{}); break;
          case SRC_AFFINE_RELU_UP: run(std::integral_constant<int, SRC_AFFINE_RELU_UP>{}, std::integral_constant<int, 4>{}); break;
          default: run(std::integral_constant<int, SRC_AFFINE_RELU>{}, std::integral_constant<int, 4>{}); break;
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
