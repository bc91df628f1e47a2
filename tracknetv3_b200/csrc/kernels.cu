// HBM-bound sm_100a kernels of the TrackNet / InpaintNet hot path. Each cites the reference lines it
// replaces; none of them falls back to a library.
#include "kernels.cuh"
#include <algorithm>
#include "prof.cuh"
#include <limits.h>

namespace tnb {

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// =============================================================================================
// input packing: NCHW fp32 (reference train.py:86 x.float().cuda()) -> NHWC fp32, channels padded to Cpad
// =============================================================================================
__global__ void pack_input_kernel(const float* __restrict__ x, float* __restrict__ out, uint8_t* __restrict__ planar,
                                  uint8_t* __restrict__ presplit, int presplit_fmt, int N, int C, int H, int W, int Cpad) {
  pdl_launch_dependents();
  pdl_wait();
  const long long npix = (long long)N * H * W;
  const long long hw = (long long)H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix;
       p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / hw, r = p - n * hw;
    const float* src = x + n * C * hw + r;
    float4* dst = reinterpret_cast<float4*>(out + p * Cpad);
    for (int c8 = 0; c8 < Cpad; c8 += 8) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (c8 + e < C) ? __ldg(src + (long long)(c8 + e) * hw) : 0.f;
      if (out != nullptr) {
        dst[c8 >> 2] = make_float4(v[0], v[1], v[2], v[3]);
        dst[(c8 >> 2) + 1] = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (planar != nullptr) {
        // [N][chunk of 32 ch][term][plane of 8 ch][H][W][8 x fp16]: consecutive threads = consecutive pixels of a
        // row = consecutive 16-byte elements of a plane (coalesced), and what a TMA box {8 W', H', planes} lands as the
        // planar halo tile [plane][row][pixel][16 B] of the tcgen05 convolution
        uint4 hi, lo;
        split8<0>(v, hi, lo);
        const int chunk = c8 >> 5, pl = (c8 >> 3) & 3;
        const size_t plane_hi = ((size_t)n * (Cpad >> 5) + chunk) * 8 + pl;
        *reinterpret_cast<uint4*>(planar + (plane_hi * hw + r) * 16) = hi;
        *reinterpret_cast<uint4*>(planar + ((plane_hi + 4) * hw + r) * 16) = lo;
      }
      if (presplit != nullptr) {  // [pixel][2 (hi, lo)][Cpad] 16-bit: the first layer's weight-gradient operand
        uint4 hi, lo;
        if (presplit_fmt == 0) split8<0>(v, hi, lo); else split8<1>(v, hi, lo);
        uint8_t* dst16 = presplit + ((size_t)p * 2 * Cpad + c8) * 2;
        *reinterpret_cast<uint4*>(dst16) = hi;
        *reinterpret_cast<uint4*>(dst16 + 2 * Cpad) = lo;
      }
    }
  }
}
int launch_pack_input(const float* x, float* out, int N, int C, int H, int W, int Cpad, cudaStream_t st, void* planar,
                      void* presplit, int presplit_fmt) {
  TNB_REQUIRE(Cpad % 8 == 0 && (planar == nullptr || Cpad % 32 == 0), "pack_input: padded channel count %d", Cpad);
  const long long npix = (long long)N * H * W;
  if (int rc = launch_pdl(pack_input_kernel, dim3(min(cdiv(npix, 256), 148 * 16)), dim3(256), 0, st, x, out,
                          (uint8_t*)planar, (uint8_t*)presplit, presplit_fmt, N, C, H, W, Cpad)) return rc;
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// BatchNorm2d statistics finalisation (reference model.py:9 nn.BatchNorm2d, eps 1e-5, momentum 0.1).
// Consumes the per-tile (sum, sumsq) partials written by the conv epilogue in a fixed order
// (deterministic), combines in fp64, emits the fused affine (scale, shift) used by every consumer,
// and updates the running statistics exactly like torch (biased var to normalise, unbiased to track).
// =============================================================================================
// Column sums of a [rows][2][C] partials buffer in fp64 for the kFinCh channels of this block: thread (x = channel,
// y = row lane of kFinLanes); fixed summation order (deterministic). Few channels per block so that a 64-channel layer
// still spreads over 8 SMs, many row lanes so that the serial part is short: these kernels sit between every pair of
// convolutions and are pure latency.
static constexpr int kFinCh = 8, kFinLanes = 128;
struct FinalizeSmem { double s1[kFinLanes][kFinCh + 1], s2[kFinLanes][kFinCh + 1]; };
// returns the two column sums in thread (x, y = 0); other threads get garbage
TNB_DEVINL void finalize_column_sums(const float* __restrict__ part, int rows, int C, int c, bool active,
                                     FinalizeSmem& sm, double& out1, double& out2) {
  double a = 0.0, b = 0.0;
  if (active) {
    float fa[4], fb[4];
    int r = threadIdx.y;
    for (; r + 3 * kFinLanes < rows; r += 4 * kFinLanes) {  // 8 independent loads in flight per thread
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        fa[u] = part[((size_t)(r + kFinLanes * u) * 2 + 0) * C + c];
        fb[u] = part[((size_t)(r + kFinLanes * u) * 2 + 1) * C + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a += (double)fa[u]; b += (double)fb[u]; }
    }
    for (; r < rows; r += kFinLanes) {
      a += (double)part[((size_t)r * 2 + 0) * C + c];
      b += (double)part[((size_t)r * 2 + 1) * C + c];
    }
  }
  sm.s1[threadIdx.y][threadIdx.x] = a;
  sm.s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  double sa = 0.0, sb = 0.0;
  if (threadIdx.y < 16)  // 128 -> 16 lanes
    for (int i = threadIdx.y; i < kFinLanes; i += 16) { sa += sm.s1[i][threadIdx.x]; sb += sm.s2[i][threadIdx.x]; }
  __syncthreads();
  if (threadIdx.y < 16) { sm.s1[threadIdx.y][threadIdx.x] = sa; sm.s2[threadIdx.y][threadIdx.x] = sb; }
  __syncthreads();
  out1 = out2 = 0.0;
  if (threadIdx.y == 0)
    for (int i = 0; i < 16; ++i) { out1 += sm.s1[i][threadIdx.x]; out2 += sm.s2[i][threadIdx.x]; }
}

__global__ void __launch_bounds__(kFinCh * kFinLanes) bn_finalize_kernel(const float* __restrict__ part, int rows, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* running_mean, float* running_var, float momentum, float eps, int training,
                                   float* scale, float* shift, float* mean_out, float* invstd_out, int C) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ FinalizeSmem sm;
  const int c = blockIdx.x * kFinCh + threadIdx.x;
  double sa, sb;
  finalize_column_sums(part, rows, C, c, training && c < C, sm, sa, sb);
  if (threadIdx.y == 0 && c < C) {
    float mean, invstd;
    if (training) {
      const double m = sa / count;
      double var = sb / count - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    } else {
      mean = running_mean[c];
      invstd = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
    }
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    mean_out[c] = mean;
    invstd_out[c] = invstd;
  }
}
int launch_bn_finalize(const float* part, int rows, double count, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, int training,
                       float* scale, float* shift, float* mean, float* invstd, int C, cudaStream_t st) {
  return launch_pdl(bn_finalize_kernel, dim3(cdiv(C, kFinCh)), dim3(kFinCh, kFinLanes), 0, st, part, rows, count, gamma, beta,
                    running_mean, running_var, momentum, eps, training, scale, shift, mean, invstd, C);
}

// =============================================================================================
// predictor: 1x1 conv (+bias) + sigmoid on relu(bn(z)) (reference model.py:54-55,71-72). fp32 FFMA: the
// 0.07% of the FLOPs that sit directly in front of the 1e-3 parity bound stay in full precision.
// Output is NCHW fp32, the layout the reference's callers consume (train.py:93, predict.py:140).
// =============================================================================================
static constexpr int kMaxPredO = 16;
// 4 threads per pixel, 16 channels each, interleaved in float4 units (thread qd owns channels 16*i + 4*qd .. +3):
// every load instruction of a warp reads whole 32-byte sectors (64 contiguous bytes per pixel); the 4 partial dot
// products are folded with two shuffles and thread qd stores outputs o = qd, qd + 4, ...
// FULL: O == O_MAX, known at compile time (the network's out_dim 8): the per-channel `o < O` tests - a quarter of the
// kernel's instructions as branches, predicates and reconvergence points (ncu source page) - disappear
template <int O_MAX, bool FULL>
__global__ void __launch_bounds__(256) predictor_fwd_kernel(SrcDesc src, int N, int H, int W,
                                                            const float* __restrict__ wp,
                                                            const float* __restrict__ bias, int O_rt, int o0, int OT,
                                                            float* __restrict__ y) {
  // this launch computes output channels [o0, o0 + O) of OT (wp / bias already point at channel o0)
  const int O = FULL ? O_MAX : O_rt;
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sw[kMaxPredO * 64 + kMaxPredO];
  for (int i = threadIdx.x; i < O * 64; i += blockDim.x) sw[i] = wp[i];
  for (int i = threadIdx.x; i < O; i += blockDim.x) sw[kMaxPredO * 64 + i] = bias[i];
  __syncthreads();
  const bool affine = src.mode != SRC_IDENTITY;
  const long long hw = (long long)H * W, npix = (long long)N * hw;
  const long long total = (npix * 4 + 31) / 32 * 32;  // whole warps: the shuffles below need every lane
  // (sample, pixel in sample) of this thread's item, advanced incrementally: a 64-bit division per item was a fifth of the
  // kernel's instructions (ncu: 706 warp instructions per 8 pixels, issue slots 64 % busy)
  const long long it0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long pstep = ((long long)gridDim.x * blockDim.x) >> 2;  // blockDim.x is a multiple of 4
  long long n = (it0 >> 2) / hw, r = (it0 >> 2) - n * hw;
  for (long long it = it0; it < total; it += (long long)gridDim.x * blockDim.x, r += pstep) {
    while (r >= hw) { r -= hw; ++n; }
    const long long p = it >> 2;
    const int qd = (int)(it & 3);
    const bool ok = p < npix;
    float a[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = 16 * i + 4 * qd;
      float4 v = ok ? __ldg(reinterpret_cast<const float4*>(src.ptr + p * 64 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (affine) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(src.scale + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(src.shift + c));
        v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
        v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
      }
      a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
    }
    float s[O_MAX];
#pragma unroll
    for (int o = 0; o < O_MAX; ++o) {
      s[o] = 0.f;
      if (o < O) {
#pragma unroll
        for (int i = 0; i < 16; ++i) s[o] = fmaf(a[i], sw[o * 64 + (i >> 2) * 16 + qd * 4 + (i & 3)], s[o]);
      }
      s[o] += __shfl_xor_sync(0xffffffffu, s[o], 1);
      s[o] += __shfl_xor_sync(0xffffffffu, s[o], 2);
    }
    if (ok) {
#pragma unroll
      for (int o = 0; o < O_MAX; ++o)
        if ((o & 3) == qd && o < O) y[(n * OT + o0 + o) * hw + r] = 1.f / (1.f + expf(-(s[o] + sw[kMaxPredO * 64 + o])));
    }
  }
}
int launch_predictor_fwd(const SrcDesc& src, int N, int H, int W, const float* wp, const float* bias, int O,
                         float* y, cudaStream_t st) {
  TNB_REQUIRE(src.C == 64 && O >= 1, "predictor: expects 64 input channels and out_dim >= 1 (got %d, %d)", src.C, O);
  TNB_REQUIRE((src.mode == SRC_IDENTITY || src.mode == SRC_AFFINE_RELU) && src.Hs == H && src.Ws == W,
              "predictor: the activation must be a same-resolution identity / BN+ReLU source (mode %d)", src.mode);
  const long long npix = (long long)N * H * W;
  ProfScope prof(PROF_PRED, st, N, H, W, 64, O);
  const int grid = (int)std::min<long long>((npix * 4 + 255) / 256, 148 * 8);
  // any out_dim (the reference accepts any seq_len, utils/general.py:66-74): groups of up to 16 output channels per launch
  for (int o0 = 0; o0 < O; o0 += kMaxPredO) {
    const int og = std::min(O - o0, kMaxPredO);
    auto go = [&](auto kern) { return launch_pdl(kern, dim3(grid), dim3(256), 0, st, src, N, H, W, wp + o0 * 64, bias + o0, og, o0, O, y); };
    if (int rc = og == 8 ? go(predictor_fwd_kernel<8, true>) : og < 8 ? go(predictor_fwd_kernel<8, false>)
               : og == kMaxPredO ? go(predictor_fwd_kernel<kMaxPredO, true>) : go(predictor_fwd_kernel<kMaxPredO, false>))
      return rc;
  }
  return 0;
}

// backward of sigmoid + 1x1 conv: dl = dy*y*(1-y); dA[p][c] = sum_o dl[o] W[o][c];
// dW[o][c] = sum_p dl[p][o] a[p][c]; db[o] = sum_p dl[p][o]   (autograd of model.py:71-72)
// Two streaming kernels without block-level barriers (a single tiled kernel with load -> barrier -> compute phases ran
// at 1.7 TB/s): dA needs only dl and W; dW / db need dl and the activation, reduced in registers along the pixels.
template <int O_MAX, bool FULL>
__global__ void __launch_bounds__(256) predictor_bwd_da_kernel(long long npix, long long hw, const float* __restrict__ wp,
                                                               int O_rt, int o0, int OT, const float* __restrict__ dy,
                                                               const float* __restrict__ y, float* __restrict__ dA) {
  // output channels [o0, o0 + O) of OT; groups after the first (o0 > 0) add to what the earlier launches wrote
  const int O = FULL ? O_MAX : O_rt;
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sw[kMaxPredO * 64];
  for (int i = threadIdx.x; i < O * 64; i += blockDim.x) sw[i] = wp[i];
  __syncthreads();
  // 4 threads per pixel, 16 channels each, interleaved in float4 units (thread qd owns channels 16*i + 4*qd .. +3,
  // i = 0..3): every store instruction of a warp then writes whole 32-byte sectors (64 contiguous bytes per pixel)
  const long long total = npix * 4;
  const long long it0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long pstep = ((long long)gridDim.x * blockDim.x) >> 2;
  long long n = (it0 >> 2) / hw, r = (it0 >> 2) - n * hw;  // advanced incrementally (no division in the loop)
  for (long long it = it0; it < total; it += (long long)gridDim.x * blockDim.x, r += pstep) {
    while (r >= hw) { r -= hw; ++n; }
    const long long p = it >> 2;
    const int qd = (int)(it & 3);
    float g[16], yv[O_MAX], dv[O_MAX];
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] = 0.f;
#pragma unroll
    for (int o = 0; o < O_MAX; ++o) {  // all 2 * O loads of the item in flight together
      yv[o] = o < O ? __ldg(y + (n * OT + o0 + o) * hw + r) : 0.f;
      dv[o] = o < O ? __ldg(dy + (n * OT + o0 + o) * hw + r) : 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(dA + p * 64 + qd * 4);
    if (o0 > 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = dst[4 * i];
        g[4 * i] = t.x; g[4 * i + 1] = t.y; g[4 * i + 2] = t.z; g[4 * i + 3] = t.w;
      }
    }
#pragma unroll
    for (int o = 0; o < O_MAX; ++o) {
      if (o < O) {
        const float d = dv[o] * yv[o] * (1.f - yv[o]);
#pragma unroll
        for (int i = 0; i < 16; ++i) g[i] = fmaf(d, sw[o * 64 + (i >> 2) * 16 + qd * 4 + (i & 3)], g[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[4 * i] = make_float4(g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
  }
}
// one warp per chunk of 32 consecutive pixels: lane l first computes dl of pixel p0 + l (coalesced plane reads), then
// the warp walks the chunk two pixels at a time - half-warp h takes pixel 2j + h, lane owns 4 channels of the
// activation (one 16-byte load; 512 contiguous bytes per warp instruction, 8 of them in flight) and the dl values
// arrive by shuffle; acc[o][4] lives in registers for the whole grid-stride loop
template <int O_MAX>
__global__ void __launch_bounds__(256) predictor_bwd_dw_kernel(SrcDesc src, long long npix, long long hw, int O, int o0,
                                                               int OT, const float* __restrict__ dy,
                                                               const float* __restrict__ y, float* __restrict__ part) {
  // output channels [o0, o0 + O) of OT. part: [gridDim.x][OT * 65] per-block partial sums (dW rows of 64, then db),
  // summed in block order by predictor_dw_final_kernel - no atomics, the gradient is bit-identical from run to run
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, c4 = (lane & 15) * 4;
  const bool affine = src.mode != SRC_IDENTITY;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (affine) {
    sc = *reinterpret_cast<const float4*>(src.scale + c4);
    sh = *reinterpret_cast<const float4*>(src.shift + c4);
  }
  float acc[O_MAX][4], dsum[O_MAX];
#pragma unroll
  for (int o = 0; o < O_MAX; ++o) { acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f; dsum[o] = 0.f; }
  const long long nchunks = (npix + 31) / 32;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  const long long ch0 = blockIdx.x * (long long)(blockDim.x >> 5) + warp;
  long long n = (ch0 * 32 + lane) / hw, r = (ch0 * 32 + lane) - n * hw;  // of pixel p, advanced incrementally
  for (long long ch = ch0; ch < nchunks; ch += wstride, r += wstride * 32) {
    while (r >= hw) { r -= hw; ++n; }
    const long long p0 = ch * 32, p = p0 + lane;
    float d[O_MAX], yv[O_MAX];
#pragma unroll
    for (int o = 0; o < O_MAX; ++o) {
      const bool ok = o < O && p < npix;
      yv[o] = ok ? __ldg(y + (n * OT + o0 + o) * hw + r) : 0.f;
      d[o] = ok ? __ldg(dy + (n * OT + o0 + o) * hw + r) : 0.f;
    }
#pragma unroll
    for (int o = 0; o < O_MAX; ++o) { d[o] = d[o] * yv[o] * (1.f - yv[o]); dsum[o] += d[o]; }
#pragma unroll 8
    for (int j = 0; j < 16; ++j) {
      const int pj = 2 * j + half;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p0 + pj < npix) {
        a = __ldg(reinterpret_cast<const float4*>(src.ptr + (p0 + pj) * 64 + c4));
        if (affine) {
          a.x = fmaxf(fmaf(a.x, sc.x, sh.x), 0.f); a.y = fmaxf(fmaf(a.y, sc.y, sh.y), 0.f);
          a.z = fmaxf(fmaf(a.z, sc.z, sh.z), 0.f); a.w = fmaxf(fmaf(a.w, sc.w, sh.w), 0.f);
        }
      }
#pragma unroll
      for (int o = 0; o < O_MAX; ++o) {
        const float dj = __shfl_sync(0xffffffffu, d[o], pj);  // dl of the pixel this half-warp is on (0 beyond npix)
        acc[o][0] = fmaf(dj, a.x, acc[o][0]); acc[o][1] = fmaf(dj, a.y, acc[o][1]);
        acc[o][2] = fmaf(dj, a.z, acc[o][2]); acc[o][3] = fmaf(dj, a.w, acc[o][3]);
      }
    }
  }
  // fold the two half-warps, then block reduction over the 8 warps, then one partial per (o, channel) and per o
  __shared__ float red[8][O_MAX][64];
  __shared__ float redb[8][O_MAX];
#pragma unroll
  for (int o = 0; o < O_MAX; ++o) {
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[o][k] += __shfl_xor_sync(0xffffffffu, acc[o][k], 16);
    if (half == 0) {
      red[warp][o][c4] = acc[o][0]; red[warp][o][c4 + 1] = acc[o][1];
      red[warp][o][c4 + 2] = acc[o][2]; red[warp][o][c4 + 3] = acc[o][3];
    }
    const float s = warp_sum(dsum[o]);
    if (lane == 0) redb[warp][o] = s;
  }
  __syncthreads();
  float* mine = part + (size_t)blockIdx.x * OT * 65;
  for (int i = threadIdx.x; i < O * 64; i += blockDim.x) {
    const int o = i >> 6, c = i & 63;
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][o][c];
    mine[(o0 + o) * 64 + c] = s;
  }
  if (threadIdx.x < O) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += redb[w][threadIdx.x];
    mine[OT * 64 + o0 + threadIdx.x] = s;
  }
}
// dW[o][c] / db[o] = sum over the blocks' partials in block order (fp64 accumulation, fixed order)
__global__ void __launch_bounds__(256) predictor_dw_final_kernel(const float* __restrict__ part, int nblocks, int OT,
                                                                 float* __restrict__ dwp, float* __restrict__ dbias) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= OT * 65) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += (double)part[(size_t)b * OT * 65 + i];
  if (i < OT * 64) dwp[i] = (float)s; else dbias[i - OT * 64] = (float)s;
}
static int predictor_dw_blocks(long long npix) {
  const long long blocks_w = ((npix + 31) / 32 + 7) / 8;
  return (int)std::min<long long>(blocks_w, 148 * 4);
}
size_t predictor_bwd_workspace_bytes(int N, int H, int W, int O) {
  return sizeof(float) * (size_t)predictor_dw_blocks((long long)N * H * W) * O * 65;
}
int launch_predictor_bwd(const SrcDesc& src, int N, int H, int W, const float* wp, int O, const float* dy,
                         const float* y, float* dA, float* dwp, float* dbias, float* part, cudaStream_t st) {
  TNB_REQUIRE(src.C == 64 && O >= 1, "predictor_bwd: expects 64 channels and out_dim >= 1 (got %d, %d)", src.C, O);
  TNB_REQUIRE((src.mode == SRC_IDENTITY || src.mode == SRC_AFFINE_RELU) && src.Hs == H && src.Ws == W,
              "predictor_bwd: the activation must be a same-resolution identity / BN+ReLU source (mode %d)", src.mode);
  TNB_REQUIRE(part != nullptr, "predictor_bwd: needs a workspace of predictor_bwd_workspace_bytes()");
  const long long npix = (long long)N * H * W, hw = (long long)H * W;
  ProfScope prof(PROF_PRED, st, N, H, W, 64, O);
  const long long blocks_a = (npix * 4 + 255) / 256;
  const int grida = (int)std::min<long long>(blocks_a, 148 * 8);
  const int gridw = predictor_dw_blocks(npix);
  for (int o0 = 0; o0 < O; o0 += kMaxPredO) {  // groups of up to 16 output channels, see launch_predictor_fwd
    const int og = std::min(O - o0, kMaxPredO);
    auto go_a = [&](auto kern) { return launch_pdl(kern, dim3(grida), dim3(256), 0, st, npix, hw, wp + o0 * 64, og, o0, O, dy, y, dA); };
    if (int rc = og == 8 ? go_a(predictor_bwd_da_kernel<8, true>) : og < 8 ? go_a(predictor_bwd_da_kernel<8, false>)
               : og == kMaxPredO ? go_a(predictor_bwd_da_kernel<kMaxPredO, true>) : go_a(predictor_bwd_da_kernel<kMaxPredO, false>))
      return rc;
    if (int rc = og <= 8 ? launch_pdl(predictor_bwd_dw_kernel<8>, dim3(gridw), dim3(256), 0, st, src, npix, hw, og, o0, O, dy, y,
                                      part)
                         : launch_pdl(predictor_bwd_dw_kernel<kMaxPredO>, dim3(gridw), dim3(256), 0, st, src, npix, hw, og, o0,
                                      O, dy, y, part))
      return rc;
  }
  return launch_pdl(predictor_dw_final_kernel, dim3((O * 65 + 255) / 256), dim3(256), 0, st, (const float*)part, gridw, O, dwp,
                    dbias);
}

// =============================================================================================
// BatchNorm + ReLU backward (autograd of model.py:13-15), fused with the gradient routing of
// MaxPool2d / Upsample / cat (autograd of model.py:59-69): each thread owns a 2x2 window x 4 channels,
// gathers dL/da from the consumers' dgrad outputs, masks by ReLU, and either reduces
// (sum dy, sum dy*xhat) or writes dz = scale * (dy - mean(dy) - xhat * mean(dy*xhat)).
// =============================================================================================
TNB_DEVINL float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
TNB_DEVINL float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// CFG 0: one GRAD_SAME consumer; CFG 1: {GRAD_POOL, GRAD_SAME}; CFG 2: anything else (generic gather)
template <bool APPLY, int CFG>
__global__ void __launch_bounds__(256) bn_bwd_kernel(const __grid_constant__ BnBwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int CQ = a.C >> 2;
  const int Hw = (a.H + 1) >> 1, Ww = (a.W + 1) >> 1;
  const unsigned items = (unsigned)a.N * Hw * Ww * CQ;  // < 2^31, checked by the launcher
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned cq_shift = 31 - __clz(CQ);             // CQ is a power of two (256 % CQ == 0)
  float4 acc1 = make_float4(0, 0, 0, 0), acc2 = make_float4(0, 0, 0, 0);
  float amax = 0.f;
  // dz_format 2: dz leaves as an fp16 (hi, lo) pair - 22 bits instead of bf16's 16 - after multiplication by a power of
  // two that brings the tensor into fp16's range. |dz| <= max_c |scale_c| * (|g| + |mean g| + |xhat| |mean g xhat|) with
  // max |g| measured by the reduction pass: the multiplier puts max_c |scale_c| * max |g| into [2^7, 2^8), which leaves
  // a factor 256 of head room for the bracket (values beyond are clamped by the split, never infinite). Every CTA
  // derives the same multiplier; block 0 publishes it for the consumers (dgrad / wgrad divide their accumulators by it).
  float dz_mul = 1.f;
  if (APPLY && a.dz_format == 2) {
    __shared__ float s_scmax[8];
    float m = 0.f;
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) m = fmaxf(m, fabsf(a.scale[c]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_scmax[threadIdx.x >> 5] = m;
    __syncthreads();
    m = 0.f;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, s_scmax[i]);
    const float bound = m * (*a.gmax);
    if (bound > 0.f && bound < 3.0e38f) {
      int e;
      frexpf(bound, &e);
      dz_mul = ldexpf(1.f, max(-100, min(100, 8 - e)));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.dz_mul = dz_mul;
  }
  for (unsigned it0 = blockIdx.x * blockDim.x + threadIdx.x; it0 < items; it0 += stride) {
    const unsigned it = it0;
    const int cq = (int)(it & (CQ - 1));
    unsigned r = it >> cq_shift;
    const int ww = (int)(r % (unsigned)Ww); r /= (unsigned)Ww;
    const int wh = (int)(r % (unsigned)Hw);
    const int n = (int)(r / (unsigned)Hw);
    const int c = cq * 4;
    const float4 sc = ld4(a.scale + c), sh = ld4(a.shift + c), mu = ld4(a.mean + c), is = ld4(a.invstd + c);
    float4 m1 = make_float4(0, 0, 0, 0), m2 = m1;
    if (APPLY) {
      m1 = ld4(a.sums + c);
      m2 = ld4(a.sums + a.C + c);
      m1 = make_float4(m1.x * a.inv_count, m1.y * a.inv_count, m1.z * a.inv_count, m1.w * a.inv_count);
      m2 = make_float4(m2.x * a.inv_count, m2.y * a.inv_count, m2.z * a.inv_count, m2.w * a.inv_count);
    }
    float4 z[4], act[4], dy[4];
    bool valid[4];
    size_t zoff[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int h = 2 * wh + (k >> 1), w = 2 * ww + (k & 1);
      valid[k] = (h < a.H) && (w < a.W);
      zoff[k] = ((size_t)(n * a.H + h) * a.W + w);
      dy[k] = make_float4(0, 0, 0, 0);
      z[k] = make_float4(0, 0, 0, 0);
    }
    float4 gp = make_float4(0, 0, 0, 0);
    bool has_pool = false;
    if (CFG == 0 || CFG == 1) {
      // fast paths (one same-resolution consumer, optionally preceded by a max-pool consumer): every global load of
      // the item is issued before the first use, so one memory latency is paid per item instead of two
      const GradSrc& gs = a.g[CFG == 1 ? 1 : 0];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (valid[k]) z[k] = ld4(a.z + zoff[k] * a.C + c);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (valid[k]) dy[k] = ld4(gs.ptr + zoff[k] * gs.C + gs.coff + c);
      if (CFG == 1) {
        const GradSrc& g = a.g[0];
        has_pool = wh < g.Hs && ww < g.Ws && valid[3];
        if (has_pool) gp = ld4(g.ptr + ((size_t)(n * g.Hs + wh) * g.Ws + ww) * g.C + g.coff + c);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (valid[k]) z[k] = ld4(a.z + zoff[k] * a.C + c);
      for (int gi = 0; gi < a.ng; ++gi) {
        const GradSrc& g = a.g[gi];
        if (g.mode == GRAD_SAME) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int h = 2 * wh + (k >> 1), w = 2 * ww + (k & 1);
            if (valid[k]) dy[k] = f4add(dy[k], ld4(g.ptr + ((size_t)(n * g.Hs + h) * g.Ws + w) * g.C + g.coff + c));
          }
        } else if (g.mode == GRAD_UP) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int h = 2 * wh + (k >> 1), w = 2 * ww + (k & 1);
            if (valid[k]) {
              float4 t[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int hh = 2 * h + (j >> 1), wv = 2 * w + (j & 1);
                t[j] = ld4(g.ptr + ((size_t)(n * g.Hs + hh) * g.Ws + wv) * g.C + g.coff + c);
              }
              dy[k] = f4add(dy[k], f4add(f4add(t[0], t[1]), f4add(t[2], t[3])));
            }
          }
        } else {  // GRAD_POOL
          has_pool = wh < g.Hs && ww < g.Ws && valid[3];
          if (has_pool) gp = ld4(g.ptr + ((size_t)(n * g.Hs + wh) * g.Ws + ww) * g.C + g.coff + c);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      act[k] = make_float4(fmaf(z[k].x, sc.x, sh.x), fmaf(z[k].y, sc.y, sh.y), fmaf(z[k].z, sc.z, sh.z),
                           fmaf(z[k].w, sc.w, sh.w));
    if (has_pool) {  // the consumer saw maxpool2x2(relu(act)); first maximum in window scan order wins
      int ix = 0, iy = 0, iz = 0, iw = 0;
      float bx = act[0].x, by = act[0].y, bz = act[0].z, bw = act[0].w;
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        if (act[k].x > bx) { bx = act[k].x; ix = k; }
        if (act[k].y > by) { by = act[k].y; iy = k; }
        if (act[k].z > bz) { bz = act[k].z; iz = k; }
        if (act[k].w > bw) { bw = act[k].w; iw = k; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (ix == k) dy[k].x += gp.x;
        if (iy == k) dy[k].y += gp.y;
        if (iz == k) dy[k].z += gp.z;
        if (iw == k) dy[k].w += gp.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!valid[k]) continue;
      float4 d = dy[k];
      d.x = act[k].x > 0.f ? d.x : 0.f;
      d.y = act[k].y > 0.f ? d.y : 0.f;
      d.z = act[k].z > 0.f ? d.z : 0.f;
      d.w = act[k].w > 0.f ? d.w : 0.f;
      const float4 xh = make_float4((z[k].x - mu.x) * is.x, (z[k].y - mu.y) * is.y, (z[k].z - mu.z) * is.z,
                                    (z[k].w - mu.w) * is.w);
      if (APPLY) {
        const int h = 2 * wh + (k >> 1), w = 2 * ww + (k & 1);
        float4 o;
        o.x = sc.x * (d.x - m1.x - xh.x * m2.x);
        o.y = sc.y * (d.y - m1.y - xh.y * m2.y);
        o.z = sc.z * (d.z - m1.z - xh.z * m2.z);
        o.w = sc.w * (d.w - m1.w - xh.w * m2.w);
        const size_t pix = (size_t)(n * a.H + h) * a.W + w;
        if (a.dz_format == 0) {
          *reinterpret_cast<float4*>(a.dz + pix * a.C + c) = o;
        } else {
          // pre-split: [pixel][2 (hi, lo)][C]; this thread owns channels c..c+3 (8 bytes of each term)
          uint32_t h01, l01, h23, l23;
          if (a.dz_format == 2) {
            split2<0>(o.x * dz_mul, o.y * dz_mul, h01, l01);
            split2<0>(o.z * dz_mul, o.w * dz_mul, h23, l23);
          } else {
            split2<1>(o.x, o.y, h01, l01);
            split2<1>(o.z, o.w, h23, l23);
          }
          uint8_t* base = reinterpret_cast<uint8_t*>(a.dz) + (pix * 2 * a.C + c) * 2;
          *reinterpret_cast<uint2*>(base) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(base + 2 * a.C) = make_uint2(l01, l23);
        }
        void* const act_out = (a.act_presplit != nullptr && !a.act_pool) ? a.act_presplit : a.act_full;
        if (act_out != nullptr) {  // the activation itself, pre-split: a later layer's wgrad operand
          uint32_t h01, l01, h23, l23;
          if (a.dz_format == 2) {
            split2<0>(fmaxf(act[k].x, 0.f), fmaxf(act[k].y, 0.f), h01, l01);
            split2<0>(fmaxf(act[k].z, 0.f), fmaxf(act[k].w, 0.f), h23, l23);
          } else {
          split2<1>(fmaxf(act[k].x, 0.f), fmaxf(act[k].y, 0.f), h01, l01);
          split2<1>(fmaxf(act[k].z, 0.f), fmaxf(act[k].w, 0.f), h23, l23);
          }
          uint8_t* base = reinterpret_cast<uint8_t*>(act_out) + (pix * 2 * a.C + c) * 2;
          *reinterpret_cast<uint2*>(base) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(base + 2 * a.C) = make_uint2(l01, l23);
        }
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
      } else {
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));  // max |g|
        acc1 = f4add(acc1, d);
        acc2.x = fmaf(d.x, xh.x, acc2.x);
        acc2.y = fmaf(d.y, xh.y, acc2.y);
        acc2.z = fmaf(d.z, xh.z, acc2.z);
        acc2.w = fmaf(d.w, xh.w, acc2.w);
      }
    }
    if (APPLY && a.act_presplit != nullptr && a.act_pool) {
      // the 2x2 window this thread holds IS a max-pool window: maxpool(relu(bn(z))) of the next layer's input view,
      // pre-split, one pixel of the half-resolution tensor (H, W even: all four elements are valid)
      float4 mx = make_float4(0.f, 0.f, 0.f, 0.f);  // relu: the maximum is at least 0
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mx.x = fmaxf(mx.x, act[k].x); mx.y = fmaxf(mx.y, act[k].y);
        mx.z = fmaxf(mx.z, act[k].z); mx.w = fmaxf(mx.w, act[k].w);
      }
      uint32_t h01, l01, h23, l23;
      if (a.dz_format == 2) { split2<0>(mx.x, mx.y, h01, l01); split2<0>(mx.z, mx.w, h23, l23); }
      else                  { split2<1>(mx.x, mx.y, h01, l01); split2<1>(mx.z, mx.w, h23, l23); }
      const size_t ppix = (size_t)(n * (a.H >> 1) + wh) * (a.W >> 1) + ww;
      uint8_t* base = reinterpret_cast<uint8_t*>(a.act_presplit) + (ppix * 2 * a.C + c) * 2;
      *reinterpret_cast<uint2*>(base) = make_uint2(h01, h23);
      *reinterpret_cast<uint2*>(base + 2 * a.C) = make_uint2(l01, l23);
    }
  }
  if (APPLY && a.amax != nullptr) {
    amax = warp_max(amax);  // non-negative floats order like their bit patterns
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(reinterpret_cast<int*>(a.amax), __float_as_int(amax));
  }
  if (!APPLY && a.gmax != nullptr) {  // max |g| over the tensor (a maximum: the order of the atomics does not matter)
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(reinterpret_cast<int*>(a.gmax), __float_as_int(amax));
  }
  if (!APPLY) {
    // every thread keeps one channel quad for its whole grid-stride loop (256 % CQ == 0, stride % CQ == 0)
    __shared__ float red[8][256];
    red[0][threadIdx.x] = acc1.x; red[1][threadIdx.x] = acc1.y; red[2][threadIdx.x] = acc1.z; red[3][threadIdx.x] = acc1.w;
    red[4][threadIdx.x] = acc2.x; red[5][threadIdx.x] = acc2.y; red[6][threadIdx.x] = acc2.z; red[7][threadIdx.x] = acc2.w;
    __syncthreads();
    for (int j = threadIdx.x; j < 8 * CQ; j += 256) {
      const int comp = j / CQ, cq = j - comp * CQ;
      float s = 0.f;
      for (int t = cq; t < 256; t += CQ) s += red[comp][t];
      const int which = comp >> 2, cc = comp & 3;
      a.part[((size_t)blockIdx.x * 2 + which) * a.C + cq * 4 + cc] = s;
    }
  }
}

int bn_bwd_num_blocks(int N, int H, int W, int C) {
  const long long items = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
  return min(cdiv(items, 256), 148 * 8);
}
static int bn_bwd_cfg(const BnBwdArgs& a) {
  const bool same_res0 = a.ng >= 1 && a.g[0].Hs == a.H && a.g[0].Ws == a.W;
  if (a.ng == 1 && a.g[0].mode == GRAD_SAME && same_res0) return 0;
  if (a.ng == 2 && a.g[0].mode == GRAD_POOL && a.g[1].mode == GRAD_SAME && a.g[1].Hs == a.H && a.g[1].Ws == a.W) return 1;
  return 2;
}
static int bn_bwd_check(const BnBwdArgs& a) {
  TNB_REQUIRE(a.C % 4 == 0 && 256 % (a.C / 4) == 0, "bn_bwd: unsupported channel count %d", a.C);
  TNB_REQUIRE(a.act_full == nullptr || a.act_presplit == nullptr || a.act_pool,
              "bn_bwd: act_full is the second output next to a POOLED act_presplit (use act_presplit alone otherwise)");
  TNB_REQUIRE(!(a.act_presplit != nullptr && a.act_pool) || (a.H % 2 == 0 && a.W % 2 == 0),
              "bn_bwd: the pooled activation needs even H, W (got %d x %d)", a.H, a.W);
  TNB_REQUIRE(a.dz_format >= 0 && a.dz_format <= 2 && (a.dz_format != 2 || (a.gmax != nullptr && a.dz_mul != nullptr)),
              "bn_bwd: dz_format %d (2 needs the gmax / dz_mul scalars)", a.dz_format);
  TNB_REQUIRE((long long)a.N * ((a.H + 1) / 2) * ((a.W + 1) / 2) * (a.C / 4) < (1ll << 31), "bn_bwd: tensor too large");
  return 0;
}
int launch_bn_bwd_reduce(const BnBwdArgs& a, cudaStream_t st) {
  if (int rc = bn_bwd_check(a)) return rc;
  ProfScope prof(PROF_BN_BWD, st, a.N, a.H, a.W, a.C, a.C);
  const int nb = bn_bwd_num_blocks(a.N, a.H, a.W, a.C);
  switch (bn_bwd_cfg(a)) {
    case 0: return launch_pdl(bn_bwd_kernel<false, 0>, dim3(nb), dim3(256), 0, st, a);
    case 1: return launch_pdl(bn_bwd_kernel<false, 1>, dim3(nb), dim3(256), 0, st, a);
    default: return launch_pdl(bn_bwd_kernel<false, 2>, dim3(nb), dim3(256), 0, st, a);
  }
}
int launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t st) {
  if (int rc = bn_bwd_check(a)) return rc;
  ProfScope prof(PROF_BN_BWD, st, a.N, a.H, a.W, a.C, a.C);
  const int nb = bn_bwd_num_blocks(a.N, a.H, a.W, a.C);
  switch (bn_bwd_cfg(a)) {
    case 0: return launch_pdl(bn_bwd_kernel<true, 0>, dim3(nb), dim3(256), 0, st, a);
    case 1: return launch_pdl(bn_bwd_kernel<true, 1>, dim3(nb), dim3(256), 0, st, a);
    default: return launch_pdl(bn_bwd_kernel<true, 2>, dim3(nb), dim3(256), 0, st, a);
  }
}
__global__ void __launch_bounds__(kFinCh * kFinLanes) bn_bwd_finalize_kernel(const float* __restrict__ part, int rows, int C,
                                                               float* sums, float* dgamma, float* dbeta) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ FinalizeSmem sm;
  const int c = blockIdx.x * kFinCh + threadIdx.x;
  double sa, sb;
  finalize_column_sums(part, rows, C, c, c < C, sm, sa, sb);
  if (threadIdx.y == 0 && c < C) {
    sums[c] = (float)sa;
    sums[C + c] = (float)sb;
    dbeta[c] = (float)sa;   // d/dbeta  = sum dy
    dgamma[c] = (float)sb;  // d/dgamma = sum dy * xhat
  }
}
int launch_bn_bwd_finalize(const float* part, int rows, int C, float* sums, float* dgamma, float* dbeta,
                           cudaStream_t st) {
  if (int rc = launch_pdl(bn_bwd_finalize_kernel, dim3(cdiv(C, kFinCh)), dim3(kFinCh, kFinLanes), 0, st, part, rows, C, sums, dgamma, dbeta)) return rc;
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// fp32 NHWC -> pre-split [pixel][2 (hi, lo)][C] (see include/tracknet_b200.h): bf16, or fp16 after multiplication by mul
__global__ void presplit_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, long long nchunks, int C, int fmt,
                                float mul) {
  const int nch = C >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nchunks;
       i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    ld8(x + i * 8, v);
    uint4 hi, lo;
    if (fmt == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= mul;
      split8<0>(v, hi, lo);
    } else {
      split8<1>(v, hi, lo);
    }
    const long long pix = i / nch;
    const int c = (int)(i - pix * nch) * 8;
    uint8_t* dst = out + ((size_t)pix * 2 * C + c) * 2;  // [pixel][2 (hi, lo)][C] 16-bit
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + 2 * C) = lo;
  }
}
// Materialise a layer's logical input view (BN affine + ReLU + MaxPool / Upsample / cat of its producers, exactly what
// the conv kernels gather on the fly) ONCE in the pre-split 16-bit format, so that kernels which would otherwise each
// redo that arithmetic several times (wgrad: 3 filter rows x Cout/128 tiles) fill their operands with plain copies.
template <int FMT>
__global__ void __launch_bounds__(256) view_presplit_kernel(const __grid_constant__ ViewDesc V, uint8_t* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int nch = V.C >> 3;
  const long long total = (long long)V.N * V.H * V.W * nch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % nch);
    const long long pix = i / nch;
    const int w = (int)(pix % V.W);
    const long long r = pix / V.W;
    const int h = (int)(r % V.H), n = (int)(r / V.H);
    const int c = ch * 8;
    float v[8];
    if (c >= V.C0) view_load8(V.s[1], view_pix_off(V.s[1], n, h, w), c - V.C0, v);
    else           view_load8(V.s[0], view_pix_off(V.s[0], n, h, w), c, v);
    uint4 hi, lo;
    split8<FMT>(v, hi, lo);
    uint8_t* dst = out + ((size_t)pix * 2 * V.C + c) * 2;  // [pixel][2 (hi, lo)][C] 16-bit
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + 2 * V.C) = lo;
  }
}
int launch_view_presplit(const ViewDesc& view, void* out, int fmt, cudaStream_t st) {
  TNB_REQUIRE(view.C % 8 == 0 && view.C0 % 8 == 0, "view_presplit: channel counts must be multiples of 8");
  const long long total = (long long)view.N * view.H * view.W * (view.C / 8);
  const int blocks = min(cdiv(total, 256), 148 * 16);
  if (int rc = fmt == 0 ? launch_pdl(view_presplit_kernel<0>, dim3(blocks), dim3(256), 0, st, view, (uint8_t*)out)
                        : launch_pdl(view_presplit_kernel<1>, dim3(blocks), dim3(256), 0, st, view, (uint8_t*)out))
    return rc;
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_presplit(const float* x, void* out, long long npixels, int C, int fmt, float mul, cudaStream_t st) {
  TNB_REQUIRE(fmt == 0 || fmt == 1, "presplit: format %d (0 = fp16, 1 = bf16)", fmt);
  TNB_REQUIRE(C % 8 == 0, "presplit: channels %d must be a multiple of 8", C);
  const long long nchunks = npixels * (C / 8);
  presplit_kernel<<<min(cdiv(nchunks, 256), 148 * 16), 256, 0, st>>>(x, (uint8_t*)out, nchunks, C, fmt, mul);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// WBCE / focal loss (reference utils/metric.py:3-20) forward + backward, one pass each.
// =============================================================================================
TNB_DEVINL float wbce_elem(float p, float y) {
  const float lp = logf(fminf(fmaxf(p, 1e-7f), 1.f));
  const float lq = logf(fminf(fmaxf(1.f - p, 1e-7f), 1.f));
  return -((1.f - p) * (1.f - p) * y * lp + p * p * (1.f - y) * lq);
}
TNB_DEVINL float wbce_grad(float p, float y) {
  const float q = 1.f - p;
  const float cp = fminf(fmaxf(p, 1e-7f), 1.f), cq = fminf(fmaxf(q, 1e-7f), 1.f);
  const float lp = logf(cp), lq = logf(cq);
  const float gp = (p >= 1e-7f && p <= 1.f) ? 1.f / cp : 0.f;   // d clamp(p)/dp, closed interval like torch
  const float gq = (q >= 1e-7f && q <= 1.f) ? -1.f / cq : 0.f;  // d log(clamp(1-p))/dp
  const float t1 = -2.f * q * y * lp + q * q * y * gp;
  const float t2 = 2.f * p * (1.f - y) * lq + p * p * (1.f - y) * gq;
  return -(t1 + t2);
}
static constexpr int kWbceBlocks = 592;  // 148 SMs x 4
int wbce_num_blocks(long long) { return kWbceBlocks; }

__global__ void __launch_bounds__(256) wbce_fwd_kernel(const float* __restrict__ p, const float* __restrict__ y,
                                                       long long per_sample, double* part) {
  pdl_launch_dependents();
  pdl_wait();
  const float* ps = p + (size_t)blockIdx.y * per_sample;
  const float* ys = y + (size_t)blockIdx.y * per_sample;
  float acc = 0.f;
  double dacc = 0.0;
  int cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample;
       i += (long long)gridDim.x * blockDim.x) {
    acc += wbce_elem(ps[i], ys[i]);
    if (++cnt == 64) { dacc += (double)acc; acc = 0.f; cnt = 0; }
  }
  dacc += (double)acc;
  __shared__ double sd[256];
  sd[threadIdx.x] = dacc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sd[threadIdx.x] += sd[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = sd[0];
}
__global__ void __launch_bounds__(256) wbce_final_kernel(const double* part, int nsamples, int nblocks,
                                                         long long per_sample, int reduce, float* out) {
  pdl_launch_dependents();
  pdl_wait();
  // one block; fixed summation order => deterministic. sample sums first, then the batch total.
  __shared__ double sd[256];
  double total = 0.0;
  for (int n = 0; n < nsamples; ++n) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) s += part[(size_t)n * nblocks + b];
    sd[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
      if (threadIdx.x < k) sd[threadIdx.x] += sd[threadIdx.x + k];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (!reduce) out[n] = (float)(sd[0] / (double)per_sample);
      total += sd[0];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && reduce) out[0] = (float)(total / ((double)per_sample * nsamples));
}
int launch_wbce_fwd(const float* p, const float* y, int nsamples, long long per_sample, int reduce, double* part,
                    float* out, cudaStream_t st) {
  if (int rc = launch_pdl(wbce_fwd_kernel, dim3(kWbceBlocks, nsamples), dim3(256), 0, st, p, y, per_sample, part)) return rc;
  TNB_CHECK_CUDA(cudaGetLastError());
  if (int rc = launch_pdl(wbce_final_kernel, dim3(1), dim3(256), 0, st, (const double*)part, nsamples, kWbceBlocks, per_sample, reduce, out)) return rc;
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
__global__ void __launch_bounds__(256) wbce_bwd_kernel(const float* __restrict__ p, const float* __restrict__ y,
                                                       const float* __restrict__ gout, long long per_sample,
                                                       int nsamples, int reduce, float* __restrict__ dp) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = per_sample * nsamples;
  const float inv = reduce ? (float)(1.0 / (double)total) : (float)(1.0 / (double)per_sample);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float g = reduce ? gout[0] : gout[i / per_sample];
    dp[i] = g * inv * wbce_grad(p[i], y[i]);
  }
}
int launch_wbce_bwd(const float* p, const float* y, const float* gout, int nsamples, long long per_sample,
                    int reduce, float* dp, cudaStream_t st) {
  if (int rc = launch_pdl(wbce_bwd_kernel, dim3(148 * 8), dim3(256), 0, st, p, y, gout, per_sample, nsamples, reduce, dp)) return rc;
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// sample mixup (reference train.py:37-38): out[n] = x[n]*lam[n] + x[perm[n]]*(1-lam[n])
// =============================================================================================
__global__ void mixup_kernel(const float* __restrict__ x, const float* __restrict__ lam,
                             const long long* __restrict__ perm, float* __restrict__ out, long long per_sample) {
  const int n = blockIdx.y;
  const float l = lam[n];
  const float* a = x + (size_t)n * per_sample;
  const float* b = x + (size_t)perm[n] * per_sample;
  float* o = out + (size_t)n * per_sample;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample;
       i += (long long)gridDim.x * blockDim.x)
    o[i] = a[i] * l + b[i] * (1.f - l);
}
int launch_mixup(const float* x, const float* lam, const long long* perm, float* out, int N, long long per_sample,
                 cudaStream_t st) {
  mixup_kernel<<<dim3(min(cdiv(per_sample, 1024), 148 * 2), N), 256, 0, st>>>(x, lam, perm, out, per_sample);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// multi-tensor Adam (reference train.py:242 torch.optim.Adam(lr), betas (0.9,0.999), eps 1e-8):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// =============================================================================================
// One flat work list over all tensors: a CTA owns one chunk of kAdamChunk consecutive elements of one tensor (found by a
// scan over the table's element counts: 53 entries for TrackNet), 128-bit loads / stores when the tensor's four pointers
// allow. A grid per tensor left the few multi-million-element tensors to 64 CTAs each (0.14 ms for 317 MB).
static constexpr int kAdamChunk = 4096;
__global__ void __launch_bounds__(256) adam_kernel(const AdamTensor* __restrict__ tab, int ntensors, float lr, float b1,
                                                   float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  __shared__ int s_tensor;
  __shared__ long long s_first;
  if (threadIdx.x == 0) {
    long long chunk = blockIdx.x;
    int ti = -1;
    for (int i = 0; i < ntensors; ++i) {
      const long long c = (tab[i].n + kAdamChunk - 1) / kAdamChunk;
      if (chunk < c) { ti = i; break; }
      chunk -= c;
    }
    s_tensor = ti;
    s_first = chunk * kAdamChunk;
  }
  __syncthreads();
  if (s_tensor < 0) return;
  const AdamTensor t = tab[s_tensor];
  const long long i0 = s_first, i1 = min(t.n, s_first + kAdamChunk);
  const float step_size = lr / bc1;
  auto update = [&](float g, float p, float& m, float& v) -> float {
    if (wd != 0.f) g = fmaf(wd, p, g);
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    return p - step_size * (m / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                     reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
  auto scalar = [&](long long j) {
    float m = t.m[j], v = t.v[j];
    t.p[j] = update(t.g[j], t.p[j], m, v);
    t.m[j] = m; t.v[j] = v;
  };
  if (vec) {
    for (long long i = i0 + threadIdx.x * 4; i + 4 <= i1; i += 256 * 4) {
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      float4 p = *reinterpret_cast<const float4*>(t.p + i);
      float4 m = *reinterpret_cast<const float4*>(t.m + i), v = *reinterpret_cast<const float4*>(t.v + i);
      p.x = update(g.x, p.x, m.x, v.x); p.y = update(g.y, p.y, m.y, v.y);
      p.z = update(g.z, p.z, m.z, v.z); p.w = update(g.w, p.w, m.w, v.w);
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
      *reinterpret_cast<float4*>(t.p + i) = p;
    }
    const long long t0 = i0 + ((i1 - i0) & ~3LL);  // the last chunk of a tensor whose size is not a multiple of 4
    if (threadIdx.x < i1 - t0) scalar(t0 + threadIdx.x);
  } else {
    for (long long j = i0 + threadIdx.x; j < i1; j += 256) scalar(j);
  }
}
int launch_adam(const AdamTensor* tab, int ntensors, long long total_n, float lr, float b1, float b2, float eps,
                float wd, int step, cudaStream_t st) {
  const float bc1 = (float)(1.0 - pow((double)b1, (double)step));
  const float bc2 = (float)(1.0 - pow((double)b2, (double)step));
  // chunks <= total_n / chunk + one partial chunk per tensor
  const long long chunks = total_n / kAdamChunk + ntensors;
  adam_kernel<<<(unsigned)chunks, 256, 0, st>>>(tab, ntensors, lr, b1, b2, eps, wd, bc1, sqrtf(bc2));
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// heatmap -> bounding box decode (reference test.py:52-79 predict_location on predict.py:35's y_pred > 0.5).
// One CTA per heatmap: threshold, 8-connected union-find labelling (root = raster-first pixel of each
// component), per-component bounding boxes, then the reference's selection rule: maximum bbox area
// w*h, ties resolved towards the component OpenCV lists first, i.e. the one whose raster-first pixel
// comes LAST (OpenCV returns external contours in reverse raster order of their start pixel).
// Components nested inside a hole of another component are dropped by RETR_EXTERNAL, but their bbox is
// strictly inside the enclosing one, so they can never win or tie the maximum.
// =============================================================================================
// parents always point to a smaller (raster-earlier) index, so concurrent path halving only ever shortens chains
TNB_DEVINL int uf_find(volatile int* parent, int x) {
  int p = parent[x];
  while (p != x) {
    const int g = parent[p];
    if (g != p) parent[x] = g;  // halve the path
    x = p; p = g;
  }
  return x;
}
TNB_DEVINL void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    const int old = atomicMin(parent + a, b);
    if (old == a) return;
    a = old;
  }
}
template <bool U8>
__global__ void __launch_bounds__(512) decode_kernel(const void* __restrict__ maps, float thresh, int H, int W,
                                                     int* ws, int* out) {
  const int HW = H * W;
  int* parent = ws + (size_t)blockIdx.x * 5 * HW;
  int* minx = parent + HW; int* maxx = minx + HW; int* miny = maxx + HW; int* maxy = miny + HW;
  const float* mf = reinterpret_cast<const float*>(maps) + (size_t)blockIdx.x * HW;
  const unsigned char* mu = reinterpret_cast<const unsigned char*>(maps) + (size_t)blockIdx.x * HW;
  __shared__ int s_any;
  __shared__ unsigned long long s_best[16];
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  int any = 0;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const bool fg = U8 ? (mu[p] != 0) : (mf[p] > thresh);
    parent[p] = fg ? p : -1;
    if (fg) { any = 1; minx[p] = INT_MAX; maxx[p] = -1; miny[p] = INT_MAX; maxy[p] = -1; }
  }
  if (any) s_any = 1;
  __syncthreads();
  if (!s_any) {
    if (threadIdx.x < 4) out[blockIdx.x * 4 + threadIdx.x] = 0;
    return;
  }
  // 8-connectivity with the minimal set of unions per pixel: the already-visited neighbours are W, NW, N, NE. N is
  // adjacent to all the others, so when N is foreground one union suffices; otherwise NE and one of {NW, W} (NW and W
  // are adjacent to each other) cover every case.
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    if (parent[p] < 0) continue;
    const int y = p / W, x = p - y * W;
    volatile int* vp = parent;
    const bool fw = x > 0 && vp[p - 1] >= 0;
    if (y == 0) {
      if (fw) uf_union(parent, p, p - 1);
      continue;
    }
    if (vp[p - W] >= 0) { uf_union(parent, p, p - W); continue; }
    if (x + 1 < W && vp[p - W + 1] >= 0) uf_union(parent, p, p - W + 1);
    if (x > 0 && vp[p - W - 1] >= 0) uf_union(parent, p, p - W - 1);
    else if (fw) uf_union(parent, p, p - 1);
  }
  __syncthreads();
  // bounding boxes from the horizontal runs: only a run's first pixel can lower min x, only its last can raise max x,
  // and one pixel per run suffices for the rows
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    if (parent[p] < 0) continue;
    const int y = p / W, x = p - y * W;
    const bool first = x == 0 || parent[p - 1] < 0, last = x + 1 == W || parent[p + 1] < 0;
    if (!first && !last) continue;
    const int r = uf_find(parent, p);
    if (first) { atomicMin(minx + r, x); atomicMin(miny + r, y); atomicMax(maxy + r, y); }
    if (last) atomicMax(maxx + r, x);
  }
  __syncthreads();
  unsigned long long best = 0ull;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    if (((volatile int*)parent)[p] != p) continue;
    const unsigned long long area = (unsigned long long)(maxx[p] - minx[p] + 1) * (unsigned)(maxy[p] - miny[p] + 1);
    const unsigned long long key = (area << 32) | (unsigned)p;
    best = key > best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) best = s_best[i] > best ? s_best[i] : best;
    const int r = (int)(best & 0xffffffffu);
    out[blockIdx.x * 4 + 0] = minx[r];
    out[blockIdx.x * 4 + 1] = miny[r];
    out[blockIdx.x * 4 + 2] = maxx[r] - minx[r] + 1;
    out[blockIdx.x * 4 + 3] = maxy[r] - miny[r] + 1;
  }
}
size_t decode_workspace_bytes(int nmaps, int H, int W) { return (size_t)nmaps * 5 * H * W * sizeof(int); }
int launch_decode(const void* maps, int is_u8, float thresh, int nmaps, int H, int W, void* ws, int* out,
                  cudaStream_t st) {
  if (nmaps == 0) return 0;
  if (is_u8) decode_kernel<true><<<nmaps, 512, 0, st>>>(maps, thresh, H, W, (int*)ws, out);
  else       decode_kernel<false><<<nmaps, 512, 0, st>>>(maps, thresh, H, W, (int*)ws, out);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// InpaintNet forward (reference model.py:100-129): the whole 1-D U-Net in ONE kernel, one CTA per
// trajectory, activations resident in shared memory (the reference launches ~30 tiny kernels).
// =============================================================================================
TNB_DEVINL void conv1d_k3(float* out, int Cout, const float* inA, int CA, const float* inB, int CB,
                          const float* __restrict__ w, const float* __restrict__ b, int L, int act) {
  const int Cin = CA + CB;
  for (int idx = threadIdx.x; idx < Cout * L; idx += blockDim.x) {
    const int l = idx % L, co = idx / L;
    const float* wr = w + (size_t)co * Cin * 3;
    float s = b[co];
#pragma unroll 4
    for (int ci = 0; ci < CA; ++ci) {
      const float* row = inA + ci * L;
      const float w0 = __ldg(wr + ci * 3), w1 = __ldg(wr + ci * 3 + 1), w2 = __ldg(wr + ci * 3 + 2);
      const float xm = l > 0 ? row[l - 1] : 0.f, xp = l + 1 < L ? row[l + 1] : 0.f;
      s = fmaf(w0, xm, s); s = fmaf(w1, row[l], s); s = fmaf(w2, xp, s);
    }
#pragma unroll 4
    for (int ci = 0; ci < CB; ++ci) {
      const float* row = inB + ci * L;
      const float* wq = wr + (size_t)(CA + ci) * 3;
      const float w0 = __ldg(wq), w1 = __ldg(wq + 1), w2 = __ldg(wq + 2);
      const float xm = l > 0 ? row[l - 1] : 0.f, xp = l + 1 < L ? row[l + 1] : 0.f;
      s = fmaf(w0, xm, s); s = fmaf(w1, row[l], s); s = fmaf(w2, xp, s);
    }
    if (act == 1) s = s > 0.f ? s : 0.01f * s;          // LeakyReLU(0.01), model.py:81
    else if (act == 2) s = 1.f / (1.f + expf(-s));      // sigmoid, model.py:127
    out[idx] = s;
  }
  __syncthreads();
}
__global__ void __launch_bounds__(1024) inpaint_fwd_kernel(const float* __restrict__ coords,
                                                          const float* __restrict__ mask, InpaintParams P, int L,
                                                          float* __restrict__ out, float coor_th) {
  extern __shared__ float sm[];
  float* in0 = sm;            // [3][L]   cat(x, m) permuted (model.py:114-115)
  float* x1 = in0 + 3 * L;    // [32][L]
  float* x2 = x1 + 32 * L;    // [64][L]
  float* x3 = x2 + 64 * L;    // [128][L]
  float* t1 = x3 + 128 * L;   // [256][L]
  float* t2 = t1 + 256 * L;   // [256][L]
  float* u1 = t2 + 256 * L;   // [128][L]
  float* u2 = u1 + 128 * L;   // [64][L]
  float* u3 = u2 + 64 * L;    // [32][L]
  float* o = u3 + 32 * L;     // [2][L]
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    in0[0 * L + i] = coords[((size_t)n * L + i) * 2 + 0];
    in0[1 * L + i] = coords[((size_t)n * L + i) * 2 + 1];
    in0[2 * L + i] = mask[(size_t)n * L + i];
  }
  __syncthreads();
  conv1d_k3(x1, 32, in0, 3, nullptr, 0, P.w[0], P.b[0], L, 1);
  conv1d_k3(x2, 64, x1, 32, nullptr, 0, P.w[1], P.b[1], L, 1);
  conv1d_k3(x3, 128, x2, 64, nullptr, 0, P.w[2], P.b[2], L, 1);
  conv1d_k3(t1, 256, x3, 128, nullptr, 0, P.w[3], P.b[3], L, 1);
  conv1d_k3(t2, 256, t1, 256, nullptr, 0, P.w[4], P.b[4], L, 1);
  conv1d_k3(u1, 128, t2, 256, x3, 128, P.w[5], P.b[5], L, 1);  // cat([x, x3]) model.py:120
  conv1d_k3(u2, 64, u1, 128, x2, 64, P.w[6], P.b[6], L, 1);    // cat([x, x2]) model.py:122
  conv1d_k3(u3, 32, u2, 64, x1, 32, P.w[7], P.b[7], L, 1);     // cat([x, x1]) model.py:124
  conv1d_k3(o, 2, u3, 32, nullptr, 0, P.w[8], P.b[8], L, 2);
  if (coor_th < 0.f) {
    for (int i = threadIdx.x; i < L * 2; i += blockDim.x) {
      const int l = i >> 1, c = i & 1;
      out[((size_t)n * L + l) * 2 + c] = o[c * L + l];
    }
    return;
  }
  // Rectification epilogue (predict.py:256-261 = test.py:401,406-408), one thread per point:
  // coor_inpaint * mask + coor_pred * (1 - mask) with torch's operation order (every product and the sum rounded to fp32,
  // no fused multiply-add), then both coordinates -> 0 where both are below COOR_TH.
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float m = in0[2 * L + l], om = __fsub_rn(1.f, m);
    const float x = __fadd_rn(__fmul_rn(o[l], m), __fmul_rn(in0[l], om));
    const float y = __fadd_rn(__fmul_rn(o[L + l], m), __fmul_rn(in0[L + l], om));
    const bool zero = x < coor_th && y < coor_th;
    *reinterpret_cast<float2*>(out + ((size_t)n * L + l) * 2) = zero ? make_float2(0.f, 0.f) : make_float2(x, y);
  }
}
int launch_inpaint_fwd(const float* coords, const float* mask, const InpaintParams& p, int N, int L, float* out,
                       float coor_th, cudaStream_t st) {
  const size_t smem = (size_t)(3 + 32 + 64 + 128 + 256 + 256 + 128 + 64 + 32 + 2) * L * sizeof(float);
  TNB_REQUIRE(smem <= 227 * 1024, "inpaint_fwd: sequence length %d too long for the fused kernel", L);
  TNB_CHECK_CUDA(cudaFuncSetAttribute(inpaint_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // latency-bound: one CTA per trajectory, as many warps as the SM takes to hide the L2 latency of the weight reads
  inpaint_fwd_kernel<<<N, 1024, smem, st>>>(coords, mask, p, L, out, coor_th);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// =============================================================================================
// InpaintNet backward (autograd of reference model.py:113-129 as used by train.py:147-166): ONE kernel,
// one CTA per trajectory. The forward is recomputed into shared memory (nothing is saved between the
// two launches), then every layer's activation gradient, weight gradient and bias gradient is formed
// from the shared-memory activations; parameter gradients are summed over trajectories with
// red.global.add.f32 into caller-zeroed buffers.
// =============================================================================================
TNB_DEVINL void act_bwd(float* g, const float* a, int n, int act) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = a[i];
    g[i] *= act == 1 ? (v > 0.f ? 1.f : 0.01f) : v * (1.f - v);  // LeakyReLU(0.01) / sigmoid, from the outputs
  }
  __syncthreads();
}
// dW[co][ci][k] += sum_l dz[co][l] * in[ci][l+k-1];  db[co] += sum_l dz[co][l]
TNB_DEVINL void conv1d_k3_wgrad(const float* dz, int Cout, const float* inA, int CA, const float* inB, int CB,
                                float* __restrict__ dw, float* __restrict__ db, int L) {
  const int Cin = CA + CB, total = Cout * Cin * 3;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int co = idx / (Cin * 3), r = idx - co * Cin * 3, ci = r / 3, k = r - ci * 3;
    const float* row = ci < CA ? inA + ci * L : inB + (ci - CA) * L;
    const float* g = dz + co * L;
    const int lo = k == 0 ? 1 : 0, hi = k == 2 ? L - 1 : L;
    float s = 0.f;
    for (int l = lo; l < hi; ++l) s = fmaf(g[l], row[l + k - 1], s);
    atomicAdd(dw + idx, s);
  }
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += dz[co * L + l];
    atomicAdd(db + co, s);
  }
}
// dIn[ci][l] (+)= sum_co sum_k dz[co][l+1-k] * w[co][ci_off+ci][k] for the C channels starting at ci_off
TNB_DEVINL void conv1d_k3_dgrad(const float* dz, int Cout, const float* __restrict__ w, int Cin, int ci_off, int C,
                                float* dIn, int L, bool accumulate) {
  for (int idx = threadIdx.x; idx < C * L; idx += blockDim.x) {
    const int ci = idx / L, l = idx - ci * L;
    const float* wq = w + (size_t)(ci_off + ci) * 3;
    float s = 0.f;
    for (int co = 0; co < Cout; ++co) {
      const float* g = dz + co * L;
      const float* wr = wq + (size_t)co * Cin * 3;
      const float gp = l + 1 < L ? g[l + 1] : 0.f, gm = l > 0 ? g[l - 1] : 0.f;
      s = fmaf(gp, __ldg(wr), s); s = fmaf(g[l], __ldg(wr + 1), s); s = fmaf(gm, __ldg(wr + 2), s);
    }
    dIn[idx] = accumulate ? dIn[idx] + s : s;
  }
}
__global__ void __launch_bounds__(1024) inpaint_bwd_kernel(const float* __restrict__ coords,
                                                          const float* __restrict__ mask, InpaintParams P,
                                                          const float* __restrict__ dout, InpaintGrads G, int L,
                                                          float* __restrict__ dcoords) {
  extern __shared__ float sm[];
  float* in0 = sm;            float* x1 = in0 + 3 * L;   float* x2 = x1 + 32 * L;  float* x3 = x2 + 64 * L;
  float* t1 = x3 + 128 * L;   float* t2 = t1 + 256 * L;  float* u1 = t2 + 256 * L;  float* u2 = u1 + 128 * L;
  float* u3 = u2 + 64 * L;    float* o = u3 + 32 * L;
  float* g_in0 = o + 2 * L;   float* g_x1 = g_in0 + 3 * L; float* g_x2 = g_x1 + 32 * L; float* g_x3 = g_x2 + 64 * L;
  float* g_t1 = g_x3 + 128 * L; float* g_t2 = g_t1 + 256 * L; float* g_u1 = g_t2 + 256 * L; float* g_u2 = g_u1 + 128 * L;
  float* g_u3 = g_u2 + 64 * L;  float* g_o = g_u3 + 32 * L;
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    in0[0 * L + i] = coords[((size_t)n * L + i) * 2 + 0];
    in0[1 * L + i] = coords[((size_t)n * L + i) * 2 + 1];
    in0[2 * L + i] = mask[(size_t)n * L + i];
  }
  for (int i = threadIdx.x; i < L * 2; i += blockDim.x) g_o[(i & 1) * L + (i >> 1)] = dout[(size_t)n * L * 2 + i];
  __syncthreads();
  conv1d_k3(x1, 32, in0, 3, nullptr, 0, P.w[0], P.b[0], L, 1);
  conv1d_k3(x2, 64, x1, 32, nullptr, 0, P.w[1], P.b[1], L, 1);
  conv1d_k3(x3, 128, x2, 64, nullptr, 0, P.w[2], P.b[2], L, 1);
  conv1d_k3(t1, 256, x3, 128, nullptr, 0, P.w[3], P.b[3], L, 1);
  conv1d_k3(t2, 256, t1, 256, nullptr, 0, P.w[4], P.b[4], L, 1);
  conv1d_k3(u1, 128, t2, 256, x3, 128, P.w[5], P.b[5], L, 1);
  conv1d_k3(u2, 64, u1, 128, x2, 64, P.w[6], P.b[6], L, 1);
  conv1d_k3(u3, 32, u2, 64, x1, 32, P.w[7], P.b[7], L, 1);
  conv1d_k3(o, 2, u3, 32, nullptr, 0, P.w[8], P.b[8], L, 2);
  // predictor
  act_bwd(g_o, o, 2 * L, 2);
  conv1d_k3_wgrad(g_o, 2, u3, 32, nullptr, 0, G.w[8], G.b[8], L);
  conv1d_k3_dgrad(g_o, 2, P.w[8], 32, 0, 32, g_u3, L, false);
  __syncthreads();
  // up_3 on cat([u2, x1])
  act_bwd(g_u3, u3, 32 * L, 1);
  conv1d_k3_wgrad(g_u3, 32, u2, 64, x1, 32, G.w[7], G.b[7], L);
  conv1d_k3_dgrad(g_u3, 32, P.w[7], 96, 0, 64, g_u2, L, false);
  conv1d_k3_dgrad(g_u3, 32, P.w[7], 96, 64, 32, g_x1, L, false);
  __syncthreads();
  // up_2 on cat([u1, x2])
  act_bwd(g_u2, u2, 64 * L, 1);
  conv1d_k3_wgrad(g_u2, 64, u1, 128, x2, 64, G.w[6], G.b[6], L);
  conv1d_k3_dgrad(g_u2, 64, P.w[6], 192, 0, 128, g_u1, L, false);
  conv1d_k3_dgrad(g_u2, 64, P.w[6], 192, 128, 64, g_x2, L, false);
  __syncthreads();
  // up_1 on cat([t2, x3])
  act_bwd(g_u1, u1, 128 * L, 1);
  conv1d_k3_wgrad(g_u1, 128, t2, 256, x3, 128, G.w[5], G.b[5], L);
  conv1d_k3_dgrad(g_u1, 128, P.w[5], 384, 0, 256, g_t2, L, false);
  conv1d_k3_dgrad(g_u1, 128, P.w[5], 384, 256, 128, g_x3, L, false);
  __syncthreads();
  // buttleneck.conv_2, conv_1
  act_bwd(g_t2, t2, 256 * L, 1);
  conv1d_k3_wgrad(g_t2, 256, t1, 256, nullptr, 0, G.w[4], G.b[4], L);
  conv1d_k3_dgrad(g_t2, 256, P.w[4], 256, 0, 256, g_t1, L, false);
  __syncthreads();
  act_bwd(g_t1, t1, 256 * L, 1);
  conv1d_k3_wgrad(g_t1, 256, x3, 128, nullptr, 0, G.w[3], G.b[3], L);
  conv1d_k3_dgrad(g_t1, 256, P.w[3], 128, 0, 128, g_x3, L, true);
  __syncthreads();
  // down_3, down_2, down_1
  act_bwd(g_x3, x3, 128 * L, 1);
  conv1d_k3_wgrad(g_x3, 128, x2, 64, nullptr, 0, G.w[2], G.b[2], L);
  conv1d_k3_dgrad(g_x3, 128, P.w[2], 64, 0, 64, g_x2, L, true);
  __syncthreads();
  act_bwd(g_x2, x2, 64 * L, 1);
  conv1d_k3_wgrad(g_x2, 64, x1, 32, nullptr, 0, G.w[1], G.b[1], L);
  conv1d_k3_dgrad(g_x2, 64, P.w[1], 32, 0, 32, g_x1, L, true);
  __syncthreads();
  act_bwd(g_x1, x1, 32 * L, 1);
  conv1d_k3_wgrad(g_x1, 32, in0, 3, nullptr, 0, G.w[0], G.b[0], L);
  if (dcoords != nullptr) {  // gradient w.r.t. the (N, L, 2) coordinates (the mask channel gets none)
    conv1d_k3_dgrad(g_x1, 32, P.w[0], 3, 0, 2, g_in0, L, false);
    __syncthreads();
    for (int i = threadIdx.x; i < L * 2; i += blockDim.x) dcoords[(size_t)n * L * 2 + i] = g_in0[(i & 1) * L + (i >> 1)];
  }
}
int launch_inpaint_bwd(const float* coords, const float* mask, const InpaintParams& p, const float* dout,
                       const InpaintGrads& g, int N, int L, float* dcoords, cudaStream_t st) {
  if (N == 0) return 0;
  const size_t smem = (size_t)2 * (3 + 32 + 64 + 128 + 256 + 256 + 128 + 64 + 32 + 2) * L * sizeof(float);
  TNB_REQUIRE(smem <= 227 * 1024, "inpaint_bwd: sequence length %d too long for the fused kernel (max 28)", L);
  TNB_CHECK_CUDA(cudaFuncSetAttribute(inpaint_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  inpaint_bwd_kernel<<<N, 1024, smem, st>>>(coords, mask, p, dout, g, L, dcoords);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// =============================================================================================
// Temporal ensemble of overlapping sliding-window predictions (reference predict.py:163-209 for heatmaps and
// :245-301 for InpaintNet coordinates, where it is a per-frame Python loop over a growing torch.cat buffer on the
// CPU after a full D2H of the heatmaps). Sample s predicts frames s..s+L-1; output frame j of this batch combines
// element [row0 + k][L-1-k] for k = 0..L-1 of the logical buffer cat(state (L-1 rows), pred (B rows), zeros).
// One pass: every prediction element is read once per output frame that uses it (HBM-bound).
// =============================================================================================
struct EnsembleArgs {
  const float* state;  // [(L-1)][L][E]  predictions of the previous L-1 samples (zeros before the first batch)
  const float* pred;   // [B][L][E]
  float* out;          // [B + n_tail][E]
  float weight[16];
  int L, B, count0, tail_base, n_tail;
  long long E;
};
__global__ void __launch_bounds__(256) temporal_ensemble_kernel(const __grid_constant__ EnsembleArgs a) {
  const int j = blockIdx.y;
  const int L = a.L, S = L - 1;
  int row0;
  float div = 0.f;  // > 0: plain sum divided by `div`; 0: weighted sum
  if (j < a.B) {
    row0 = j;
    const int s = a.count0 + j;
    if (s < S) div = (float)(s + 1);               // incomplete buffer, predict.py:181-183
  } else {
    const int f = j - a.B + 1;                     // last input sequence, predict.py:197-201
    row0 = a.tail_base + f;
    div = (float)(L - f);
  }
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < a.E; e += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < L; ++k) {
      const int r = row0 + k;
      float v = 0.f;
      if (r < S) v = a.state[((size_t)r * L + (L - 1 - k)) * a.E + e];
      else if (r < S + a.B) v = a.pred[((size_t)(r - S) * L + (L - 1 - k)) * a.E + e];
      // torch evaluates (buffer * weight).sum(0): a rounded product, then a rounded add - no fused multiply-add
      acc = __fadd_rn(acc, div > 0.f ? v : __fmul_rn(v, a.weight[k]));
    }
    a.out[(size_t)j * a.E + e] = div > 0.f ? __fdiv_rn(acc, div) : acc;
  }
}
int launch_temporal_ensemble(const float* state, const float* pred, float* out, const float* weight_host, int L,
                             long long E, int B, int count0, int tail_base, int n_tail, cudaStream_t st) {
  TNB_REQUIRE(L >= 1 && L <= 16, "temporal_ensemble: sequence length %d not in 1..16", L);
  TNB_REQUIRE(B >= 0 && E > 0 && (n_tail == 0 || n_tail == L - 1), "temporal_ensemble: bad batch / tail (%d, %d)", B, n_tail);
  if (B + n_tail == 0) return 0;
  EnsembleArgs a;
  a.state = state; a.pred = pred; a.out = out; a.L = L; a.B = B; a.count0 = count0; a.tail_base = tail_base;
  a.n_tail = n_tail; a.E = E;
  for (int k = 0; k < 16; ++k) a.weight[k] = k < L ? weight_host[k] : 0.f;
  const int bx = (int)std::min<long long>((E + 255) / 256, 148 * 4);
  temporal_ensemble_kernel<<<dim3(bx, B + n_tail), 256, 0, st>>>(a);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// =============================================================================================
// Per-map statistics for the evaluation bookkeeping (reference test.py:159-169): the detection confidence = maximum
// of the heatmap inside the predicted bounding box (0 when the box is empty), and whether the ground-truth map has
// any non-zero value (np.amax(y_t) > 0). One CTA per map; replaces a D2H copy of both full heatmaps.
// =============================================================================================
__global__ void __launch_bounds__(256) eval_stats_kernel(const float* __restrict__ y_pred, const float* __restrict__ y_true,
                                                         const int* __restrict__ boxes, int H, int W,
                                                         float* __restrict__ conf, int* __restrict__ true_any) {
  const int m = blockIdx.x;
  const int bx = boxes[m * 4 + 0], by = boxes[m * 4 + 1], bw = boxes[m * 4 + 2], bh = boxes[m * 4 + 3];
  const float* yp = y_pred + (size_t)m * H * W;
  float best = -INFINITY;
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) best = fmaxf(best, yp[(size_t)(by + i / bw) * W + bx + i % bw]);
  float tmax = -INFINITY;
  if (y_true != nullptr) {
    const float* yt = y_true + (size_t)m * H * W;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) tmax = fmaxf(tmax, yt[i]);
  }
  __shared__ float s1[8], s2[8];
  best = warp_max(best); tmax = warp_max(tmax);
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = best; s2[threadIdx.x >> 5] = tmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { best = fmaxf(best, s1[i]); tmax = fmaxf(tmax, s2[i]); }
    conf[m] = (bw > 0 && bh > 0) ? fmaxf(best, s1[0]) : 0.f;
    if (true_any != nullptr) true_any[m] = fmaxf(tmax, s2[0]) > 0.f ? 1 : 0;
  }
}
int launch_eval_stats(const float* y_pred, const float* y_true, const int* boxes, int nmaps, int H, int W, float* conf,
                      int* true_any, cudaStream_t st) {
  if (nmaps == 0) return 0;
  eval_stats_kernel<<<nmaps, 256, 0, st>>>(y_pred, y_true, boxes, H, W, conf, true_any);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// =============================================================================================
// Frame preprocessing (reference dataset.py:435-461, 611-647, 783-812): `img.resize((W, H))` of every uint8 frame -
// Pillow's antialiased BICUBIC in 22-bit fixed point, horizontal pass first, rounded to uint8 between the passes
// (Pillow 10.0.0 src/libImaging/Resample.c, the reference's pinned dependency) - then HWC -> CHW, / 255 and stacking
// into the (N, C, H, W) float tensor the network consumes. Integer arithmetic throughout: bit-exact against Pillow.
// The coefficient tables are built on the host exactly as Resample.c builds them (tracknetv3_b200/frames.py).
// =============================================================================================
TNB_DEVINL int clip8_fixed(int acc) {
  const int v = acc >> 22;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}
// horizontal pass: src [nimg][hs][ws][C] u8 -> tmp [nimg][hs][wd][C] u8
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, int nimg, int hs, int ws, int wd, int C,
                                                       const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                       uint8_t* __restrict__ tmp) {
  const long long total = (long long)nimg * hs * wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % wd);
    const long long row = i / wd;  // image * hs + y
    const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    const uint8_t* p = src + (row * ws + xmin) * C;
    const int* k = kk + (size_t)xx * ksize;
    int acc[4] = {1 << 21, 1 << 21, 1 << 21, 1 << 21};
    for (int x = 0; x < cnt; ++x) {
      const int w = k[x];
      for (int c = 0; c < C; ++c) acc[c] += (int)p[x * C + c] * w;
    }
    for (int c = 0; c < C; ++c) tmp[i * C + c] = (uint8_t)clip8_fixed(acc[c]);
  }
}
// vertical pass fused with HWC -> CHW, / 255 and the channel stacking: tmp [nimg][hs][wd][C] u8 ->
// out[(img / per_sample) * sample_stride + ((img % per_sample) * frame_stride + chan_off + c) * hd * wd + yy * wd + xx]
__global__ void __launch_bounds__(256) resize_v_stack_kernel(const uint8_t* __restrict__ tmp, int nimg, int hs, int wd, int hd,
                                                             int C, const int* __restrict__ bounds, const int* __restrict__ kk,
                                                             int ksize, float* __restrict__ out, int per_sample,
                                                             long long sample_stride, int chan_off, int frame_stride) {
  const long long total = (long long)nimg * hd * wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % wd);
    const long long r = i / wd;
    const int yy = (int)(r % hd);
    const int img = (int)(r / hd);
    const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    const uint8_t* p = tmp + (((long long)img * hs + ymin) * wd + xx) * C;
    const int* k = kk + (size_t)yy * ksize;
    int acc[4] = {1 << 21, 1 << 21, 1 << 21, 1 << 21};
    for (int y = 0; y < cnt; ++y) {
      const int w = k[y];
      for (int c = 0; c < C; ++c) acc[c] += (int)p[(long long)y * wd * C + c] * w;
    }
    float* o = out + (long long)(img / per_sample) * sample_stride +
               ((long long)(img % per_sample) * frame_stride + chan_off) * hd * wd + (long long)yy * wd + xx;
    // the reference divides the float64 array by 255. and casts to float32 at train.py:86 / predict.py:171
    for (int c = 0; c < C; ++c) o[(long long)c * hd * wd] = __double2float_rn((double)clip8_fixed(acc[c]) / 255.0);
  }
}
// Background subtraction of bg_mode 'subtract' / 'subtract_concat' (dataset.py:438, 442, 797, 801):
// np.sum(np.absolute(img - median), 2).astype('uint8') - float64 arithmetic, truncation toward zero, and the wrap-around
// of numpy's double -> uint8 cast for sums above 255 (the reference feeds that image to PIL as it is).
__global__ void __launch_bounds__(256) bg_subtract_kernel(const uint8_t* __restrict__ frames, const double* __restrict__ median,
                                                          long long nimg, long long hw, uint8_t* __restrict__ out) {
  const long long total = nimg * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % hw;
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) v += fabs((double)frames[i * 3 + c] - median[p * 3 + c]);
    out[i] = (uint8_t)((long long)v & 0xff);
  }
}
int launch_bg_subtract(const uint8_t* frames, const double* median, long long nimg, int hs, int ws, uint8_t* out,
                       cudaStream_t st) {
  if (nimg == 0) return 0;
  const long long total = nimg * hs * ws;
  bg_subtract_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(frames, median, nimg,
                                                                                            (long long)hs * ws, out);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_resize_frames(const uint8_t* src, int nimg, int hs, int ws, int C, const int* hbounds, const int* hkk, int hksize,
                         const int* vbounds, const int* vkk, int vksize, int hd, int wd, uint8_t* tmp, float* out,
                         int per_sample, long long sample_stride, int chan_off, int frame_stride, cudaStream_t st) {
  TNB_REQUIRE(C >= 1 && C <= 4 && per_sample >= 1 && frame_stride >= C, "resize_frames: channels %d / frames per sample %d / stride %d",
              C, per_sample, frame_stride);
  if (nimg == 0) return 0;
  const uint8_t* vsrc = src;
  if (hbounds != nullptr) {  // widths differ: horizontal pass first (Resample.c ImagingResampleInner)
    const long long total = (long long)nimg * hs * wd;
    resize_h_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(src, nimg, hs, ws, wd, C,
                                                                                           hbounds, hkk, hksize, tmp);
    TNB_CHECK_CUDA(cudaGetLastError());
    vsrc = tmp;
  } else {
    TNB_REQUIRE(ws == wd, "resize_frames: no horizontal table but widths differ (%d vs %d)", ws, wd);
  }
  const long long total = (long long)nimg * hd * wd;
  resize_v_stack_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(
      vsrc, nimg, hs, wd, hd, C, vbounds, vkk, vksize, out, per_sample, sample_stride, chan_off, frame_stride);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tnb
