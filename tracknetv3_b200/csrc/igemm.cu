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
#include <algorithm>

namespace tnb {


// =============================================================================================
// weight packing: OIHW fp32 -> per (n-tile, k-chunk, tap) smem images [term][plane(4)][BN][8] (uint16)
// (64-wide tiles: [plane(4)][term][BN][8], see conv3x3_merged)
// mode 0 (forward):  B[n = co][k = ci] = W[co][ci][dy][dx]
// mode 1 (dgrad):    B[n = ci][k = co] = W[co][ci][2-dy][2-dx]      (180-degree rotated, transposed)
// =============================================================================================
// One launch packs every tensor of a table (the 17 forward operands and, when training, the 16 dgrad operands of a
// TrackNet step): 33 launches of 3-7 us collapse into one.
__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ PackTable t) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < t.total;
       i += (long long)gridDim.x * blockDim.x) {
    int ei = 0;
    while (ei + 1 < t.n && i >= t.e[ei + 1].first) ++ei;
    const PackEntry& E = t.e[ei];
    const int BN = E.BN, nchunks = E.Kpad / 32;
    long long r = i - E.first;
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
      if (E.mode == 0) {
        if (n < E.Co && k < E.Ci) x = E.w[(((size_t)n * E.Ci + k) * 3 + dy) * 3 + dx];
      } else {
        if (n < E.Ci && k < E.Co) x = E.w[(((size_t)k * E.Ci + n) * 3 + (2 - dy)) * 3 + (2 - dx)];
      }
      v[e] = x;
    }
    uint4 hi, lo;
    if (E.fmt == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= kW16Mul;  // keeps the lo halves normal fp16 numbers, see common.cuh
      split8<0>(v, hi, lo);
    } else {
      split8<1>(v, hi, lo);
    }
    // stage base (uint16 elements): ((ntile*nchunks + chunk)*9 + tap) * (2*4*BN*8)
    const size_t stage = (((size_t)ntile * nchunks + chunk) * 9 + tap) * (size_t)(64 * BN);
    // [term][plane][BN rows][8], or for the merged 64-wide tiles [plane][term][BN rows][8] (conv3x3_merged): hi and lo
    // rows of a plane are then one operand of 2 * BN rows
    size_t off_hi, off_lo;
    if (E.layout == 1) {
      off_hi = ((size_t)plane * 2 * BN + nl) * 8;
      off_lo = ((size_t)(plane * 2 + 1) * BN + nl) * 8;
    } else {
      off_hi = ((size_t)plane * BN + nl) * 8;
      off_lo = (size_t)32 * BN + off_hi;
    }
    *reinterpret_cast<uint4*>(E.out + stage + off_hi) = hi;
    *reinterpret_cast<uint4*>(E.out + stage + off_lo) = lo;
  }
}

size_t conv3x3_wpack_elems(int Kside, int Nside) {
  const int Kpad = (Kside + 31) / 32 * 32;
  return (size_t)Nside * Kpad * 9 * 2;
}

int pack_table_add(PackTable* t, const float* w, uint16_t* out, int Co, int Ci, int mode, int fmt, int BN) {
  TNB_REQUIRE(t->n < kMaxPack, "pack_weights: more than %d tensors in one table", kMaxPack);
  const int Nside = mode == 0 ? Co : Ci;
  const int Kside = mode == 0 ? Ci : Co;
  const int Kpad = (Kside + 31) / 32 * 32;
  TNB_REQUIRE(BN > 0 && Nside % BN == 0, "pack_weights: N side %d not divisible by BN %d", Nside, BN);
  PackEntry& E = t->e[t->n++];
  E.w = w; E.out = out; E.Co = Co; E.Ci = Ci; E.Kpad = Kpad; E.BN = BN; E.mode = mode; E.fmt = fmt;
  E.layout = conv3x3_weight_layout(BN);  // a function of the tile width alone
  E.first = t->total;
  t->total += (long long)(Nside / BN) * (Kpad / 32) * 9 * 4 * BN;
  return 0;
}
int launch_pack_table(const PackTable& t, cudaStream_t st) {
  if (t.n == 0) return 0;
  const int blocks = (int)std::min<long long>((t.total + 255) / 256, 148 * 16);
  return launch_pdl(pack_weights_kernel, dim3(blocks), dim3(256), 0, st, t);
}
int launch_pack_weights(const float* w, uint16_t* out, int Co, int Ci, int mode, int fmt, int BN,
                        cudaStream_t st) {
  PackTable t;
  t.n = 0; t.total = 0;
  if (int rc = pack_table_add(&t, w, out, Co, Ci, mode, fmt, BN)) return rc;
  return launch_pack_table(t, st);
}

}  // namespace tnb
