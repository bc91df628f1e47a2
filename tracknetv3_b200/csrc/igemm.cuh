// Shared declarations for the tcgen05 implicit-GEMM convolution kernels (igemm.cu) and the
// elementwise kernels that consume the same "view" description of a layer input.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "../../include/tracknet_b200.h"

namespace tnb {

// How a layer's logical input (what the reference's Conv2d sees, model.py:57-73) is produced from
// the tensors we actually keep in HBM. We never materialise BN-apply / ReLU / MaxPool / Upsample /
// concat: each consumer recomputes them on the fly from the producer's raw conv output z.
enum SrcMode : int {
  SRC_IDENTITY = TNB_SRC_IDENTITY,                  // plain NHWC fp32 tensor (packed network input, or dz)
  SRC_AFFINE_RELU = TNB_SRC_AFFINE_RELU,            // relu(z*scale + shift)                  (model.py:12-16)
  SRC_AFFINE_RELU_POOL = TNB_SRC_AFFINE_RELU_POOL,  // maxpool2x2(relu(z*scale+shift)), z is 2H x 2W (model.py:59,61,63)
  SRC_AFFINE_RELU_UP = TNB_SRC_AFFINE_RELU_UP,      // nearest x2 upsample of relu(z*scale+shift)  (model.py:65,67,69)
  SRC_PRESPLIT = TNB_SRC_PRESPLIT,                  // already (hi, lo) 16-bit, [pixel][2][C]: pure copy
  SRC_PRESPLIT_UP = TNB_SRC_PRESPLIT_UP,            // the same at half resolution, read as its nearest x2 upsampling (wgrad)
  SRC_PLANAR16 = TNB_SRC_PLANAR16                   // planar fp16 (hi, lo) planes [N][plane][H][W][8]: staged by tensor-TMA
};
using SrcDesc = tnb_src_t;    // see include/tracknet_b200.h
using ViewDesc = tnb_view_t;

// pixel offset (in pixels, not elements) of view pixel (n,h,w) inside source s
TNB_DEVINL int view_pix_off(const SrcDesc& s, int n, int h, int w) {
  if (s.mode == SRC_AFFINE_RELU_POOL) return (n * s.Hs + 2 * h) * s.Ws + 2 * w;
  if (s.mode == SRC_AFFINE_RELU_UP) return (n * s.Hs + (h >> 1)) * s.Ws + (w >> 1);
  return (n * s.Hs + h) * s.Ws + w;
}

TNB_DEVINL void ld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// 8 consecutive channels [c, c+8) of the view at a pixel whose source pixel offset is `poff`.
TNB_DEVINL void view_load8(const SrcDesc& s, int poff, int c, float (&v)[8]) {
  const float* p = s.ptr + (size_t)poff * s.C + c;
  if (s.mode == SRC_IDENTITY) {
    ld8(p, v);
    return;
  }
  float sc[8], sh[8];
  ld8(s.scale + c, sc);
  ld8(s.shift + c, sh);
  if (s.mode == SRC_AFFINE_RELU_POOL) {
    float a[8], m[8];
    ld8(p, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = fmaf(a[i], sc[i], sh[i]);
    ld8(p + s.C, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], fmaf(a[i], sc[i], sh[i]));
    ld8(p + (size_t)s.Ws * s.C, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], fmaf(a[i], sc[i], sh[i]));
    ld8(p + (size_t)s.Ws * s.C + s.C, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaxf(m[i], fmaf(a[i], sc[i], sh[i])), 0.f);
  } else {
    float a[8];
    ld8(p, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaf(a[i], sc[i], sh[i]), 0.f);
  }
}

// ---- batched gather: issue all global loads of a batch first, then convert/store (memory-level parallelism)
struct Raw8 { float4 a, b; };
TNB_DEVINL Raw8 ld_raw8(const float* p) {
  Raw8 r;
  r.a = __ldg(reinterpret_cast<const float4*>(p));
  r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  return r;
}
TNB_DEVINL void raw_to_arr(const Raw8& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w;
  v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
// number of raw 8-float loads one view element needs
template <int MODE> struct RawCount { static constexpr int value = (MODE == SRC_AFFINE_RELU_POOL) ? 4 : 1; };
template <int MODE>
TNB_DEVINL void view_issue(const SrcDesc& s, int poff, int c, Raw8 (&raw)[RawCount<MODE>::value]) {
  if (MODE == SRC_PRESPLIT) {
    // [pixel][2 (hi, lo)][C] 16-bit: raw.a = the 8 hi values, raw.b = the 8 lo values of channels [c, c+8)
    const uint8_t* b = reinterpret_cast<const uint8_t*>(s.ptr) + ((size_t)poff * 2 * s.C + c) * 2;
    const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(b));
    const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(b + 2 * s.C));
    raw[0].a = *reinterpret_cast<const float4*>(&x0);
    raw[0].b = *reinterpret_cast<const float4*>(&x1);
    return;
  }
  const float* p = s.ptr + (size_t)poff * s.C + c;
  raw[0] = ld_raw8(p);
  if (MODE == SRC_AFFINE_RELU_POOL) {
    raw[1 % RawCount<MODE>::value] = ld_raw8(p + s.C);
    raw[2 % RawCount<MODE>::value] = ld_raw8(p + (size_t)s.Ws * s.C);
    raw[3 % RawCount<MODE>::value] = ld_raw8(p + (size_t)s.Ws * s.C + s.C);
  }
}
// finish: BN affine + ReLU (+ max over the 2x2 window); IDENTITY multiplies by `mul` (power-of-two scaling of dz)
template <int MODE>
TNB_DEVINL void view_finish(const Raw8 (&raw)[RawCount<MODE>::value], const float (&sc)[8], const float (&sh)[8],
                            float mul, float (&v)[8]) {
  float a[8];
  raw_to_arr(raw[0], a);
  if (MODE == SRC_IDENTITY || MODE == SRC_PRESPLIT) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = a[i] * mul;
    return;
  }
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaf(a[i], sc[i], sh[i]);
  if (MODE == SRC_AFFINE_RELU_POOL) {
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      raw_to_arr(raw[k % RawCount<MODE>::value], a);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], fmaf(a[i], sc[i], sh[i]));
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = fmaxf(m[i], 0.f);
}
// power-of-two multiplier that brings a tensor with max |x| = amax into [2^9, 2^10): safe for fp16 hi/lo splitting
TNB_DEVINL float pow2_scale_for(const float* amax_ptr) {
  if (amax_ptr == nullptr) return 1.f;
  const float amax = *amax_ptr;
  if (!(amax > 0.f) || !(amax < 3.0e38f)) return 1.f;
  int e;
  frexpf(amax, &e);
  return ldexpf(1.f, 10 - e);
}

// ---- host-side launchers (igemm.cu) ----------------------------------------------------------
__host__ __device__ inline int pad_px(int px) {  // plane stride = 2 (mod 8) pixels -> conflict-free fill stores
  int r = px & 7;
  return px + ((2 - r) & 7);
}

struct ConvPlan {
  int BN, MT, SA, SB, G, nbuf, tmem_cols, merged;
  size_t smem_bytes;
  int tiles_h, tiles_w;
  int tall;  // tile orientation, see conv3x3_plan
};
// fmt: 0 = fp16 split, 1 = bf16 split. nterms: 1 (single pass) or 3 (hi*hi + lo*hi + hi*lo).
// bn_bwd_fused: reserve the per-channel constant table of the fused BatchNorm-backward reduction (dgrad epilogue)
int conv3x3_weight_layout(int BN);    // 0 plain [term][plane][BN], 1 merged [plane][term][BN]
bool conv3x3_merged(int BN);  // weights of this tile width are packed [plane][hi | lo][BN] (one MMA for x_hi * [w_hi | w_lo])
// copy_fill: the view is a pre-split tensor (dgrad) whose halo tiles are filled by cp.async copies - only the number of
// halo-tile stages (and with it the shared-memory size) depends on it, never the tiling
int conv3x3_plan(int N, int H, int W, int Cin, int Cout, int nterms, ConvPlan* plan, bool bn_bwd_fused = false,
                 bool copy_fill = false);
size_t conv3x3_wpack_elems(int Kside, int Nside);  // uint16 elements of a packed weight buffer
int launch_pack_weights(const float* w_oihw, uint16_t* out, int Co, int Ci, int mode /*0 fwd, 1 dgrad*/,
                        int fmt, int BN, cudaStream_t st);
// many tensors in one launch: pack_table_add() each, then launch_pack_table()
struct PackEntry { const float* w; uint16_t* out; long long first; int Co, Ci, Kpad, BN, mode, fmt, layout; };
constexpr int kMaxPack = 36;
struct PackTable { PackEntry e[kMaxPack]; long long total; int n; };
int pack_table_add(PackTable* t, const float* w_oihw, uint16_t* out, int Co, int Ci, int mode, int fmt, int BN);
int launch_pack_table(const PackTable& t, cudaStream_t st);
// Optional fusion for dgrad launches: `out` is dL/d(activation) of a producer layer whose raw conv output is `z`; the
// per-tile partials become (sum g, sum g * xhat) with g = out masked by that layer's ReLU - the reduction pass of its
// BatchNorm backward - instead of (sum out, sum out^2).
struct BnBwdFuse { const float *z, *scale, *shift, *mean, *invstd; };
// the two MMA-issue-loop variants of the kernel (conv.cu / conv_lean.cu); launch_conv3x3 plans and picks
int launch_conv3x3_generic(const ViewDesc& view, const uint16_t* wpack, float* out, float* stat_part, int Cout, int nterms,
                           int fmt, int variant, const ConvPlan& plan, cudaStream_t st, const BnBwdFuse* fuse = nullptr);
int launch_conv3x3_lean(const ViewDesc& view, const uint16_t* wpack, float* out, float* stat_part, int Cout, int nterms,
                        int fmt, int variant, const ConvPlan& plan, cudaStream_t st, const BnBwdFuse* fuse = nullptr);
int num_sms();
// Tensor map of a TNB_SRC_PLANAR16 tensor ([N][C / 32][2][4][H][W][8] fp16) for the convolution's halo tiles: dims
// {8 W, H, planes, N}, box {8 box_w, box_h, 8 planes, 1}, no swizzle, out-of-range elements read as zero. The encoder
// (cuTensorMapEncodeTiled) is fetched from the driver through the runtime: no link-time dependency on libcuda.
int make_planar16_tmap(CUtensorMap* out, const void* base, int N, int H, int W, int C, int box_w, int box_h);
#ifdef TNB_TRACE
extern long long* g_conv_trace;  // debug build (make TRACE=1): set by tnb_debug_set_trace, read by the conv launchers
#endif
int launch_conv3x3(const ViewDesc& view, const uint16_t* wpack, float* out, float* stat_part, int Cout,
                   int nterms, int fmt, int variant, cudaStream_t st, const BnBwdFuse* fuse = nullptr);
int conv3x3_num_stat_rows(int N, int H, int W, int Cin, int Cout, int nterms, bool bn_bwd_fused = false);

// ws: optional scratch of wgrad3x3_ws_floats(view, Cout) floats. With it every split-K CTA stores its partial tap-major
// into its own slab (lanes of a warp hit consecutive addresses) and a small kernel sums the slabs in split order into
// dw_oihw: deterministic, plain stores, dw need not be zeroed. Without it the partials are added into dw_oihw with
// atomics (the caller zeroes it; summation order, hence the last bits, vary from run to run).
size_t wgrad3x3_ws_floats(const ViewDesc& view, int Cout);
// fmt: 16-bit format of the pre-split operands (0 fp16, 1 bf16; the view operand, if split on the fly, follows it);
// dz_mul: optional device scalar dz was stored multiplied by (tnb_bnbwd_t.dz_format 2) - the result is divided by it
int launch_wgrad3x3(const ViewDesc& view, const void* dz_presplit, float* dw_oihw, int Cout, int CinReal, int nterms,
                    int variant, cudaStream_t st, float* ws = nullptr, int fmt = 1, const float* dz_mul = nullptr);

}  // namespace tnb
