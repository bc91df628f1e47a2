// TrackNet.forward and its autograd backward as one C-ABI call each (reference model.py:44-73).
// This file is host-side orchestration only: it wires the 17 Conv2DBlocks' tensors into the view /
// gradient-source descriptors the kernels consume and lays out the caller-provided workspace.
#include "kernels.cuh"
#include "prof.cuh"
#include <string.h>
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace tnb {

namespace {

constexpr int kLayers = 17;
// Layer wiring of TrackNet (reference model.py:47-53 constructor, :58-70 forward).
struct LayerDef {
  int cout, level;  // level: spatial size = (H >> level, W >> level)
  int src0, mode0;  // producer layer (-1 = packed network input) and how the consumer sees it
  int src1, mode1;  // second (skip) source for the concat layers, -1 if none
};
const LayerDef kDefs[kLayers] = {
    {64, 0, -1, SRC_IDENTITY, -1, 0},                         //  0 down_block_1.conv_1
    {64, 0, 0, SRC_AFFINE_RELU, -1, 0},                       //  1 down_block_1.conv_2   -> x1
    {128, 1, 1, SRC_AFFINE_RELU_POOL, -1, 0},                 //  2 down_block_2.conv_1   (MaxPool, model.py:59)
    {128, 1, 2, SRC_AFFINE_RELU, -1, 0},                      //  3 down_block_2.conv_2   -> x2
    {256, 2, 3, SRC_AFFINE_RELU_POOL, -1, 0},                 //  4 down_block_3.conv_1   (model.py:61)
    {256, 2, 4, SRC_AFFINE_RELU, -1, 0},                      //  5 down_block_3.conv_2
    {256, 2, 5, SRC_AFFINE_RELU, -1, 0},                      //  6 down_block_3.conv_3   -> x3
    {512, 3, 6, SRC_AFFINE_RELU_POOL, -1, 0},                 //  7 bottleneck.conv_1     (model.py:63)
    {512, 3, 7, SRC_AFFINE_RELU, -1, 0},                      //  8 bottleneck.conv_2
    {512, 3, 8, SRC_AFFINE_RELU, -1, 0},                      //  9 bottleneck.conv_3
    {256, 2, 9, SRC_AFFINE_RELU_UP, 6, SRC_AFFINE_RELU},      // 10 up_block_1.conv_1     cat[up(x), x3] (model.py:65)
    {256, 2, 10, SRC_AFFINE_RELU, -1, 0},                     // 11 up_block_1.conv_2
    {256, 2, 11, SRC_AFFINE_RELU, -1, 0},                     // 12 up_block_1.conv_3
    {128, 1, 12, SRC_AFFINE_RELU_UP, 3, SRC_AFFINE_RELU},     // 13 up_block_2.conv_1     cat[up(x), x2] (model.py:67)
    {128, 1, 13, SRC_AFFINE_RELU, -1, 0},                     // 14 up_block_2.conv_2
    {64, 0, 14, SRC_AFFINE_RELU_UP, 1, SRC_AFFINE_RELU},      // 15 up_block_3.conv_1     cat[up(x), x1] (model.py:69)
    {64, 0, 15, SRC_AFFINE_RELU, -1, 0},                      // 16 up_block_3.conv_2
};

struct LayerBuf {
  int H, W, cin, cin_real, cout;
  float *z, *scale, *shift, *mean, *invstd, *stat_part;
  int stat_rows;
  float *dz, *din, *bwd_part, *bwd_sums, *amax;
  int bwd_rows;
  int fused_rows;  // > 0: the BatchNorm-backward reduction of THIS layer is produced by the dgrad epilogue of layer + 1
  uint16_t *wf, *wd;
  uint8_t* cat16;  // decoder concat layers (training): their own pre-split wgrad operand [up half at half resolution | skip
                   // half], filled by two apply passes many layers apart (whole-pass backward), else nullptr
};

// The reduction pass of layer p's BatchNorm backward can ride on the dgrad of layer p + 1 when that convolution is
// the only consumer of p's activation and sees it at the same resolution through a plain BN + ReLU view.
bool bn_reduce_fusable(int p) {
  if (p + 1 >= kLayers) return false;
  const LayerDef& d = kDefs[p + 1];
  if (d.src0 != p || d.src1 >= 0 || d.mode0 != SRC_AFFINE_RELU) return false;
  for (int j = 0; j < kLayers; ++j)
    if (j != p + 1 && (kDefs[j].src0 == p || kDefs[j].src1 == p)) return false;
  return true;
}
struct Plan {
  LayerBuf L[kLayers];
  float* wgrad_ws;  // scratch: split-K slabs of the current layer's weight gradient (deterministic ordered sum)
  float* pred_ws;   // scratch: per-block partials of the predictor's weight / bias gradient
  uint8_t* vsplit;  // scratch: the current layer's input view materialised as pre-split bf16 (largest: 192 ch @ full res)
  float* amax_all;  // [2 * kLayers]: max |g| per layer (raised by the BatchNorm-backward reduction pass), then the power of
                    // two each layer's dz is stored multiplied by (written by the apply pass, read by dgrad / wgrad)
  float* xin;       // network input, fp32 NHWC padded to cpad channels (first layer's wgrad view; its forward view when
                    // the tensor-TMA path does not apply)
  uint8_t* xin_split;  // training with the tensor-TMA input path: the input pre-split [pixel][hi | lo][cpad] in the
                       // backward pass's 16-bit format - the first layer's weight-gradient operand, written by the same
                       // pack launch (the fp32 tensor and a view pass over it are then not needed at all)
  uint8_t* xin16;   // the same as planar fp16 (hi, lo) planes (TNB_SRC_PLANAR16): the first convolution stages its halo
                    // tiles from it with tensor-TMA; nullptr when the layer's tile plan is not row-major / <= 30 columns
  float* dA_pred;
  int cpad;
  size_t bytes;
};

struct Bump {
  uint8_t* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

// Layer l reads the activation of layer l - 1 - as it is, through MaxPool2d, or (first half of a decoder concat) through
// the x2 upsampling: that operand of its weight gradient is written by the BatchNorm-backward apply pass of layer l - 1,
// which recomputes the activation anyway (tnb_bnbwd_t.act_presplit: at its own resolution - the upsampled half is read
// through (h/2, w/2) addressing - or max-pooled, act_pool), instead of a separate view pass over z; the wgrad of layer l
// is launched right after that pass. The skip half of a concat still comes from a view pass: its producer's apply pass
// runs many layers later.
// tnb_tracknet_cfg_t.training: 0 = running statistics, nothing kept (inference); 1 = batch statistics, running statistics
// and counters advanced, everything a backward pass needs kept (model.train()); 2 = running statistics like 0 AND the
// backward state of 1 (a model in eval() called with gradients enabled: fine-tuning with frozen BatchNorm layers, which
// the reference's autograd path allows, model.py:4-16). `c.training != 0` sizes and fills the backward state.
bool batch_stats(const tnb_tracknet_cfg_t& c) { return c.training == 1; }

bool wgrad_operand_from_bn_bwd(const tnb_tracknet_cfg_t& c, int l) {
  if (c.variant & 128) return false;  // variant bit 128: always materialise with tnb_view_presplit (ablation)
  if (l <= 0 || kDefs[l].src0 != l - 1) return false;
  const LayerDef& d = kDefs[l];
  if (c.variant & 8192) return d.src1 < 0 && d.mode0 == SRC_AFFINE_RELU;  // bit 8192: plain layers only (ablation)
  if (d.src1 < 0) return d.mode0 == SRC_AFFINE_RELU || d.mode0 == SRC_AFFINE_RELU_POOL;
  return d.mode0 == SRC_AFFINE_RELU_UP && d.mode1 == SRC_AFFINE_RELU;
}

// Operand format of the backward pass's tensor-core kernels. 3-term products (the default precision): bf16 (hi, lo) pairs,
// 16 bits, fp32's exponent range, no scaling. Single-pass backward (bwd_terms 1): fp16, 11 bits instead of bf16's 8 -
// every dz tensor is then stored multiplied by a power of two derived from the measured max |g| (bn_bwd_kernel,
// dz_format 2) and the consumers divide it out. fp16 (hi, lo) pairs (22 bits) for the 3-term products are an ablation
// (variant bit 2048 / TNB_BWD_FMT=fp16): measured 2-4x closer to fp64 on single kernels with short accumulation chains,
// no different at network level (the truncating fp32 accumulation of tcgen05.mma dominates) and 1.2 % slower.
// The fused dgrad + BatchNorm-reduction epilogue (variant bit 64) does not measure max |g|: bf16 only.
int bwd_fmt(const tnb_tracknet_cfg_t& c) {
  static const int env = [] {  // ablation, read once: TNB_BWD_FMT=fp16 | bf16
    const char* e = getenv("TNB_BWD_FMT");
    return e == nullptr ? -1 : (strcmp(e, "fp16") == 0 ? 0 : 1);
  }();
  if (c.variant & 64) return 1;
  if (env >= 0) return env;
  if (c.variant & 2048) return 0;
  if (c.variant & 1024) return 1;
  return c.bwd_terms == 1 ? 0 : 1;
}

int layer_cin(const tnb_tracknet_cfg_t& c, int l) {
  if (l == 0) return (c.in_dim + 31) / 32 * 32;
  const LayerDef& d = kDefs[l];
  int cin = kDefs[d.src0].cout;
  if (d.src1 >= 0) cin += kDefs[d.src1].cout;
  return cin;
}

// The weight-gradient kernel's view of layer l's input, materialised pre-split (bf16 hi/lo) in `vsplit`: one tensor, or
// for a decoder concat [up(x), skip] the up-sampled half at ITS OWN (half) resolution - read by the wgrad fill through
// (h/2, w/2) addressing, a quarter of the bytes of the up-sampled tensor - followed by the skip half.
ViewDesc wgrad_view(const tnb_tracknet_cfg_t& c, int l, const uint8_t* vsplit) {
  const LayerDef& d = kDefs[l];
  const int H = c.h >> d.level, W = c.w >> d.level, cin = layer_cin(c, l);
  ViewDesc pv;
  memset(&pv, 0, sizeof(pv));
  pv.N = c.n; pv.H = H; pv.W = W; pv.C = cin;
  if (d.src1 >= 0 && d.mode0 == SRC_AFFINE_RELU_UP && d.mode1 == SRC_AFFINE_RELU) {
    const int c0 = kDefs[d.src0].cout, hs = c.h >> kDefs[d.src0].level, ws = c.w >> kDefs[d.src0].level;
    const uint8_t* skip = vsplit + (size_t)c.n * hs * ws * c0 * 4;
    pv.s[0] = SrcDesc{reinterpret_cast<const float*>(vsplit), nullptr, nullptr, c0, hs, ws, SRC_PRESPLIT_UP};
    pv.s[1] = SrcDesc{reinterpret_cast<const float*>(skip), nullptr, nullptr, cin - c0, H, W, SRC_PRESPLIT};
    pv.C0 = c0;
  } else {
    pv.s[0] = SrcDesc{reinterpret_cast<const float*>(vsplit), nullptr, nullptr, cin, H, W, SRC_PRESPLIT};
    pv.s[1] = pv.s[0];
    pv.C0 = cin;
  }
  return pv;
}

int build_plan(const tnb_tracknet_cfg_t& c, void* ws, Plan* P) {
  TNB_REQUIRE(c.n > 0 && c.h > 0 && c.w > 0 && c.h % 8 == 0 && c.w % 8 == 0,
              "tracknet: input %dx%d must be divisible by 8 (reference model.py:59-69 pools 3x then concatenates)",
              c.h, c.w);
  TNB_REQUIRE(c.in_dim > 0 && c.out_dim > 0, "tracknet: bad in_dim/out_dim %d/%d", c.in_dim, c.out_dim);
  TNB_REQUIRE(c.training >= 0 && c.training <= 2, "tracknet: training must be 0, 1 or 2 (got %d)", c.training);
  TNB_REQUIRE((c.fwd_terms == 1 || c.fwd_terms == 3) && (c.bwd_terms == 1 || c.bwd_terms == 3),
              "tracknet: terms must be 1 or 3");
  Bump b{reinterpret_cast<uint8_t*>(ws), 0};
  P->cpad = (c.in_dim + 31) / 32 * 32;
  const size_t npix0 = (size_t)c.n * c.h * c.w;
  P->amax_all = b.take<float>(2 * kLayers);
  P->vsplit = nullptr;
  if (c.training) {
    size_t vmax = 0;
    for (int l = 0; l < kLayers; ++l) {
      const size_t e = (size_t)c.n * (c.h >> kDefs[l].level) * (c.w >> kDefs[l].level) * layer_cin(c, l);
      vmax = e > vmax ? e : vmax;
    }
    P->vsplit = b.take<uint8_t>(vmax * 4);
    size_t wmax = 0;
    for (int l = 0; l < kLayers; ++l) {
      const ViewDesc pv = wgrad_view(c, l, nullptr);
      const size_t e = wgrad3x3_ws_floats(pv, kDefs[l].cout);
      if (e == 0) return -2;
      wmax = std::max(wmax, e);
    }
    P->wgrad_ws = b.take<float>(wmax);
    P->pred_ws = b.take<float>(predictor_bwd_workspace_bytes(c.n, c.h, c.w, c.out_dim) / sizeof(float));
  }
  P->xin = b.take<float>(npix0 * P->cpad);
  P->xin16 = nullptr;
  {
    // tensor-TMA staging of the network input: 3-term forward on row-major tiles whose halo row fits a TMA box (256
    // elements = 32 pixels). TNB_INPUT_TMA=0 / variant bit 4096: the gather path on the fp32 tensor (ablation).
    static const int env = [] { const char* e = getenv("TNB_INPUT_TMA"); return e ? atoi(e) : 1; }();
    ConvPlan cp;
    if (int rc = conv3x3_plan(c.n, c.h, c.w, P->cpad, kDefs[0].cout, c.fwd_terms, &cp)) return rc;
    if (env != 0 && !(c.variant & 4096) && c.fwd_terms == 3 && !cp.tall && (8 * cp.MT + 2) * 8 <= 256)
      P->xin16 = b.take<uint8_t>(npix0 * P->cpad * 4);
  }
  P->xin_split = (P->xin16 != nullptr && c.training && !(c.variant & 128)) ? b.take<uint8_t>(npix0 * P->cpad * 4) : nullptr;
  P->dA_pred = b.take<float>(npix0 * 64);
  for (int l = 0; l < kLayers; ++l) {
    LayerBuf& B = P->L[l];
    const LayerDef& d = kDefs[l];
    B.H = c.h >> d.level; B.W = c.w >> d.level;
    B.cout = d.cout; B.cin = layer_cin(c, l); B.cin_real = (l == 0) ? c.in_dim : B.cin;
    const size_t npix = (size_t)c.n * B.H * B.W;
    B.z = b.take<float>(npix * B.cout);
    B.scale = b.take<float>(B.cout); B.shift = b.take<float>(B.cout);
    B.mean = b.take<float>(B.cout); B.invstd = b.take<float>(B.cout);
    B.stat_rows = conv3x3_num_stat_rows(c.n, B.H, B.W, B.cin, B.cout, c.fwd_terms);
    if (B.stat_rows < 0) return -2;
    B.stat_part = b.take<float>((size_t)B.stat_rows * 2 * B.cout);
    B.wf = b.take<uint16_t>(conv3x3_wpack_elems(B.cin, B.cout));
    if (c.training) {
      B.dz = b.take<float>(npix * B.cout);
      B.din = (l > 0) ? b.take<float>(npix * B.cin) : nullptr;
      B.bwd_rows = bn_bwd_num_blocks(c.n, B.H, B.W, B.cout);
      B.fused_rows = 0;
      // variant bit 64 switches the fusion ON. It is off by default: measured neutral on the bs-10 step (the dgrad
      // epilogue gets slower by about what the stand-alone reduction pass costs, profiles/r1_summary.md 6)
      if ((c.variant & 64) && bn_reduce_fusable(l)) {
        // dgrad of layer l + 1: K side = its cout, N side = its cin = this layer's cout, same resolution
        B.fused_rows = conv3x3_num_stat_rows(c.n, B.H, B.W, kDefs[l + 1].cout, B.cout, c.bwd_terms, true);
        if (B.fused_rows < 0) return -2;
      }
      B.bwd_part = b.take<float>((size_t)(B.fused_rows > B.bwd_rows ? B.fused_rows : B.bwd_rows) * 2 * B.cout);
      B.bwd_sums = b.take<float>(2 * B.cout);
      B.amax = P->amax_all + l;
      B.wd = (l > 0) ? b.take<uint16_t>(conv3x3_wpack_elems(B.cout, B.cin)) : nullptr;
      B.cat16 = nullptr;
      if (d.src1 >= 0 && d.mode0 == SRC_AFFINE_RELU_UP && d.mode1 == SRC_AFFINE_RELU) {
        const size_t up = (size_t)c.n * (c.h >> kDefs[d.src0].level) * (c.w >> kDefs[d.src0].level) * kDefs[d.src0].cout;
        B.cat16 = b.take<uint8_t>((up + npix * kDefs[d.src1].cout) * 4);
      }
    } else {
      B.dz = B.din = B.bwd_part = B.bwd_sums = B.amax = nullptr; B.wd = nullptr; B.bwd_rows = 0; B.fused_rows = 0;
      B.cat16 = nullptr;
    }
  }
  P->bytes = (b.off + 255) & ~(size_t)255;
  return 0;
}

SrcDesc make_src(const Plan& P, const tnb_tracknet_cfg_t& c, int layer, int mode) {
  SrcDesc s;
  if (layer < 0) {
    s.ptr = P.xin; s.scale = nullptr; s.shift = nullptr; s.C = P.cpad; s.Hs = c.h; s.Ws = c.w; s.mode = SRC_IDENTITY;
    if (mode == SRC_PLANAR16) { s.ptr = reinterpret_cast<const float*>(P.xin16); s.mode = SRC_PLANAR16; }
  } else {
    const LayerBuf& B = P.L[layer];
    s.ptr = B.z; s.scale = B.scale; s.shift = B.shift; s.C = B.cout; s.Hs = B.H; s.Ws = B.W; s.mode = mode;
  }
  return s;
}
ViewDesc make_view(const Plan& P, const tnb_tracknet_cfg_t& c, int l) {
  const LayerDef& d = kDefs[l];
  const LayerBuf& B = P.L[l];
  ViewDesc v;
  memset(&v, 0, sizeof(v));
  v.s[0] = make_src(P, c, d.src0, d.mode0);
  v.C0 = v.s[0].C;
  if (d.src1 >= 0) v.s[1] = make_src(P, c, d.src1, d.mode1);
  else v.s[1] = v.s[0];
  v.C = B.cin; v.N = c.n; v.H = B.H; v.W = B.W;
  return v;
}

struct CounterTable { long long* p[kLayers]; };
__global__ void inc_counters_kernel2(CounterTable t) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x < kLayers) *t.p[threadIdx.x] += 1;
}

}  // namespace

size_t tracknet_workspace_bytes(const tnb_tracknet_cfg_t& c) {
  Plan P;
  if (build_plan(c, nullptr, &P)) return 0;
  return P.bytes;
}

static int forward_enqueue(const tnb_tracknet_cfg_t& c, const float* x, void* const* params, float* y, void* ws,
                           size_t ws_bytes, cudaStream_t st) {
  Plan P;
  if (int rc = build_plan(c, ws, &P)) return rc;
  TNB_REQUIRE(ws_bytes >= P.bytes, "tracknet_forward: workspace too small (%zu < %zu)", ws_bytes, P.bytes);
  // fp32 NHWC for the first layer's weight gradient (training) or its gather path; planar fp16 pairs for tensor-TMA
  float* xin32 = ((c.training && P.xin_split == nullptr) || P.xin16 == nullptr) ? P.xin : nullptr;
  if (int rc = launch_pack_input(x, xin32, c.n, c.in_dim, c.h, c.w, P.cpad, st, P.xin16, P.xin_split, bwd_fmt(c))) return rc;
  // every weight operand of the step in ONE launch: the 17 forward images (fp16 hi/lo) and, when a backward will
  // follow, the 16 dgrad images (bf16 hi/lo, rotated / transposed) - the parameters do not change in between
  PackTable pt;
  pt.n = 0; pt.total = 0;
  for (int l = 0; l < kLayers; ++l) {
    LayerBuf& B = P.L[l];
    const float* w = (const float*)params[l * 6 + 0];
    ConvPlan cp;
    if (int rc = conv3x3_plan(c.n, B.H, B.W, B.cin, B.cout, c.fwd_terms, &cp)) return rc;
    if (int rc = pack_table_add(&pt, w, B.wf, B.cout, B.cin_real, 0, 0, cp.BN)) return rc;
    if (c.training && l > 0) {
      const bool fused = P.L[l - 1].fused_rows > 0;
      if (int rc = conv3x3_plan(c.n, B.H, B.W, B.cout, B.cin, c.bwd_terms, &cp, fused)) return rc;
      if (int rc = pack_table_add(&pt, w, B.wd, B.cout, B.cin, 1, bwd_fmt(c), cp.BN)) return rc;
    }
  }
  if (int rc = launch_pack_table(pt, st)) return rc;
  for (int l = 0; l < kLayers; ++l) {
    LayerBuf& B = P.L[l];
    ViewDesc v = make_view(P, c, l);
    if (l == 0 && P.xin16 != nullptr) { v.s[0] = make_src(P, c, -1, SRC_PLANAR16); v.s[1] = v.s[0]; }
    if (int rc = launch_conv3x3(v, B.wf, B.z, batch_stats(c) ? B.stat_part : nullptr, B.cout, c.fwd_terms, 0, c.variant & 3, st))
      return rc;
    if (int rc = launch_bn_finalize(B.stat_part, B.stat_rows, (double)c.n * B.H * B.W, (const float*)params[l * 6 + 1],
                                    (const float*)params[l * 6 + 2], (float*)params[l * 6 + 3],
                                    (float*)params[l * 6 + 4], c.bn_momentum, c.bn_eps, batch_stats(c) ? 1 : 0, B.scale, B.shift,
                                    B.mean, B.invstd, B.cout, st))
      return rc;
  }
  if (batch_stats(c)) {
    CounterTable t;
    for (int l = 0; l < kLayers; ++l) t.p[l] = (long long*)params[l * 6 + 5];
    if (int rc = launch_pdl(inc_counters_kernel2, dim3(1), dim3(32), 0, st, t)) return rc;
  }
  const SrcDesc last = make_src(P, c, kLayers - 1, SRC_AFFINE_RELU);
  return launch_predictor_fwd(last, c.n, c.h, c.w, (const float*)params[kLayers * 6 + 0],
                              (const float*)params[kLayers * 6 + 1], c.out_dim, y, st);
}

// Layers hi .. lo of the backward pass (hi == kLayers - 1: preceded by the predictor's backward). A whole pass is
// (kLayers - 1, 0); data-parallel training runs it as two ranges so that the gradients of the first one (bottleneck and
// decoder: 85 % of the parameters) can be all-reduced while the second one still computes.
static int backward_enqueue(const tnb_tracknet_cfg_t& c, const float* dy, const float* y, void* const* params,
                            void* const* grads, void* ws, size_t ws_bytes, cudaStream_t st, int hi = kLayers - 1,
                            int lo = 0) {
  TNB_REQUIRE(c.training != 0, "tracknet_backward: the forward call ran with training = 0 and kept no state for a backward pass "
                               "(training = 2 is the running-statistics forward that does)");
  TNB_REQUIRE(0 <= lo && lo <= hi && hi < kLayers, "tracknet_backward: layer range %d..%d", hi, lo);
  Plan P;
  if (int rc = build_plan(c, ws, &P)) return rc;
  TNB_REQUIRE(ws_bytes >= P.bytes, "tracknet_backward: workspace too small (%zu < %zu)", ws_bytes, P.bytes);
  const int bf = bwd_fmt(c);
  float* gmax_all = P.amax_all;            // [kLayers]
  float* mul_all = P.amax_all + kLayers;   // [kLayers]
  if (hi == kLayers - 1) {
    const SrcDesc last = make_src(P, c, kLayers - 1, SRC_AFFINE_RELU);
    if (int rc = launch_predictor_bwd(last, c.n, c.h, c.w, (const float*)params[kLayers * 6 + 0], c.out_dim, dy, y,
                                      P.dA_pred, (float*)grads[kLayers * 3 + 0], (float*)grads[kLayers * 3 + 1],
                                      P.pred_ws, st))
      return rc;
    if (bf == 0) TNB_CHECK_CUDA(cudaMemsetAsync(gmax_all, 0, sizeof(float) * kLayers, st));
  }
  auto run_wgrad = [&](int layer, const ViewDesc& pv) -> int {
    LayerBuf& W = P.L[layer];
    float* dw = (float*)grads[layer * 3 + 0];
    const float* mul = bf == 0 ? mul_all + layer : nullptr;
    if (c.variant & 512) {  // variant bit 512: split-K partials added straight into dw with atomics (ablation; not deterministic)
      TNB_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)W.cout * W.cin_real * 9, st));
      return launch_wgrad3x3(pv, W.dz, dw, W.cout, W.cin_real, c.bwd_terms, (c.variant >> 2) & 3, st, nullptr, bf, mul);
    }
    return launch_wgrad3x3(pv, W.dz, dw, W.cout, W.cin_real, c.bwd_terms, (c.variant >> 2) & 3, st, P.wgrad_ws, bf, mul);
  };
  // layer whose wgrad waits for its operand from the next BatchNorm-backward apply pass. Nothing waits across the end
  // of a range: the last layer of a range that is not the last range materialises its view with a view pass instead
  // (same values bit for bit), so that all gradients of the range are final when the call returns
  int pending = -1;
  // Whole-pass backward: the wgrad of a decoder concat layer j waits until the apply pass of its SKIP producer (an
  // encoder layer, many layers later) has written the second half of its operand too - no view pass at all for it; its
  // operand lives in the layer's own buffer (cat16) meanwhile. A ranged backward (data parallel) keeps the view pass for
  // the skip half: every gradient of a range has to be final when the range returns. Variant bit 16384: never defer.
  const bool defer_skip = hi == kLayers - 1 && lo == 0 && !(c.variant & (128 | 8192 | 16384));
  int deferred_for[kLayers];  // skip producer layer -> concat layer waiting for it (-1: none)
  for (int i = 0; i < kLayers; ++i) deferred_for[i] = -1;
  for (int l = hi; l >= lo; --l) {
    LayerBuf& B = P.L[l];
    BnBwdArgs a;
    memset(&a, 0, sizeof(a));
    // consumers of this layer's activation = every later layer that lists it as a source (+ the predictor)
    if (l == kLayers - 1) {
      a.g[a.ng++] = GradSrc{P.dA_pred, 64, 0, GRAD_SAME, c.h, c.w};
    }
    for (int j = l + 1; j < kLayers; ++j) {
      const LayerDef& d = kDefs[j];
      const LayerBuf& J = P.L[j];
      if (d.src0 == l) {
        const int m = d.mode0 == SRC_AFFINE_RELU_POOL ? GRAD_POOL : d.mode0 == SRC_AFFINE_RELU_UP ? GRAD_UP : GRAD_SAME;
        TNB_REQUIRE(a.ng < 2, "tracknet_backward: too many consumers");
        a.g[a.ng++] = GradSrc{J.din, J.cin, 0, m, J.H, J.W};
      }
      if (d.src1 == l) {
        TNB_REQUIRE(a.ng < 2, "tracknet_backward: too many consumers");
        a.g[a.ng++] = GradSrc{J.din, J.cin, kDefs[d.src0].cout, GRAD_SAME, J.H, J.W};
      }
    }
    a.z = B.z; a.scale = B.scale; a.shift = B.shift; a.mean = B.mean; a.invstd = B.invstd;
    a.N = c.n; a.H = B.H; a.W = B.W; a.C = B.cout;
    a.part = B.bwd_part; a.sums = B.bwd_sums; a.dz = B.dz; a.amax = nullptr;
    a.dz_format = bf == 0 ? 2 : 1;  // dz -> pre-split fp16 (scaled) / bf16
    a.gmax = bf == 0 ? gmax_all + l : nullptr; a.dz_mul = bf == 0 ? mul_all + l : nullptr;
    // running statistics (training == 2): mean and variance are constants of the layer, so the two batch-statistics terms
    // of dz = scale * (g - mean(g) - xhat * mean(g * xhat)) vanish - the apply pass multiplies both sums by inv_count
    a.inv_count = batch_stats(c) ? (float)(1.0 / ((double)c.n * B.H * B.W)) : 0.f;
    if (B.fused_rows == 0)
      if (int rc = launch_bn_bwd_reduce(a, st)) return rc;
    if (int rc = launch_bn_bwd_finalize(B.bwd_part, B.fused_rows > 0 ? B.fused_rows : B.bwd_rows, B.cout, B.bwd_sums,
                                        (float*)grads[l * 3 + 1], (float*)grads[l * 3 + 2], st))
      return rc;
    const bool pend_cat = pending == l + 1 && kDefs[pending].src1 >= 0;  // the waiting layer is a decoder concat
    a.act_presplit = (pending == l + 1) ? ((pend_cat && defer_skip) ? P.L[pending].cat16 : P.vsplit) : nullptr;
    a.act_pool = (pending == l + 1 && kDefs[pending].mode0 == SRC_AFFINE_RELU_POOL) ? 1 : 0;
    const int waiting = deferred_for[l];  // concat layer whose skip half is this layer's activation
    if (waiting >= 0) {
      const ViewDesc wv = wgrad_view(c, waiting, P.L[waiting].cat16);
      a.act_full = const_cast<float*>(wv.s[1].ptr);
      TNB_REQUIRE(a.act_presplit == nullptr || a.act_pool, "tracknet_backward: skip producer %d feeds a plain layer too", l);
    }
    if (int rc = launch_bn_bwd_apply(a, st)) return rc;
    if (pending == l + 1) {
      if (pend_cat && defer_skip) {
        deferred_for[kDefs[pending].src1] = pending;  // up half written; the skip half follows with that layer's apply pass
      } else {
        const ViewDesc pv = wgrad_view(c, pending, P.vsplit);
        if (pv.C0 < pv.C) {  // decoder concat: the skip half (second source) from a view pass, behind the half just written
          const ViewDesc v = make_view(P, c, pending);
          ViewDesc v1 = v;
          v1.s[0] = v.s[1]; v1.s[1] = v.s[1]; v1.C0 = v1.C = v.C - v.C0;
          if (int rc = launch_view_presplit(v1, const_cast<float*>(pv.s[1].ptr), bf, st)) return rc;
        }
        if (int rc = run_wgrad(pending, pv)) return rc;
      }
      pending = -1;
    }
    if (waiting >= 0) {
      if (int rc = run_wgrad(waiting, wgrad_view(c, waiting, P.L[waiting].cat16))) return rc;
      deferred_for[l] = -1;
    }
    if (l > 0) {
      ViewDesc dv;
      memset(&dv, 0, sizeof(dv));
      dv.s[0] = SrcDesc{B.dz, bf == 0 ? mul_all + l : nullptr, nullptr, B.cout, B.H, B.W, SRC_PRESPLIT};
      dv.s[1] = dv.s[0];
      dv.C0 = dv.C = B.cout; dv.N = c.n; dv.H = B.H; dv.W = B.W;
      const LayerBuf* Pp = (l > 0 && P.L[l - 1].fused_rows > 0) ? &P.L[l - 1] : nullptr;  // producer whose reduction rides along
      BnBwdFuse fuse{};
      if (Pp != nullptr) fuse = BnBwdFuse{Pp->z, Pp->scale, Pp->shift, Pp->mean, Pp->invstd};
      // B.wd: the dgrad weight image, packed by the forward call of this step
      if (int rc = launch_conv3x3(dv, B.wd, B.din, Pp != nullptr ? Pp->bwd_part : nullptr, B.cin, c.bwd_terms, bf,
                                  c.variant & 3, st, Pp != nullptr ? &fuse : nullptr))
        return rc;
    }
    if (wgrad_operand_from_bn_bwd(c, l) && !(l == lo && lo > 0)) { pending = l; continue; }
    if (l == 0 && P.xin_split != nullptr) {  // the network input, pre-split by the forward's pack launch
      if (int rc = run_wgrad(0, wgrad_view(c, 0, P.xin_split))) return rc;
      continue;
    }
    // materialise the input view once (bf16 hi/lo), then both wgrad operands are plain copies
    const ViewDesc v = make_view(P, c, l);
    const ViewDesc pv = wgrad_view(c, l, P.vsplit);
    if (pv.C0 < pv.C) {  // decoder concat [up(x), skip]: two tensors, see wgrad_view
      ViewDesc v0 = v, v1 = v;
      v0.s[0].mode = SRC_AFFINE_RELU; v0.s[1] = v0.s[0]; v0.C0 = v0.C = v.C0; v0.H = v.s[0].Hs; v0.W = v.s[0].Ws;
      v1.s[0] = v.s[1]; v1.s[1] = v.s[1]; v1.C0 = v1.C = v.C - v.C0;
      if (int rc = launch_view_presplit(v0, const_cast<float*>(pv.s[0].ptr), bf, st)) return rc;
      if (int rc = launch_view_presplit(v1, const_cast<float*>(pv.s[1].ptr), bf, st)) return rc;
    } else {
      if (int rc = launch_view_presplit(v, P.vsplit, bf, st)) return rc;
    }
    if (int rc = run_wgrad(l, pv)) return rc;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// CUDA-graph replay of the launch sequences. One train step is ~60 (forward) + ~110 (backward) dependent launches
// of 3-400 us kernels plus memsets; issued one by one the stream leaves ~2 ms of gaps per 24 ms step. Every argument
// of every launch is a function of (cfg, the pointers passed in), so the sequence is captured once per distinct
// argument set - on its second sighting, the first run stays eager and warms the function attributes - and replayed
// while the caller keeps handing in the same buffers (training loops do: parameters are fixed, the caching allocator
// returns the same workspace / gradient blocks every step). A small LRU keeps the alternating input buffers of a
// double-buffered loader resident. TNB_GRAPHS=0 or active per-launch profiling bypass all of it.
// ---------------------------------------------------------------------------------------------
namespace {
struct GraphEntry { uint64_t key; cudaGraphExec_t exec; uint64_t stamp; };
struct GraphCache {
  std::mutex mu;
  std::vector<GraphEntry> entries;   // LRU, at most kMaxGraphs
  std::vector<uint64_t> seen;        // keys that ran eagerly once (ring)
  uint64_t clock = 0;
  int failures = 0;                  // capture failures: after a few, stop trying
  long long captures = 0, replays = 0, eager = 0;
  cudaStream_t side[64] = {};        // per-device capture stream (the caller's stream may be the legacy default stream,
                                     // which cannot be captured); the instantiated graph is launched into the caller's
};
constexpr int kMaxGraphs = 16, kMaxSeen = 64;
GraphCache g_graphs;

int g_graph_switch = [] { const char* e = getenv("TNB_GRAPHS"); return e ? (atoi(e) != 0 ? 1 : 0) : 1; }();
bool graphs_enabled() { return g_graph_switch != 0 && !prof_enabled() && g_graphs.failures < 3; }
uint64_t fnv(uint64_t h, const void* data, size_t n) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}
template <typename F>
int run_maybe_graphed(uint64_t key, cudaStream_t st, F&& enqueue) {
  if (!graphs_enabled()) return enqueue(st);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread &&
      (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)) {
    cudaGetLastError();
    return enqueue(st);  // the caller is capturing this stream itself
  }
  std::lock_guard<std::mutex> lock(g_graphs.mu);
  GraphCache& G = g_graphs;
  ++G.clock;
  for (auto& e : G.entries)
    if (e.key == key) {
      e.stamp = G.clock;
      ++G.replays;
      TNB_CHECK_CUDA(cudaGraphLaunch(e.exec, st));
      return 0;
    }
  bool second = false;
  for (uint64_t k : G.seen) second |= (k == key);
  if (!second) {
    if ((int)G.seen.size() < kMaxSeen) G.seen.push_back(key); else G.seen[G.clock % kMaxSeen] = key;
    ++G.eager;
    return enqueue(st);
  }
  // second sighting: capture on this device's side stream, instantiate, launch into the caller's stream
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); ++G.eager; return enqueue(st); }
  cudaStream_t& side = G.side[dev];
  if (side == nullptr && cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError(); ++G.failures; ++G.eager; side = nullptr;
    return enqueue(st);
  }
  if (cudaStreamBeginCapture(side, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError(); ++G.failures; ++G.eager;
    return enqueue(st);
  }
  const int rc = enqueue(side);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(side, &graph);
  if (rc != 0 || ce != cudaSuccess || graph == nullptr) {
    cudaGetLastError(); ++G.failures;
    if (graph) cudaGraphDestroy(graph);
    if (rc != 0) return rc;
    ++G.eager;
    return enqueue(st);  // nothing was executed during the failed capture
  }
  cudaGraphExec_t exec = nullptr;
  if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
    cudaGetLastError(); ++G.failures; ++G.eager;
    cudaGraphDestroy(graph);
    return enqueue(st);
  }
  ++G.captures;
  cudaGraphDestroy(graph);
  if ((int)G.entries.size() >= kMaxGraphs) {
    size_t lru = 0;
    for (size_t i = 1; i < G.entries.size(); ++i) if (G.entries[i].stamp < G.entries[lru].stamp) lru = i;
    cudaGraphExecDestroy(G.entries[lru].exec);
    G.entries.erase(G.entries.begin() + lru);
  }
  G.entries.push_back(GraphEntry{key, exec, G.clock});
  TNB_CHECK_CUDA(cudaGraphLaunch(exec, st));
  return 0;
}
uint64_t key_common(const tnb_tracknet_cfg_t& c, int kind, void* ws, size_t ws_bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  uint64_t h = 1469598103934665603ull;
  h = fnv(h, &c, sizeof(c)); h = fnv(h, &kind, sizeof(kind)); h = fnv(h, &dev, sizeof(dev));
  h = fnv(h, &ws, sizeof(ws)); h = fnv(h, &ws_bytes, sizeof(ws_bytes));
  return h;
}
}  // namespace

int tracknet_forward(const tnb_tracknet_cfg_t& c, const float* x, void* const* params, float* y, void* ws,
                     size_t ws_bytes, cudaStream_t st) {
  uint64_t key = key_common(c, 0, ws, ws_bytes);
  key = fnv(key, &x, sizeof(x)); key = fnv(key, &y, sizeof(y));
  key = fnv(key, params, sizeof(void*) * (kLayers * 6 + 2));
  return run_maybe_graphed(key, st, [&](cudaStream_t s) { return forward_enqueue(c, x, params, y, ws, ws_bytes, s); });
}

int tracknet_backward(const tnb_tracknet_cfg_t& c, const float* dy, const float* y, void* const* params,
                      void* const* grads, void* ws, size_t ws_bytes, cudaStream_t st, int hi, int lo) {
  TNB_REQUIRE(0 <= lo && lo <= hi && hi < kLayers, "tracknet_backward: layer range %d..%d", hi, lo);
  uint64_t key = key_common(c, 1 + (hi << 8) + (lo << 16), ws, ws_bytes);
  key = fnv(key, &dy, sizeof(dy)); key = fnv(key, &y, sizeof(y));
  key = fnv(key, params, sizeof(void*) * (kLayers * 6 + 2));
  key = fnv(key, grads, sizeof(void*) * (kLayers * 3 + 2));
  return run_maybe_graphed(key, st, [&](cudaStream_t s) {
    return backward_enqueue(c, dy, y, params, grads, ws, ws_bytes, s, hi, lo);
  });
}

// First layer of the backward pass's second range in data-parallel training (tnb_tracknet_backward_range): layers
// 7..16 + the predictor hold 9.59 M of TrackNet's 11.34 M parameters.
int tracknet_grad_split_layer() { return 7; }

void graph_stats(long long* out4) {
  std::lock_guard<std::mutex> lock(g_graphs.mu);
  out4[0] = g_graphs.captures; out4[1] = g_graphs.replays; out4[2] = g_graphs.eager; out4[3] = g_graphs.failures;
}

int set_graph_replay(int on) {
  const int prev = g_graph_switch;
  g_graph_switch = on ? 1 : 0;
  return prev;
}

int tracknet_debug_layer(const tnb_tracknet_cfg_t& c, void* ws, int layer, void** out_ptr8, int* out_dim5) {
  TNB_REQUIRE(layer >= 0 && layer < kLayers, "tracknet_debug_layer: layer %d", layer);
  Plan P;
  if (int rc = build_plan(c, ws, &P)) return rc;
  const LayerBuf& B = P.L[layer];
  const int bf = bwd_fmt(c);
  void* ptrs[8] = {B.z, B.scale, B.shift, B.mean, B.invstd, B.dz, B.din,
                   (c.training && bf == 0) ? (void*)(P.amax_all + kLayers + layer) : nullptr};
  for (int i = 0; i < 8; ++i) out_ptr8[i] = ptrs[i];
  out_dim5[0] = B.H; out_dim5[1] = B.W; out_dim5[2] = B.cin; out_dim5[3] = B.cout; out_dim5[4] = bf;
  return 0;
}

int tracknet_num_launches(const tnb_tracknet_cfg_t& c, int backward) {
  const int pred_groups = (c.out_dim + 15) / 16;  // the predictor kernels take 16 output channels per launch
  // pack_input, pack_weights (all tensors), (conv, bn_finalize) x17, counters, predictor
  if (!backward) return 1 + 1 + kLayers * 2 + (batch_stats(c) ? 1 : 0) + pred_groups;
  int fused = 0;
  for (int l = 0; l < kLayers; ++l) fused += ((c.variant & 64) && bn_reduce_fusable(l)) ? 1 : 0;
  int emitted = 0;  // wgrad operands written by the BatchNorm-backward apply pass: no view_presplit launch
  for (int l = 0; l < kLayers; ++l) emitted += wgrad_operand_from_bn_bwd(c, l) ? 1 : 0;
  const int scatter = (c.variant & 512) ? 0 : kLayers;  // split-K slabs -> OIHW, one ordered-sum launch per layer
  // predictor backward (dA + dW per group, final sum), (reduce [unless fused into the next layer's dgrad], finalize,
  // apply, wgrad) x17, view_presplit for the layers whose operand the apply pass did not emit (+ the second tensor of
  // the 3 decoder concat layers), dgrad x16, scatter
  // (whole pass: the skip halves of the 3 decoder concats come out of apply passes as well unless a variant bit says no)
  const int skip_views = (c.variant & (128 | 8192 | 16384)) ? 3 : 0;
  Plan P;  // does the pack launch of the forward pre-split the network input (no view pass for the first layer)?
  const bool input_presplit = build_plan(c, nullptr, &P) == 0 && P.xin_split != nullptr;
  return 2 * pred_groups + 1 + kLayers * 4 - fused + (kLayers - emitted) - (input_presplit ? 1 : 0) + skip_views +
         (kLayers - 1) + scatter;
}

}  // namespace tnb
