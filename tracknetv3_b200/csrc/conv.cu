// Persistent tcgen05 implicit-GEMM 3x3 convolution (forward and dgrad) for sm_100a.
//
// Replaces cuDNN's nn.Conv2d(3x3, padding='same', bias=False) forward (reference model.py:8,13) and its
// autograd dgrad (reference train.py:95), with BatchNorm-apply + ReLU (model.py:14-15), MaxPool2d
// (model.py:59,61,63) and Upsample + torch.cat (model.py:65,67,69) of the producing layers fused into the
// operand gather, and the BatchNorm statistics of the output fused into the epilogue.
//
// One CTA per SM loops over (pixel tile, channel tile) work items:
//   warp 0      one elected thread issues tcgen05.mma (kind::f16, M=128, N=BN, K=16; 3 MMAs per K step in the
//               fp32-faithful hi/lo split mode) into one of up to two TMEM accumulator buffers
//   warp 1      weight tiles: 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx) of pre-packed smem images
//   warps 2-5   epilogue: tcgen05.ld -> fp32 NHWC stores + per-channel (sum, sumsq) partials; overlaps the
//               next tile's MMAs when two accumulator buffers fit in TMEM (2*MT*BN <= 512 columns)
//   warps 6-11  operand producers: batched global gathers -> BN affine/ReLU/pool/upsample/concat ->
//               16-bit hi/lo split -> planar smem halo tile -> fence.proxy.async -> mbarrier
// The planar tile [plane = 8 channels][pixel][16 B] is a SWIZZLE_NONE K-major operand in which every 3x3 tap
// is just a different 16-byte aligned start address: one halo tile serves all 9 taps.
#include "igemm.cuh"
#include <cstdlib>
#include "prof.cuh"
#include <type_traits>

namespace tnb {

#define TNB_CK_NAME conv3x3_kernel
#define TNB_CK_ARGS ConvArgs
#define TNB_CK_LAUNCH launch_conv3x3_generic
#define TNB_CK_LEAN 0
#include "conv_kernel.inc"

static int pick_bn(int nside) {
  if (nside % 256 == 0) return 256;
  if (nside % 192 == 0) return 192;
  if (nside % 128 == 0) return 128;
  if (nside % 64 == 0) return 64;
  if (nside % 32 == 0) return 32;
  return 0;
}
static int pow2_cols(int c) {
  int p = 32;
  while (p < c) p <<= 1;
  return p;
}
int make_planar16_tmap(CUtensorMap* out, const void* base, int N, int H, int W, int C, int box_w, int box_h) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    EncodeFn f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return f;
  }();
  TNB_REQUIRE(encode != nullptr, "tensor-TMA: the driver does not export cuTensorMapEncodeTiled");
  TNB_REQUIRE(C % 32 == 0 && box_w * 8 <= 256 && box_h <= 256 && (reinterpret_cast<uintptr_t>(base) & 15) == 0,
              "tensor-TMA: bad planar16 tensor (C %d, box %d x %d)", C, box_w, box_h);
  const cuuint64_t planes = (cuuint64_t)C / 32 * 8;  // per chunk: hi 4 | lo 4 planes of 8 channels
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, planes, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * 16 * planes};  // bytes, dims 1..3
  cuuint32_t box[4] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_h, 8, 1}, es[4] = {1, 1, 1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TNB_REQUIRE(r == CUDA_SUCCESS, "tensor-TMA: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

// 64-wide output tiles: an M = 128, N = 64 MMA reads (128 + 64) x 32 B of shared memory (48 clocks at 128 B/clk) for 32
// clocks of math, so the three MMAs of a split product cost 144 clocks per K step. With the weights packed
// [plane][hi | lo][64 rows], x_hi * [w_hi | w_lo] is ONE MMA of width 128 (64 clocks of math = 64 of operand reads)
// and x_lo * w_hi a second one: 112 clocks, at the price of 2 x 64 accumulator columns per tile (MT <= 2).
// TNB_CONV_MERGE=0 restores the three-MMA form (ablation).
bool conv3x3_merged(int BN) {
  static const int on = [] { const char* e = getenv("TNB_CONV_MERGE"); return e ? atoi(e) : 1; }();
  return on && BN == 64;
}
int conv3x3_weight_layout(int BN) { return conv3x3_merged(BN) ? 1 : 0; }

int conv3x3_plan(int N, int H, int W, int Cin, int Cout, int nterms, ConvPlan* plan, bool bn_bwd_fused, bool copy_fill) {
  TNB_REQUIRE(Cin % 32 == 0, "conv3x3: view channels %d must be a multiple of 32", Cin);
  const int BN = pick_bn(Cout);
  TNB_REQUIRE(BN >= 32 && BN % 16 == 0, "conv3x3: unsupported output channel count %d", Cout);
  const int TP = nterms > 1 ? 2 : 1;
  const bool merged = conv3x3_merged(BN);
  const int ACCW = (merged && nterms > 1) ? 2 * BN : BN;  // accumulator columns per M tile
  int MT = 512 / ACCW;
  if (MT > 4) MT = 4;
  // Two accumulator buffers in TMEM (2 * MT * BN <= 512 columns) let the epilogue of tile i overlap the MMAs of tile
  // i + 1; with one buffer it is exposed (7-17 % of the layer, measured: profiles/r1_summary.md 6). Narrower tiles
  // re-stream the weights twice as often; with the tile orientation below removing the padded tile rows that still
  // pays even for the 768-channel layer (0.669 -> 0.604 ms). TNB_CONV_PLAN=0 restores the widest tile, =2 keeps the
  // double buffering but forces the row-major orientation (tools/ablate_plan.py).
  static const int plan_mode = [] { const char* e = getenv("TNB_CONV_PLAN"); return e ? atoi(e) : 1; }();
  if (plan_mode != 0) while (MT > 1 && 2 * MT * ACCW > 512) MT >>= 1;
  // Orientation: M = 128 rows of the MMA are 16 groups of 8 consecutive pixels. Groups along W stacked over 16 image rows
  // give a 16 x 8*MT tile; groups along H stacked over 16 image columns give an 8*MT x 16 tile. Take the one that pads
  // the image less: at 72 x 128 and 36 x 64 (H = 4.5 and 2.25 tiles of 16 rows) the tall-group tile wastes 0 / 10 %
  // of the MMAs instead of 10 / 25 %. TNB_CONV_PLAN=2 forces the first form.
  auto padded = [&](int mt, bool tall) {
    const long long th = tall ? (H + 8 * mt - 1) / (8 * mt) * (8 * mt) : (H + 15) / 16 * 16;
    const long long tw = tall ? (W + 15) / 16 * 16 : (W + 8 * mt - 1) / (8 * mt) * (8 * mt);
    return th * tw;
  };
  const bool tall = plan_mode != 2 && padded(MT, true) < padded(MT, false);
  while (MT > 1 && 8 * (MT - 1) >= (tall ? H : W)) --MT;  // do not tile wider (taller) than the image
  // Halo-tile stages: 2 for the gathering (forward) views, 3 for the copy-filled dgrad views. A/B on one box, two runs
  // each (profiles/r2_sa_summary.txt): the third stage lets the producers start the next tile's second chunk before the
  // current tile's has retired - dgrad 5.40 -> 5.18 ms per step, every launch 3-9 % faster; on the forward launches it
  // costs the gathers 42 KB of L1 (64 -> 64 +5 %, 64 -> 128 +10 %, the rest unchanged). TNB_CONV_SA=<forward><dgrad>,
  // two digits 2..4, overrides (ablation).
  static const int sa_env = [] { const char* e = getenv("TNB_CONV_SA"); return e ? atoi(e) : 0; }();
  const int sa_pick = copy_fill ? sa_env % 10 : sa_env / 10;
  const int sa_want = (sa_pick >= 2 && sa_pick <= 4) ? sa_pick : (copy_fill ? 3 : 2);
  // The tiling (MT, taps per weight stage) is chosen with 2 stages and never depends on the wish for more: per-tile
  // partial rows and the packed-weight layout are computed by callers that do not know the fill kind. Extra stages are
  // taken afterwards if they fit next to at least two weight stages.
  int SA = 2, SB = 0, G = 1;
  size_t smem = 0;
  const int mt_max = MT;
  bool found = false;
  // pass 0: narrow tiles (BN <= 128) get weight stages holding a whole filter row (3 taps) so that 18+ MMAs chain into
  // one accumulator; pass 1: one tap per stage, as many stages as fit
  for (int pass = (BN <= 128 ? 0 : 1); pass < 2 && !found; ++pass) {
    const int g = pass == 0 ? 3 : 1;
    for (MT = mt_max; MT >= 1; --MT) {
      const int pitch = 8 * MT + 2, halo = 18 * pitch;
      const size_t a_stage = (size_t)TP * 4 * pad_px(halo) * 16;
      const size_t b_stage = (size_t)g * (merged ? 2 : TP) * 64 * BN;
      const size_t fixed = kHdrBytes + ((halo * 8 + 127) & ~127) + (size_t)4 * 2 * BN * 4 + (bn_bwd_fused ? BN * 16 : 0) +
                           2 * a_stage;
      if (fixed + 2 * b_stage <= (size_t)kMaxSmem) {
        G = g;
        SB = (int)((kMaxSmem - fixed) / b_stage);
        // deeper rings measured no better (cap 5: round 1, and again with the lean issue loop, TNB_CONV_SB_CAP), one tap
        // per slot clearly worse
        static const int cap_env = [] { const char* e = getenv("TNB_CONV_SB_CAP"); return e ? atoi(e) : 0; }();
        const int cap = g == 3 ? (cap_env >= 2 && cap_env <= 8 ? cap_env : 3) : 8;
        while (SA < sa_want && fixed + (SA + 1 - 2) * a_stage + 2 * b_stage <= (size_t)kMaxSmem) ++SA;
        const size_t fixed_sa = fixed + (SA - 2) * a_stage;
        SB = (int)((kMaxSmem - fixed_sa) / b_stage);
        if (SB > cap) SB = cap;
        smem = fixed_sa + SB * b_stage;
        found = true;
        break;
      }
    }
  }
  TNB_REQUIRE(found, "conv3x3: no shared-memory plan for Cin=%d Cout=%d", Cin, Cout);
  plan->BN = BN; plan->MT = MT; plan->SA = SA; plan->SB = SB; plan->G = G;
  plan->nbuf = (2 * MT * ACCW <= 512) ? 2 : 1;
  plan->tmem_cols = pow2_cols(plan->nbuf * MT * ACCW);
  plan->merged = merged ? 1 : 0;
  plan->smem_bytes = smem;
  plan->tall = tall ? 1 : 0;
  plan->tiles_h = tall ? (H + 8 * MT - 1) / (8 * MT) : (H + 15) / 16;
  plan->tiles_w = tall ? (W + 15) / 16 : (W + 8 * MT - 1) / (8 * MT);
  (void)N;
  return 0;
}

int conv3x3_num_stat_rows(int N, int H, int W, int Cin, int Cout, int nterms, bool bn_bwd_fused) {
  ConvPlan p;
  if (conv3x3_plan(N, H, W, Cin, Cout, nterms, &p, bn_bwd_fused)) return -1;
  return N * p.tiles_h * p.tiles_w;
}

int launch_conv3x3(const ViewDesc& view, const uint16_t* wpack, float* out, float* stat_part, int Cout,
                   int nterms, int fmt, int variant, cudaStream_t st, const BnBwdFuse* fuse) {
  ConvPlan p;
  const bool dgrad = view.s[0].mode == SRC_PRESPLIT;  // copy fill (pre-split gradients); the forward gathers and splits
  int rc = conv3x3_plan(view.N, view.H, view.W, view.C, Cout, nterms, &p, fuse != nullptr, dgrad);
  if (rc) return rc;
  // Which MMA-issue loop (conv_kernel.inc): the lean one for the forward pass and wherever the tile is 64 wide (short MMAs:
  // the generic loop's ~20 instructions per MMA fall behind), the generic one on the wide dgrad tiles, where it measured
  // ~5 % faster (per-launch A/B on one box, profiles/r2_experiments.md). TNB_CONV_LEAN=0 / 1 forces one of them (ablation).
  static const int lean_env = [] { const char* e = getenv("TNB_CONV_LEAN"); return e ? atoi(e) : -1; }();
  const bool lean = !dgrad || (lean_env >= 0 ? lean_env != 0 : p.BN == 64);  // the switch acts on the dgrad launches
  return lean ? launch_conv3x3_lean(view, wpack, out, stat_part, Cout, nterms, fmt, variant, p, st, fuse)
              : launch_conv3x3_generic(view, wpack, out, stat_part, Cout, nterms, fmt, variant, p, st, fuse);
}

}  // namespace tnb
