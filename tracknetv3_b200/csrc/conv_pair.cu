// CTA-pair variant of the persistent tcgen05 implicit-GEMM 3x3 convolution (forward and dgrad) - EXPERIMENTAL, off by
// default (TNB_CONV_PAIR=1), not yet run on a GPU. Both kernels are conv_kernel.inc: conv.cu includes it with TNB_CK_PAIR 0
// (the validated single-CTA kernel, SASS unchanged), this file with TNB_CK_PAIR 1, which switches on the pair protocol
// of wgrad3x3_pair_kernel (wgrad.cu):
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

#define TNB_CK_NAME conv3x3_pair_kernel
#define TNB_CK_ARGS ConvPairArgs
#define TNB_CK_PAIR 1
#define TNB_CK_LEAN 0
#include "conv_kernel.inc"

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
#define TNB_CONVP_CASE(F, A, B) if (fmt == F && m0 == A && m1 == B) rc2 = go(conv3x3_pair_kernel<F, A, B, false>); else
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
