// CTA-pair kernel with the lean MMA-issue loop (conv_kernel.inc with TNB_CK_PAIR 1 and TNB_CK_LEAN 1) - EXPERIMENTAL,
// selected by TNB_CONV_PAIR=1 together with TNB_CONV_LEAN=1; see conv_pair.cu and conv_lean.cu.
#include "igemm.cuh"
#include <cstdlib>
#include "prof.cuh"
#include <type_traits>

namespace tnb {

#define TNB_CK_NAME conv3x3_pair_lean_kernel
#define TNB_CK_ARGS ConvPairLeanArgs
#define TNB_CK_PAIR 1
#define TNB_CK_LEAN 1
#include "conv_kernel.inc"

int launch_conv3x3_pair_lean(const ViewDesc& view, const uint16_t* wpack, float* out, float* stat_part, int Cout, int nterms,
                             int fmt, int variant, const ConvPlan& p, cudaStream_t st) {
  const int m0 = view.s[0].mode, m1 = (view.C0 < view.C) ? view.s[1].mode : view.s[0].mode;
  ConvPairLeanArgs a;
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
#define TNB_CONVPL_CASE(F, A, B) if (fmt == F && m0 == A && m1 == B) rc2 = go(conv3x3_pair_lean_kernel<F, A, B, false>); else
  TNB_CONVPL_CASE(0, SRC_IDENTITY, SRC_IDENTITY)
  TNB_CONVPL_CASE(0, SRC_AFFINE_RELU, SRC_AFFINE_RELU)
  TNB_CONVPL_CASE(0, SRC_AFFINE_RELU_POOL, SRC_AFFINE_RELU_POOL)
  TNB_CONVPL_CASE(0, SRC_AFFINE_RELU_UP, SRC_AFFINE_RELU)
  TNB_CONVPL_CASE(1, SRC_PRESPLIT, SRC_PRESPLIT)
  { tnb::set_last_error("conv3x3 (pair, lean): unsupported (fmt %d, source modes %d/%d) combination", fmt, m0, m1); return -2; }
#undef TNB_CONVPL_CASE
  if (rc2) return rc2;
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tnb
