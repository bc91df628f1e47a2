// Lean-issue variant of the single-CTA tcgen05 3x3 convolution kernel - EXPERIMENTAL, off by default (TNB_CONV_LEAN=1),
// not yet run on a GPU. Same kernel body as conv.cu (conv_kernel.inc) with TNB_CK_LEAN 1: the loop in which the one
// elected thread issues the MMAs is specialised on the chain length and the product form and steps 32-bit descriptor
// words. Why: on the 64-wide layers an MMA lasts 32-64 clocks, the generic issue loop costs ~115 clocks per K step for
// 2-3 of them, and the measured times say the two ADD instead of overlapping (profiles/r1_final.md section 10,
// DESIGN.md 6b). Its own translation unit keeps the shipped kernel's code byte-identical until this one is measured.
#include "igemm.cuh"
#include <cstdlib>
#include "prof.cuh"
#include <type_traits>

namespace tnb {

#define TNB_CK_NAME conv3x3_lean_kernel
#define TNB_CK_ARGS ConvLeanArgs
#define TNB_CK_PAIR 0
#define TNB_CK_LEAN 1
#include "conv_kernel.inc"

int launch_conv3x3_lean(const ViewDesc& view, const uint16_t* wpack, float* out, float* stat_part, int Cout, int nterms,
                        int fmt, int variant, const ConvPlan& p, cudaStream_t st) {
  const int m0 = view.s[0].mode, m1 = (view.C0 < view.C) ? view.s[1].mode : view.s[0].mode;
  ConvLeanArgs a;
  a.view = view; a.wpack = wpack; a.out = out; a.stat_part = stat_part;
  a.bz = a.bsc = a.bsh = a.bmu = a.bis = nullptr;
  a.Cout = Cout; a.BN = p.BN; a.MT = p.MT; a.SA = p.SA; a.SB = p.SB; a.G = p.G; a.nbuf = p.nbuf; a.nterms = nterms;
  a.variant = variant | (cp_async_ca_env() ? 256 : 0); a.tmem_cols = p.tmem_cols; a.tiles_h = p.tiles_h; a.tiles_w = p.tiles_w;
  a.tall = p.tall; a.merged = p.merged;
  a.ntiles = view.N * p.tiles_h * p.tiles_w;
  a.nwork = a.ntiles * (Cout / p.BN);
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.nwork < sms ? a.nwork : sms;
  ProfScope prof(view.s[0].mode == SRC_PRESPLIT ? PROF_CONV_DGRAD : PROF_CONV_FWD, st, view.N, view.H, view.W, view.C, Cout);
  auto go = [&](auto kern) -> int {
    TNB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    kern<<<grid, kThreads, p.smem_bytes, st>>>(a);
    return 0;
  };
  int rc2 = -2;
#define TNB_CONVL_CASE(F, A, B) if (fmt == F && m0 == A && m1 == B) rc2 = go(conv3x3_lean_kernel<F, A, B, false>); else
  TNB_CONVL_CASE(0, SRC_IDENTITY, SRC_IDENTITY)
  TNB_CONVL_CASE(0, SRC_AFFINE_RELU, SRC_AFFINE_RELU)
  TNB_CONVL_CASE(0, SRC_AFFINE_RELU_POOL, SRC_AFFINE_RELU_POOL)
  TNB_CONVL_CASE(0, SRC_AFFINE_RELU_UP, SRC_AFFINE_RELU)
  TNB_CONVL_CASE(1, SRC_PRESPLIT, SRC_PRESPLIT)
  { tnb::set_last_error("conv3x3 (lean): unsupported (fmt %d, source modes %d/%d) combination", fmt, m0, m1); return -2; }
#undef TNB_CONVL_CASE
  if (rc2) return rc2;
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tnb
