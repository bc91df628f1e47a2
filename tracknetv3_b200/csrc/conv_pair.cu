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
#define TNB_CK_LAUNCH launch_conv3x3_pair
#define TNB_CK_PAIR 1
#define TNB_CK_LEAN 0
#include "conv_kernel.inc"

}  // namespace tnb
