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
#define TNB_CK_LAUNCH launch_conv3x3_lean
#define TNB_CK_PAIR 0
#define TNB_CK_LEAN 1
#include "conv_kernel.inc"

}  // namespace tnb
