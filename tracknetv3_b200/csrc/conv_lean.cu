// Lean-issue variant of the tcgen05 3x3 convolution kernel: the same kernel body as conv.cu (conv_kernel.inc) with
// TNB_CK_LEAN 1 - the loop in which the one elected thread issues the MMAs is specialised on the chain length and the
// product form and steps 32-bit descriptor words (3-4 instructions per MMA instead of ~20). launch_conv3x3 (conv.cu) uses
// it for the forward pass and for every 64-wide tile: measured on one box against the generic loop, 192 -> 64 forward
// 1.097 -> 0.902 ms, 64 -> 64 forward 0.443 -> 0.422, 64-wide dgrad 0.385 -> 0.373; the wide dgrad tiles are ~5 % faster
// with the generic loop and keep it (profiles/r2_experiments.md). Its own translation unit: the two variants compile in
// parallel and neither changes the other's register allocation.
#include "igemm.cuh"
#include <cstdlib>
#include "prof.cuh"
#include <type_traits>

namespace tnb {

#define TNB_CK_NAME conv3x3_lean_kernel
#define TNB_CK_ARGS ConvLeanArgs
#define TNB_CK_LAUNCH launch_conv3x3_lean
#define TNB_CK_LEAN 1
#include "conv_kernel.inc"

}  // namespace tnb
