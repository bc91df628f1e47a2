// CTA-pair kernel with the lean MMA-issue loop (conv_kernel.inc with TNB_CK_PAIR 1 and TNB_CK_LEAN 1) - EXPERIMENTAL,
// selected by TNB_CONV_PAIR=1 together with TNB_CONV_LEAN=1; see conv_pair.cu and conv_lean.cu.
#include "igemm.cuh"
#include <cstdlib>
#include "prof.cuh"
#include <type_traits>

namespace tnb {

#define TNB_CK_NAME conv3x3_pair_lean_kernel
#define TNB_CK_ARGS ConvPairLeanArgs
#define TNB_CK_LAUNCH launch_conv3x3_pair_lean
#define TNB_CK_PAIR 1
#define TNB_CK_LEAN 1
#include "conv_kernel.inc"

}  // namespace tnb
