// Optional per-launch timing (CUDA events on the launching stream) used by bench.py's roofline leg.
#pragma once
#include <cuda_runtime.h>
namespace tnb {
enum ProfKind : int { PROF_CONV_FWD = 0, PROF_CONV_DGRAD = 1, PROF_WGRAD = 2, PROF_BN_BWD = 3, PROF_PRED = 4,
                      PROF_BN_FIN = 5, PROF_PACK = 6 };
bool prof_enabled();  // per-launch timing is on: launch sequences must not be replayed from a CUDA graph
struct ProfScope {
  int idx;
  cudaStream_t st;
  ProfScope(int kind, cudaStream_t st, int n, int h, int w, int cin, int cout);
  ~ProfScope();
};
}  // namespace tnb
