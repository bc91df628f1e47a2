// tools/mma_align_probe.cu - does the ALIGNMENT of the shared-memory operand matter for tcgen05.mma throughput?
// The conv kernels read their A operand (128 pixels x 16 channels per MMA) from an un-swizzled planar halo tile: a core
// matrix is 8 consecutive pixels x 16 B = 128 contiguous bytes, but it starts wherever the tap puts it: at
// (row * PITCH + dx + 8 j) * 16 bytes with PITCH = 18 pixels (288 B), i.e. 16-byte aligned, rarely 128-byte aligned.
// One CTA per SM issues a long chain of M = 128 MMAs (kind::f16, K = 16, both operands K-major SWIZZLE_NONE) and times
// it with clock64 for: N in {64, 128, 256}; A row pitch (SBO) 256 / 288 / 384 B; A start offset 0 / 16 / 48 B; and the 9-tap
// address pattern of the real kernel. Output: clocks per MMA next to the math floor N / 2 and the operand bytes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I tracknetv3_b200/csrc -o tools/mma_align_probe tools/mma_align_probe.cu
#include "common.cuh"
#include <cstdarg>
#include <vector>

namespace tnb { void set_last_error(const char*, ...) {} }
using namespace tnb;

struct Args {
  int N, a_sbo, a_off, a_lbo, taps, iters;  // taps: 1 = same address every MMA, 9 = (dy, dx) pattern with row pitch a_sbo
  unsigned long long* clocks;
  int merged;  // 0: uniform MMAs of width N. The 64-wide conv tiles issue x_hi * [w_hi | w_lo] (N = 128) and x_lo * w_hi
               // (N = 64) into one accumulator: 1 = alternating per K step (what the kernel does), 2 = the N = 128 MMAs
               // of a 3-tap weight stage first, then its N = 64 MMAs (two shape switches per stage instead of twelve)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 16);
  uint8_t* a_base = smem + 1024;
  uint8_t* b_base = smem + 1024 + 112 * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (112 + 64) * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0x3c003c00u;  // fp16 1.0
  if (warp == 0) {
    if (elect_one()) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0) {
    const bool lead = elect_one();
    const uint32_t idesc = make_idesc(128, a.N, 0, 0, 0);
    const uint64_t a_desc0 = make_smem_desc(smem_u32(a_base) + a.a_off, a.a_lbo, a.a_sbo);
    const uint64_t b_desc0 = make_smem_desc(smem_u32(b_base), a.N * 16, 128);
    long long t0 = 0;
    if (lead && a.merged) {
      const uint32_t id128 = make_idesc(128, 128, 0, 0, 0), id64 = make_idesc(128, 64, 0, 0, 0);
      const uint64_t b128 = make_smem_desc(smem_u32(b_base), 128 * 16, 128);  // [plane][hi 64 | lo 64 rows]
      t0 = clock64();
      for (int it = 0; it < a.iters; ++it) {
        for (int g = 0; g < 3; ++g) {  // one weight stage = one filter row: 3 taps x 2 K steps, two M tiles
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t d = tmem_base + mt * 128;
            if (a.merged == 1) {
              for (int t = 0; t < 3; ++t)
                for (int ks = 0; ks < 2; ++ks) {
                  const uint32_t off = (uint32_t)(g * a.a_sbo + t * 16 + mt * 128 + ks * 2 * a.a_lbo);
                  umma_f16(d, a_desc0 + (off >> 4), b128 + ((ks * 2 * 128 * 16) >> 4), id128, 1);
                  umma_f16(d, a_desc0 + ((off + 4 * a.a_lbo) >> 4), b128 + ((ks * 2 * 128 * 16) >> 4), id64, 1);
                }
            } else {
              for (int t = 0; t < 3; ++t)
                for (int ks = 0; ks < 2; ++ks) {
                  const uint32_t off = (uint32_t)(g * a.a_sbo + t * 16 + mt * 128 + ks * 2 * a.a_lbo);
                  umma_f16(d, a_desc0 + (off >> 4), b128 + ((ks * 2 * 128 * 16) >> 4), id128, 1);
                }
              for (int t = 0; t < 3; ++t)
                for (int ks = 0; ks < 2; ++ks) {
                  const uint32_t off = (uint32_t)(g * a.a_sbo + t * 16 + mt * 128 + ks * 2 * a.a_lbo);
                  umma_f16(d, a_desc0 + ((off + 4 * a.a_lbo) >> 4), b128 + ((ks * 2 * 128 * 16) >> 4), id64, 1);
                }
            }
          }
        }
      }
      umma_commit(bar);
    } else if (lead) {
      t0 = clock64();
      for (int it = 0; it < a.iters; ++it) {
        for (int tap = 0; tap < a.taps; ++tap) {
          const uint32_t off = (uint32_t)((tap / 3) * a.a_sbo + (tap % 3) * 16);  // (dy, dx) of a 3x3 filter
          // two K steps (the two 16-channel halves of a 32-channel chunk), accumulators alternate like M tiles
          umma_f16(tmem_base, a_desc0 + (off >> 4), b_desc0, idesc, 1);
          umma_f16(tmem_base + 256, a_desc0 + ((off + 2 * a.a_lbo) >> 4), b_desc0 + ((2 * a.N * 16) >> 4), idesc, 1);
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (lead) a.clocks[blockIdx.x] = (unsigned long long)(clock64() - t0);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned long long* clocks;
  CK(cudaMalloc(&clocks, sizeof(unsigned long long) * sms));
  const size_t smem = 1024 + (112 + 64) * 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("M = 128, K = 16, kind::f16, SWIZZLE_NONE K-major operands; clocks per MMA (max over %d SMs), math floor = N / 2\n", sms);
  const int Ns[] = {64, 128, 256};
  struct Cfg { int sbo, off, lbo, taps; const char* what; };
  const Cfg cfgs[] = {
      {256, 0, 6144, 1, "rows 256 B apart, start aligned, same address"},
      {256, 16, 6144, 1, "rows 256 B apart, start + 16 B"},
      {256, 48, 6144, 1, "rows 256 B apart, start + 48 B"},
      {288, 0, 6144, 1, "rows 288 B apart (the 18-pixel halo pitch), start aligned"},
      {288, 0, 5280, 1, "rows 288 B apart, plane stride 5280 B (the kernel's padded plane)"},
      {384, 0, 6912, 1, "rows 384 B apart (24-pixel pitch), start aligned"},
      {288, 0, 5280, 9, "the kernel's pattern: 9 taps, 288 B pitch, padded planes"},
      {384, 0, 6912, 9, "9 taps, 384 B pitch (dx = 0 taps aligned)"},
      {256, 0, 6144, 9, "9 taps, 256 B pitch"},
  };
  for (int N : Ns)
    for (const Cfg& c : cfgs) {
      Args a{N, c.sbo, c.off, c.lbo, c.taps, c.taps == 9 ? 400 : 3600, clocks, 0};
      unsigned long long best = ~0ull;
      for (int rep = 0; rep < 3; ++rep) {
        probe_kernel<<<sms, 128, smem>>>(a);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(sms);
        CK(cudaMemcpy(h.data(), clocks, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
        unsigned long long mx = 0;
        for (auto v : h) mx = v > mx ? v : mx;
        if (mx < best) best = mx;
      }
      const double per = (double)best / (2.0 * a.iters * a.taps);
      printf("N %3d  %-70s %7.1f clk/MMA  (floor %3d, A 4096 + B %5d bytes -> %5.1f B/clk)\n", N, c.what, per, N / 2, N * 32,
             (4096.0 + N * 32) / per);
    }
  printf("merged 64-wide form: pairs of one N = 128 and one N = 64 MMA (floor 64 + 32 = 96 clocks of math, 14 KB of operands = 112 clocks)\n");
  for (int merged = 1; merged <= 2; ++merged) {
    Args a{64, 288, 0, 5280, 9, 300, clocks, merged};
    unsigned long long best = ~0ull;
    for (int rep = 0; rep < 3; ++rep) {
      probe_kernel<<<sms, 128, smem>>>(a);
      CK(cudaDeviceSynchronize());
      std::vector<unsigned long long> h(sms);
      CK(cudaMemcpy(h.data(), clocks, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
      unsigned long long mx = 0;
      for (auto v : h) mx = v > mx ? v : mx;
      if (mx < best) best = mx;
    }
    printf("  %-60s %7.1f clocks per pair\n", merged == 1 ? "alternating N = 128 / N = 64 (issue order of the kernel)" : "grouped: six N = 128 MMAs, then six N = 64 MMAs per weight stage",
           (double)best / (a.iters * 3 * 2 * 6));
  }
  return 0;
}