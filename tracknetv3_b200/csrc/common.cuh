// Common device helpers for the sm_100a kernels: mbarrier, bulk-TMA, tcgen05 (UMMA/TMEM)
// PTX wrappers, shared-memory matrix descriptors, operand splitting.
//
// Nothing here is derived from the reference (which has no native code);
// descriptor bit layouts follow the PTX ISA tcgen05 matrix / instruction descriptor tables.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define TNB_DEVINL __device__ __forceinline__

namespace tnb {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define TNB_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      tnb::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                          cudaGetErrorString(_e));                                        \
      return (int)_e ? (int)_e : -1;                                                      \
    }                                                                                     \
  } while (0)
#define TNB_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      tnb::set_last_error(__VA_ARGS__);                                                   \
      return -2;                                                                          \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol). The network's passes are ~40 + ~115 dependent launches, many of them
// 3-10 us elementwise kernels between the tensor-core kernels. Launched with the attribute below, the NEXT kernel's CTAs
// are scheduled as soon as every CTA of the running one has passed pdl_launch_dependents() (or exited) and an SM has
// room; they run their prologue and block in pdl_wait() until the running grid has completed and its memory is visible.
// Contract: a kernel launched through launch_pdl() calls pdl_wait() before its first global-memory access (reads of a
// predecessor's output AND writes to anything a predecessor may still read).
// Measured on one box (profiles/r2_experiments.md): releasing the dependents early from the SMALL kernels only and letting
// the persistent tcgen05 kernels release theirs at exit is 2-3 % faster at batch 1 (1.07 -> 1.03 ms eval forward,
// 3.70 -> 3.61 ms train step; 9 % without CUDA graphs) and neutral at batch 10; an early release from the persistent
// kernels as well costs 0.8 ms per bs-10 step and is not done. TNB_PDL=0 launches everything fully serialised (ablation).
// ---------------------------------------------------------------------------------------------
inline bool pdl_enabled() {
  static const int v = [] { const char* e = getenv("TNB_PDL"); return e ? atoi(e) : 1; }();
  return v != 0;
}
TNB_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
TNB_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  TNB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
TNB_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
TNB_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
TNB_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
TNB_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TNB_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
TNB_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
TNB_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
TNB_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("tnb: mbarrier wait timeout (block %d,%d thread %d bar %p parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, (void*)bar, parity);
      __trap();
    }
  }
}

// Same with a sleep between polls (sleep_ns = 0: plain spin): for roles that have slack in the pipeline.
TNB_DEVINL void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (sleep_ns) __nanosleep(sleep_ns);
    if (++spins > (1u << 26)) {
      printf("tnb: mbarrier wait timeout (block %d,%d thread %d bar %p parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, (void*)bar, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// async-proxy fences and bulk TMA (1-D cp.async.bulk; SASS: UBLKCP)
// ---------------------------------------------------------------------------------------------
TNB_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
TNB_DEVINL void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Tensor-TMA: one 4-D box (cp.async.bulk.tensor, SASS UTMALDG) from the tensor described by `tmap` to shared memory;
// out-of-range coordinates (the convolution's padding ring) arrive as zeros; completion as bytes on the mbarrier.
TNB_DEVINL void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
TNB_DEVINL void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// 16-byte asynchronous global->shared copy (LDGSTS). src_bytes = 0 zero-fills the destination without reading.
// ca = true allocates the line in L1 (.ca) instead of bypassing it (.cg). With .cg every 16-byte copy fetches its own
// 32-byte sector from L2 even when the neighbouring thread copies the other half (measured with ncu on the wgrad fill:
// 31.7 sectors per 512-byte warp request, 2.38 GB moved for 1.19 GB used); with .ca the two halves share one fetch.
// wgrad 7.72 -> 6.52 ms per step, dgrad 5.80 -> 5.63 (profiles/r1_summary.md 6).
TNB_DEVINL void cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes, bool ca = false) {
  if (ca)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
  else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
inline int cp_async_ca_env() {  // TNB_CPASYNC_CA=0 restores the L1-bypassing copies (ablation)
  static const int v = [] { const char* e = getenv("TNB_CPASYNC_CA"); return e ? atoi(e) : 1; }();
  return v;
}
// The mbarrier receives one arrival (counted against its initial expected count) once all cp.async operations issued
// so far by this thread have landed: producers never wait for their own loads.
TNB_DEVINL void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ---------------------------------------------------------------------------------------------
TNB_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
TNB_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
TNB_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TNB_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
TNB_DEVINL void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, descriptors given as (low, high) 32-bit words: callers that step the start address do 32-bit arithmetic only.
TNB_DEVINL void umma_f16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
TNB_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp receives row (lane_base+i), columns [c, c+32).
TNB_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
TNB_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
TNB_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2): cluster rank / barrier, remote mbarrier arrive,
// paired TMEM allocation, MMA and multicast commit
// ---------------------------------------------------------------------------------------------
TNB_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
TNB_DEVINL void cluster_sync_all() {  // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
TNB_DEVINL void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {  // the barrier at the same offset in CTA `cta`
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
TNB_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // acquire at cluster scope (remote arrivals)
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 26)) {
      printf("tnb: cluster mbarrier wait timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
TNB_DEVINL void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // warp 0 of BOTH CTAs, same smem offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
TNB_DEVINL void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
TNB_DEVINL void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
TNB_DEVINL void umma_f16_w_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                uint32_t accumulate) {  // umma_f16_w for a CTA pair
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
TNB_DEVINL void umma_commit_pair(uint64_t* bar) {  // arrives on the barrier at this offset in both CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}


// ---------------------------------------------------------------------------------------------
// Descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts, in 16-byte units:
//   K-major  : ((8,n),2)      : ((1,SBO),LBO)      8 rows x 16B core matrices; SBO between 8-row
//                                                  groups, LBO between the two K halves of one MMA.
//   MN-major : ((1,n),(8,2))  : ((.,SBO),(1,LBO))  8 k-rows x 16B (8 MN elems); SBO between MN chunks
//                                                  of 8 elements, LBO between 8-row K groups.
// Both are served by ONE physical "planar" tile layout [plane = 8 channels][pixel][16 bytes].
TNB_DEVINL uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = 0 (no swizzle)
}
// Instruction descriptor for kind::f16: fp32 accumulate, A/B format 0 = fp16 / 1 = bf16.
TNB_DEVINL uint32_t make_idesc2(int M, int N, int a_format, int b_format, int a_mn_major, int b_mn_major);
TNB_DEVINL uint32_t make_idesc(int M, int N, int ab_format, int a_mn_major, int b_mn_major) {
  return make_idesc2(M, N, ab_format, ab_format, a_mn_major, b_mn_major);
}
// A and B formats are independent fields of the descriptor (0 = fp16, 1 = bf16)
TNB_DEVINL uint32_t make_idesc2(int M, int N, int a_format, int b_format, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                          // D format f32
  d |= (uint32_t)a_format << 7;          // A format
  d |= (uint32_t)b_format << 10;         // B format
  d |= (uint32_t)(a_mn_major & 1) << 15; // A major
  d |= (uint32_t)(b_mn_major & 1) << 16; // B major
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) 16-bit split. x ~= hi + lo with hi = rn16(x), lo = rn16(x - hi).
// FMT 0: fp16 (clamped to the finite range), FMT 1: bf16.
// ---------------------------------------------------------------------------------------------
// fp16 (hi, lo) pairs carry 22 bits only while lo = fp16(x - hi) ~ 2^-11 |x| stays a NORMAL fp16 number, i.e. for
// |x| >= 2^-3. Convolution weights are ~1e-2 (Kaiming bound 1 / sqrt(9 Cin)): unscaled, their lo halves are subnormal and
// the pair keeps ~18 bits (measured: that alone was the forward's 4e-5 heatmap distance and 10-20x the fp32 rate of ReLU
// mask flips in the backward pass). Weights packed in fp16 are therefore stored multiplied by 2^10 (22 bits for every
// |w| >= 2^-13, clamped beyond |w| = 63.97) and the convolution epilogue multiplies the accumulator by 2^-10: exact.
constexpr float kW16Mul = 1024.f;
constexpr float kW16Inv = 1.f / 1024.f;

template <int FMT>
TNB_DEVINL void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (FMT == 0) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
  }
}
template <int FMT>
TNB_DEVINL void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  split2<FMT>(v[0], v[1], hi.x, lo.x);
  split2<FMT>(v[2], v[3], hi.y, lo.y);
  split2<FMT>(v[4], v[5], hi.z, lo.z);
  split2<FMT>(v[6], v[7], hi.w, lo.w);
}

TNB_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
TNB_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace tnb
