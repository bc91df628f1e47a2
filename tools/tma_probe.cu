// tools/tma_probe.cu - how fast does ONE SM fill a planar halo tile [plane = 8 ch][pixel][16 B] (the operand layout of the
// tcgen05 conv / wgrad kernels) from a pre-split 16-bit activation tensor, for different global layouts and copy engines?
//   mode 0  LDGSTS: 192 threads issue 16-byte cp.async copies (what conv3x3_kernel's dgrad producers do today)
//   mode 1  tensor-TMA, global layout [N][H][W][2C/8 planes][8] (today's pre-split layout), box {8, PW, PH, 4 planes}
//   mode 2  tensor-TMA, global layout [N][2C/8 planes][H][W][8] (planar in HBM too), box {8, PW, PH, 4 planes}
//   mode 3  tensor-TMA, planar layout, rows merged: dims {W * 8, H, plane, N}, box {PW * 8, PH, 4}
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu   (no -lcuda: the encoder
// is fetched with cudaGetDriverEntryPoint). Run: tools/tma_probe  -> one line per (mode, shape).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
    if (++spins > (1u << 26)) { printf("probe: mbarrier timeout\n"); __trap(); }
  }
}
__device__ __forceinline__ void tma5(void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(s32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma4(void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(s32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void cp16(void* dst, const void* src, uint32_t n) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(s32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_arrive(uint64_t* b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(b)) : "memory"); }

struct Args {
  const uint8_t* src;
  int N, H, W, C;       // C = channels; the tensor holds 2C 16-bit values per pixel (hi C | lo C)
  int PW, PH;           // halo tile: PW x PH pixels (tile + 2)
  int tiles_w, tiles_h, ntiles, mode, stages;
  unsigned long long* clocks;
};

// One CTA per SM; each walks tiles blockIdx.x, +gridDim.x, ...; per tile C/32 chunk stages of [2 terms][4 planes][PH*PW px][16 B].
// Thread 0 (TMA modes) or 192 threads (LDGSTS) fill; one consumer thread waits `full` and immediately releases `empty`.
__global__ void __launch_bounds__(256, 1) probe_kernel(const __grid_constant__ CUtensorMap tmap, const Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 8;
  uint8_t* stage0 = smem + 128;
  const int px = a.PW * a.PH;
  const int plane = px * 16;
  const int stage_bytes = 2 * 4 * plane;
  const int tid = threadIdx.x;
  const bool tma = a.mode != 0;
  if (tid == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], tma ? 1 : 192); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nchunks = a.C / 32;
  const long long t0 = clock64();
  if (tid == 224) {  // consumer
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x)
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full[s], ph);
        mbar_arrive(&empty[s]);
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
  } else if (tma && tid == 0) {
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tw = t % a.tiles_w; t /= a.tiles_w;
      const int th = t % a.tiles_h; const int n = t / a.tiles_h;
      const int w0 = tw * (a.PW - 2) - 1, h0 = th * (a.PH - 2) - 1;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect(&full[s], (uint32_t)stage_bytes);
        uint8_t* dst = stage0 + (size_t)s * stage_bytes;
        for (int term = 0; term < 2; ++term) {
          const int p0 = term * (a.C / 8) + c * 4;
          if (a.mode == 1) tma5(dst + term * 4 * plane, &tmap, 0, w0, h0, p0, n, &full[s]);        // dims {8, W, H, P, N}
          else if (a.mode == 2) tma5(dst + term * 4 * plane, &tmap, 0, w0, h0, p0, n, &full[s]);   // dims {8, W, H, P, N}, planar strides
          else tma4(dst + term * 4 * plane, &tmap, w0 * 8, h0, p0, n, &full[s]);                   // dims {W*8, H, P, N}
        }
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (!tma && tid < 192) {
    const int j = tid & 3, pbase = tid >> 2;
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      int t = tile;
      const int tw = t % a.tiles_w; t /= a.tiles_w;
      const int th = t % a.tiles_h; const int n = t / a.tiles_h;
      const int w0 = tw * (a.PW - 2) - 1, h0 = th * (a.PH - 2) - 1;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* dst = stage0 + (size_t)s * stage_bytes + j * plane;
        const uint8_t* sb = a.src + (size_t)(c * 32 + j * 8) * 2;
        for (int p = pbase; p < px; p += 48) {
          const int hr = p / a.PW, hc = p - hr * a.PW;
          const int h = h0 + hr, w = w0 + hc;
          const bool ok = h >= 0 && h < a.H && w >= 0 && w < a.W;
          const uint8_t* q = ok ? sb + ((size_t)(n * a.H + h) * a.W + w) * a.C * 4 : sb;
          cp16(dst + p * 16, q, ok ? 16u : 0u);
          cp16(dst + p * 16 + 4 * plane, q + a.C * 2, ok ? 16u : 0u);
        }
        cp_arrive(&full[s]);
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
    }
  }
  __syncthreads();
  if (tid == 0) a.clocks[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (encode == nullptr) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned long long* clocks;
  CK(cudaMalloc(&clocks, sizeof(unsigned long long) * sms));
  struct Shape { int N, H, W, C, TW, TH; };
  // the dgrad launches of the step: 288x512x64 (16 x 32 tiles), 144x256x128, 72x128x256 (tall: 32 x 16), 36x64x512
  Shape shapes[] = {{10, 288, 512, 64, 32, 16}, {10, 288, 512, 64, 16, 16}, {10, 144, 256, 128, 16, 16}, {10, 72, 128, 256, 16, 8}, {10, 72, 128, 256, 16, 32}};
  for (const Shape& sh : shapes) {
    const size_t bytes = (size_t)sh.N * sh.H * sh.W * sh.C * 4;
    uint8_t* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    for (int mode = 0; mode < 4; ++mode) {
      Args a;
      a.src = buf; a.N = sh.N; a.H = sh.H; a.W = sh.W; a.C = sh.C; a.PW = sh.TW + 2; a.PH = sh.TH + 2;
      a.tiles_w = (sh.W + sh.TW - 1) / sh.TW; a.tiles_h = (sh.H + sh.TH - 1) / sh.TH; a.ntiles = sh.N * a.tiles_w * a.tiles_h;
      a.mode = mode; a.clocks = clocks;
      const int stage_bytes = 2 * 4 * a.PW * a.PH * 16;
      a.stages = (220 * 1024 - 128) / stage_bytes; if (a.stages > 4) a.stages = 4;
      if (a.stages < 2) { printf("shape too large\n"); continue; }
      CUtensorMap tmap = {};
      const cuuint64_t P = 2 * sh.C / 8;
      CUresult r = CUDA_SUCCESS;
      if (mode == 1) {
        cuuint64_t dims[5] = {8, (cuuint64_t)sh.W, (cuuint64_t)sh.H, P, (cuuint64_t)sh.N};
        cuuint64_t strides[4] = {(cuuint64_t)sh.C * 4, (cuuint64_t)sh.W * sh.C * 4, 16, (cuuint64_t)sh.H * sh.W * sh.C * 4};
        cuuint32_t box[5] = {8, (cuuint32_t)a.PW, (cuuint32_t)a.PH, 4, 1}, es[5] = {1, 1, 1, 1, 1};
        r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else if (mode == 2) {
        cuuint64_t dims[5] = {8, (cuuint64_t)sh.W, (cuuint64_t)sh.H, P, (cuuint64_t)sh.N};
        cuuint64_t strides[4] = {16, (cuuint64_t)sh.W * 16, (cuuint64_t)sh.H * sh.W * 16, (cuuint64_t)sh.H * sh.W * sh.C * 4};
        cuuint32_t box[5] = {8, (cuuint32_t)a.PW, (cuuint32_t)a.PH, 4, 1}, es[5] = {1, 1, 1, 1, 1};
        r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else if (mode == 3) {
        if (a.PW * 8 > 256) { printf("mode 3 shape %dx%dx%d tile %dx%d: box row %d > 256 elements, skipped\n", sh.H, sh.W, sh.C, sh.TH, sh.TW, a.PW * 8); continue; }
        cuuint64_t dims[4] = {(cuuint64_t)sh.W * 8, (cuuint64_t)sh.H, P, (cuuint64_t)sh.N};
        cuuint64_t strides[3] = {(cuuint64_t)sh.W * 16, (cuuint64_t)sh.H * sh.W * 16, (cuuint64_t)sh.H * sh.W * sh.C * 4};
        cuuint32_t box[4] = {(cuuint32_t)a.PW * 8, (cuuint32_t)a.PH, 4, 1}, es[4] = {1, 1, 1, 1};
        r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (r != CUDA_SUCCESS) { printf("mode %d: cuTensorMapEncodeTiled failed (%d)\n", mode, (int)r); continue; }
      const size_t smem = 128 + (size_t)a.stages * stage_bytes;
      CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        probe_kernel<<<sms, 256, smem>>>(tmap, a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      std::vector<unsigned long long> hc(sms);
      CK(cudaMemcpy(hc.data(), clocks, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
      unsigned long long mx = 0; for (auto v : hc) mx = v > mx ? v : mx;
      const double moved = (double)a.ntiles * (sh.C / 32) * stage_bytes;   // bytes landed in shared memory, halo included
      printf("mode %d  %3dx%3dx%3d tile %2dx%2d stages %d: %.4f ms  %.1f GB/s into smem  %.1f B/clk/SM (max %llu clk)  [algorithmic %.1f GB/s]\n",
             mode, sh.H, sh.W, sh.C, sh.TH, sh.TW, a.stages, best, moved / best / 1e6, moved / sms / (double)mx, mx,
             (double)bytes / best / 1e6);
    }
    CK(cudaFree(buf));
  }
  return 0;
}
