// Video-side kernels of the input contract: the per-pixel median background over all
// frames of a clip (reference dataset.py:102-107) and the binary-disc training labels (dataset.py:400-410).
// Both are HBM-bound byte work: coalesced 32-bit loads of 4 neighbouring byte positions per thread, grid sized from the
// byte count.
#include "kernels.cuh"
#include <algorithm>

namespace tnb {

// np.median(frame_arr, 0) for uint8 frames: per byte position p the two middle order statistics a <= b of the T values
// frames[t][p] (a == b for odd T); numpy returns their float64 mean. Selection by bisection on the VALUE (8 rounds, each
// counting `v <= mid` over the T frames for both ranks at once): no per-thread histogram, no sort scratch; every
// round re-reads the T bytes of the position, which for neighbouring threads are neighbouring bytes (coalesced), and for
// clips up to ~40 frames of 720p come from the 126 MB L2 after the first round.
// out_f64 (optional): the float64 median (bg_mode 'subtract' / 'subtract_concat' keep it, dataset.py:108-109);
// out_u8 (optional): median.astype('uint8') = floor((a + b) / 2) (bg_mode 'concat', dataset.py:105).
__global__ void __launch_bounds__(256) median_u8_kernel(const uint8_t* __restrict__ frames, int T, long long P,
                                                        double* __restrict__ out_f64, uint8_t* __restrict__ out_u8) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // group of 4 byte positions
  const long long p0 = q * 4;
  if (p0 >= P) return;
  const bool vec = (P % 4 == 0) && p0 + 4 <= P;                           // 32-bit loads need P % 4 == 0 (row alignment)
  const int nb = (int)(P - p0 < 4 ? P - p0 : 4);
  const int k_lo = (T - 1) / 2, k_hi = T / 2;                             // 0-based ranks of the two middle elements
  int lo_a[4], hi_a[4], lo_b[4], hi_b[4];                                 // bisection brackets [lo, hi] per byte, both ranks
#pragma unroll
  for (int j = 0; j < 4; ++j) { lo_a[j] = 0; hi_a[j] = 255; lo_b[j] = 0; hi_b[j] = 255; }
  for (int round = 0; round < 8; ++round) {
    int mid_a[4], mid_b[4], cnt_a[4], cnt_b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { mid_a[j] = (lo_a[j] + hi_a[j]) >> 1; mid_b[j] = (lo_b[j] + hi_b[j]) >> 1; cnt_a[j] = 0; cnt_b[j] = 0; }
    for (int t = 0; t < T; ++t) {
      const uint8_t* src = frames + (size_t)t * P + p0;
      uint32_t word = 0;
      if (vec) word = *reinterpret_cast<const uint32_t*>(src);
      else for (int j = 0; j < nb; ++j) word |= (uint32_t)src[j] << (8 * j);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int v = (word >> (8 * j)) & 0xff;
        cnt_a[j] += v <= mid_a[j];
        cnt_b[j] += v <= mid_b[j];
      }
    }
    // smallest value v with count(<= v) >= k + 1 is the rank-k order statistic
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (cnt_a[j] >= k_lo + 1) hi_a[j] = mid_a[j]; else lo_a[j] = mid_a[j] + 1;
      if (cnt_b[j] >= k_hi + 1) hi_b[j] = mid_b[j]; else lo_b[j] = mid_b[j] + 1;
    }
  }
  for (int j = 0; j < nb; ++j) {
    const int a = lo_a[j], b = lo_b[j];
    if (out_f64 != nullptr) out_f64[p0 + j] = 0.5 * (double)(a + b);
    if (out_u8 != nullptr) out_u8[p0 + j] = (uint8_t)((a + b) >> 1);
  }
}
int launch_median_u8(const uint8_t* frames, int T, long long P, double* out_f64, uint8_t* out_u8, cudaStream_t st) {
  TNB_REQUIRE(T >= 1 && P >= 1, "median_u8: empty clip (%d frames of %lld bytes)", T, P);
  const long long groups = (P + 3) / 4;
  median_u8_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(frames, T, P, out_f64, out_u8);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Training labels (dataset.py:400-410 `_get_heatmap`): map m of H x W is 1 where (x - (cx + 1))^2 + (y - (cy + 1))^2 <=
// sigma^2 on the 1-based grid x = 1..W, y = 1..H - i.e. (i - cx)^2 + (j - cy)^2 <= sigma^2 for 0-based pixel (j, i) -
// and all-zero when cx == cy == 0. centers: int32 [nmaps][2] = (cx, cy). One float4 store per thread.
__global__ void __launch_bounds__(256) label_disc_kernel(const int* __restrict__ centers, int nmaps, int H, int W, float r2,
                                                         float* __restrict__ out) {
  const long long total4 = (long long)nmaps * H * W / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    const int m = (int)(e / ((long long)H * W));
    const int rem = (int)(e - (long long)m * H * W);
    const int y = rem / W, x0 = rem - y * W;
    const int cx = centers[2 * m], cy = centers[2 * m + 1];
    float v[4];
    const float dy2 = (float)((y - cy) * (y - cy));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int dx = x0 + j - cx;
      v[j] = (cx == 0 && cy == 0) ? 0.f : ((float)(dx * dx) + dy2 <= r2 ? 1.f : 0.f);
    }
    *reinterpret_cast<float4*>(out + e) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
int launch_label_discs(const int* centers, int nmaps, int H, int W, float sigma, float* out, cudaStream_t st) {
  TNB_REQUIRE(W % 4 == 0, "label_discs: width %d not a multiple of 4", W);
  if (nmaps == 0) return 0;
  const long long total4 = (long long)nmaps * H * W / 4;
  label_disc_kernel<<<(int)std::min<long long>((total4 + 255) / 256, 148 * 16), 256, 0, st>>>(centers, nmaps, H, W,
                                                                                             sigma * sigma, out);
  TNB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tnb
