// HBM-bound kernels of the TrackNet hot path (everything that is not a 3x3 convolution).
#pragma once
#include "igemm.cuh"

namespace tnb {

enum GradMode : int { GRAD_SAME = TNB_GRAD_SAME, GRAD_POOL = TNB_GRAD_POOL, GRAD_UP = TNB_GRAD_UP };
using GradSrc = tnb_gradsrc_t;  // see include/tracknet_b200.h
using BnBwdArgs = tnb_bnbwd_t;

int launch_view_presplit(const ViewDesc& view, void* out, int fmt, cudaStream_t st);
int launch_presplit(const float* x, void* out, long long npixels, int C, int fmt, float mul, cudaStream_t st);
int launch_pack_input(const float* x_nchw, float* out_nhwc, int N, int C, int H, int W, int Cpad, cudaStream_t st,
                      void* out_planar16 = nullptr, void* out_presplit = nullptr, int presplit_fmt = 1);
int launch_bn_finalize(const float* part, int rows, double count, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, int training,
                       float* scale, float* shift, float* mean, float* invstd, int C, cudaStream_t st);
int launch_predictor_fwd(const SrcDesc& src, int N, int H, int W, const float* wp, const float* bias, int O,
                         float* y_nchw, cudaStream_t st);
size_t predictor_bwd_workspace_bytes(int N, int H, int W, int O);
int launch_predictor_bwd(const SrcDesc& src, int N, int H, int W, const float* wp, int O, const float* dy,
                         const float* y, float* dA, float* dwp, float* dbias, float* part, cudaStream_t st);
int bn_bwd_num_blocks(int N, int H, int W, int C);
int launch_bn_bwd_reduce(const BnBwdArgs& a, cudaStream_t st);
int launch_bn_bwd_finalize(const float* part, int rows, int C, float* sums, float* dgamma, float* dbeta,
                           cudaStream_t st);
int launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t st);

int wbce_num_blocks(long long per_sample);
int launch_wbce_fwd(const float* p, const float* y, int nsamples, long long per_sample, int reduce,
                    double* part, float* out, cudaStream_t st);
int launch_wbce_bwd(const float* p, const float* y, const float* gout, int nsamples, long long per_sample,
                    int reduce, float* dp, cudaStream_t st);
int launch_mixup(const float* x, const float* lam, const long long* perm, float* out, int N, long long per_sample,
                 cudaStream_t st);

struct AdamTensor { float* p; const float* g; float* m; float* v; long long n; };
int launch_adam(const AdamTensor* table_dev, int ntensors, long long total_n, float lr, float b1, float b2, float eps,
                float wd, int step, cudaStream_t st);

size_t decode_workspace_bytes(int nmaps, int H, int W);
int launch_decode(const void* maps, int is_u8, float thresh, int nmaps, int H, int W, void* ws, int* out_xywh,
                  cudaStream_t st);

struct InpaintParams { const float* w[9]; const float* b[9]; };
// coor_th < 0: out = InpaintNet(coords, mask); coor_th >= 0: the rectified trajectory of predict.py:256-261 (blend with the
// input where mask == 0, zero the points below the threshold)
int launch_inpaint_fwd(const float* coords, const float* mask, const InpaintParams& p, int N, int L, float* out,
                       float coor_th, cudaStream_t st);
int launch_median_u8(const uint8_t* frames, int T, long long P, double* out_f64, uint8_t* out_u8, cudaStream_t st);
int launch_label_discs(const int* centers, int nmaps, int H, int W, float sigma, float* out, cudaStream_t st);
int launch_resize_frames(const uint8_t* src, int nimg, int hs, int ws, int C, const int* hbounds, const int* hkk, int hksize,
                         const int* vbounds, const int* vkk, int vksize, int hd, int wd, uint8_t* tmp, float* out,
                         int per_sample, long long sample_stride, int chan_off, int frame_stride, cudaStream_t st);
int launch_bg_subtract(const uint8_t* frames, const double* median, long long nimg, int hs, int ws, uint8_t* out,
                       cudaStream_t st);
int launch_eval_stats(const float* y_pred, const float* y_true, const int* boxes, int nmaps, int H, int W, float* conf,
                      int* true_any, cudaStream_t st);
int launch_temporal_ensemble(const float* state, const float* pred, float* out, const float* weight_host, int L,
                             long long E, int B, int count0, int tail_base, int n_tail, cudaStream_t st);
struct InpaintGrads { float* w[9]; float* b[9]; };
int launch_inpaint_bwd(const float* coords, const float* mask, const InpaintParams& p, const float* dout,
                       const InpaintGrads& g, int N, int L, float* dcoords, cudaStream_t st);

}  // namespace tnb
