/* tracknet_b200 — C ABI of the B200-native (sm_100a) TrackNetV3 hot path.
 *
 * The reference (qaz812345/TrackNetV3) is pure Python and has no FFI of its own: its hot path is the
 * torch.nn modules in model.py and the functions in utils/metric.py / test.py, which dispatch to
 * cuDNN/ATen. This header is the boundary a maintainer binds instead (ctypes stub in INTEGRATION.md).
 * Every entry point cites the reference code it replaces as file:line into the reference tree.
 *
 * Conventions
 *   - plain C types only: raw DEVICE pointers, ints, a cudaStream_t passed as void*;
 *   - no ownership transfer: every buffer (inputs, outputs, saved-for-backward, workspace) is allocated
 *     by the caller (PyTorch in the shipped host code) and passed in;
 *   - return 0 on success, non-zero on error; tnb_last_error() gives the message (thread-local);
 *   - all activations are fp32; internal activation layout is NHWC, model inputs/outputs are NCHW
 *     exactly as the reference's modules take/return them;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef TRACKNET_B200_H
#define TRACKNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNB_ABI_VERSION 2

/* ---- descriptors -------------------------------------------------------------------------- */

/* How one channel-slice of a conv layer's logical input is produced from a stored tensor. */
enum { TNB_SRC_IDENTITY = 0, TNB_SRC_AFFINE_RELU = 1, TNB_SRC_AFFINE_RELU_POOL = 2, TNB_SRC_AFFINE_RELU_UP = 3,
       TNB_SRC_PRESPLIT = 4 /* ptr holds the "pre-split" 16-bit format, see tnb_presplit_bf16 */,
       TNB_SRC_PRESPLIT_UP = 5 /* pre-split tensor at HALF resolution read through nearest x2 upsampling (wgrad only) */,
       TNB_SRC_PLANAR16 = 6 /* planar fp16 (hi, lo) pairs, see tnb_pack_nchw_to_planar16: the forward convolution stages
                               its halo tiles with tensor-TMA (cp.async.bulk.tensor), padding ring zero-filled by the
                               hardware; single-source views, forward only */ };
typedef struct {
  const float* ptr;   /* [N, Hs, Ws, C] fp32 NHWC */
  const float* scale; /* [C] fused BatchNorm scale (gamma * invstd). IDENTITY: NULL, or a pointer to ONE float =
                         max|tensor| -> the kernel pre-scales by a power of two before its fp16 hi/lo split and
                         un-scales the result (used for gradients, whose magnitude is ~1e-8) */
  const float* shift; /* [C] fused BatchNorm shift (beta - mean * scale) */
  int C, Hs, Ws, mode;
} tnb_src_t;

/* Logical input of a conv layer = channel concat of up to two sources (torch.cat order, model.py:65). */
typedef struct {
  tnb_src_t s[2];
  int C0; /* channels taken from s[0]; the rest come from s[1] */
  int C;  /* total channels, multiple of 32 */
  int N, H, W;
} tnb_view_t;

enum { TNB_GRAD_SAME = 0, TNB_GRAD_POOL = 1, TNB_GRAD_UP = 2 };
typedef struct {
  const float* ptr; /* consumer's input-gradient tensor [N, Hs, Ws, C] */
  int C, coff, mode, Hs, Ws;
} tnb_gradsrc_t;

typedef struct {
  tnb_gradsrc_t g[2];
  int ng;
  const float* z;
  const float *scale, *shift, *mean, *invstd;
  int N, H, W, C;
  float* part;       /* reduce: [tnb_bn_bwd_blocks()][2][C] */
  const float* sums; /* apply:  [2][C] */
  float* dz;         /* apply:  [N,H,W,C] */
  float inv_count;
  float* amax;       /* apply, optional: device scalar, atomically raised to max|dz| (zero it first) */
  int dz_format;     /* apply: 0 = fp32 [N,H,W,C]; 1 = pre-split bf16 (same byte size, see tnb_presplit_bf16); 2 = pre-split
                        fp16 after multiplication by a power of two chosen from gmax (22-bit operands for the tensor
                        cores instead of 16; the multiplier is published in *dz_mul) */
  void* act_presplit; /* apply, optional: also write relu(scale*z+shift) - this layer's activation, the wgrad operand of
                         the next layer - in the pre-split format (bf16, or fp16 with dz_format 2; saves a
                         tnb_view_presplit pass); NULL = off */
  float* gmax;       /* reduce, optional: device scalar, atomically raised to max |g| of the masked incoming gradient (zero
                        it first); apply with dz_format 2: read */
  float* dz_mul;     /* apply with dz_format 2: device scalar that receives the power-of-two multiplier dz was stored with */
  int act_pool;      /* apply with act_presplit: 0 = the activation at this layer's resolution [N,H,W][2][C]; 1 = its
                        2x2 max-pool [N,H/2,W/2][2][C] (H, W even) - the wgrad operand of a next layer that reads this one
                        through MaxPool2d (model.py:59,61,63) */
  void* act_full;    /* apply, optional: the activation at this layer's resolution [N,H,W][2][C] pre-split, in addition to
                        act_presplit (an encoder block's last layer feeds the next block through the pool AND a decoder
                        concat as the skip connection: both wgrad operands come out of one pass) */
} tnb_bnbwd_t;

typedef struct {
  int n, h, w;     /* batch, height, width (h, w multiples of 8 like the reference, model.py:59-69) */
  int in_dim;      /* TrackNet(in_dim, out_dim), reference model.py:45 */
  int out_dim;
  int training;    /* 1: BatchNorm uses batch statistics and updates running stats (model.train()), the state of a backward
                      pass is kept in the workspace; 0: running statistics, nothing kept (model.eval() under no_grad);
                      2: running statistics AND the backward state (model.eval() with gradients enabled - autograd through
                      frozen BatchNorm layers, which the reference's modules allow, model.py:4-16): tnb_tracknet_backward
                      then returns dz = gamma / sqrt(running_var + eps) * g without the batch-statistics terms */
  int fwd_terms;   /* 3 = fp16 hi/lo split, fp32-faithful (default); 1 = single fp16 pass (TF32-class) */
  int bwd_terms;   /* 3 = bf16 hi/lo split (default; gradients need fp32's exponent range); 1 = single bf16 pass */
  int variant;     /* bring-up probe bits; 0 in production */
  float bn_eps;    /* 1e-5 */
  float bn_momentum; /* 0.1 */
} tnb_tracknet_cfg_t;

/* ---- misc ---------------------------------------------------------------------------------- */
const char* tnb_last_error(void);
int tnb_abi_version(void);

/* ---- operator level (each is one or two kernel launches on `stream`) ----------------------- */

/* NCHW fp32 -> NHWC fp32 with channels zero-padded to cpad. Replaces the layout the reference feeds
 * to Conv2d directly (train.py:86 `x.float().cuda()`, model.py:58). */
int tnb_pack_nchw_to_nhwc(const float* x_nchw, float* out_nhwc, int n, int c, int h, int w, int cpad, void* stream);
/* The same input in the layout the first convolution stages with tensor-TMA (TNB_SRC_PLANAR16): fp16 (hi, lo) pairs
 * (x ~ hi + lo to 2^-22), planar - out is [N][cpad / 32 chunks][2 (hi, lo)][4 planes of 8 channels][H][W][8] 16-bit,
 * n * h * w * cpad * 4 bytes like the fp32 NHWC tensor. With out_nhwc != NULL the fp32 NHWC tensor of
 * tnb_pack_nchw_to_nhwc is written by the same launch (the weight gradient of the first layer reads it). */
int tnb_pack_nchw_to_planar16(const float* x_nchw, void* out_planar16, float* out_nhwc, int n, int c, int h, int w,
                              int cpad, void* stream);

/* "Pre-split" tensor format: every fp32 value x is stored as two bf16 numbers hi = rn(x), lo = rn(x - hi)
 * (x ~ hi + lo to 2^-17), laid out [pixel][2 (hi, lo)][C] 16-bit: the hi terms of a pixel's channels are contiguous
 * (so 16-byte cp.async copies of neighbouring channel chunks coalesce into full sectors), the lo terms follow at
 * +2*C bytes. Same byte size as the fp32 tensor. Gradient tensors
 * (dz) are produced in this format by tnb_bn_relu_bwd_apply(dz_format = 1) and consumed by dgrad
 * (TNB_SRC_PRESPLIT view) and wgrad without any per-element arithmetic in the consumers. */
int tnb_presplit_bf16(const float* x_nhwc, void* out, long long npixels, int c, void* stream);
/* The same layout with fp16 (hi, lo) pairs of x * mul (x * mul ~ hi + lo to 2^-22; mul = a power of two that brings the
 * tensor into fp16's range, values beyond +-65504 are clamped). What tnb_bn_relu_bwd_apply(dz_format = 2) writes; the
 * consumers divide their results by mul (tnb_src_t.scale of a TNB_SRC_PRESPLIT source / the dz_mul arguments). */
int tnb_presplit_fp16(const float* x_nhwc, void* out, long long npixels, int c, float mul, void* stream);
/* Materialise a whole logical view (BN affine + ReLU + MaxPool / Upsample / cat of the producers) in the pre-split
 * format: out is [N,H,W][2 (hi, lo)][C] 16-bit, fmt 0 = fp16 (values clamped to +-65504), 1 = bf16. Used by the backward
 * pass so that the weight-gradient kernel's operand fills are plain copies (TNB_SRC_PRESPLIT). */
int tnb_view_presplit(const tnb_view_t* view, void* out, int fmt, void* stream);

/* Weight pre-packing for the tcgen05 kernels. mode 0: forward operand, mode 1: dgrad operand (rotated,
 * transposed). fmt 0: fp16 split of w * 2^10 (so that the lo halves of ~1e-2 weights stay normal fp16 numbers and the
 * pair keeps 22 bits; tnb_conv3x3_fwd with fmt 0 multiplies its accumulators by 2^-10), 1: bf16 split. Source is the
 * reference's canonical OIHW parameter
 * (`<block>.conv.weight`, model.py:8). The packed buffer is opaque: it is the shared-memory image the consuming
 * tnb_conv3x3_fwd launch streams with bulk TMA, one image per (output-channel tile, 32-channel chunk, filter tap), and
 * its inner layout follows the tile width the launcher will pick for this N side ([hi | lo][plane][rows], or
 * [plane][hi | lo][rows] for the 64-wide tiles whose two weight terms feed one MMA). Pack and convolve in the same
 * process (the layout also follows the TNB_CONV_MERGE ablation switch). */
size_t tnb_conv3x3_wpack_elems(int k_side, int n_side); /* number of uint16 elements */
int tnb_conv3x3_pack_weights(const float* w_oihw, uint16_t* out, int cout, int cin, int mode, int fmt, void* stream);

/* 3x3 'same' convolution, bias-free (nn.Conv2d in Conv2DBlock, model.py:8,13), with BatchNorm/ReLU/
 * MaxPool/Upsample/cat of the PRODUCING layers fused into the operand load (view), and per-tile
 * BatchNorm (sum, sumsq) partials of the OUTPUT emitted by the epilogue (stat_part may be NULL).
 * Also used for dgrad (view = dz, weights packed with mode 1). out: [N,H,W,cout] fp32. */
int tnb_conv3x3_stat_rows(int n, int h, int w, int cin, int cout, int terms);
/* The launch plan tnb_conv3x3_fwd will use for this shape (pure host arithmetic, no device needed): out[0..11] = output-
 * channel tile BN, M tiles per CTA tile MT, halo-tile stages, weight-ring slots, taps per weight stage, TMEM accumulator
 * buffers, TMEM columns, dynamic shared memory in bytes, merged-weights flag, tile orientation, 0 (reserved),
 * weight-pack layout (0 [term][plane][rows], 1 [plane][term][rows]). Returns 0 or an error code. */
int tnb_conv3x3_plan_query(int n, int h, int w, int cin, int cout, int terms, int* out12);
int tnb_conv3x3_fwd(const tnb_view_t* view, const uint16_t* wpack, float* out, float* stat_part, int cout,
                    int terms, int fmt, int variant, void* stream);

/* dgrad (view = pre-split dz, weights packed with mode 1, fmt 1) fused with the reduction pass of the BatchNorm
 * backward of the layer whose activation gradient it produces (autograd of model.py:13-15): `out` = dL/da of that
 * layer, z/scale/shift/mean/invstd = that layer's raw conv output and BatchNorm constants (cout channels). The
 * epilogue writes per-tile partials part[row][2][cout] = (sum g, sum g * xhat), g = out * [scale * z + shift > 0],
 * xhat = (z - mean) * invstd, which tnb_bn_relu_bwd_finalize consumes; rows = tnb_conv3x3_dgrad_bnreduce_rows(). */
int tnb_conv3x3_dgrad_bnreduce_rows(int n, int h, int w, int cin, int cout, int terms);
int tnb_conv3x3_dgrad_bnreduce(const tnb_view_t* view, const uint16_t* wpack, float* out, float* part, int cout,
                               int terms, const float* z, const float* scale, const float* shift, const float* mean,
                               const float* invstd, void* stream);

/* Weight gradient of the same convolution (autograd of model.py:13 via train.py:95):
 * dw[cout][cin_real][3][3] += sum dz * view. dw must be zeroed by the caller. dz is in the pre-split bf16 format
 * ([N,H,W,cout] logical) or, with fmt = 0, fp16 pairs of dz * (*dz_mul) (dz_mul: device scalar, may be NULL = 1; the
 * result is divided by it); both operands must be of the SAME 16-bit format (fmt: 0 fp16, 1 bf16), a view operand that is
 * not pre-split is split to that format on the fly. With a pre-split view three kernels exist:
 * CTA pairs (tcgen05 cta_group::2) when cout % 256 == 0 and the input tile is 128 channels, the tap-stacked kernel for
 * cout == 64, the single-CTA kernel otherwise; variant bit 32 forces the generic kernels, bit 64 the single-CTA one. */
int tnb_conv3x3_wgrad(const tnb_view_t* view, const void* dz_presplit, float* dw_oihw, int cout, int cin_real,
                      int terms, int variant, int fmt, const float* dz_mul, void* stream);
/* The same, DETERMINISTIC, with a scratch buffer of tnb_conv3x3_wgrad_ws_elems(view, cout) floats: every split-K CTA
 * stores its partial tap-major into its own slab (plain coalesced stores, no atomics) and a second small kernel sums the
 * slabs in split order into dw_oihw (dw need not be zeroed). Gradients are bit-identical from run to run, which is what
 * the reference asks of cuDNN with `torch.backends.cudnn.deterministic = True` (train.py:205). What
 * tnb_tracknet_backward uses. */
size_t tnb_conv3x3_wgrad_ws_elems(const tnb_view_t* view, int cout);
int tnb_conv3x3_wgrad_ws(const tnb_view_t* view, const void* dz_presplit, float* dw_oihw, int cout, int cin_real,
                         int terms, int variant, float* scratch, int fmt, const float* dz_mul, void* stream);

/* BatchNorm2d statistics -> fused affine + running-stat update (model.py:9; torch defaults eps 1e-5,
 * momentum 0.1, unbiased running_var). training==0 uses the running statistics (model.eval()). */
int tnb_bn_finalize(const float* stat_part, int rows, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, int training,
                    float* scale, float* shift, float* mean, float* invstd, int c, void* stream);

/* BatchNorm+ReLU backward with MaxPool/Upsample/cat gradient routing (autograd of model.py:13-15,59-69). */
int tnb_bn_bwd_blocks(int n, int h, int w, int c);
int tnb_bn_relu_bwd_reduce(const tnb_bnbwd_t* args, void* stream);
int tnb_bn_relu_bwd_finalize(const float* part, int rows, int c, float* sums, float* dgamma, float* dbeta, void* stream);
int tnb_bn_relu_bwd_apply(const tnb_bnbwd_t* args, void* stream);

/* predictor: 1x1 conv + bias + sigmoid (model.py:54-55,71-72). y / dy are NCHW [n,out_dim,h,w]; any out_dim >= 1 (the
 * reference takes any seq_len, utils/general.py:66-74). The backward needs a workspace of
 * tnb_conv1x1_bias_sigmoid_bwd_workspace_bytes() bytes: per-block partial sums of dweight / dbias, added in block order
 * (deterministic; dweight and dbias are overwritten, not accumulated). */
int tnb_conv1x1_bias_sigmoid_fwd(const tnb_src_t* src, int n, int h, int w, const float* weight, const float* bias,
                                 int out_dim, float* y_nchw, void* stream);
size_t tnb_conv1x1_bias_sigmoid_bwd_workspace_bytes(int n, int h, int w, int out_dim);
int tnb_conv1x1_bias_sigmoid_bwd(const tnb_src_t* src, int n, int h, int w, const float* weight, int out_dim,
                                 const float* dy_nchw, const float* y_nchw, float* d_act_nhwc, float* dweight,
                                 float* dbias, void* workspace, void* stream);

/* WBCELoss(y_pred, y, reduce) (utils/metric.py:3-20). out: 1 float (reduce) or nsamples floats.
 * part: workspace of tnb_wbce_workspace_bytes(nsamples) bytes. gout: upstream gradient (1 or nsamples). */
size_t tnb_wbce_workspace_bytes(int nsamples);
int tnb_wbce_fwd(const float* y_pred, const float* y, int nsamples, long long per_sample, int reduce, void* part,
                 float* out, void* stream);
int tnb_wbce_bwd(const float* y_pred, const float* y, const float* gout, int nsamples, long long per_sample,
                 int reduce, float* d_y_pred, void* stream);

/* mixup (train.py:37-38): out[i] = x[i]*lam[i] + x[perm[i]]*(1-lam[i]). */
int tnb_mixup(const float* x, const float* lam, const long long* perm, float* out, int n, long long per_sample,
              void* stream);

/* torch.optim.Adam step over many tensors in one launch (train.py:96,242). table: device array of
 * {float* p; const float* g; float* m; float* v; long long n;}; total_n = the sum of n over the table (sizes the grid: one
 * CTA per 4096-element chunk of a tensor). step counts from 1. */
int tnb_adam_multi(const void* table_dev, int ntensors, long long total_n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, void* stream);

/* predict_location (test.py:52-79) for a batch of maps on the GPU. maps: nmaps x h x w, float (foreground =
 * value > thresh, predict.py:35) or uint8 (foreground = non-zero, as to_img() output). out: nmaps x 4 int32
 * (x, y, w, h), all zeros for an empty map. */
size_t tnb_heatmap_decode_workspace_bytes(int nmaps, int h, int w);
int tnb_heatmap_decode(const void* maps, int is_u8, float thresh, int nmaps, int h, int w, void* workspace,
                       int* out_xywh, void* stream);

/* InpaintNet.forward(x, m) (model.py:113-129) in one kernel. params: 18 device pointers in state_dict
 * order (down_1.conv.weight, down_1.conv.bias, ..., predictor.weight, predictor.bias). */
int tnb_inpaintnet_fwd(const float* coords, const float* mask, const void* const* params, int n, int l, float* out,
                       void* stream);

/* The rectified trajectory of the InpaintNet inference loops (predict.py:256-261, test.py:401,406-408) in the same
 * launch: out = InpaintNet(coords, mask) * mask + coords * (1 - mask), then both coordinates of a point set to 0 where
 * both are < coor_th (COOR_TH, utils/general.py:19). Operation order and roundings are torch's (no fused multiply-add). */
int tnb_inpaintnet_rectify(const float* coords, const float* mask, const void* const* params, int n, int l, float coor_th,
                           float* out, void* stream);

/* Autograd of the above, as train.py:147-166 needs it (loss.backward() through InpaintNet): one kernel that
 * recomputes the forward in shared memory. dout: (n, l, 2) gradient w.r.t. the output. grads: 18 device pointers
 * (same order as params) the parameter gradients are ADDED into (caller zeroes them). dcoords: optional (n, l, 2)
 * gradient w.r.t. coords, may be NULL. l <= 28. */
int tnb_inpaintnet_bwd(const float* coords, const float* mask, const void* const* params, const float* dout,
                       void* const* grads, int n, int l, float* dcoords, void* stream);

/* Frame preprocessing (dataset.py:435-461, 783-812): Pillow's antialiased BICUBIC `img.resize((wd, hd))` of nimg uint8
 * images [nimg][hs][ws][c] (22-bit fixed point, horizontal pass then vertical pass, bit-exact against Pillow), then
 * HWC -> CHW, / 255 and stacking: image i lands in out + (i / per_sample) * sample_stride +
 * ((i % per_sample) * frame_stride + chan_off) * hd * wd, as c float planes (frame_stride = channels every frame
 * occupies in the stack: c, or 4 for bg_mode 'subtract_concat'). Coefficient tables (device memory) as Pillow's
 * precompute_coeffs / normalize_coeffs_8bpc build them: bounds = (first tap, tap count) per output sample, kk = ksize
 * ints per output sample; hbounds/hkk NULL when ws == wd (Pillow skips that pass); the vertical table must always
 * be given (identity table when hs == hd). tmp: nimg * hs * wd * c bytes of scratch. */
int tnb_resize_frames(const uint8_t* src, int nimg, int hs, int ws, int c, const int* hbounds, const int* hkk, int hksize,
                      const int* vbounds, const int* vkk, int vksize, int hd, int wd, uint8_t* tmp, float* out,
                      int per_sample, long long sample_stride, int chan_off, int frame_stride, void* stream);

/* Median background of a clip (dataset.py:102-107: `np.median(self.frame_arr, 0)`): frames uint8 [nframes][frame_bytes]
 * (frame_bytes = hs * ws * 3). Per byte position the two middle order statistics a <= b of the nframes values;
 * out_f64[p] = (a + b) / 2 - numpy's float64 median, kept by bg_mode 'subtract' / 'subtract_concat' (dataset.py:108-109) -
 * and out_u8[p] = floor((a + b) / 2) = `median.astype('uint8')` (bg_mode 'concat', dataset.py:105). Either may be NULL. */
int tnb_median_u8(const uint8_t* frames, int nframes, long long frame_bytes, double* out_f64, uint8_t* out_u8,
                  void* stream);

/* Training labels (dataset.py:400-410 `_get_heatmap`, called with integer centres at :632): map m = 1 where
 * (i - cx)^2 + (j - cy)^2 <= sigma^2 for pixel (row j, column i), all zeros when cx == cy == 0. centers_xy: device int32
 * [nmaps][2]. out: [nmaps][h][w] fp32, w % 4 == 0. */
int tnb_label_discs(const int* centers_xy, int nmaps, int h, int w, float sigma, float* out, void* stream);

/* Background-difference image of bg_mode 'subtract' / 'subtract_concat' (dataset.py:438, 442):
 * out[i][y][x] = uint8 cast (numpy semantics: truncate, wrap modulo 256) of sum_c |frames[i][y][x][c] - median[y][x][c]|,
 * frames uint8 [nimg][hs][ws][3], median float64 [hs][ws][3]. The result goes through tnb_resize_frames with c = 1. */
int tnb_bg_subtract_u8(const uint8_t* frames, const double* median, long long nimg, int hs, int ws, uint8_t* out,
                       void* stream);

/* Per-map statistics of the evaluation bookkeeping (test.py:159-169): conf[m] = max of y_pred map m inside
 * boxes[m] = (x, y, w, h) (0 for an empty box); true_any[m] = 1 iff map m of y_true has a value > 0 (y_true and
 * true_any may be NULL). boxes: int32 as written by tnb_heatmap_decode. */
int tnb_eval_stats(const float* y_pred, const float* y_true, const int* boxes_xywh, int nmaps, int h, int w,
                   float* conf, int* true_any, void* stream);

/* Temporal ensemble of sliding-window predictions (predict.py:163-209 heatmaps, :245-301 coordinates). state: the
 * previous seq_len-1 samples' predictions [(seq_len-1)][seq_len][frame_elems] (zeros before the first batch); pred:
 * this batch [batch][seq_len][frame_elems]; weight_host: seq_len floats in HOST memory (test.py:25-50). sample_count:
 * samples consumed before this batch. n_tail = seq_len-1 when this batch contains the last sample (index tail_base in
 * the batch): the remaining frames are appended, else 0. out: [batch + n_tail][frame_elems]. */
int tnb_temporal_ensemble(const float* state, const float* pred, float* out, const float* weight_host, int seq_len,
                          long long frame_elems, int batch, int sample_count, int tail_base, int n_tail, void* stream);

/* ---- network level: TrackNet.forward / its autograd backward (model.py:57-73) --------------- */

/* params: 104 device pointers in state_dict order: for each of the 17 Conv2DBlocks
 *   conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked (int64),
 * then predictor.weight, predictor.bias.
 * grads:  53 device pointers in parameters() order: per block conv.weight, bn.weight, bn.bias grads, then
 * predictor.weight, predictor.bias grads (all overwritten, not accumulated). */
size_t tnb_tracknet_workspace_bytes(const tnb_tracknet_cfg_t* cfg);
int tnb_tracknet_forward(const tnb_tracknet_cfg_t* cfg, const float* x_nchw, void* const* params, float* y_nchw,
                         void* workspace, size_t workspace_bytes, void* stream);
int tnb_tracknet_backward(const tnb_tracknet_cfg_t* cfg, const float* dy_nchw, const float* y_nchw,
                          void* const* params, void* const* grads, void* workspace, size_t workspace_bytes,
                          void* stream);
/* The backward pass in pieces, for data-parallel training (the reference has no multi-GPU code; torch DDP overlaps its
 * bucketed allreduce with the backward the same way): layers layer_hi .. layer_lo of the 17 convolution blocks in
 * backward order (state_dict numbering, 16 = up_block_3.conv_2 ... 0 = down_block_1.conv_1); layer_hi = 16 also runs the
 * predictor's backward. Calling (16, s) and then (s - 1, 0) equals tnb_tracknet_backward; after the first call returns,
 * the gradients of layers >= s and of the predictor are final on `stream` and can be reduced across ranks on another
 * stream while the second call computes. tnb_tracknet_grad_split_layer() = the s this library recommends (7: 85 % of the
 * parameters lie behind it and no weight gradient is pending across that boundary). */
int tnb_tracknet_backward_range(const tnb_tracknet_cfg_t* cfg, const float* dy_nchw, const float* y_nchw,
                                void* const* params, void* const* grads, void* workspace, size_t workspace_bytes,
                                int layer_hi, int layer_lo, void* stream);
int tnb_tracknet_grad_split_layer(void);
/* tnb_tracknet_forward / tnb_tracknet_backward replay their launch sequence from a CUDA graph once the same
 * argument set (cfg and every pointer) is seen again; per-launch profiling and a caller-side stream capture bypass it.
 * This switch turns the replay off (0) or on (1, default; env TNB_GRAPHS=0 also disables); returns the old value. */
int tnb_set_graph_replay(int on);
/* out4 = {graphs captured, graph replays, eager (stream-launched) calls while replay was allowed, capture failures} */
int tnb_graph_stats(long long* out4);
/* number of kernels launched by one forward / backward call (for bench.py's gpu_launches) */
int tnb_tracknet_num_launches(const tnb_tracknet_cfg_t* cfg, int backward);
/* Debugging aid (tools/diag_layers.py): where layer `layer` (0..16, state_dict order) keeps its tensors inside a
 * workspace laid out for cfg. out_ptr[0..7] = z (raw conv output, fp32 NHWC), BatchNorm scale, shift, mean, invstd
 * ([cout] fp32 each), dz (gradient w.r.t. z in the pre-split 16-bit format the backward pass uses), din (gradient w.r.t.
 * the layer's input view, fp32 NHWC [N,H,W,cin]; NULL for layer 0), dz multiplier (device scalar; fp16 pairs only, else
 * NULL); out_dim[0..4] = H, W, cin, cout, 16-bit format of dz (0 fp16 pairs of dz * multiplier, 1 bf16 pairs). Pointers
 * are valid after tnb_tracknet_forward (z, scale ...) / tnb_tracknet_backward (dz, din) of the same cfg and workspace.
 * Pure host arithmetic: with workspace = NULL the "pointers" are the tensors' byte offsets inside a workspace. */
int tnb_tracknet_debug_layer(const tnb_tracknet_cfg_t* cfg, void* workspace, int layer, void** out_ptr8, int* out_dim5);

/* ---- measurement support (bench.py roofline leg) -------------------------------------------- */
/* Per-launch CUDA-event timing of the tensor-core and BN-backward kernels, recorded on the launching
 * stream. enable(1) clears and starts recording, enable(0) stops. collect() synchronises and returns the
 * number of records: desc[6*i..] = {kind (0 conv fwd, 1 dgrad, 2 wgrad, 3 bn_bwd, 4 predictor), n, h, w,
 * cin, cout}, ms[i] = device duration of that launch. */
int tnb_profile_enable(int on);
int tnb_profile_collect(int max_records, int* desc, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* TRACKNET_B200_H */
