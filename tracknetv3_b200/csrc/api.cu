// extern "C" surface of libtracknet_b200.so (declared in include/tracknet_b200.h).
#include "kernels.cuh"
#include <stdarg.h>

namespace tnb {
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
size_t tracknet_workspace_bytes(const tnb_tracknet_cfg_t& c);
int tracknet_forward(const tnb_tracknet_cfg_t& c, const float* x, void* const* params, float* y, void* ws,
                     size_t ws_bytes, cudaStream_t st);
int tracknet_backward(const tnb_tracknet_cfg_t& c, const float* dy, const float* y, void* const* params,
                      void* const* grads, void* ws, size_t ws_bytes, cudaStream_t st, int hi, int lo);
int tracknet_grad_split_layer();
int tracknet_num_launches(const tnb_tracknet_cfg_t& c, int backward);
int tracknet_debug_layer(const tnb_tracknet_cfg_t& c, void* ws, int layer, void** out_ptr8, int* out_dim5);
int set_graph_replay(int on);
void graph_stats(long long* out4);
}  // namespace tnb

using namespace tnb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

#ifdef TNB_TRACE
namespace tnb { long long* g_conv_trace = nullptr; }
#endif

extern "C" {

const char* tnb_last_error(void) { return g_err; }
int tnb_abi_version(void) { return TNB_ABI_VERSION; }

int tnb_pack_nchw_to_nhwc(const float* x, float* out, int n, int c, int h, int w, int cpad, void* stream) {
  TNB_REQUIRE(cpad % 4 == 0 && cpad >= c, "pack_nchw_to_nhwc: bad cpad %d for c %d", cpad, c);
  return launch_pack_input(x, out, n, c, h, w, cpad, ST(stream));
}
int tnb_pack_nchw_to_planar16(const float* x, void* out_planar16, float* out_nhwc, int n, int c, int h, int w, int cpad,
                              void* stream) {
  return launch_pack_input(x, out_nhwc, n, c, h, w, cpad, ST(stream), out_planar16);
}
int tnb_presplit_bf16(const float* x, void* out, long long npixels, int c, void* stream) {
  return launch_presplit(x, out, npixels, c, 1, 1.f, ST(stream));
}
int tnb_presplit_fp16(const float* x, void* out, long long npixels, int c, float mul, void* stream) {
  return launch_presplit(x, out, npixels, c, 0, mul, ST(stream));
}
int tnb_view_presplit(const tnb_view_t* view, void* out, int fmt, void* stream) {
  return launch_view_presplit(*view, out, fmt, ST(stream));
}
size_t tnb_conv3x3_wpack_elems(int k_side, int n_side) { return conv3x3_wpack_elems(k_side, n_side); }
int tnb_conv3x3_pack_weights(const float* w, uint16_t* out, int cout, int cin, int mode, int fmt, void* stream) {
  // the tile width must match what the conv launcher will pick for this N side
  ConvPlan p;
  const int nside = mode == 0 ? cout : cin, kside = mode == 0 ? cin : cout;
  if (int rc = conv3x3_plan(1, 16, 16, (kside + 31) / 32 * 32, nside, 3, &p)) return rc;
  return launch_pack_weights(w, out, cout, cin, mode, fmt, p.BN, ST(stream));
}
int tnb_conv3x3_plan_query(int n, int h, int w, int cin, int cout, int terms, int* out12) {
  ConvPlan p;
  if (int rc = conv3x3_plan(n, h, w, (cin + 31) / 32 * 32, cout, terms, &p)) return rc;
  const int v[12] = {p.BN, p.MT, p.SA, p.SB, p.G, p.nbuf, p.tmem_cols, (int)p.smem_bytes, p.merged, p.tall, 0 /* (was: CTA-pair flag) */,
                     conv3x3_weight_layout(p.BN)};
  for (int i = 0; i < 12; ++i) out12[i] = v[i];
  return 0;
}
int tnb_conv3x3_stat_rows(int n, int h, int w, int cin, int cout, int terms) {
  return conv3x3_num_stat_rows(n, h, w, cin, cout, terms);
}
int tnb_conv3x3_fwd(const tnb_view_t* view, const uint16_t* wpack, float* out, float* stat_part, int cout, int terms,
                    int fmt, int variant, void* stream) {
  return launch_conv3x3(*view, wpack, out, stat_part, cout, terms, fmt, variant, ST(stream));
}
int tnb_conv3x3_dgrad_bnreduce_rows(int n, int h, int w, int cin, int cout, int terms) {
  return conv3x3_num_stat_rows(n, h, w, cin, cout, terms, true);
}
int tnb_conv3x3_dgrad_bnreduce(const tnb_view_t* view, const uint16_t* wpack, float* out, float* part, int cout,
                               int terms, const float* z, const float* scale, const float* shift, const float* mean,
                               const float* invstd, void* stream) {
  const BnBwdFuse fuse{z, scale, shift, mean, invstd};
  return launch_conv3x3(*view, wpack, out, part, cout, terms, 1, 0, ST(stream), &fuse);
}
int tnb_conv3x3_wgrad(const tnb_view_t* view, const void* dz_presplit, float* dw, int cout, int cin_real, int terms,
                      int variant, int fmt, const float* dz_mul, void* stream) {
  return launch_wgrad3x3(*view, dz_presplit, dw, cout, cin_real, terms, variant, ST(stream), nullptr, fmt, dz_mul);
}
size_t tnb_conv3x3_wgrad_ws_elems(const tnb_view_t* view, int cout) { return wgrad3x3_ws_floats(*view, cout); }
int tnb_conv3x3_wgrad_ws(const tnb_view_t* view, const void* dz_presplit, float* dw, int cout, int cin_real, int terms,
                         int variant, float* scratch, int fmt, const float* dz_mul, void* stream) {
  return launch_wgrad3x3(*view, dz_presplit, dw, cout, cin_real, terms, variant, ST(stream), scratch, fmt, dz_mul);
}
int tnb_bn_finalize(const float* part, int rows, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, int training, float* scale,
                    float* shift, float* mean, float* invstd, int c, void* stream) {
  return launch_bn_finalize(part, rows, count, gamma, beta, running_mean, running_var, momentum, eps, training, scale,
                            shift, mean, invstd, c, ST(stream));
}
int tnb_bn_bwd_blocks(int n, int h, int w, int c) { return bn_bwd_num_blocks(n, h, w, c); }
int tnb_bn_relu_bwd_reduce(const tnb_bnbwd_t* a, void* stream) { return launch_bn_bwd_reduce(*a, ST(stream)); }
int tnb_bn_relu_bwd_finalize(const float* part, int rows, int c, float* sums, float* dgamma, float* dbeta,
                             void* stream) {
  return launch_bn_bwd_finalize(part, rows, c, sums, dgamma, dbeta, ST(stream));
}
int tnb_bn_relu_bwd_apply(const tnb_bnbwd_t* a, void* stream) { return launch_bn_bwd_apply(*a, ST(stream)); }

int tnb_conv1x1_bias_sigmoid_fwd(const tnb_src_t* src, int n, int h, int w, const float* weight, const float* bias,
                                 int out_dim, float* y, void* stream) {
  return launch_predictor_fwd(*src, n, h, w, weight, bias, out_dim, y, ST(stream));
}
size_t tnb_conv1x1_bias_sigmoid_bwd_workspace_bytes(int n, int h, int w, int out_dim) {
  return predictor_bwd_workspace_bytes(n, h, w, out_dim);
}
int tnb_conv1x1_bias_sigmoid_bwd(const tnb_src_t* src, int n, int h, int w, const float* weight, int out_dim,
                                 const float* dy, const float* y, float* d_act, float* dweight, float* dbias,
                                 void* workspace, void* stream) {
  return launch_predictor_bwd(*src, n, h, w, weight, out_dim, dy, y, d_act, dweight, dbias, (float*)workspace, ST(stream));
}

size_t tnb_wbce_workspace_bytes(int nsamples) { return sizeof(double) * (size_t)nsamples * wbce_num_blocks(0); }
int tnb_wbce_fwd(const float* p, const float* y, int nsamples, long long per_sample, int reduce, void* part,
                 float* out, void* stream) {
  return launch_wbce_fwd(p, y, nsamples, per_sample, reduce, (double*)part, out, ST(stream));
}
int tnb_wbce_bwd(const float* p, const float* y, const float* gout, int nsamples, long long per_sample, int reduce,
                 float* dp, void* stream) {
  return launch_wbce_bwd(p, y, gout, nsamples, per_sample, reduce, dp, ST(stream));
}
int tnb_mixup(const float* x, const float* lam, const long long* perm, float* out, int n, long long per_sample,
              void* stream) {
  return launch_mixup(x, lam, perm, out, n, per_sample, ST(stream));
}
int tnb_adam_multi(const void* table, int ntensors, long long total_n, float lr, float b1, float b2, float eps, float wd,
                   int step, void* stream) {
  TNB_REQUIRE(step >= 1, "adam_multi: step counts from 1");
  return launch_adam((const AdamTensor*)table, ntensors, total_n, lr, b1, b2, eps, wd, step, ST(stream));
}
size_t tnb_heatmap_decode_workspace_bytes(int nmaps, int h, int w) { return decode_workspace_bytes(nmaps, h, w); }
int tnb_heatmap_decode(const void* maps, int is_u8, float thresh, int nmaps, int h, int w, void* workspace, int* out,
                       void* stream) {
  return launch_decode(maps, is_u8, thresh, nmaps, h, w, workspace, out, ST(stream));
}
int tnb_inpaintnet_fwd(const float* coords, const float* mask, const void* const* params, int n, int l, float* out,
                       void* stream) {
  InpaintParams p;
  for (int i = 0; i < 9; ++i) { p.w[i] = (const float*)params[2 * i]; p.b[i] = (const float*)params[2 * i + 1]; }
  return launch_inpaint_fwd(coords, mask, p, n, l, out, -1.f, ST(stream));
}
int tnb_inpaintnet_rectify(const float* coords, const float* mask, const void* const* params, int n, int l, float coor_th,
                           float* out, void* stream) {
  TNB_REQUIRE(coor_th >= 0.f, "inpaintnet_rectify: negative threshold %f", coor_th);
  InpaintParams p;
  for (int i = 0; i < 9; ++i) { p.w[i] = (const float*)params[2 * i]; p.b[i] = (const float*)params[2 * i + 1]; }
  return launch_inpaint_fwd(coords, mask, p, n, l, out, coor_th, ST(stream));
}
int tnb_median_u8(const uint8_t* frames, int nframes, long long frame_bytes, double* out_f64, uint8_t* out_u8,
                  void* stream) {
  return launch_median_u8(frames, nframes, frame_bytes, out_f64, out_u8, ST(stream));
}
int tnb_label_discs(const int* centers_xy, int nmaps, int h, int w, float sigma, float* out, void* stream) {
  return launch_label_discs(centers_xy, nmaps, h, w, sigma, out, ST(stream));
}
int tnb_inpaintnet_bwd(const float* coords, const float* mask, const void* const* params, const float* dout,
                       void* const* grads, int n, int l, float* dcoords, void* stream) {
  InpaintParams p;
  InpaintGrads g;
  for (int i = 0; i < 9; ++i) {
    p.w[i] = (const float*)params[2 * i]; p.b[i] = (const float*)params[2 * i + 1];
    g.w[i] = (float*)grads[2 * i];        g.b[i] = (float*)grads[2 * i + 1];
  }
  return launch_inpaint_bwd(coords, mask, p, dout, g, n, l, dcoords, ST(stream));
}

int tnb_resize_frames(const uint8_t* src, int nimg, int hs, int ws, int c, const int* hbounds, const int* hkk, int hksize,
                      const int* vbounds, const int* vkk, int vksize, int hd, int wd, uint8_t* tmp, float* out,
                      int per_sample, long long sample_stride, int chan_off, int frame_stride, void* stream) {
  return launch_resize_frames(src, nimg, hs, ws, c, hbounds, hkk, hksize, vbounds, vkk, vksize, hd, wd, tmp, out, per_sample,
                              sample_stride, chan_off, frame_stride, ST(stream));
}
int tnb_bg_subtract_u8(const uint8_t* frames, const double* median, long long nimg, int hs, int ws, uint8_t* out,
                       void* stream) {
  return launch_bg_subtract(frames, median, nimg, hs, ws, out, ST(stream));
}
int tnb_eval_stats(const float* y_pred, const float* y_true, const int* boxes_xywh, int nmaps, int h, int w,
                   float* conf, int* true_any, void* stream) {
  return launch_eval_stats(y_pred, y_true, boxes_xywh, nmaps, h, w, conf, true_any, ST(stream));
}
int tnb_temporal_ensemble(const float* state, const float* pred, float* out, const float* weight_host, int seq_len,
                          long long frame_elems, int batch, int sample_count, int tail_base, int n_tail, void* stream) {
  return launch_temporal_ensemble(state, pred, out, weight_host, seq_len, frame_elems, batch, sample_count, tail_base,
                                  n_tail, ST(stream));
}

size_t tnb_tracknet_workspace_bytes(const tnb_tracknet_cfg_t* cfg) { return tracknet_workspace_bytes(*cfg); }
int tnb_tracknet_forward(const tnb_tracknet_cfg_t* cfg, const float* x, void* const* params, float* y, void* ws,
                         size_t ws_bytes, void* stream) {
  return tracknet_forward(*cfg, x, params, y, ws, ws_bytes, ST(stream));
}
int tnb_tracknet_backward(const tnb_tracknet_cfg_t* cfg, const float* dy, const float* y, void* const* params,
                          void* const* grads, void* ws, size_t ws_bytes, void* stream) {
  return tracknet_backward(*cfg, dy, y, params, grads, ws, ws_bytes, ST(stream), 16, 0);
}
int tnb_tracknet_backward_range(const tnb_tracknet_cfg_t* cfg, const float* dy, const float* y, void* const* params,
                                void* const* grads, void* ws, size_t ws_bytes, int layer_hi, int layer_lo, void* stream) {
  return tracknet_backward(*cfg, dy, y, params, grads, ws, ws_bytes, ST(stream), layer_hi, layer_lo);
}
int tnb_tracknet_grad_split_layer(void) { return tracknet_grad_split_layer(); }
int tnb_set_graph_replay(int on) { return set_graph_replay(on); }
int tnb_graph_stats(long long* out4) { graph_stats(out4); return 0; }
int tnb_tracknet_num_launches(const tnb_tracknet_cfg_t* cfg, int backward) {
  return tracknet_num_launches(*cfg, backward);
}
int tnb_tracknet_debug_layer(const tnb_tracknet_cfg_t* cfg, void* workspace, int layer, void** out_ptr8, int* out_dim5) {
  return tracknet_debug_layer(*cfg, workspace, layer, out_ptr8, out_dim5);
}

#ifdef TNB_TRACE
// debug build only (make TRACE=1 -> libtracknet_b200_trace.so; not part of the ABI): per-role clock64 stamps of CTA 0 of
// the conv kernels into a device buffer of 4 * 24 * 32 int64 (tools/trace_conv.py)
int tnb_debug_set_trace(void* dev_buf) { tnb::g_conv_trace = (long long*)dev_buf; return 0; }
#endif
}  // extern "C"
