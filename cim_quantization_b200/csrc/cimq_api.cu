// extern "C" entry points of libcimq.so (see include/cimq.h).  Argument checking and dispatch only;
// every call enqueues kernels on the caller's stream and returns.
#include <stdarg.h>
#include <string.h>

#include "cimq_common.cuh"
#include "cim_v2.cuh"

namespace cimq {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace cimq

using namespace cimq;

extern "C" {

int cimq_version(void) { return CIMQ_VERSION; }

const char *cimq_last_error(void) { return g_error; }

/* development aid (not declared in cimq.h): device buffer of 16 int64 that CTA 0 of the tcgen05 forward kernel
 * fills with per-role cycle counters; NULL disables */
int cimq_debug_set_timers(long long *buf) { g_tc_debug = buf; return 0; }

int cimq_layer_info(const cimq_layer_t *layer, cimq_info_t *info) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  CIMQ_REQUIRE(info != nullptr, "info is NULL");
  memset(info, 0, sizeof(*info));
  info->out_hw = g.OH; info->L = g.L; info->M = g.M; info->F = g.F;
  info->NX = g.NX; info->NSW = g.NSW; info->NSA = g.NSA; info->pairs = g.pairs;
  info->state_words = g.state_words;
  info->tc_forward = tc_forward_supported(g) ? 1 : 0;
  info->tc_backward = tc_backward_supported(g) ? 1 : 0;
  info->state_bytes = (int64_t)g.NX * g.Cout * g.state_words * g.M * 4;
  info->table_bytes = table_total_bytes(g);
  info->wdigits_bytes = (int64_t)g.NSW * g.Cout * g.F * 4;
  info->wtiles_bytes = wtiles_bytes(g);
  info->bwd_workspace_bytes = conv_backward_ws_bytes(g);
  info->psum_count = (int64_t)g.B * g.NX * g.NSW * g.NSA * g.L * g.Cout;
  info->tc_v2 = (v2_forward_supported(g) && v2_backward_supported(g)) ? 1 : 0;
  info->state_v2_bytes = info->tc_v2 ? v2::state_bytes(g) : 0;
  return 0;
}

int cimq_step_sizes(const float *alpha_act, const float *alpha_weight, float ga, float gw, float *s_out,
                    void *stream) {
  CIMQ_REQUIRE(alpha_act && alpha_weight && s_out, "step_sizes: NULL argument");
  return launch_step_sizes(alpha_act, alpha_weight, ga, gw, s_out, as_stream(stream));
}

int cimq_lsq_quantize(const float *x, int64_t n, const float *s, int32_t qn, int32_t qp, void *codes,
                      void *stream) {
  CIMQ_REQUIRE(n >= 0 && (n == 0 || (x && s && codes)), "lsq_quantize: NULL argument");
  CIMQ_REQUIRE(qn <= qp && qn >= -128 && qp <= 255 && (qn >= 0 || qp <= 127), "lsq_quantize: codes must fit a byte");
  return launch_lsq_quantize(x, n, s, qn, qp, codes, false, as_stream(stream));
}

int cimq_codes_from_fakequant(const float *xq, int64_t n, const float *s, int32_t qn, int32_t qp, void *codes,
                              void *stream) {
  CIMQ_REQUIRE(n >= 0 && (n == 0 || (xq && s && codes)), "codes_from_fakequant: NULL argument");
  CIMQ_REQUIRE(qn <= qp && qn >= -128 && qp <= 255 && (qn >= 0 || qp <= 127),
               "codes_from_fakequant: codes must fit a byte");
  return launch_lsq_quantize(xq, n, s, qn, qp, codes, true, as_stream(stream));
}

int cimq_lsq_fakequant(const float *x, int64_t n, const float *s, int32_t qn, int32_t qp, int32_t rescale, float *y,
                       void *stream) {
  CIMQ_REQUIRE(n >= 0 && (n == 0 || (x && s && y)), "lsq_fakequant: NULL argument");
  return launch_lsq_fakequant(x, n, s, qn, qp, rescale, y, as_stream(stream));
}

int64_t cimq_lsq_backward_workspace_bytes(int64_t n) { return lsq_backward_ws_bytes(n); }

int cimq_lsq_backward(const float *grad_xq, const float *x, int64_t n, const float *s, int32_t qn, int32_t qp,
                      float g, float *grad_x, float *grad_alpha, void *workspace, void *stream) {
  CIMQ_REQUIRE(n > 0 && grad_xq && x && s && grad_x && grad_alpha, "lsq_backward: NULL argument");
  return launch_lsq_backward(grad_xq, x, n, s, qn, qp, g, grad_x, grad_alpha, workspace, as_stream(stream));
}

int cimq_alpha_quantize(const float *alpha, int64_t n, int32_t qn, int32_t qp, float *alpha_q, float *aux,
                        void *stream) {
  CIMQ_REQUIRE(alpha && alpha_q && aux, "alpha_quantize: NULL argument");
  return launch_alpha_quant(alpha, n, qn, qp, nullptr, alpha_q, aux, as_stream(stream));
}

int cimq_alpha_quantize_backward(const float *alpha, const float *grad_alpha_q, int64_t n, int32_t qn, int32_t qp,
                                 const float *aux, float *grad_alpha, void *stream) {
  CIMQ_REQUIRE(alpha && grad_alpha_q && aux && grad_alpha, "alpha_quantize_backward: NULL argument");
  return launch_alpha_quant(alpha, n, qn, qp, grad_alpha_q, grad_alpha, const_cast<float *>(aux), as_stream(stream));
}

int cimq_adc_table(const cimq_layer_t *layer, const float *s, const float *alpha_q, const int8_t *binary_mask,
                   void *table, int32_t *status, void *stream) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  CIMQ_REQUIRE(s && binary_mask && table, "adc_table: NULL argument");
  return launch_adc_table(g, s, alpha_q, nullptr, binary_mask, table, status, as_stream(stream));
}

int cimq_adc_table2(const cimq_layer_t *layer, const float *s, const float *alpha_q, const float *alpha_scale,
                    const int8_t *binary_mask, void *table, int32_t *status, void *stream) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  CIMQ_REQUIRE(s && binary_mask && table, "adc_table2: NULL argument");
  return launch_adc_table(g, s, alpha_q, alpha_scale, binary_mask, table, status, as_stream(stream));
}

int cimq_layer_prepare(const cimq_layer_t *layer, const float *weight, const float *alpha_act,
                       const float *alpha_weight, float ga, float gw, const float *alpha_cim, int32_t aq_qn,
                       int32_t aq_qp, const int8_t *binary_mask, float *s_out, int8_t *wcodes, float *alpha_q,
                       float *aux, void *table, void *wtiles, int32_t *status, void *stream) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  return launch_layer_prepare(g, weight, alpha_act, alpha_weight, ga, gw, alpha_cim, aq_qn, aq_qp, binary_mask, s_out,
                              wcodes, alpha_q, aux, table, wtiles, status, as_stream(stream));
}

int cimq_weight_prepare(const cimq_layer_t *layer, const int8_t *wcodes, float *wdigits, void *wtiles,
                        void *stream) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  CIMQ_REQUIRE(wcodes != nullptr, "weight_prepare: wcodes is NULL");
  return launch_weight_prepare(g, wcodes, wdigits, wtiles, as_stream(stream));
}

int cimq_conv_forward(const cimq_layer_t *layer, const uint8_t *xcodes, const int8_t *wcodes, const void *wtiles,
                      const void *table, const float *s, const int8_t *binary_mask, float *out, void *state,
                      uint32_t flags, void *stream) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  CIMQ_REQUIRE(xcodes && wcodes && table && s && out, "conv_forward: NULL argument");
  if (flags & CIMQ_FLAG_V2) {
    CIMQ_REQUIRE(!(flags & CIMQ_FLAG_FORCE_SIMT) && wtiles != nullptr && v2_forward_supported(g),
                 "conv_forward: CIMQ_FLAG_V2 on a layer the v2 kernels do not cover (see cimq_info_t.tc_v2)");
    return launch_conv_v2_forward(g, xcodes, wtiles, table, out, reinterpret_cast<uint8_t *>(state),
                                  as_stream(stream));
  }
  uint32_t *state1 = reinterpret_cast<uint32_t *>(state);
  if (!(flags & CIMQ_FLAG_FORCE_SIMT) && wtiles != nullptr && tc_forward_supported(g))
    return launch_conv_tc_forward(g, xcodes, wtiles, table, s, binary_mask, out, state1, as_stream(stream));
  return launch_conv_simt(g, SIMT_FORWARD, xcodes, wcodes, table, s, binary_mask, out, state1, nullptr, nullptr,
                          as_stream(stream));
}

int cimq_conv_forward_stochastic(const cimq_layer_t *layer, const uint8_t *xcodes, const int8_t *wcodes,
                                 const void *table, const float *s, const float *alpha_q, float *out,
                                 uint32_t *state, uint64_t seed, void *stream) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  CIMQ_REQUIRE(xcodes && wcodes && table && s && alpha_q && out, "conv_forward_stochastic: NULL argument");
  return launch_conv_simt(g, SIMT_FORWARD_STOCH, xcodes, wcodes, table, s, nullptr, out, state, nullptr, nullptr,
                          as_stream(stream), alpha_q, (unsigned long long)seed);
}

int cimq_conv_backward(const cimq_layer_t *layer, const float *grad_out, const uint8_t *xcodes,
                       const float *wdigits, const void *wtiles, const void *state, const float *s,
                       const int8_t *binary_mask, float *grad_xq, float *grad_wq, float *grad_alpha_q,
                       void *workspace, uint32_t flags, void *stream) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  return launch_conv_backward(g, grad_out, xcodes, wdigits, wtiles, reinterpret_cast<const uint32_t *>(state), s,
                              binary_mask, grad_xq, grad_wq,
                              grad_alpha_q, workspace, flags, as_stream(stream));
}

int cimq_conv_psums(const cimq_layer_t *layer, const uint8_t *xcodes, const int8_t *wcodes, int32_t *psums,
                    void *stream) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  CIMQ_REQUIRE(xcodes && wcodes, "conv_psums: NULL argument");
  return launch_conv_simt(g, SIMT_PSUMS, xcodes, wcodes, nullptr, nullptr, nullptr, nullptr, nullptr, psums,
                          nullptr, as_stream(stream));
}

int cimq_conv_psum_abs_sums(const cimq_layer_t *layer, const uint8_t *xcodes, const int8_t *wcodes,
                            unsigned long long *sums, void *stream) {
  Geo g;
  if (make_geo(layer, &g)) return 1;
  CIMQ_REQUIRE(xcodes && wcodes, "conv_psum_abs_sums: NULL argument");
  return launch_conv_simt(g, SIMT_ABS_SUMS, xcodes, wcodes, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                          sums, as_stream(stream));
}

int64_t cimq_bn_workspace_bytes(int32_t batch, int32_t channels) { return bn_workspace_bytes(batch, channels); }

int cimq_bn_forward(const float *x, const float *residual, const float *weight, const float *bias,
                    float *running_mean, float *running_var, int32_t training, float momentum, float eps,
                    int32_t relu, int32_t batch, int32_t channels, int32_t hw, float *y, float *save_mean,
                    float *save_invstd, void *workspace, void *stream) {
  return launch_bn_forward(x, residual, weight, bias, running_mean, running_var, training, momentum, eps, relu, batch,
                           channels, hw, y, save_mean, save_invstd, workspace, as_stream(stream));
}

int cimq_bn_forward_quant(const float *x, const float *residual, const float *weight, const float *bias,
                          float *running_mean, float *running_var, int32_t training, float momentum, float eps,
                          int32_t relu, int32_t batch, int32_t channels, int32_t hw, float *y, float *save_mean,
                          float *save_invstd, void *workspace, const float *next_alpha_act, float next_grad_scale,
                          int32_t next_qp, uint8_t *next_codes, void *stream) {
  return launch_bn_forward(x, residual, weight, bias, running_mean, running_var, training, momentum, eps, relu, batch,
                           channels, hw, y, save_mean, save_invstd, workspace, as_stream(stream), next_alpha_act,
                           next_grad_scale, next_qp, next_codes);
}

int cimq_bn_backward(const float *grad_y, const float *x, const float *y, const float *weight, const float *mean,
                     const float *invstd, int32_t training, int32_t relu, int32_t batch, int32_t channels, int32_t hw,
                     float *grad_x, float *grad_residual, float *grad_weight, float *grad_bias, void *workspace,
                     void *stream) {
  return launch_bn_backward(grad_y, x, y, weight, mean, invstd, training, relu, batch, channels, hw, grad_x,
                            grad_residual, grad_weight, grad_bias, workspace, as_stream(stream));
}

}  // extern "C"
