// Shared declarations of libcimq (internal).  See include/cimq.h for the public C ABI.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cimq.h"

namespace cimq {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define CIMQ_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      ::cimq::set_error(__VA_ARGS__);  \
      return 1;                        \
    }                                  \
  } while (0)

#define CIMQ_CUDA_OK(expr)                                                                \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::cimq::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                        \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

// Multi-bit ADC STE clip bounds (lsq.py:310-311): clipped where ps >= Qp + 1e-5 or ps <= Qn - 1e-5, compared in
// fp32.  chi = smallest integer partial sum that is clipped high, clo = largest one clipped low.  For small ranges
// that is qp+1 / qn-1; where 1e-5 is below half an ulp of the bound (Qp >= 512, Qn <= -256) the bound itself clips.
inline void multibit_clip_bounds(int qn, int qp, int *chi, int *clo) {
  const float hi = (float)((double)qp + 1e-5), lo = (float)((double)qn - 1e-5);
  int c = qp;
  while ((float)c < hi) ++c;
  *chi = c;
  c = qn;
  while ((float)c > lo) --c;
  *clo = c;
}

// ---- derived layer geometry -------------------------------------------------------------------
// Plain-old-data copy of the layer with everything the kernels need; passed by value.
struct Geo {
  int B, Cin, H, W, Cout, K, stride, pad;
  int OH, OW, L, M, F, KK;          // KK = K*K
  int NX, NSW, NSA, pairs;          // crossbars, weight / act digit planes, NSW*NSA
  int abits, abs_, wbits, wbs;      // activation / weight bits and bits per slice
  int amask, wmask;                 // (1<<abs_)-1, (1<<wbs)-1
  int xbar;
  int adc_mode, qn, qp;
  int clip_hi, clip_lo;             // multi-bit ADC: the STE mask is off where psum >= clip_hi or psum <= clip_lo
  int state_bits, state_words;      // bits of ADC state per (crossbar, channel, pixel)
};

inline int make_geo(const cimq_layer_t *l, Geo *g) {
  CIMQ_REQUIRE(l != nullptr, "layer is NULL");
  CIMQ_REQUIRE(l->batch > 0 && l->in_channels > 0 && l->in_hw > 0 && l->out_channels > 0, "bad layer shape");
  CIMQ_REQUIRE(l->kernel > 0 && l->stride > 0 && l->padding >= 0, "bad kernel/stride/padding");
  CIMQ_REQUIRE(l->abitslice > 0 && l->wbitslice > 0 && l->nbits_a >= l->abitslice && l->nbits_w >= l->wbitslice,
               "bad bit widths");
  CIMQ_REQUIRE(l->nbits_a <= 8 && l->nbits_w <= 8, "codes are one byte: nbits_a, nbits_w <= 8");
  CIMQ_REQUIRE(l->xbar > 0, "xbar must be positive");
  CIMQ_REQUIRE(l->adc_mode >= 0 && l->adc_mode <= 2, "bad adc_mode");
  g->B = l->batch; g->Cin = l->in_channels; g->H = l->in_hw; g->W = l->in_hw;
  g->Cout = l->out_channels; g->K = l->kernel; g->stride = l->stride; g->pad = l->padding;
  g->OH = (g->H + 2 * g->pad - g->K) / g->stride + 1;
  g->OW = g->OH;
  CIMQ_REQUIRE(g->OH > 0, "empty output");
  g->L = g->OH * g->OW; g->M = g->B * g->L;
  g->KK = g->K * g->K; g->F = g->Cin * g->KK;
  g->xbar = l->xbar;
  g->NX = (g->F + g->xbar - 1) / g->xbar;
  g->abits = l->nbits_a; g->abs_ = l->abitslice; g->wbits = l->nbits_w; g->wbs = l->wbitslice;
  g->NSA = g->abits / g->abs_; g->NSW = g->wbits / g->wbs;   // int(bits/bit_slice), _quan_base.py:203-204
  g->pairs = g->NSA * g->NSW;
  g->amask = (1 << g->abs_) - 1; g->wmask = (1 << g->wbs) - 1;
  g->adc_mode = l->adc_mode; g->qn = l->adc_qn; g->qp = l->adc_qp;
  multibit_clip_bounds(g->qn, g->qp, &g->clip_hi, &g->clip_lo);
  g->state_bits = (g->adc_mode == CIMQ_ADC_MULTIBIT ? 1 : 3) * g->pairs;
  g->state_words = (g->state_bits + 31) / 32;
  return 0;
}

// ---- ADC state bit layout ------------------------------------------------------------------------
// Per (crossbar i, channel c, pixel m) the forward stores `state_words` uint32 at
//   state[((i*Cout + c)*state_words + w)*M + m]
// holding, for weight slice k and activation slice j, with the state pair index sq = j*NSW + k:
//   binary / ternary ADC : bit sq = code is +1, bit pairs+sq = code is -1, bit 2*pairs+sq = clipped (STE mask off)
//   multi-bit ADC        : bit sq = clipped
// (sq is activation-slice major so that the forward epilogue, which walks the weight slices of one
//  activation digit plane in a rolled loop, writes consecutive bits.)
__host__ __device__ inline int state_pair(const Geo &g, int k, int j) { return j * g.NSW + k; }
__host__ __device__ inline int state_clip_bit(const Geo &g, int sq) {
  return g.adc_mode == CIMQ_ADC_MULTIBIT ? sq : 2 * g.pairs + sq;
}

// fp16 round trip of an integer partial sum (the reference stores psums as fp16, lsq.py:169).
__device__ __forceinline__ float psum_as_stored(int p) {
  return __half2float(__float2half_rn((float)p));
}

// ---- kernel launchers (one per .cu file) ---------------------------------------------------------
int launch_step_sizes(const float *aa, const float *aw, float ga, float gw, float *s, cudaStream_t st);
int launch_lsq_quantize(const float *x, int64_t n, const float *s, int qn, int qp, void *codes, bool from_fq,
                        cudaStream_t st);
int launch_lsq_fakequant(const float *x, int64_t n, const float *s, int qn, int qp, int rescale, float *y,
                         cudaStream_t st);
int launch_lsq_backward(const float *gq, const float *x, int64_t n, const float *s, int qn, int qp, float g,
                        float *gx, float *galpha, void *ws, cudaStream_t st);
int64_t lsq_backward_ws_bytes(int64_t n);

int launch_adc_table(const Geo &g, const float *s, const float *alpha_q, const float *alpha_scale, const int8_t *mask,
                     void *table, int32_t *status, cudaStream_t st);
int launch_alpha_quant(const float *alpha, int64_t n, int qn, int qp, const float *gaq, float *out, float *aux,
                       cudaStream_t st);
int launch_weight_prepare(const Geo &g, const int8_t *wcodes, float *wdigits, void *wtiles, cudaStream_t st);
int64_t wtiles_bytes(const Geo &g);
bool tc_forward_supported(const Geo &g);
int tc_channel_tile_for(const Geo &g);          // output channels per CTA of the tcgen05 kernel (0: unsupported)
int launch_im2col_lut(const Geo &g, void *lut, cudaStream_t st);
int64_t bn_workspace_bytes(int B, int C);
int launch_bn_forward(const float *x, const float *residual, const float *weight, const float *bias,
                      float *running_mean, float *running_var, int training, float momentum, float eps, int relu,
                      int B, int C, int HW, float *y, float *save_mean, float *save_invstd, void *workspace,
                      cudaStream_t st, const float *q_alpha = nullptr, float q_g = 0.0f, int q_qp = 0,
                      uint8_t *codes = nullptr);
int launch_bn_backward(const float *gy, const float *x, const float *y, const float *weight, const float *save_mean,
                       const float *save_invstd, int training, int relu, int B, int C, int HW, float *gx, float *gres,
                       float *gweight, float *gbias, void *workspace, cudaStream_t st);
extern long long *g_tc_debug;
bool tc_backward_supported(const Geo &g);
int64_t wtiles_bwd_bytes(const Geo &g);
int launch_weight_tiles_bwd(const Geo &g, const int8_t *wcodes, void *tiles, cudaStream_t st);
int64_t bwd_tc_partial_bytes(const Geo &g);
bool bwd_input_tc_can_fold(const Geo &g);
// fold = 1: `out` is grad_x [B,Cin,H,W], zero-filled, and the kernel folds into it; fold = 0: `out` is the
// unfolded gradient gxu[b][f][l] for launch_col2im
// v2 = true: `state` is the v2 byte planes (cim_v2.cuh) and `wtb` the pre-scaled dgrad tiles (WtLayout::bwd2)
int launch_bwd_input_tc(const Geo &g, const float *go, const uint32_t *state, const void *wtb, const float *s,
                        const int8_t *mask, float *out, int fold, bool v2, cudaStream_t st);
// v2 = true also needs `scales` (launch_go_scales)
int launch_bwd_weight_tc(const Geo &g, const float *go, const uint8_t *xcodes, const uint32_t *state,
                         const float *s, const int8_t *mask, float *partial, float *gw, bool v2, const void *scales,
                         cudaStream_t st);
bool v2_backward_supported(const Geo &g);
// ---- v2 backward (cim_bwd_v2.cu + the V2 instance of the wgrad kernel in cim_bwd_tc.cu) ----
int launch_weight_tiles_bwd2(const Geo &g, const int8_t *wcodes, void *tiles, cudaStream_t st);
// grad_out scale pre-pass: `scales` = float rowscale[M] (power of two per pixel row), then (256-byte aligned)
// uint32 chmax[Cout] (bit pattern of the largest |grad_out| of the channel); bwd_v2_scales_bytes() bytes
int64_t bwd_v2_scales_bytes(const Geo &g);
int launch_go_scales(const Geo &g, const float *go, void *scales, cudaStream_t st);
// grad_alpha from plane C, staged through shared memory (cim_alpha_v2.cu); writes alpha_v3_blocks() partials
bool alpha_v3_supported(const Geo &g);
int alpha_v3_blocks(const Geo &g);
int launch_alpha_v3(const Geo &g, const float *go, const uint8_t *cplanes, float *partial, cudaStream_t st);
int launch_bwd_input_v2(const Geo &g, const float *go, const uint8_t *state, const void *wtb, const float *s,
                        const void *scales, float *out, int fold, cudaStream_t st);

// Sections of the prepared-weights buffer ("wtiles"): forward int8 digit tiles, im2col LUT (int2 per crossbar
// row), backward bf16 digit tiles.  A section the layer does not support has zero bytes.
struct WtLayout {
  int64_t fwd_off, fwd_bytes, lut_off, lut_bytes, bwd_off, bwd_bytes;
  int64_t fwd8_off, fwd8_bytes;  // v2 forward: e4m3 digit tiles (+-0.5), same tiling as the int8 ones
  int64_t bwd2_off, bwd2_bytes;  // v2 dgrad: fp16 digit tiles 2^k * digit, columns in dgrad_col_f order (cim_v2.cuh)
  int64_t total;
};
WtLayout wt_layout(const Geo &g);

enum SimtMode { SIMT_FORWARD = 0, SIMT_PSUMS = 1, SIMT_ABS_SUMS = 2, SIMT_FORWARD_STOCH = 3 };
int launch_conv_simt(const Geo &g, int mode, const uint8_t *xcodes, const int8_t *wcodes, const void *table,
                     const float *s, const int8_t *binary_mask, float *out, uint32_t *state, int32_t *psums,
                     unsigned long long *sums, cudaStream_t st, const float *alpha_q = nullptr,
                     unsigned long long seed = 0);
int launch_conv_tc_forward(const Geo &g, const uint8_t *xcodes, const void *wtiles, const void *table,
                           const float *s, const int8_t *mask, float *out, uint32_t *state, cudaStream_t st);

int64_t conv_backward_ws_bytes(const Geo &g);
int launch_conv_backward(const Geo &g, const float *go, const uint8_t *xcodes, const float *wdigits,
                         const void *wtiles, const uint32_t *state, const float *s, const int8_t *mask, float *gxq,
                         float *gwq, float *galpha, void *ws, uint32_t flags, cudaStream_t st);

// table = NX*pairs*Cout entries of 16 bytes: int4 {tp, tg, amp (fp32 bits), 0}, followed (256-byte aligned) by the
// same values tiled for the tcgen05 epilogue: uint32 [Cout/CT][NX][pairs][tp | tg | amp][CT]
__host__ __device__ inline int64_t table_entries(const Geo &g) { return (int64_t)g.NX * g.pairs * g.Cout; }
__host__ __device__ inline int64_t table_tiled_offset(const Geo &g) { return (table_entries(g) * 16 + 255) & ~(int64_t)255; }
// ... followed (256-byte aligned) by the v2 section: header {o0, o1} + constants blocks (cim_v2.cuh)
int64_t table_v2_offset(const Geo &g);
int64_t table_total_bytes(const Geo &g);

int launch_layer_prepare(const Geo &g, const float *weight, const float *alpha_act, const float *alpha_weight, float ga,
                         float gw, const float *alpha_cim, int aq_qn, int aq_qp, const int8_t *mask, float *s,
                         int8_t *wcodes, float *alpha_q, float *aux, void *table, void *wtiles, int32_t *status,
                         cudaStream_t st);
bool v2_forward_supported(const Geo &g);
int launch_conv_v2_forward(const Geo &g, const uint8_t *xcodes, const void *wtiles, const void *table, float *out,
                           uint8_t *state, cudaStream_t st);

}  // namespace cimq
