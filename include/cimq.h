/*
 * cimq.h -- C ABI of libcimq.so: the B200 (sm_100a) implementation of the CiM-aware quantized
 * convolution path of UtkarshSaxena1/CiM_Quantization (models/_modules/lsq.py).
 *
 * The reference has no FFI: its seam is the Python class / autograd.Function surface
 * (SURVEY.md section 8b).  The host mirror of that surface (cim_quantization_b200/modules,
 * cim_quantization_b200/functional.py) binds these entry points with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.  Each entry point below names the
 * reference code (file:line, relative to the reference tree) it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; plain pointers and sizes
 *     only, no framework types;
 *   - buffers (including workspaces) are owned by the caller; nothing here allocates or frees;
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*): no host
 *     synchronisation, no allocation -> CUDA-graph capturable;
 *   - return value 0 = success, non-zero = error; cimq_last_error() returns a thread-local
 *     message for the last failing call of the calling thread;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Tensor layouts (all contiguous)
 *   activations / activation codes  [B, Cin, H, W]            (NCHW)
 *   weights / weight codes          [Cout, F], F = Cin*k*k, f = c*k*k + ky*k + kx  (nn.Unfold order)
 *   conv output, grad_out           [B, Cout, L], L = OH*OW   (NCHW; the reference's [B,L,Cout] is its transpose)
 *   per-psum tables                 [NX, NSW, NSA, Cout]      (same flattening as alpha_cim [1,NX,NSW,NSA,1,Cout])
 *   binary_mask                     int8 [NSW, NSA]           (_quan_base.py:207-214, int8 wrap-around included)
 *   ADC state                       uint32 [NX, Cout, NWORDS, M], M = B*L; for weight slice k, activation slice j,
 *                                   sq = j*NSW + k:
 *                                     binary/ternary: bit sq = code +1, bit pairs+sq = code -1, bit 2*pairs+sq = clipped
 *                                     multi-bit     : bit sq = clipped (STE mask off)
 *   ADC state v2 (CIMQ_FLAG_V2)     uint8, channel fastest, 1-bit slices with NSW = NSA <= 3:
 *                                     plane D [NX][M][Cout]      sum_k 4^k * #{j: (i,k,j) not clipped}
 *                                     plane W [NX][M][Cout]      sum_j 4^j * #{k: (i,k,j) not clipped}
 *                                     plane C [NX][NSA][M][Cout] sum_k 4^k * (code(i,k,j) + 1)   (not for multi-bit)
 *                                   i.e. exactly the per-slice sums of the STE mask that dgrad / wgrad need
 *                                   (lsq.py:298-313, 362-376) and the ternary codes grad_alpha needs (lsq.py:321-334).
 */
#ifndef CIMQ_H_
#define CIMQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CIMQ_VERSION 200

/* ADC modes: lsq.py:197-230 */
enum {
  CIMQ_ADC_MULTIBIT = 0, /* adcbits > 1.5: clamp(psum, qn, qp), no scale factor (lsq.py:226-230) */
  CIMQ_ADC_BINARY = 1,   /* adcbits == 1  : sign(psum) * alpha (lsq.py:200-202) */
  CIMQ_ADC_TERNARY = 2   /* adcbits == 1.5: clamp(round(psum/alpha), -1, 1) * alpha (lsq.py:222-225) */
};

/* flags for cimq_conv_forward / backward */
enum {
  CIMQ_FLAG_FORCE_SIMT = 1,   /* use the CUDA-core kernels even where a tcgen05 kernel exists (tests) */
  CIMQ_FLAG_DETERMINISTIC = 2, /* backward: fold grad_x with a separate fixed-order pass instead of fp32
                                  reductions in the dgrad epilogue (bit-reproducible run to run, slower) */
  CIMQ_FLAG_V2 = 4             /* second-generation kernels and ADC-state format ("v2", below).  Only where
                                  cimq_info_t.tc_v2 is set; the table must come from cimq_adc_table2 with
                                  alpha_scale (binary / ternary ADC) and the flag must be passed to BOTH
                                  cimq_conv_forward and cimq_conv_backward of a step */
};

/* Geometry + quantisation of one Conv2dLSQCiM layer (_quan_base.py:174-237, lsq.py:512-531).
 * Square images/kernels, dilation 1, groups 1 -- the envelope the reference supports. */
typedef struct cimq_layer {
  int32_t batch, in_channels, in_hw;
  int32_t out_channels, kernel, stride, padding;
  int32_t nbits_a, abitslice, nbits_w, wbitslice;
  int32_t xbar;            /* crossbar rows (arr) */
  int32_t adc_mode;        /* CIMQ_ADC_* */
  int32_t adc_qn, adc_qp;  /* clamp range for CIMQ_ADC_MULTIBIT (lsq.py:125-126) */
} cimq_layer_t;

/* Sizes derived from a layer; filled by cimq_layer_info. */
typedef struct cimq_info {
  int32_t out_hw, L, M, F, NX, NSW, NSA, pairs;
  int32_t state_words;          /* uint32 words of ADC state per (crossbar, channel, pixel) */
  int32_t tc_forward;           /* 1 if the tcgen05 forward kernel covers this layer */
  int32_t tc_backward;          /* 1 if the tcgen05 dgrad / wgrad kernels cover this layer */
  int32_t tc_v2;                /* 1 if the v2 kernels (CIMQ_FLAG_V2) cover this layer */
  int64_t state_bytes;          /* NX*Cout*state_words*M*4 */
  int64_t table_bytes;          /* ADC table: NX*pairs*Cout entries {tp, tg, amp, 0} + a tiled copy for tcgen05 */
  int64_t wdigits_bytes;        /* fp32 weight digit planes [NSW, Cout, F] */
  int64_t wtiles_bytes;         /* int8 weight digit tiles in tcgen05 shared-memory order (0 if !tc_forward) */
  int64_t bwd_workspace_bytes;  /* scratch for cimq_conv_backward */
  int64_t psum_count;           /* B*NX*NSW*NSA*L*Cout */
  int64_t state_v2_bytes;       /* ADC state in the v2 format (0 if !tc_v2) */
} cimq_info_t;

int cimq_version(void);
const char *cimq_last_error(void);
int cimq_layer_info(const cimq_layer_t *layer, cimq_info_t *info_host);

/* ---- LSQ fake-quantiser (lsq.py:23-32, 547-555) ------------------------------------------ */

/* s[0] = grad_scale(alpha_act, ga), s[1] = grad_scale(alpha_weight, gw): value (a - a*g) + a*g,
 * one IEEE fp32 op at a time (lsq.py:23-26, 548, 554). */
int cimq_step_sizes(const float *alpha_act, const float *alpha_weight, float ga, float gw, float *s_out,
                    void *stream);

/* codes[i] = rint(clamp(x[i] / *s, qn, qp)) as one byte (uint8 when qn >= 0, else int8).
 * Replaces the div/clamp/round chain of lsq.py:549 and lsq.py:555; the fake-quant float
 * x_q = code * s is never materialised. */
int cimq_lsq_quantize(const float *x, int64_t n, const float *s, int32_t qn, int32_t qp, void *codes,
                      void *stream);

/* codes[i] = rint(xq[i] / *s) clamped to [qn,qp]: recovers the integer codes from fake-quant
 * floats for the 17-argument Function API (lsq.py:97-98).  Exact-recovery semantics. */
int cimq_codes_from_fakequant(const float *xq, int64_t n, const float *s, int32_t qn, int32_t qp, void *codes,
                              void *stream);

/* y[i] = rint(clamp(x[i] / *s, qn, qp)) * (rescale ? *s : 1) in fp32: the float outputs of the plain
 * LSQ modules ActLSQ (codes, lsq.py:654), LinearLSQ (lsq.py:609) and Conv2dLSQ (lsq.py:427). */
int cimq_lsq_fakequant(const float *x, int64_t n, const float *s, int32_t qn, int32_t qp, int32_t rescale, float *y,
                       void *stream);

/* Backward of x_q = round_pass(clamp(x/s)) * s with s = grad_scale(alpha, g) (autograd of
 * lsq.py:547-555): grad_x = grad_xq * 1[qn <= x/s <= qp];
 * *grad_alpha = g * sum(grad_xq * (q - (x/s)*1[...])).  workspace: cimq_lsq_backward_workspace_bytes. */
int64_t cimq_lsq_backward_workspace_bytes(int64_t n);
int cimq_lsq_backward(const float *grad_xq, const float *x, int64_t n, const float *s, int32_t qn, int32_t qp,
                      float g, float *grad_x, float *grad_alpha, void *workspace, void *stream);

/* nbits_alpha range quantiser of alpha_cim (lsq.py:566-571):
 *   scale = (max(alpha) - min(alpha)) / (qp - qn);  alpha_q = clamp(rint(alpha / scale), qn, qp) * scale
 * aux (5 floats: scale, max, min, #max, #min) is kept for the backward, which reproduces torch autograd:
 * straight-through round, inclusive clamp mask, and the scale's gradient flowing into the max / min elements
 * (shared evenly between ties). */
int cimq_alpha_quantize(const float *alpha, int64_t n, int32_t qn, int32_t qp, float *alpha_q, float *aux,
                        void *stream);
int cimq_alpha_quantize_backward(const float *alpha, const float *grad_alpha_q, int64_t n, int32_t qn, int32_t qp,
                                 const float *aux, float *grad_alpha, void *stream);

/* ---- per-step tables ----------------------------------------------------------------------- */

/* ADC decision thresholds and amplitudes for every (crossbar, w-slice, a-slice, channel).
 * The ADC code of an integer partial sum p is a monotone function of p
 *   v(p) = fl(fl(fl(fp16(p) * s_w) * s_a) / alpha_q)            (lsq.py:169, 195, 223)
 * so code and STE clip mask are integer compares against
 *   tp = min{p >= 1 : rint(v(p)) >= 1}        (ternary code, lsq.py:224)
 *   tg = min{p >= 1 : v(p) >= 1 + 1e-5}       (clip mask, lsq.py:310-311)
 * and amp = alpha_q * binary_mask (lsq.py:225, 233).  alpha_q may be NULL for CIMQ_ADC_MULTIBIT
 * (amp = binary_mask).  s = {s_a, s_w} from cimq_step_sizes.  status (int32, optional) gets bit 0
 * set if a step size or alpha_q is not finite and positive. */
int cimq_adc_table(const cimq_layer_t *layer, const float *s, const float *alpha_q, const int8_t *binary_mask,
                   void *table, int32_t *status, void *stream);

/* The same, plus the constants of the v2 kernels.  alpha_scale (device, 1 float; aux[0] of cimq_alpha_quantize) is the
 * step of the nbits_alpha quantiser of lsq.py:566-571: alpha_q = n * alpha_scale with integer n in [1, 2^nbits_alpha),
 * which lets the tensor core do the shift-and-add of lsq.py:233 on exact integers
 * (out = alpha_scale * sum code * n * binary_mask).  status bit 1 is set if some alpha_q is not such a multiple
 * (n < 2048 is required).  alpha_scale may be NULL for CIMQ_ADC_MULTIBIT. */
int cimq_adc_table2(const cimq_layer_t *layer, const float *s, const float *alpha_q, const float *alpha_scale,
                    const int8_t *binary_mask, void *table, int32_t *status, void *stream);

/* Everything of a training / inference step that depends only on the layer's PARAMETERS, in ONE launch (layers with
 * cimq_info_t.tc_v2 only): s = {s_a, s_w} (cimq_step_sizes), the weight codes (cimq_lsq_quantize of the fp32
 * weights [Cout, F]), alpha_q / aux (cimq_alpha_quantize of alpha_cim with range [aq_qn, aq_qp]; skipped for
 * CIMQ_ADC_MULTIBIT, pass NULL), the ADC table including the v2 constants (cimq_adc_table2) and the v2 sections of
 * the weight tiles (cimq_weight_prepare).  Bit-identical to calling those entry points one after the other; ten
 * launches fewer per layer per step.  status as for cimq_adc_table2. */
int cimq_layer_prepare(const cimq_layer_t *layer, const float *weight, const float *alpha_act,
                       const float *alpha_weight, float ga, float gw, const float *alpha_cim, int32_t aq_qn,
                       int32_t aq_qp, const int8_t *binary_mask, float *s_out, int8_t *wcodes, float *alpha_q,
                       float *aux, void *table, void *wtiles, int32_t *status, void *stream);

/* Sign-magnitude digit planes of the weight codes (slicing_weights_signed, lsq.py:438-464):
 * wdigits fp32 [NSW, Cout, F] (CUDA-core backward; may be NULL); wtiles (may be NULL): operand tiles in
 * tcgen05 shared-memory order -- int8 digit tiles + im2col LUT for the forward, bf16 digit tiles for dgrad. */
int cimq_weight_prepare(const cimq_layer_t *layer, const int8_t *wcodes, float *wdigits, void *wtiles,
                        void *stream);

/* ---- the CiM convolution --------------------------------------------------------------------- */

/* get_cim_output_signed.forward (lsq.py:92-237): im2col, bit slicing, one integer contraction per
 * (crossbar, w-slice, a-slice), ADC quantisation of every partial sum, shift-and-add.
 * out [B,Cout,L] fp32; state (optional, needed for backward) records code and clip bit of every
 * partial sum. */
int cimq_conv_forward(const cimq_layer_t *layer, const uint8_t *xcodes, const int8_t *wcodes, const void *wtiles,
                      const void *table, const float *s, const int8_t *binary_mask, float *out, void *state,
                      uint32_t flags, void *stream);

/* The same with the stochastic near-ADC-less read-out of lsq.py:205-220 (adcbits 1.5 only): every partial sum is
 * read through two sigmoids of sharpness 0.01 around +-alpha_q/2, 50 Bernoulli draws each, code =
 * clamp(round(n1/50 + n2/50 - 1), -1, 1).  alpha_q [NX,NSW,NSA,Cout].  The draws come from a counter-based
 * generator keyed by (seed, index of the partial sum): reproducible for a seed, not bit-comparable with the
 * reference's torch RNG (statistical parity).  `state` (for the backward) is the deterministic one, as in the
 * reference, whose backward does not see the sampled code.  CUDA-core kernel. */
int cimq_conv_forward_stochastic(const cimq_layer_t *layer, const uint8_t *xcodes, const int8_t *wcodes,
                                 const void *table, const float *s, const float *alpha_q, float *out,
                                 uint32_t *state, uint64_t seed, void *stream);

/* get_cim_output_signed.backward (lsq.py:244-386): grad_xq [B,Cin,H,W], grad_wq [Cout,F],
 * grad_alpha_q [NX,NSW,NSA,Cout] (NULL for CIMQ_ADC_MULTIBIT).  wdigits is needed by the CUDA-core kernels,
 * wtiles by the tcgen05 kernels; pass both to let the library choose. */
int cimq_conv_backward(const cimq_layer_t *layer, const float *grad_out, const uint8_t *xcodes,
                       const float *wdigits, const void *wtiles, const void *state, const float *s,
                       const int8_t *binary_mask, float *grad_xq, float *grad_wq, float *grad_alpha_q,
                       void *workspace, uint32_t flags, void *stream);

/* Raw integer partial sums int32 [B,NX,NSW,NSA,L,Cout] (= ctx.ps_int, lsq.py:192).  Test / debug. */
int cimq_conv_psums(const cimq_layer_t *layer, const uint8_t *xcodes, const int8_t *wcodes, int32_t *psums,
                    void *stream);

/* sum over (B,L) of |psum| as uint64 [NX,NSW,NSA,Cout] (must be zeroed by the caller): the exact
 * integer statistic behind the alpha_cim initialisation (lsq.py:35-87, 557-563). */
int cimq_conv_psum_abs_sums(const cimq_layer_t *layer, const uint8_t *xcodes, const int8_t *wcodes,
                            unsigned long long *sums, void *stream);

/* ---- batch norm (+ residual) (+ ReLU): the element-wise step either side of the CiM convolution ------
 * The reference model is conv -> nn.BatchNorm2d -> [+ shortcut] -> F.relu (models/cifar10/resnet.py:60-66,
 * 110-112); SURVEY 8 f-2.  fp32 NCHW, x/y/grad [B,C,HW]; torch.nn.BatchNorm2d semantics: batch statistics with
 * biased variance in training (running buffers updated with `momentum` and the unbiased variance), running
 * statistics otherwise.  workspace: cimq_bn_workspace_bytes(B, C) bytes. */
int64_t cimq_bn_workspace_bytes(int32_t batch, int32_t channels);

/* y = BN(x) [+ residual] [ReLU].  residual / weight / bias may be NULL; running_mean / running_var may be NULL in
 * training (no running statistics); save_mean / save_invstd [C] receive the batch statistics in training. */
int cimq_bn_forward(const float *x, const float *residual, const float *weight, const float *bias,
                    float *running_mean, float *running_var, int32_t training, float momentum, float eps,
                    int32_t relu, int32_t batch, int32_t channels, int32_t hw, float *y, float *save_mean,
                    float *save_invstd, void *workspace, void *stream);

/* The same, and -- SURVEY 8 f-2 -- the NEXT layer's activation quantiser (lsq.py:547-549) fused into the epilogue:
 * next_codes [B,C,HW] uint8 = rint(clamp(y / s, 0, next_qp)) with s = grad_scale(next_alpha_act[0], next_grad_scale),
 * byte-identical to cimq_lsq_quantize(y, cimq_step_sizes(...)[0], 0, next_qp); y is still written (the residual path and
 * the quantiser's backward need it).  Replaces resnet.py:83-86 + the first line of the next Conv2dLSQCiM.forward. */
int cimq_bn_forward_quant(const float *x, const float *residual, const float *weight, const float *bias,
                          float *running_mean, float *running_var, int32_t training, float momentum, float eps,
                          int32_t relu, int32_t batch, int32_t channels, int32_t hw, float *y, float *save_mean,
                          float *save_invstd, void *workspace, const float *next_alpha_act, float next_grad_scale,
                          int32_t next_qp, uint8_t *next_codes, void *stream);

/* Gradients of the above: grad_x [B,C,HW], grad_residual (NULL if there was none; = grad_y where the ReLU passed),
 * grad_weight / grad_bias [C] (may be NULL).  y is the forward output (ReLU mask; may be NULL if relu == 0);
 * mean / invstd are save_mean / save_invstd of the forward (training) or the running statistics (inference). */
int cimq_bn_backward(const float *grad_y, const float *x, const float *y, const float *weight, const float *mean,
                     const float *invstd, int32_t training, int32_t relu, int32_t batch, int32_t channels, int32_t hw,
                     float *grad_x, float *grad_residual, float *grad_weight, float *grad_bias, void *workspace,
                     void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CIMQ_H_ */
