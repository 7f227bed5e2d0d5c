// The LSQ quantiser's integer code, shared by the quantiser kernels (lsq_quant.cu) and the fused batch-norm epilogue
// that emits the next layer's activation codes (bn_fused.cu): both must produce the same bytes.
#pragma once

#include "cimq_common.cuh"

namespace cimq {

// The reference computes `(x / s).clamp(qn, qp).round()` (lsq.py:549) with an IEEE fp32 division.  A
// correctly-rounded division per element (MUFU.RCP + Newton + range check) makes the kernel
// instruction-bound, so the quotient is first approximated with the step size's reciprocal hoisted out of
// the loop (x*r refined by one FMA residual step: error << 1e-3 for |x/s| < 65536) and the exact division is
// only evaluated when the approximation is within 1e-3 of a point where the result could change (a
// rounding tie k+0.5, or -- for the backward's inclusive clamp mask -- a clamp bound).  Results are
// bit-identical to the IEEE chain.
struct StepSize {
  float s, r;
  bool ok;  // s is a positive normal number whose reciprocal is finite: the fast path is valid
};
__device__ __forceinline__ StepSize load_step(const float *sp) {
  StepSize st;
  st.s = __ldg(sp);
  st.r = __frcp_rn(st.s);
  st.ok = st.s >= 1e-30f && st.s <= 1e30f;
  return st;
}
__device__ __forceinline__ float approx_quotient(float x, const StepSize &st) {
  const float q0 = x * st.r;
  const float u = fmaf(fmaf(-q0, st.s, x), st.r, q0);
  return fabsf(q0) < 65536.0f ? u : q0;  // far outside the code range only the sign matters
}
// rint (half to even) of a value in (-2^22, 2^22) without the conversion unit
__device__ __forceinline__ float rint_magic(float u) { return __fadd_rn(__fadd_rn(u, 12582912.0f), -12582912.0f); }

// integer code rint(clamp(x / s, qn, qp)), as the low byte of the result
__device__ __forceinline__ uint32_t lsq_code(float x, const StepSize &st, float qn, float qp) {
  float uc = fminf(fmaxf(approx_quotient(x, st), qn), qp);
  float biased = __fadd_rn(uc, 12582912.0f);  // low mantissa bits = rint(uc) in two's complement
  const float t = uc - __fadd_rn(biased, -12582912.0f);
  if (fabsf(fabsf(t) - 0.5f) < 1e-3f || !(fabsf(x) <= 3.0e38f) || !st.ok) {  // near a tie / non-finite: exact
    uc = fminf(fmaxf(__fdiv_rn(x, st.s), qn), qp);
    biased = __fadd_rn(uc, 12582912.0f);
  }
  return __float_as_uint(biased) & 0xffu;
}
__device__ __forceinline__ float lsq_code_f(float x, const StepSize &st, float qn, float qp) {
  float uc = fminf(fmaxf(approx_quotient(x, st), qn), qp);
  float c = rint_magic(uc);
  if (fabsf(fabsf(uc - c) - 0.5f) < 1e-3f || !(fabsf(x) <= 3.0e38f) || !st.ok)
    c = rint_magic(fminf(fmaxf(__fdiv_rn(x, st.s), qn), qp));
  return c;
}


// forward value of grad_scale(alpha, g): y - y_grad + y_grad with y_grad = alpha * g, each op rounded (lsq.py:23-26)
__device__ __forceinline__ float grad_scale_value(float a, float g) {
  const float ag = __fmul_rn(a, g);
  return __fadd_rn(__fsub_rn(a, ag), ag);
}

}  // namespace cimq
