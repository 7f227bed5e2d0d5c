// LSQ fake-quantiser kernels (HBM-bound): fp32 -> one-byte integer codes, and the STE backward
// with the learned-step-size gradient.  Reference: models/_modules/lsq.py:23-32, 547-555.
//
// Roofline: forward moves 5 B/element (4 read + 1 written), backward 12 B/element
// (grad + x read, grad written).  Access is 128-bit vectorised and fully coalesced; each thread
// keeps several independent 16-byte loads in flight.
#include "cimq_common.cuh"
#include "lsq_code.cuh"

namespace cimq {

namespace {

constexpr int kThreads = 256;
constexpr int kSMs = 148;

__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__global__ void step_sizes_kernel(const float *__restrict__ aa, const float *__restrict__ aw, float ga, float gw,
                                  float *__restrict__ s) {
  if (threadIdx.x == 0) {
    // grad_scale value: y - y_grad + y_grad with y_grad = a * g, each op rounded (lsq.py:23-26)
    s[0] = grad_scale_value(aa[0], ga);
    s[1] = grad_scale_value(aw[0], gw);
  }
}

// 16 elements per thread per iteration: four 16-byte loads, one 16-byte store of codes.
__global__ void __launch_bounds__(kThreads) lsq_quantize_vec_kernel(const float4 *__restrict__ x, int64_t n16,
                                                                    const float *__restrict__ sp, float qn,
                                                                    float qp, uint4 *__restrict__ codes) {
  const StepSize s = load_step(sp);
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n16; i += (int64_t)gridDim.x * kThreads) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ldg_stream(x + 4 * i + k);
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t c0 = lsq_code(v[k].x, s, qn, qp);
      uint32_t c1 = lsq_code(v[k].y, s, qn, qp);
      uint32_t c2 = lsq_code(v[k].z, s, qn, qp);
      uint32_t c3 = lsq_code(v[k].w, s, qn, qp);
      w[k] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
    }
    codes[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void lsq_quantize_scalar_kernel(const float *__restrict__ x, int64_t begin, int64_t n,
                                           const float *__restrict__ sp, float qn, float qp,
                                           uint8_t *__restrict__ codes) {
  const StepSize s = load_step(sp);
  for (int64_t i = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    codes[i] = (uint8_t)lsq_code(x[i], s, qn, qp);
}

// Float outputs for the plain LSQ modules (ActLSQ returns codes, LinearLSQ / Conv2dLSQ need
// round_pass(clamp(x/s)) optionally times s; lsq.py:427, 609, 654).  8 B/element.
__global__ void __launch_bounds__(kThreads) lsq_fakequant_kernel(const float *__restrict__ x, int64_t n,
                                                                 const float *__restrict__ sp, float qn, float qp,
                                                                 int rescale, float *__restrict__ y, int vec_ok) {
  const StepSize s = load_step(sp);
  const float m = rescale ? s.s : 1.0f;
  const int64_t n4 = vec_ok ? n / 4 : 0;
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  float4 *y4 = reinterpret_cast<float4 *>(y);
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    float4 v = ldg_stream(x4 + i), o;
    o.x = __fmul_rn(lsq_code_f(v.x, s, qn, qp), m);
    o.y = __fmul_rn(lsq_code_f(v.y, s, qn, qp), m);
    o.z = __fmul_rn(lsq_code_f(v.z, s, qn, qp), m);
    o.w = __fmul_rn(lsq_code_f(v.w, s, qn, qp), m);
    y4[i] = o;
  }
  for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kThreads)
    y[i] = __fmul_rn(lsq_code_f(x[i], s, qn, qp), m);
}

__device__ __forceinline__ void lsq_bwd_elem(float g, float x, const StepSize &st, float qn, float qp, float &gx,
                                             float &acc) {
  float u = approx_quotient(x, st);
  {
    const float uc = fminf(fmaxf(u, qn), qp);
    const float t = uc - rint_magic(uc);
    // exact division near a rounding tie, near a clamp bound (inclusive mask; x == 0 is exact) or for odd inputs
    const bool near_bound = (fabsf(u - qn) < 1e-3f || fabsf(u - qp) < 1e-3f) && x != 0.0f;
    if (fabsf(fabsf(t) - 0.5f) < 1e-3f || near_bound || !(fabsf(x) <= 3.0e38f) || !st.ok) u = __fdiv_rn(x, st.s);
  }
  bool inside = (u >= qn) && (u <= qp);  // torch.clamp backward mask is inclusive
  float q = rint_magic(fminf(fmaxf(u, qn), qp));
  gx = inside ? g : 0.0f;
  acc = fmaf(g, q - (inside ? u : 0.0f), acc);
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sm[kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  v = (threadIdx.x < kThreads / 32) ? sm[threadIdx.x] : 0.0;
  if (threadIdx.x < 32)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // valid in thread 0
}

// 8 elements per thread per iteration.  Writes one fp64 partial of sum(g * (q - u*mask)) per block.
__global__ void __launch_bounds__(kThreads) lsq_backward_kernel(const float *__restrict__ gq,
                                                                const float *__restrict__ x, int64_t n,
                                                                const float *__restrict__ sp, float qn, float qp,
                                                                float *__restrict__ gx,
                                                                double *__restrict__ partials, int vec_ok) {
  const StepSize s = load_step(sp);
  float acc = 0.0f;
  const int64_t n8 = vec_ok ? n / 8 : 0;
  const float4 *g4 = reinterpret_cast<const float4 *>(gq);
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  float4 *o4 = reinterpret_cast<float4 *>(gx);
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n8; i += (int64_t)gridDim.x * kThreads) {
    float4 ga = ldg_stream(g4 + 2 * i), gb = ldg_stream(g4 + 2 * i + 1);
    float4 xa = ldg_stream(x4 + 2 * i), xb = ldg_stream(x4 + 2 * i + 1);
    float4 oa, ob;
    lsq_bwd_elem(ga.x, xa.x, s, qn, qp, oa.x, acc);
    lsq_bwd_elem(ga.y, xa.y, s, qn, qp, oa.y, acc);
    lsq_bwd_elem(ga.z, xa.z, s, qn, qp, oa.z, acc);
    lsq_bwd_elem(ga.w, xa.w, s, qn, qp, oa.w, acc);
    lsq_bwd_elem(gb.x, xb.x, s, qn, qp, ob.x, acc);
    lsq_bwd_elem(gb.y, xb.y, s, qn, qp, ob.y, acc);
    lsq_bwd_elem(gb.z, xb.z, s, qn, qp, ob.z, acc);
    lsq_bwd_elem(gb.w, xb.w, s, qn, qp, ob.w, acc);
    o4[2 * i] = oa;
    o4[2 * i + 1] = ob;
  }
  for (int64_t i = n8 * 8 + blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kThreads) {
    float o;
    lsq_bwd_elem(gq[i], x[i], s, qn, qp, o, acc);
    gx[i] = o;
  }
  double tot = block_sum((double)acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
  // (Measured: letting the block that finishes last add the partials -- a ticket counter instead of the one-block
  // kernel below -- saved nothing inside a captured step and cost 4 us here: 1184 same-address atomics at the tail.)
}

__global__ void __launch_bounds__(kThreads) lsq_backward_finish_kernel(const double *__restrict__ partials,
                                                                       int nparts, float g,
                                                                       float *__restrict__ galpha) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nparts; i += kThreads) v += partials[i];
  v = block_sum(v);
  if (threadIdx.x == 0) galpha[0] = (float)((double)g * v);
}

inline int grid_for(int64_t work_items) {
  int64_t blocks = (work_items + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)kSMs * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

constexpr int kMaxPartials = kSMs * 8;

}  // namespace

int launch_step_sizes(const float *aa, const float *aw, float ga, float gw, float *s, cudaStream_t st) {
  step_sizes_kernel<<<1, 32, 0, st>>>(aa, aw, ga, gw, s);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_lsq_quantize(const float *x, int64_t n, const float *s, int qn, int qp, void *codes, bool,
                        cudaStream_t st) {
  if (n == 0) return 0;
  bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(codes)) & 15u) == 0;
  int64_t n16 = aligned ? n / 16 : 0;
  if (n16 > 0) {
    lsq_quantize_vec_kernel<<<grid_for(n16), kThreads, 0, st>>>(reinterpret_cast<const float4 *>(x), n16, s,
                                                               (float)qn, (float)qp,
                                                               reinterpret_cast<uint4 *>(codes));
    CIMQ_CUDA_OK(cudaGetLastError());
  }
  if (n16 * 16 < n) {
    lsq_quantize_scalar_kernel<<<grid_for(n - n16 * 16), kThreads, 0, st>>>(
        x, n16 * 16, n, s, (float)qn, (float)qp, reinterpret_cast<uint8_t *>(codes));
    CIMQ_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

int launch_lsq_fakequant(const float *x, int64_t n, const float *s, int qn, int qp, int rescale, float *y,
                         cudaStream_t st) {
  if (n == 0) return 0;
  int vec_ok = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
  lsq_fakequant_kernel<<<grid_for((n + 3) / 4), kThreads, 0, st>>>(x, n, s, (float)qn, (float)qp, rescale, y,
                                                                  vec_ok);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int64_t lsq_backward_ws_bytes(int64_t) { return (int64_t)kMaxPartials * sizeof(double); }

int launch_lsq_backward(const float *gq, const float *x, int64_t n, const float *s, int qn, int qp, float g,
                        float *gx, float *galpha, void *ws, cudaStream_t st) {
  CIMQ_REQUIRE(ws != nullptr, "lsq_backward: workspace is NULL");
  int vec_ok = ((reinterpret_cast<uintptr_t>(gq) | reinterpret_cast<uintptr_t>(x) |
                 reinterpret_cast<uintptr_t>(gx)) & 15u) == 0;
  int grid = grid_for((n + 7) / 8);
  double *partials = reinterpret_cast<double *>(ws);
  lsq_backward_kernel<<<grid, kThreads, 0, st>>>(gq, x, n, s, (float)qn, (float)qp, gx, partials, vec_ok);
  CIMQ_CUDA_OK(cudaGetLastError());
  lsq_backward_finish_kernel<<<1, kThreads, 0, st>>>(partials, grid, g, galpha);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cimq
