// Batch norm (+ residual add) (+ ReLU), training and inference, forward and backward, fp32 NCHW -- the
// element-wise step either side of the CiM convolution (the reference model is
// models/cifar10/resnet.py:60-66, 110-112: conv -> nn.BatchNorm2d -> [+ shortcut] -> F.relu).
//
// HBM-bound.  A channel's data is B runs of HW contiguous floats; the work of a channel is split over `splits`
// blocks (grid = (C, splits)) so that narrow layers (16 channels) still fill 148 SMs -- cuDNN's spatial
// batch-norm kernels use one block per channel and take 134 / 267 us (forward / backward) on [256,16,32,32],
// 16.8 MB.  Two launches per direction: per-block partial sums (fp32 in the block, double across blocks, fixed
// order: deterministic), then the element-wise pass, whose blocks each re-reduce the few partials of their channel.
#include "cimq_common.cuh"
#include "lsq_code.cuh"

namespace cimq {

namespace {

constexpr int kBnThreads = 256;

__device__ __forceinline__ float block_sum(float v, float *red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < kBnThreads / 32 ? red[threadIdx.x] : 0.0f;
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

// The float4 items of block (c, s) -- images b = s, s + splits, ... of channel c, HW / 4 items each -- as ONE flat range:
// with one image per loop iteration a block of 256 threads had a single load in flight per thread (16x16 and 8x8 layers
// used 64 / 16 of its threads) and the kernels were latency bound at 1.3 TB/s.  kBnUnroll items per thread are loaded
// before the first one is used.
constexpr int kBnUnroll = 4;
struct BnItems {
  int total, hw4, shift, s, splits, C, c;
  __device__ __forceinline__ BnItems(int B, int C_, int HW, int c_, int s_, int splits_) {
    hw4 = HW >> 2; s = s_; splits = splits_; C = C_; c = c_;
    shift = (hw4 & (hw4 - 1)) == 0 ? 31 - __clz(hw4) : -1;
    const int nb = s_ < B ? (B - s_ + splits_ - 1) / splits_ : 0;
    total = nb * hw4;
  }
  // float4 index (into the [B, C, HW/4] tensor) of flat item t
  __device__ __forceinline__ int64_t at(int t) const {
    const int bi = shift >= 0 ? t >> shift : t / hw4, i = t - bi * hw4;
    return ((int64_t)(s + bi * splits) * C + c) * hw4 + i;
  }
};

// forward pass 1: partial[c][s] = {sum x, sum x^2} over the images b = s, s+splits, ...
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float *__restrict__ x, int B, int C, int HW,
                                                              double2 *__restrict__ partial) {
  __shared__ float red[kBnThreads / 32];
  const int c = blockIdx.x, s = blockIdx.y, splits = gridDim.y;
  float s1 = 0.0f, s2 = 0.0f;
  if (VEC) {
    const BnItems it(B, C, HW, c, s, splits);
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for (int t0 = threadIdx.x; t0 < it.total; t0 += kBnUnroll * kBnThreads) {
      float4 v[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int t = t0 + u * kBnThreads;
        v[u] = t < it.total ? __ldg(x4 + it.at(t)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        s1 += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        s2 += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
      }
    }
  }
  for (int b = s; b < B && !VEC; b += splits) {
    const float *p = x + ((int64_t)b * C + c) * HW;
    {
      for (int i = threadIdx.x; i < HW; i += kBnThreads) {
        const float v = __ldg(p + i);
        s1 += v;
        s2 += v * v;
      }
    }
  }
  const float t1 = block_sum(s1, red), t2 = block_sum(s2, red);
  if (threadIdx.x == 0) partial[(int64_t)c * splits + s] = make_double2((double)t1, (double)t2);
}

// forward pass 2: statistics of the channel from the partials (training) or the running buffers (inference),
// y = (x - mean) * invstd * gamma + beta [+ residual] [ReLU]; block (c, 0) publishes mean / invstd and updates the
// running statistics (momentum, unbiased variance -- torch.nn.BatchNorm2d semantics).
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(
    const float *__restrict__ x, const float *__restrict__ residual, const float *__restrict__ weight,
    const float *__restrict__ bias, float *__restrict__ running_mean, float *__restrict__ running_var,
    const double2 *__restrict__ partial, int training, float momentum, float eps, int relu, int B, int C, int HW,
    float *__restrict__ y, float *__restrict__ save_mean, float *__restrict__ save_invstd,
    const float *__restrict__ q_alpha, float q_g, float q_qp, uint8_t *__restrict__ codes) {
  // codes != NULL: also emit the NEXT layer's activation codes rint(clamp(y / s, 0, qp)), s = grad_scale(q_alpha, q_g)
  // (lsq.py:547-549) -- the same bytes cimq_lsq_quantize would compute from y, without reading y back.
  StepSize qs;
  qs.s = 1.0f; qs.r = 1.0f; qs.ok = true;
  if (codes != nullptr) {
    qs.s = grad_scale_value(__ldg(q_alpha), q_g);
    qs.r = __frcp_rn(qs.s);
    qs.ok = qs.s >= 1e-30f && qs.s <= 1e30f;
  }
  __shared__ float sh[2];
  const int c = blockIdx.x, s = blockIdx.y, splits = gridDim.y;
  if (threadIdx.x == 0) {
    float mean, invstd;
    if (training) {
      double a = 0.0, q = 0.0;
      for (int t = 0; t < splits; ++t) {
        const double2 v = partial[(int64_t)c * splits + t];
        a += v.x;
        q += v.y;
      }
      const double n = (double)B * HW;
      const double m = a / n;
      double var = q / n - m * m;
      var = var < 0.0 ? 0.0 : var;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      if (s == 0) {
        save_mean[c] = mean;
        save_invstd[c] = invstd;
        if (running_mean != nullptr) {
          running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
          running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)(var * n / (n > 1.0 ? n - 1.0 : 1.0));
        }
      }
    } else {
      mean = running_mean[c];
      invstd = rsqrtf(running_var[c] + eps);
    }
    sh[0] = mean;
    sh[1] = invstd;
  }
  __syncthreads();
  const float mean = sh[0];
  const float g = (weight != nullptr ? weight[c] : 1.0f) * sh[1], be = bias != nullptr ? bias[c] : 0.0f;
  if (VEC) {
    const BnItems it(B, C, HW, c, s, splits);
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const float4 *r4 = reinterpret_cast<const float4 *>(residual);
    float4 *y4 = reinterpret_cast<float4 *>(y);
    uint32_t *c4 = reinterpret_cast<uint32_t *>(codes);
    for (int t0 = threadIdx.x; t0 < it.total; t0 += kBnUnroll * kBnThreads) {
      float4 v[kBnUnroll], r[kBnUnroll];
      int64_t idx[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int t = t0 + u * kBnThreads;
        idx[u] = t < it.total ? it.at(t) : -1;
        v[u] = idx[u] >= 0 ? __ldg(x4 + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
        r[u] = (idx[u] >= 0 && r4 != nullptr) ? __ldg(r4 + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        if (idx[u] < 0) continue;
        float4 o = make_float4(fmaf(v[u].x - mean, g, be), fmaf(v[u].y - mean, g, be), fmaf(v[u].z - mean, g, be),
                               fmaf(v[u].w - mean, g, be));
        if (r4 != nullptr) { o.x += r[u].x; o.y += r[u].y; o.z += r[u].z; o.w += r[u].w; }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        y4[idx[u]] = o;
        if (codes != nullptr)
          c4[idx[u]] = lsq_code(o.x, qs, 0.0f, q_qp) | (lsq_code(o.y, qs, 0.0f, q_qp) << 8) |
                       (lsq_code(o.z, qs, 0.0f, q_qp) << 16) | (lsq_code(o.w, qs, 0.0f, q_qp) << 24);
      }
    }
  }
  for (int b = s; b < B && !VEC; b += splits) {
    const int64_t off = ((int64_t)b * C + c) * HW;
    {
      for (int i = threadIdx.x; i < HW; i += kBnThreads) {
        float o = fmaf(__ldg(x + off + i) - mean, g, be);
        if (residual != nullptr) o += __ldg(residual + off + i);
        if (relu) o = fmaxf(o, 0.f);
        y[off + i] = o;
        if (codes != nullptr) codes[off + i] = (uint8_t)lsq_code(o, qs, 0.0f, q_qp);
      }
    }
  }
}

// backward pass 1: partial[c][s] = {sum dy', sum dy' * (x - mean)} with dy' = dy where the ReLU passed
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_stats_kernel(const float *__restrict__ gy,
                                                                  const float *__restrict__ x,
                                                                  const float *__restrict__ y,
                                                                  const float *__restrict__ save_mean, int relu, int B,
                                                                  int C, int HW, double2 *__restrict__ partial) {
  __shared__ float red[kBnThreads / 32];
  const int c = blockIdx.x, s = blockIdx.y, splits = gridDim.y;
  const float mean = save_mean[c];
  float s1 = 0.0f, s2 = 0.0f;
  if (VEC) {
    const BnItems it(B, C, HW, c, s, splits);
    const float4 *g4 = reinterpret_cast<const float4 *>(gy);
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const float4 *y4 = reinterpret_cast<const float4 *>(y);
    for (int t0 = threadIdx.x; t0 < it.total; t0 += kBnUnroll * kBnThreads) {
      float4 d[kBnUnroll], v[kBnUnroll], o[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int t = t0 + u * kBnThreads;
        const bool ok = t < it.total;
        const int64_t idx = ok ? it.at(t) : 0;
        d[u] = ok ? __ldg(g4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[u] = ok ? __ldg(x4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        o[u] = (ok && relu) ? __ldg(y4 + idx) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        float4 dd = d[u];
        dd.x = o[u].x > 0.f ? dd.x : 0.f; dd.y = o[u].y > 0.f ? dd.y : 0.f;
        dd.z = o[u].z > 0.f ? dd.z : 0.f; dd.w = o[u].w > 0.f ? dd.w : 0.f;
        s1 += (dd.x + dd.y) + (dd.z + dd.w);
        s2 += (dd.x * (v[u].x - mean) + dd.y * (v[u].y - mean)) + (dd.z * (v[u].z - mean) + dd.w * (v[u].w - mean));
      }
    }
  }
  for (int b = s; b < B && !VEC; b += splits) {
    const int64_t off = ((int64_t)b * C + c) * HW;
    {
      for (int i = threadIdx.x; i < HW; i += kBnThreads) {
        float d = __ldg(gy + off + i);
        if (relu && !(__ldg(y + off + i) > 0.f)) d = 0.f;
        s1 += d;
        s2 += d * (__ldg(x + off + i) - mean);
      }
    }
  }
  const float t1 = block_sum(s1, red), t2 = block_sum(s2, red);
  if (threadIdx.x == 0) partial[(int64_t)c * splits + s] = make_double2((double)t1, (double)t2);
}

// backward pass 2: dx = gamma * invstd * (dy' - mean(dy') - xhat * mean(dy' * xhat)), d_residual = dy';
// block (c, 0) writes grad_weight = sum dy' * xhat and grad_bias = sum dy'.  Inference mode (training == 0):
// the statistics are constants, dx = gamma * invstd * dy'.
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(
    const float *__restrict__ gy, const float *__restrict__ x, const float *__restrict__ y,
    const float *__restrict__ weight, const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
    const double2 *__restrict__ partial, int training, int relu, int B, int C, int HW, float *__restrict__ gx,
    float *__restrict__ gres, float *__restrict__ gweight, float *__restrict__ gbias) {
  __shared__ float sh[2];
  const int c = blockIdx.x, s = blockIdx.y, splits = gridDim.y;
  const float mean = save_mean[c], invstd = save_invstd[c];
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    for (int t = 0; t < splits; ++t) {
      const double2 v = partial[(int64_t)c * splits + t];
      a += v.x;
      q += v.y;
    }
    if (s == 0) {
      if (gbias != nullptr) gbias[c] = (float)a;
      if (gweight != nullptr) gweight[c] = (float)(q * (double)invstd);
    }
    const double n = (double)B * HW;
    sh[0] = training ? (float)(a / n) : 0.0f;                                        // mean of dy'
    sh[1] = training ? (float)(q / n * (double)invstd * (double)invstd) : 0.0f;      // mean(dy' * xhat) * invstd
  }
  __syncthreads();
  const float m1 = sh[0], m2 = sh[1];
  const float k = (weight != nullptr ? weight[c] : 1.0f) * invstd;
  if (VEC) {
    const BnItems it(B, C, HW, c, s, splits);
    const float4 *g4 = reinterpret_cast<const float4 *>(gy);
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const float4 *y4 = reinterpret_cast<const float4 *>(y);
    float4 *o4 = reinterpret_cast<float4 *>(gx);
    float4 *r4 = reinterpret_cast<float4 *>(gres);
    for (int t0 = threadIdx.x; t0 < it.total; t0 += kBnUnroll * kBnThreads) {
      float4 d[kBnUnroll], v[kBnUnroll], o[kBnUnroll];
      int64_t idx[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int t = t0 + u * kBnThreads;
        idx[u] = t < it.total ? it.at(t) : -1;
        const bool ok = idx[u] >= 0;
        d[u] = ok ? __ldg(g4 + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[u] = ok ? __ldg(x4 + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
        o[u] = (ok && relu) ? __ldg(y4 + idx[u]) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        if (idx[u] < 0) continue;
        float4 dd = d[u];
        dd.x = o[u].x > 0.f ? dd.x : 0.f; dd.y = o[u].y > 0.f ? dd.y : 0.f;
        dd.z = o[u].z > 0.f ? dd.z : 0.f; dd.w = o[u].w > 0.f ? dd.w : 0.f;
        if (r4 != nullptr) r4[idx[u]] = dd;
        o4[idx[u]] = make_float4(k * (dd.x - m1 - (v[u].x - mean) * m2), k * (dd.y - m1 - (v[u].y - mean) * m2),
                                 k * (dd.z - m1 - (v[u].z - mean) * m2), k * (dd.w - m1 - (v[u].w - mean) * m2));
      }
    }
  }
  for (int b = s; b < B && !VEC; b += splits) {
    const int64_t off = ((int64_t)b * C + c) * HW;
    {
      for (int i = threadIdx.x; i < HW; i += kBnThreads) {
        float d = __ldg(gy + off + i);
        if (relu && !(__ldg(y + off + i) > 0.f)) d = 0.f;
        if (gres != nullptr) gres[off + i] = d;
        gx[off + i] = k * (d - m1 - (__ldg(x + off + i) - mean) * m2);
      }
    }
  }
}

inline int bn_splits(int B, int C) {
  int s = (148 * 4 + C - 1) / C;  // about four blocks per SM in total
  s = s < 1 ? 1 : s;
  return s > B ? B : s;
}

inline bool aligned16(const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

int64_t bn_workspace_bytes(int B, int C) { return (int64_t)C * bn_splits(B, C) * (int64_t)sizeof(double2); }

int launch_bn_forward(const float *x, const float *residual, const float *weight, const float *bias,
                      float *running_mean, float *running_var, int training, float momentum, float eps, int relu,
                      int B, int C, int HW, float *y, float *save_mean, float *save_invstd, void *workspace,
                      cudaStream_t st, const float *q_alpha, float q_g, int q_qp, uint8_t *codes) {
  CIMQ_REQUIRE(codes == nullptr || (q_alpha != nullptr && q_qp > 0 && q_qp <= 255), "bn_forward: bad quantiser arguments");
  CIMQ_REQUIRE(x != nullptr && y != nullptr && B > 0 && C > 0 && HW > 0, "bn_forward: bad argument");
  CIMQ_REQUIRE(training ? (save_mean && save_invstd && workspace) : (running_mean && running_var),
               "bn_forward: missing statistics buffers");
  const dim3 grid(C, bn_splits(B, C));
  const bool vec = HW % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(residual) &&
                   (reinterpret_cast<uintptr_t>(codes) & 3u) == 0;
  double2 *part = reinterpret_cast<double2 *>(workspace);
  if (training) {
    if (vec) bn_stats_kernel<true><<<grid, kBnThreads, 0, st>>>(x, B, C, HW, part);
    else bn_stats_kernel<false><<<grid, kBnThreads, 0, st>>>(x, B, C, HW, part);
    CIMQ_CUDA_OK(cudaGetLastError());
  }
  if (vec)
    bn_apply_kernel<true><<<grid, kBnThreads, 0, st>>>(x, residual, weight, bias, running_mean, running_var, part,
                                                       training, momentum, eps, relu, B, C, HW, y, save_mean,
                                                       save_invstd, q_alpha, q_g, (float)q_qp, codes);
  else
    bn_apply_kernel<false><<<grid, kBnThreads, 0, st>>>(x, residual, weight, bias, running_mean, running_var, part,
                                                        training, momentum, eps, relu, B, C, HW, y, save_mean,
                                                        save_invstd, q_alpha, q_g, (float)q_qp, codes);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_bn_backward(const float *gy, const float *x, const float *y, const float *weight, const float *save_mean,
                       const float *save_invstd, int training, int relu, int B, int C, int HW, float *gx, float *gres,
                       float *gweight, float *gbias, void *workspace, cudaStream_t st) {
  CIMQ_REQUIRE(gy && x && save_mean && save_invstd && gx && workspace && (!relu || y), "bn_backward: bad argument");
  const dim3 grid(C, bn_splits(B, C));
  const bool vec = HW % 4 == 0 && aligned16(gy) && aligned16(x) && aligned16(y) && aligned16(gx) && aligned16(gres);
  double2 *part = reinterpret_cast<double2 *>(workspace);
  if (vec) {
    bn_bwd_stats_kernel<true><<<grid, kBnThreads, 0, st>>>(gy, x, y, save_mean, relu, B, C, HW, part);
    bn_bwd_apply_kernel<true><<<grid, kBnThreads, 0, st>>>(gy, x, y, weight, save_mean, save_invstd, part, training,
                                                           relu, B, C, HW, gx, gres, gweight, gbias);
  } else {
    bn_bwd_stats_kernel<false><<<grid, kBnThreads, 0, st>>>(gy, x, y, save_mean, relu, B, C, HW, part);
    bn_bwd_apply_kernel<false><<<grid, kBnThreads, 0, st>>>(gy, x, y, weight, save_mean, save_invstd, part, training,
                                                            relu, B, C, HW, gx, gres, gweight, gbias);
  }
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cimq
