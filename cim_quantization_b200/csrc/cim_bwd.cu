// Backward of the CiM convolution (get_cim_output_signed.backward, lsq.py:244-386): the alpha-grad kernels, the
// fold (col2im), the CUDA-core dgrad / wgrad for shapes outside the tcgen05 envelope, and the dispatch to the
// tcgen05 kernels of cim_bwd_tc.cu.
//
// The reference broadcasts grad_out to the 6-D partial-sum shape, zeroes clipped entries and runs
// 2*NX*NSA*NSW small GEMMs.  Here the STE clip mask comes from the ADC state the forward stored
// (1 clip bit per partial sum) and the slice loops are pre-reduced algebraically:
//
//   grad_xunf[m, f] = s_w/NSA * sum_{k,co} go[m,co] * PW[m,i(f),k,co] * wdig_k[f,co]
//                     PW = sum_j pass[k,j] * mask[k,j] * 2^(-abs*j)
//   grad_w[f, co]   = s_a/NSW * sum_{m,j} xdig_j[m,f] * go[m,co] * PV[m,i(f),j,co]
//                     PV = sum_k pass[k,j] * mask[k,j] * 2^(-wbs*k)
//   grad_alpha[i,k,j,co] = mask[k,j]/sqrt(numel*Qp) * sum_m code[m,i,k,j,co] * go[m,co]
//
// which is the same arithmetic in a different summation order (results agree to fp32 rounding).
#include "cimq_common.cuh"
#include "cim_v2.cuh"

namespace cimq {

namespace {

constexpr int kMaxPairs = 64;

__device__ __forceinline__ uint32_t state_bit(const uint32_t *__restrict__ state, const Geo &g, int i, int c,
                                              int64_t m, int bit) {
  return (state[(((int64_t)i * g.Cout + c) * g.state_words + (bit >> 5)) * g.M + m] >> (bit & 31)) & 1u;
}

// ---------------------------------------------------------------------------------------------
// grad_alpha: partial[ms][e] = sum over a range of pixels of code * go
// grid (Cout, NX, MS), block 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bwd_alpha_partial_kernel(Geo g, int m_per_split,
                                                                const float *__restrict__ go,
                                                                const uint32_t *__restrict__ state,
                                                                float *__restrict__ partial) {
  __shared__ float red[8][8];
  const int c = blockIdx.x, i = blockIdx.y, ms = blockIdx.z;
  const int64_t mbeg = (int64_t)ms * m_per_split;
  const int64_t mend = min((int64_t)g.M, mbeg + m_per_split);
  for (int q0 = 0; q0 < g.pairs; q0 += 8) {
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.0f;
    for (int64_t m = mbeg + threadIdx.x; m < mend; m += 256) {
      int b = (int)(m / g.L), l = (int)(m % g.L);
      float gv = go[((int64_t)b * g.Cout + c) * g.L + l];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        int q = q0 + t;
        if (q < g.pairs) {
          const int sq = state_pair(g, q / g.NSA, q % g.NSA);
          uint32_t pos = state_bit(state, g, i, c, m, sq), neg = state_bit(state, g, i, c, m, g.pairs + sq);
          acc[t] += pos ? gv : (neg ? -gv : 0.0f);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float v = acc[t];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8 && q0 + threadIdx.x < g.pairs) {
      float v = 0.0f;
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      int64_t e = ((int64_t)i * g.pairs + q0 + threadIdx.x) * g.Cout + c;
      partial[(int64_t)ms * table_entries(g) + e] = v;
    }
    __syncthreads();
  }
}

// Fast path (compile-time slice counts): one pass over (state words, go) per (pixel, crossbar, channel),
// PAIRS register accumulators, predicated adds; HBM-bound (state + go bytes).
template <int NSW, int NSA>
__global__ void __launch_bounds__(256) bwd_alpha_partial_w1_kernel(Geo g, int m_per_split,
                                                                   const float *__restrict__ go,
                                                                   const uint32_t *__restrict__ state,
                                                                   float *__restrict__ partial) {
  constexpr int PAIRS = NSW * NSA;
  constexpr int SWORDS = (3 * PAIRS + 31) / 32;  // state words per (crossbar, channel, pixel)
  constexpr int NW = (2 * PAIRS + 31) / 32;      // words holding the +1 / -1 code bits
  constexpr int U = PAIRS <= 16 ? 4 : 2;         // independent (state, go) loads in flight per thread
  __shared__ float red[8][PAIRS];
  const int c = blockIdx.x, i = blockIdx.y, ms = blockIdx.z;
  const int64_t mbeg = (int64_t)ms * m_per_split;
  const int64_t mend = min((int64_t)g.M, mbeg + m_per_split);
  const uint32_t *st = state + ((int64_t)i * g.Cout + c) * SWORDS * g.M;
  float acc[PAIRS];
#pragma unroll
  for (int q = 0; q < PAIRS; ++q) acc[q] = 0.0f;
  // the image index advances incrementally (no per-element division)
  int64_t m = mbeg + threadIdx.x;
  int b = (int)(m / g.L), l = (int)(m % g.L);
  const float *gbase = go + (int64_t)c * g.L;
  const int64_t gimg = (int64_t)g.Cout * g.L;
  for (; m < mend; m += U * 256) {
    float gv[U];
    uint32_t w[U][NW];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = m + u * 256 < mend;
      gv[u] = ok ? __ldg(gbase + (int64_t)b * gimg + l) : 0.0f;
#pragma unroll
      for (int t = 0; t < NW; ++t) w[u][t] = ok ? __ldg(st + (int64_t)t * g.M + m + u * 256) : 0u;
      l += 256;
      while (l >= g.L) { l -= g.L; ++b; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int q = 0; q < PAIRS; ++q) {
        // table pair q = k*NSA + j lives at state pair sq = j*NSW + k: +gv if bit sq, -gv if bit PAIRS+sq
        const int sq = (q % NSA) * NSW + q / NSA;
        const int bp = sq, bn = PAIRS + sq;
        if (w[u][bp >> 5] & (1u << (bp & 31))) acc[q] += gv[u];  // bit test into a predicate + predicated add
        if (w[u][bn >> 5] & (1u << (bn & 31))) acc[q] -= gv[u];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PAIRS; ++q) {
    float v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < PAIRS) {
    float v = 0.0f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    partial[(int64_t)ms * table_entries(g) + ((int64_t)i * PAIRS + threadIdx.x) * g.Cout + c] = v;
  }
}

// Vectorised variant: a thread handles four consecutive pixels per load (one 128-bit load of the state word(s) and
// of grad_out, one address computation per four pixels -- the scalar version is bound by the integer pipe, and a
// third of its integer work is addressing).  Needs L % 4 == 0 and split boundaries that are multiples of 4.
template <int NSW, int NSA>
__global__ void __launch_bounds__(256) bwd_alpha_partial_w4_kernel(Geo g, int m_per_split,
                                                                   const float *__restrict__ go,
                                                                   const uint32_t *__restrict__ state,
                                                                   float *__restrict__ partial) {
  constexpr int PAIRS = NSW * NSA;
  constexpr int SWORDS = (3 * PAIRS + 31) / 32;
  constexpr int NW = (2 * PAIRS + 31) / 32;
  constexpr int U = PAIRS <= 16 ? 2 : 1;  // independent 4-pixel groups in flight per thread
  __shared__ float red[8][PAIRS];
  const int c = blockIdx.x, i = blockIdx.y, ms = blockIdx.z;
  const int64_t mbeg = (int64_t)ms * m_per_split;
  const int64_t mend = min((int64_t)g.M, mbeg + m_per_split);
  const uint32_t *st = state + ((int64_t)i * g.Cout + c) * SWORDS * g.M;
  float acc[PAIRS];
#pragma unroll
  for (int q = 0; q < PAIRS; ++q) acc[q] = 0.0f;
  int64_t m = mbeg + 4 * threadIdx.x;
  int b = (int)(m / g.L), l = (int)(m % g.L);
  const float *gbase = go + (int64_t)c * g.L;
  const int64_t gimg = (int64_t)g.Cout * g.L;
  for (; m < mend; m += U * 1024) {
    float4 gv[U];
    uint4 w[U][NW];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = m + u * 1024 < mend;
      gv[u] = ok ? __ldg(reinterpret_cast<const float4 *>(gbase + (int64_t)b * gimg + l))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < NW; ++t)
        w[u][t] = ok ? __ldg(reinterpret_cast<const uint4 *>(st + (int64_t)t * g.M + m + u * 1024))
                     : make_uint4(0u, 0u, 0u, 0u);
      l += 1024;
      while (l >= g.L) { l -= g.L; ++b; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float g4[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) {
          const int sq = (q % NSA) * NSW + q / NSA;
          const int bp = sq, bn = PAIRS + sq;
          const uint4 wp = w[u][bp >> 5], wn = w[u][bn >> 5];
          const uint32_t wpe = e == 0 ? wp.x : e == 1 ? wp.y : e == 2 ? wp.z : wp.w;
          const uint32_t wne = e == 0 ? wn.x : e == 1 ? wn.y : e == 2 ? wn.z : wn.w;
          if (wpe & (1u << (bp & 31))) acc[q] += g4[e];
          if (wne & (1u << (bn & 31))) acc[q] -= g4[e];
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PAIRS; ++q) {
    float v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < PAIRS) {
    float v = 0.0f;
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8][threadIdx.x];
    partial[(int64_t)ms * table_entries(g) + ((int64_t)i * PAIRS + threadIdx.x) * g.Cout + c] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// grad_alpha from the v2 state (cim_v2.cuh): plane C [NX][NSA][M][Cout], byte = sum_k 4^k * (code_k + 1).
//   sum_m code * go = sum_m (code + 1) * go - sum_m go
// A field is multiplied in WITHOUT an integer-to-float conversion: the masked bits, read as an fp32 subnormal, are
// field * 2^(pos - 149) exactly, and go is pre-scaled by 2^100 so that the product is a normal number; the
// power-of-two factors are undone once at the end.  2 instructions (mask, FMA) per partial sum.
// Block (i*NSA + j, split): 8 warps; a warp-iteration covers 512 contiguous bytes of the plane
// (32 / LPP pixels x Cout channels, LPP = Cout / 16 lanes per pixel; lane = (pixel, 16-channel group)).
// ---------------------------------------------------------------------------------------------
template <int NSW>
__global__ void __launch_bounds__(256, 3) bwd_alpha_v2_kernel(Geo g, int m_per_split, int cblock,
                                                           const float *__restrict__ go,
                                                           const uint8_t *__restrict__ cplanes,
                                                           float *__restrict__ partial) {
  // item = (channel quad, 8-pixel group): 8 state words (one per pixel: 4 channels x NSW two-bit code fields) and
  // 4 x 8 grad_out values (two 16-byte loads per channel).  lane = quad + nq * (pixel group inside the warp);
  // 4 x NSW + 4 accumulators per thread.  blockIdx = (i*NSA + j, pixel split, block of <= 128 channels).
  extern __shared__ float red[];  // [8 warps][NSW * cblock]
  const int ij = blockIdx.x, i = ij / g.NSA, j = ij % g.NSA, ms = blockIdx.y, c0 = blockIdx.z * cblock;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nq = cblock >> 2, gpw = 32 / nq;  // quads per pixel, pixel groups per warp
  const int q = lane % nq, pgl = lane / nq;
  const int mbeg = ms * m_per_split;
  const int mend = min(g.M, mbeg + m_per_split);
  const uint8_t *plane = cplanes + (int64_t)ij * g.M * g.Cout + c0 + 4 * q;
  const float *gbase = go + (int64_t)(c0 + 4 * q) * g.L;
  const int64_t gimg = (int64_t)g.Cout * g.L;
  const int step = 8 * gpw * 8;  // pixels per block iteration
  float acc[4][NSW], accg[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    accg[c] = 0.0f;
#pragma unroll
    for (int k = 0; k < NSW; ++k) acc[c][k] = 0.0f;
  }
  int m = mbeg + (warp * gpw + pgl) * 8;
  int b = m / g.L, l = m % g.L;
  for (; m < mend; m += step) {
    uint32_t w[8];
    float gv[4][8];
    const uint8_t *wp = plane + (int64_t)m * g.Cout;
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = __ldg(reinterpret_cast<const uint32_t *>(wp + (int64_t)e * g.Cout));
    const float *gp = gbase + b * gimg + l;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gp + (int64_t)c * g.L));
      const float4 g1 = __ldg(reinterpret_cast<const float4 *>(gp + (int64_t)c * g.L) + 1);
      gv[c][0] = g0.x; gv[c][1] = g0.y; gv[c][2] = g0.z; gv[c][3] = g0.w;
      gv[c][4] = g1.x; gv[c][5] = g1.y; gv[c][6] = g1.z; gv[c][7] = g1.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t lo = w[e], hi = w[e] >> 16;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t src = c < 2 ? lo : hi;
        accg[c] += gv[c][e];
        const float gsc = gv[c][e] * 1.2676506002282294e30f;  // 2^100
#pragma unroll
        for (int k = 0; k < NSW; ++k)
          acc[c][k] = fmaf(gsc, __uint_as_float(src & (3u << (8 * (c & 1) + 2 * k))), acc[c][k]);
      }
    }
    l += step;
    while (l >= g.L) { l -= g.L; ++b; }
  }
  // undo the scaling, subtract sum go, reduce over the lanes that share a quad, then over warps
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < NSW; ++k) {
      float v = acc[c][k] * exp2f((float)(49 - 8 * (c & 1) - 2 * k)) - accg[c];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1)
        if (o >= nq) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (pgl == 0) red[(warp * NSW + k) * cblock + 4 * q + c] = v;
    }
  __syncthreads();
  for (int t = threadIdx.x; t < NSW * cblock; t += 256) {
    float v = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8 * NSW * cblock + t];
    const int k = t / cblock, c = c0 + t % cblock;
    partial[(int64_t)ms * table_entries(g) + ((int64_t)i * g.pairs + k * g.NSA + j) * g.Cout + c] = v;
  }
}

// Block = 32 consecutive table entries x 8 slices of the split range, combined through shared memory in a fixed
// order (deterministic).
__global__ void __launch_bounds__(1024) bwd_alpha_finish_kernel(Geo g, int nsplit, float gfac,
                                                                const int8_t *__restrict__ mask,
                                                                const float *__restrict__ partial,
                                                                float *__restrict__ galpha) {
  __shared__ float red[32][32];
  const int64_t n = table_entries(g);
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5, nslice = blockDim.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * 32 + lane;
  float v = 0.0f;
  if (e < n) {
    // four loads in flight (fixed summation order: deterministic)
    float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
    int sidx = slice;
    for (; sidx + 3 * nslice < nsplit; sidx += 4 * nslice) {
      const float a0 = __ldg(partial + (int64_t)sidx * n + e), a1 = __ldg(partial + (int64_t)(sidx + nslice) * n + e);
      const float a2 = __ldg(partial + (int64_t)(sidx + 2 * nslice) * n + e), a3 = __ldg(partial + (int64_t)(sidx + 3 * nslice) * n + e);
      v0 += a0; v1 += a1; v2 += a2; v3 += a3;
    }
    for (; sidx < nsplit; sidx += nslice) v0 += __ldg(partial + (int64_t)sidx * n + e);
    v = (v0 + v1) + (v2 + v3);
  }
  red[slice][lane] = v;
  __syncthreads();
  if (slice == 0 && e < n) {
    float t = red[0][lane];
    for (int w = 1; w < nslice; ++w) t += red[w][lane];
    const int q = (int)((e / g.Cout) % g.pairs);
    galpha[e] = t * gfac * (float)mask[q];  // lsq.py:306, 323-325 / 330-332
  }
}

// ---------------------------------------------------------------------------------------------
// grad wrt the unfolded input, image-major and transposed: gxu[b][f][l]
// grid (ceil(M/32), NX, ceil(rowsmax/128)), block 128; smem Ap[NSW*Cout][32]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bwd_input_kernel(Geo g, const float *__restrict__ go,
                                                        const float *__restrict__ wdigits,
                                                        const uint32_t *__restrict__ state,
                                                        const float *__restrict__ s,
                                                        const int8_t *__restrict__ mask,
                                                        float *__restrict__ gxuT) {
  extern __shared__ __align__(16) float Ap[];  // [NSW*Cout][32]
  __shared__ float wx[kMaxPairs];              // mask[k][j] * 2^(-abs*j)
  const int i = blockIdx.y;
  const int lo = i * g.xbar, hi = min(lo + g.xbar, g.F);
  const int f = lo + blockIdx.z * 128 + threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * 32;
  if (lo + blockIdx.z * 128 >= hi) return;  // block-uniform
  if (threadIdx.x < g.pairs) {
    int j = threadIdx.x % g.NSA;
    wx[threadIdx.x] = (float)mask[threadIdx.x] * exp2f(-(float)(g.abs_ * j));
  }
  __syncthreads();
  const int nkc = g.NSW * g.Cout;
  for (int idx = threadIdx.x; idx < nkc * 32; idx += 128) {
    int p = idx & 31, kc = idx >> 5;
    int k = kc / g.Cout, co = kc % g.Cout;
    int64_t m = m0 + p;
    float v = 0.0f;
    if (m < g.M) {
      float pw = 0.0f;
      for (int j = 0; j < g.NSA; ++j) {
        int q = k * g.NSA + j;
        if (!state_bit(state, g, i, co, m, state_clip_bit(g, state_pair(g, k, j)))) pw += wx[q];
      }
      int b = (int)(m / g.L), l = (int)(m % g.L);
      v = go[((int64_t)b * g.Cout + co) * g.L + l] * pw;
    }
    Ap[kc * 32 + p] = v;
  }
  __syncthreads();
  float acc[32];
#pragma unroll
  for (int p = 0; p < 32; ++p) acc[p] = 0.0f;
  const bool active = f < hi;
  for (int kc = 0; kc < nkc; ++kc) {
    float bv = active ? __ldg(&wdigits[(int64_t)kc * g.F + f]) : 0.0f;  // wdigits is [NSW][Cout][F]
    const float4 *a4 = reinterpret_cast<const float4 *>(Ap + kc * 32);
#pragma unroll
    for (int p4 = 0; p4 < 8; ++p4) {
      float4 a = a4[p4];
      acc[4 * p4 + 0] = fmaf(a.x, bv, acc[4 * p4 + 0]);
      acc[4 * p4 + 1] = fmaf(a.y, bv, acc[4 * p4 + 1]);
      acc[4 * p4 + 2] = fmaf(a.z, bv, acc[4 * p4 + 2]);
      acc[4 * p4 + 3] = fmaf(a.w, bv, acc[4 * p4 + 3]);
    }
  }
  if (active) {
    const float scale = s[1] / (float)g.NSA;  // w_sl * s_w (lsq.py:252), mean over act slices (lsq.py:376)
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      const int64_t m = m0 + p;
      if (m < g.M) gxuT[((m / g.L) * g.F + f) * g.L + m % g.L] = acc[p] * scale;  // gxu[b][f][l]
    }
  }
}

// col2im (nn.Fold, lsq.py:380-382) as a gather: one thread per input element, the K*K tap loads are
// predicated and independent (K is a template parameter so they unroll); reads of one tap are coalesced
// (consecutive input columns -> consecutive output pixels of one gxuT row).  K == 0: runtime kernel size.
template <int KT>
__global__ void __launch_bounds__(256) col2im_kernel(Geo g, const float *__restrict__ gxuT,
                                                     float *__restrict__ gx) {
  // block = one (image, input channel) plane: no per-element division by Cin / B, 32-bit offsets inside the
  // channel's K*K rows of gxuT (K*K*M < 2^31 is checked by the launcher)
  const int K = KT > 0 ? KT : g.K;
  const int ci = blockIdx.x % g.Cin, b = blockIdx.x / g.Cin;
  const float *base = gxuT + ((int64_t)b * g.F + ci * K * K) * g.L;  // gxu[b][f][l]
  float *dst = gx + (int64_t)blockIdx.x * g.H * g.W;
  const int HW = g.H * g.W;
  const int M = g.L;  // stride between the K*K tap rows of one channel
  for (int idx = blockIdx.y * 256 + threadIdx.x; idx < HW; idx += gridDim.y * 256) {
    const int iy = idx / g.W, ix = idx - iy * g.W;
    float v = 0.0f;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int ty = iy + g.pad - ky;
      int oy = ty;
      bool oky = ty >= 0;
      if (g.stride != 1) { oky = oky && (ty % g.stride == 0); oy = ty / g.stride; }
      oky = oky && oy < g.OH;
      const int rowo = oy * g.OW + ky * K * M;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int tx = ix + g.pad - kx;
        int ox = tx;
        bool ok = oky && tx >= 0;
        if (g.stride != 1) { ok = ok && (tx % g.stride == 0); ox = tx / g.stride; }
        ok = ok && ox < g.OW;
        const float t = ok ? __ldg(base + (rowo + kx * M + ox)) : 0.0f;
        v += t;
      }
    }
    dst[idx] = v;
  }
}

inline int launch_col2im(const Geo &g, const float *gxuT, float *gx, cudaStream_t st) {
  CIMQ_REQUIRE((int64_t)g.KK * g.L < (1ll << 31), "col2im: K*K*L must be below 2^31");
  const int hw = g.H * g.W;
  int by = (hw + 1023) / 1024;  // ~4 pixels per thread
  if (by < 1) by = 1;
  dim3 grid(g.B * g.Cin, by);
  if (g.K == 3) col2im_kernel<3><<<grid, 256, 0, st>>>(g, gxuT, gx);
  else if (g.K == 1) col2im_kernel<1><<<grid, 256, 0, st>>>(g, gxuT, gx);
  else if (g.K == 5) col2im_kernel<5><<<grid, 256, 0, st>>>(g, gxuT, gx);
  else col2im_kernel<0><<<grid, 256, 0, st>>>(g, gxuT, gx);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// grad wrt the weight codes: partial[ms][f][co]
// grid (NX * ceil(rowsmax/128), ceil(Cout/32), SPLITS), block 128; thread <-> crossbar row f
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bwd_weight_kernel(Geo g, int ftiles, int m_per_split,
                                                         const float *__restrict__ go,
                                                         const uint8_t *__restrict__ xcodes,
                                                         const uint32_t *__restrict__ state,
                                                         const float *__restrict__ s,
                                                         const int8_t *__restrict__ mask,
                                                         float *__restrict__ partial) {
  extern __shared__ __align__(16) float Bs[];  // [32 pixels][NSA][32 channels]
  __shared__ float wv[kMaxPairs];              // mask[k][j] * 2^(-wbs*k)
  const int i = blockIdx.x / ftiles, ft = blockIdx.x % ftiles;
  const int lo = i * g.xbar, hi = min(lo + g.xbar, g.F);
  if (lo + ft * 128 >= hi) return;  // block-uniform
  const int f = lo + ft * 128 + threadIdx.x;
  const bool active = f < hi;
  const int co0 = blockIdx.y * 32;
  const int64_t mbeg = (int64_t)blockIdx.z * m_per_split;
  const int64_t mend = min((int64_t)g.M, mbeg + m_per_split);
  if (threadIdx.x < g.pairs) {
    int k = threadIdx.x / g.NSA;
    wv[threadIdx.x] = (float)mask[threadIdx.x] * exp2f(-(float)(g.wbs * k));
  }
  int ci = 0, ky = 0, kx = 0;
  if (active) { ci = f / g.KK; int tap = f % g.KK; ky = tap / g.K; kx = tap % g.K; }
  float acc[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) acc[t] = 0.0f;

  for (int64_t mb = mbeg; mb < mend; mb += 32) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * g.NSA * 32; idx += 128) {
      int p = idx & 31, col = (idx >> 5) & 31, j = idx >> 10;  // pixel fastest: coalesced go / state reads
      int64_t m = mb + p;
      int co = co0 + col;
      float v = 0.0f;
      if (m < mend && co < g.Cout) {
        float pv = 0.0f;
        for (int k = 0; k < g.NSW; ++k) {
          int q = k * g.NSA + j;
          if (!state_bit(state, g, i, co, m, state_clip_bit(g, state_pair(g, k, j)))) pv += wv[q];
        }
        int b = (int)(m / g.L), l = (int)(m % g.L);
        v = go[((int64_t)b * g.Cout + co) * g.L + l] * pv;
      }
      Bs[(p * g.NSA + j) * 32 + col] = v;
    }
    __syncthreads();
    if (active) {
      for (int p = 0; p < 32; ++p) {
        int64_t m = mb + p;
        if (m >= mend) break;
        int b = (int)(m / g.L), l = (int)(m % g.L), oy = l / g.OW, ox = l % g.OW;
        int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
        int code = 0;
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
          code = xcodes[(((int64_t)b * g.Cin + ci) * g.H + iy) * g.W + ix];
        if (code == 0) continue;
        for (int j = 0; j < g.NSA; ++j) {
          float d = (float)((code >> (g.abs_ * j)) & g.amask);
          if (d == 0.0f) continue;
          const float4 *b4 = reinterpret_cast<const float4 *>(Bs + (p * g.NSA + j) * 32);
#pragma unroll
          for (int t4 = 0; t4 < 8; ++t4) {
            float4 bv = b4[t4];
            acc[4 * t4 + 0] = fmaf(d, bv.x, acc[4 * t4 + 0]);
            acc[4 * t4 + 1] = fmaf(d, bv.y, acc[4 * t4 + 1]);
            acc[4 * t4 + 2] = fmaf(d, bv.z, acc[4 * t4 + 2]);
            acc[4 * t4 + 3] = fmaf(d, bv.w, acc[4 * t4 + 3]);
          }
        }
      }
    }
  }
  if (active) {
    const float scale = s[0] / (float)g.NSW;  // x_sl * s_a (lsq.py:295), mean over weight slices (lsq.py:366)
    float *dst = partial + ((int64_t)blockIdx.z * g.F + f) * g.Cout + co0;
#pragma unroll
    for (int t = 0; t < 32; ++t)
      if (co0 + t < g.Cout) dst[t] = acc[t] * scale;
  }
}

__global__ void bwd_weight_finish_kernel(Geo g, int nsplit, const float *__restrict__ partial,
                                         float *__restrict__ gw) {
  const int64_t n = (int64_t)g.Cout * g.F;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int f = (int)(idx % g.F), co = (int)(idx / g.F);
    float v = 0.0f;
    for (int sidx = 0; sidx < nsplit; ++sidx) v += partial[((int64_t)sidx * g.F + f) * g.Cout + co];
    gw[idx] = v;  // [Cout, F] == weight layout, lsq.py:369
  }
}

struct BwdPlan {
  int alpha_splits, alpha_m_per_split, alpha_splits_v2;
  int w_splits, w_m_per_split, ftiles;
  int64_t off_gxu, off_wpart, off_apart, off_scales, total;
};

inline BwdPlan make_plan(const Geo &g) {
  BwdPlan p;
  const int target_blocks = 148 * 4;
  // alpha: grid (Cout, NX, splits)
  // ~16 pixels per thread (256 threads/block): short dependent chains, many blocks
  int as = (g.M + 4095) / 4096;
  (void)target_blocks;
  as = as < 1 ? 1 : as; if (as > 64) as = 64;
  p.alpha_splits = as;
  p.alpha_m_per_split = (g.M + as - 1) / as;
  int rowsmax = g.xbar < g.F ? g.xbar : g.F;
  p.ftiles = (rowsmax + 127) / 128;
  int base = g.NX * p.ftiles * ((g.Cout + 31) / 32);
  int ws = (target_blocks + base - 1) / base;
  int max_ws = (g.M + 255) / 256;
  ws = ws < 1 ? 1 : ws; if (ws > max_ws) ws = max_ws; if (ws > 256) ws = 256;
  int mps = (g.M + ws - 1) / ws;
  mps = (mps + 31) & ~31;
  p.w_m_per_split = mps;
  p.w_splits = (g.M + mps - 1) / mps;
  auto align = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
  p.off_gxu = 0;
  p.off_wpart = align((int64_t)g.F * g.M * 4);
  int64_t wpart = (int64_t)p.w_splits * g.F * g.Cout * 4;
  if (tc_backward_supported(g) && bwd_tc_partial_bytes(g) > wpart) wpart = bwd_tc_partial_bytes(g);
  p.off_apart = p.off_wpart + align(wpart);
  p.alpha_splits_v2 = (148 * 6 + g.NX * g.NSA - 1) / (g.NX * g.NSA) + 1;
  if (alpha_v3_supported(g) && alpha_v3_blocks(g) > p.alpha_splits_v2) p.alpha_splits_v2 = alpha_v3_blocks(g);
  const int asmax = p.alpha_splits > p.alpha_splits_v2 ? p.alpha_splits : p.alpha_splits_v2;
  p.off_scales = p.off_apart + align((int64_t)asmax * table_entries(g) * 4);
  p.total = p.off_scales + (v2_backward_supported(g) ? align(bwd_v2_scales_bytes(g)) : 0);
  return p;
}

}  // namespace

int64_t conv_backward_ws_bytes(const Geo &g) { return make_plan(g).total; }

int launch_conv_backward(const Geo &g, const float *go, const uint8_t *xcodes, const float *wdigits,
                         const void *wtiles, const uint32_t *state, const float *s, const int8_t *mask, float *gxq,
                         float *gwq, float *galpha, void *ws, uint32_t flags, cudaStream_t st) {
  CIMQ_REQUIRE(go && xcodes && state && s && mask && ws, "conv_backward: NULL argument");
  const bool use_tc = !(flags & CIMQ_FLAG_FORCE_SIMT) && wtiles != nullptr && tc_backward_supported(g);
  CIMQ_REQUIRE(use_tc || wdigits != nullptr, "conv_backward: wdigits is NULL");
  CIMQ_REQUIRE(g.pairs <= kMaxPairs, "too many slice pairs");
  const BwdPlan p = make_plan(g);
  char *base = reinterpret_cast<char *>(ws);
  float *gxuT = reinterpret_cast<float *>(base + p.off_gxu);
  float *wpart = reinterpret_cast<float *>(base + p.off_wpart);
  float *apart = reinterpret_cast<float *>(base + p.off_apart);

  const bool v2s = (flags & CIMQ_FLAG_V2) != 0;
  CIMQ_REQUIRE(!v2s || (use_tc && v2::supported(g) && v2_backward_supported(g)),
               "conv_backward: CIMQ_FLAG_V2 on a layer the v2 kernels do not cover");
  const uint8_t *state2 = reinterpret_cast<const uint8_t *>(state);
  void *scales = base + p.off_scales;
  if (v2s && (gxq != nullptr || gwq != nullptr)) {
    if (launch_go_scales(g, go, scales, st)) return 1;
  }
  if (v2s && galpha != nullptr && alpha_v3_supported(g)) {
    const uint8_t *cplanes = state2 + 2 * v2::plane_bytes(g);
    if (launch_alpha_v3(g, go, cplanes, apart, st)) return 1;
    const double numel = (double)g.B * g.NX * g.NSW * g.NSA * g.L * g.Cout;
    const int64_t n = table_entries(g);
    // (hundreds of partials per entry: 32 slices of the partial range per block instead of 8)
    bwd_alpha_finish_kernel<<<(int)((n + 31) / 32), alpha_v3_blocks(g) > 64 ? 1024 : 256, 0, st>>>(
        g, alpha_v3_blocks(g), (float)(1.0 / sqrt(numel)), mask, apart, galpha);
    CIMQ_CUDA_OK(cudaGetLastError());
  } else if (v2s && galpha != nullptr && g.adc_mode != CIMQ_ADC_MULTIBIT) {
    const int cblock = g.Cout > 128 ? (g.Cout % 128 == 0 ? 128 : 64) : g.Cout;  // channels per block (Cout % 16 == 0)
    const int nq = cblock / 4, pb = 8 * (32 / nq) * 8;  // pixels per block-iteration
    CIMQ_REQUIRE(32 % nq == 0, "conv_backward: alpha kernel needs 4, 8, 16 or 32 channel quads per block");
    int splits = (148 * 6 + g.NX * g.NSA - 1) / (g.NX * g.NSA * (g.Cout / cblock));
    if (splits < 1) splits = 1;
    int mps = (g.M + splits - 1) / splits;
    mps = (mps + pb - 1) / pb * pb;
    splits = (g.M + mps - 1) / mps;
    CIMQ_REQUIRE(splits <= p.alpha_splits_v2, "conv_backward: alpha workspace too small");
    dim3 grid(g.NX * g.NSA, splits, g.Cout / cblock);
    const size_t smem = (size_t)8 * g.NSW * cblock * sizeof(float);
    const uint8_t *cplanes = state2 + 2 * v2::plane_bytes(g);
    if (g.NSW == 3) bwd_alpha_v2_kernel<3><<<grid, 256, smem, st>>>(g, mps, cblock, go, cplanes, apart);
    else bwd_alpha_v2_kernel<2><<<grid, 256, smem, st>>>(g, mps, cblock, go, cplanes, apart);
    CIMQ_CUDA_OK(cudaGetLastError());
    double numel = (double)g.B * g.NX * g.NSW * g.NSA * g.L * g.Cout;
    float gfac = (float)(1.0 / sqrt(numel));
    int64_t n = table_entries(g);
    bwd_alpha_finish_kernel<<<(int)((n + 31) / 32), 256, 0, st>>>(g, splits, gfac, mask, apart, galpha);
    CIMQ_CUDA_OK(cudaGetLastError());
  } else if (galpha != nullptr && g.adc_mode != CIMQ_ADC_MULTIBIT) {
    dim3 grid(g.Cout, g.NX, p.alpha_splits);
    // four pixels per load when the pixel runs allow 128-bit loads (state rows and grad_out rows 16-byte aligned)
    const bool vec4 = g.L % 4 == 0 && p.alpha_m_per_split % 4 == 0 && g.M % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(go) & 15u) == 0 && (reinterpret_cast<uintptr_t>(state) & 15u) == 0;
    if (g.NSW == 3 && g.NSA == 3 && vec4)
      bwd_alpha_partial_w4_kernel<3, 3><<<grid, 256, 0, st>>>(g, p.alpha_m_per_split, go, state, apart);
    else if (g.NSW == 2 && g.NSA == 2 && vec4)
      bwd_alpha_partial_w4_kernel<2, 2><<<grid, 256, 0, st>>>(g, p.alpha_m_per_split, go, state, apart);
    else if (g.NSW == 4 && g.NSA == 4 && vec4)
      bwd_alpha_partial_w4_kernel<4, 4><<<grid, 256, 0, st>>>(g, p.alpha_m_per_split, go, state, apart);
    else if (g.NSW == 3 && g.NSA == 3)
      bwd_alpha_partial_w1_kernel<3, 3><<<grid, 256, 0, st>>>(g, p.alpha_m_per_split, go, state, apart);
    else if (g.NSW == 2 && g.NSA == 2)
      bwd_alpha_partial_w1_kernel<2, 2><<<grid, 256, 0, st>>>(g, p.alpha_m_per_split, go, state, apart);
    else if (g.NSW == 4 && g.NSA == 4)
      bwd_alpha_partial_w1_kernel<4, 4><<<grid, 256, 0, st>>>(g, p.alpha_m_per_split, go, state, apart);
    else if (g.NSW == 8 && g.NSA == 8)
      bwd_alpha_partial_w1_kernel<8, 8><<<grid, 256, 0, st>>>(g, p.alpha_m_per_split, go, state, apart);
    else
      bwd_alpha_partial_kernel<<<grid, 256, 0, st>>>(g, p.alpha_m_per_split, go, state, apart);
    CIMQ_CUDA_OK(cudaGetLastError());
    // 1/sqrt(ps.numel() * Qp_adc) with Qp_adc = 1 (lsq.py:323, 330)
    double numel = (double)g.B * g.NX * g.NSW * g.NSA * g.L * g.Cout;
    float gfac = (float)(1.0 / sqrt(numel));
    int64_t n = table_entries(g);
    bwd_alpha_finish_kernel<<<(int)((n + 31) / 32), 256, 0, st>>>(g, p.alpha_splits, gfac, mask, apart, galpha);
    CIMQ_CUDA_OK(cudaGetLastError());
  }
  if (gxq != nullptr && use_tc) {
    const WtLayout wl = wt_layout(g);
    const uint8_t *wtb = reinterpret_cast<const uint8_t *>(wtiles) + wl.bwd_off;
    const uint8_t *wtb2 = reinterpret_cast<const uint8_t *>(wtiles) + wl.bwd2_off;
    if (!(flags & CIMQ_FLAG_DETERMINISTIC) && bwd_input_tc_can_fold(g)) {
      // fused fold: the dgrad epilogue reduces into grad_x (fp32 atomics in L2; summation order varies run to run)
      CIMQ_CUDA_OK(cudaMemsetAsync(gxq, 0, (size_t)g.B * g.Cin * g.H * g.W * sizeof(float), st));
      if (v2s) { if (launch_bwd_input_v2(g, go, state2, wtb2, s, scales, gxq, 1, st)) return 1; }
      else if (launch_bwd_input_tc(g, go, state, wtb, s, mask, gxq, 1, false, st)) return 1;
    } else {
      if (v2s) { if (launch_bwd_input_v2(g, go, state2, wtb2, s, scales, gxuT, 0, st)) return 1; }
      else if (launch_bwd_input_tc(g, go, state, wtb, s, mask, gxuT, 0, false, st)) return 1;
      if (launch_col2im(g, gxuT, gxq, st)) return 1;
    }
  } else if (gxq != nullptr) {
    size_t smem = (size_t)g.NSW * g.Cout * 32 * sizeof(float);
    CIMQ_REQUIRE(smem <= 200 * 1024, "conv_backward: NSW*Cout too large for the SIMT kernel");
    CIMQ_CUDA_OK(cudaFuncSetAttribute(bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((g.M + 31) / 32, g.NX, p.ftiles);
    bwd_input_kernel<<<grid, 128, smem, st>>>(g, go, wdigits, state, s, mask, gxuT);
    CIMQ_CUDA_OK(cudaGetLastError());
    if (launch_col2im(g, gxuT, gxq, st)) return 1;
  }
  if (gwq != nullptr && use_tc) {
    if (launch_bwd_weight_tc(g, go, xcodes, state, s, mask, wpart, gwq, v2s, scales, st)) return 1;
  } else if (gwq != nullptr) {
    size_t smem = (size_t)32 * g.NSA * 32 * sizeof(float);
    CIMQ_CUDA_OK(cudaFuncSetAttribute(bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(g.NX * p.ftiles, (g.Cout + 31) / 32, p.w_splits);
    bwd_weight_kernel<<<grid, 128, smem, st>>>(g, p.ftiles, p.w_m_per_split, go, xcodes, state, s, mask, wpart);
    CIMQ_CUDA_OK(cudaGetLastError());
    int64_t n = (int64_t)g.Cout * g.F;
    bwd_weight_finish_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(g, p.w_splits, wpart, gwq);
    CIMQ_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

}  // namespace cimq
