// CUDA-core (SIMT) CiM convolution: the general-shape kernel behind cimq_conv_forward for layers
// the tcgen05 kernel does not cover, behind cimq_conv_psums (raw partial sums, tests) and
// cimq_conv_psum_abs_sums (alpha_cim initialisation statistic).  Reference: lsq.py:92-237, 35-87.
//
// One block = 32 output pixels x 8 output channels.  Per crossbar chunk the block stages the
// activation digit planes [NSA][32][K] and the weight digit planes [NSW][8][K] in shared memory as
// bytes; every thread owns one (pixel, channel) and evaluates its NSW*NSA partial sums with
// dp4a.u32.s32 (unsigned activation digits x signed weight digits), applying the ADC to each.
#include "cimq_common.cuh"

namespace cimq {

namespace {

constexpr int kPix = 32;   // pixels per block (threadIdx.x)
constexpr int kCh = 8;     // channels per block (threadIdx.y)
constexpr int kMaxStateWords = 6;  // 3 bits x 64 slice pairs

__device__ __forceinline__ int dp4a_u8s8(uint32_t a, int32_t b, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
  return d;
}

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so the random stream of a partial sum depends only on
// (seed, its index, draw number) -- reproducible for a given seed whatever the launch geometry.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// Number of successes of `n` Bernoulli(prob) draws, drawn the way the reference does (lsq.py:215-216):
// ceil(prob - U) with U uniform in [0, 1), i.e. success iff U < prob.  `stream` selects the sigmoid (0 / 1).
__device__ __forceinline__ int bernoulli_count(float prob, int n, unsigned long long elem, uint32_t stream, uint2 key) {
  if (prob >= 1.0f) return n;
  if (prob <= 0.0f) return 0;
  int cnt = 0;
  for (int d = 0; d < n; d += 4) {
    const uint4 r = philox4x32(make_uint4((uint32_t)elem, (uint32_t)(elem >> 32), (uint32_t)(d >> 2), stream), key);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (d + t < n) cnt += ((float)(rr[t] >> 8) * 5.9604644775390625e-8f) < prob;  // 24-bit uniform
  }
  return cnt;
}

template <int MODE>
__global__ void __launch_bounds__(kPix * kCh) conv_simt_kernel(
    Geo g, int KS, const uint8_t *__restrict__ xcodes, const int8_t *__restrict__ wcodes,
    const int4 *__restrict__ table, const float *__restrict__ s, float *__restrict__ out,
    uint32_t *__restrict__ state, int32_t *__restrict__ psums, unsigned long long *__restrict__ sums,
    const float *__restrict__ alpha_q, unsigned long long seed) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t *adig = smem;                                  // [NSA][kPix][KS]
  int8_t *wdig = reinterpret_cast<int8_t *>(smem + (size_t)g.NSA * kPix * KS);  // [NSW][kCh][KS]

  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kPix + tx;
  const int m0 = blockIdx.x * kPix, c0 = blockIdx.y * kCh;
  const int m = m0 + tx, c = c0 + ty;
  const bool valid = (m < g.M) && (c < g.Cout);
  const int b = m / g.L, l = m % g.L;
  float sa = 1.f, sw = 1.f;
  if (MODE == SIMT_FORWARD || MODE == SIMT_FORWARD_STOCH) { sa = s[0]; sw = s[1]; }

  float acc = 0.0f;
  for (int i = 0; i < g.NX; ++i) {
    const int lo = i * g.xbar, hi = min(lo + g.xbar, g.F);
    const int rows = hi - lo, rows4 = (rows + 3) & ~3;
    __syncthreads();
    // activation digit planes: consecutive threads -> consecutive pixels (coalesced byte reads)
    for (int idx = tid; idx < kPix * rows4; idx += kPix * kCh) {
      int pix = idx % kPix, kk = idx / kPix;
      int mm = m0 + pix, code = 0;
      if (kk < rows && mm < g.M) {
        int f = lo + kk;
        int ci = f / g.KK, tap = f % g.KK, ky = tap / g.K, kx = tap % g.K;  // nn.Unfold order, lsq.py:141
        int bb = mm / g.L, ll = mm % g.L, oy = ll / g.OW, ox = ll % g.OW;
        int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
          code = xcodes[(((int64_t)bb * g.Cin + ci) * g.H + iy) * g.W + ix];
      }
      for (int j = 0; j < g.NSA; ++j)  // LSB-first digits, slicing_act lsq.py:466-480
        adig[((size_t)j * kPix + pix) * KS + kk] = (uint8_t)((code >> (g.abs_ * j)) & g.amask);
    }
    for (int idx = tid; idx < kCh * rows4; idx += kPix * kCh) {
      int cl = idx / rows4, kk = idx % rows4;
      int code = 0;
      if (kk < rows && c0 + cl < g.Cout) code = wcodes[(int64_t)(c0 + cl) * g.F + lo + kk];
      int mag = code < 0 ? -code : code;
      for (int k = 0; k < g.NSW; ++k) {  // slicing_weights_signed lsq.py:438-464
        int d = (mag >> (g.wbs * k)) & g.wmask;
        wdig[((size_t)k * kCh + cl) * KS + kk] = (int8_t)(code < 0 ? -d : d);
      }
    }
    __syncthreads();

    uint32_t st[kMaxStateWords];
#pragma unroll
    for (int w = 0; w < kMaxStateWords; ++w) st[w] = 0u;

    for (int k = 0; k < g.NSW; ++k) {
      const int32_t *w32 = reinterpret_cast<const int32_t *>(wdig + ((size_t)k * kCh + ty) * KS);
      for (int j = 0; j < g.NSA; ++j) {
        const uint32_t *a32 = reinterpret_cast<const uint32_t *>(adig + ((size_t)j * kPix + tx) * KS);
        int p = 0;
        for (int t = 0; t < rows4 / 4; ++t) p = dp4a_u8s8(a32[t], w32[t], p);
        const int q = k * g.NSA + j;
        const int64_t e = ((int64_t)i * g.pairs + q) * g.Cout + c;
        if (MODE == SIMT_PSUMS) {
          if (valid)
            psums[(((((int64_t)b * g.NX + i) * g.NSW + k) * g.NSA + j) * g.L + l) * g.Cout + c] = p;
        } else if (MODE == SIMT_ABS_SUMS) {
          // lanes of a warp are 32 pixels of the same (crossbar, pair, channel)
          int v = valid ? (p < 0 ? -p : p) : 0;
          int tot = __reduce_add_sync(0xffffffffu, v);
          if (tx == 0 && c < g.Cout) atomicAdd(&sums[e], (unsigned long long)tot);
        } else if (valid) {
          const int4 te = __ldg(&table[e]);
          const float amp = __int_as_float(te.z);
          int clip;
          if (g.adc_mode == CIMQ_ADC_MULTIBIT) {
            float ph = psum_as_stored(p);
            float cf = fminf(fmaxf(ph, (float)g.qn), (float)g.qp);  // lsq.py:228-229
            // lsq.py:310-311: ps >= Qp + 1e-5 or ps <= Qn - 1e-5 in fp32 (the bound itself clips where 1e-5 rounds away)
            clip = (ph >= (float)g.clip_hi) || (ph <= (float)g.clip_lo);
            acc += __fmul_rn(__fmul_rn(cf, sw), sa) * amp;           // lsq.py:230, 233
            if (clip) { int bit = state_pair(g, k, j); st[bit >> 5] |= 1u << (bit & 31); }
          } else {
            int pos = p >= te.x, neg = p <= -te.x;                   // ternary / sign code
            clip = (p >= te.y) || (p <= -te.y);
            if (MODE == SIMT_FORWARD_STOCH) {
              // stochastic near-ADC-less read-out (lsq.py:205-220): two sigmoids of sharpness 0.01 around
              // +-alpha/2, 50 Bernoulli draws each, code = clamp(round(n1/50 + n2/50 - 1), -1, 1).  Only the
              // output is stochastic: the backward uses the deterministic state below (lsq.py:310-332).
              const float a = __ldg(&alpha_q[e]);
              const float v = __fmul_rn(__fmul_rn(psum_as_stored(p), sw), sa);               // lsq.py:169, 195
              const float s1 = 1.0f / (1.0f + expf(-__fdiv_rn(v - 0.5f * a, 0.01f)));
              const float s2 = 1.0f / (1.0f + expf(-__fdiv_rn(v + 0.5f * a, 0.01f)));
              const unsigned long long elem =
                  (((((unsigned long long)b * g.NX + i) * g.NSW + k) * g.NSA + j) * g.L + l) * g.Cout + c;
              const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
              const int n12 = bernoulli_count(s1, 50, elem, 0u, key) + bernoulli_count(s2, 50, elem, 1u, key);
              // round half to even of (n12 - 50) / 50: +1 above 0.5, -1 below -0.5
              acc += n12 > 75 ? amp : (n12 < 25 ? -amp : 0.0f);
            } else {
              acc += pos ? amp : (neg ? -amp : 0.0f);
            }
            const int sq = state_pair(g, k, j);
            if (pos) { int bit = sq; st[bit >> 5] |= 1u << (bit & 31); }
            if (neg) { int bit = g.pairs + sq; st[bit >> 5] |= 1u << (bit & 31); }
            if (clip) { int bit = 2 * g.pairs + sq; st[bit >> 5] |= 1u << (bit & 31); }
          }
        }
      }
    }
    if ((MODE == SIMT_FORWARD || MODE == SIMT_FORWARD_STOCH) && valid && state != nullptr) {
      for (int w = 0; w < g.state_words; ++w)
        state[(((int64_t)i * g.Cout + c) * g.state_words + w) * g.M + m] = st[w];
    }
  }
  if ((MODE == SIMT_FORWARD || MODE == SIMT_FORWARD_STOCH) && valid) out[((int64_t)b * g.Cout + c) * g.L + l] = acc;
}

}  // namespace

int launch_conv_simt(const Geo &g, int mode, const uint8_t *xcodes, const int8_t *wcodes, const void *table,
                     const float *s, const int8_t *, float *out, uint32_t *state, int32_t *psums,
                     unsigned long long *sums, cudaStream_t st, const float *alpha_q, unsigned long long seed) {
  CIMQ_REQUIRE(g.state_words <= kMaxStateWords, "too many slice pairs (%d) for the ADC state", g.pairs);
  int rowsmax = g.xbar < g.F ? g.xbar : g.F;
  int ks4 = (rowsmax + 3) / 4;
  if ((ks4 & 1) == 0) ks4 += 1;  // odd word stride: conflict-free row reads
  int KS = ks4 * 4;
  size_t smem = (size_t)(g.NSA * kPix + g.NSW * kCh) * KS;
  CIMQ_REQUIRE(smem <= 200 * 1024, "crossbar depth %d needs %zu bytes of shared memory", g.xbar, smem);
  dim3 block(kPix, kCh), grid((g.M + kPix - 1) / kPix, (g.Cout + kCh - 1) / kCh);
  const int4 *tab = reinterpret_cast<const int4 *>(table);
#define CIMQ_LAUNCH_SIMT(MODE)                                                                              \
  do {                                                                                                      \
    CIMQ_CUDA_OK(cudaFuncSetAttribute(conv_simt_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      (int)smem));                                                          \
    conv_simt_kernel<MODE><<<grid, block, smem, st>>>(g, KS, xcodes, wcodes, tab, s, out, state, psums, sums,  \
                                                      alpha_q, seed);                                           \
  } while (0)
  if (mode == SIMT_FORWARD) {
    CIMQ_REQUIRE(table != nullptr && s != nullptr && out != nullptr, "conv_forward: NULL argument");
    CIMQ_LAUNCH_SIMT(SIMT_FORWARD);
  } else if (mode == SIMT_FORWARD_STOCH) {
    CIMQ_REQUIRE(table != nullptr && s != nullptr && out != nullptr && alpha_q != nullptr,
                 "conv_forward_stochastic: NULL argument");
    CIMQ_REQUIRE(g.adc_mode == CIMQ_ADC_TERNARY, "the stochastic read-out exists for adcbits 1.5 only (lsq.py:203-220)");
    CIMQ_LAUNCH_SIMT(SIMT_FORWARD_STOCH);
  } else if (mode == SIMT_PSUMS) {
    CIMQ_REQUIRE(psums != nullptr, "conv_psums: NULL output");
    CIMQ_LAUNCH_SIMT(SIMT_PSUMS);
  } else {
    CIMQ_REQUIRE(sums != nullptr, "conv_psum_abs_sums: NULL output");
    CIMQ_LAUNCH_SIMT(SIMT_ABS_SUMS);
  }
#undef CIMQ_LAUNCH_SIMT
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cimq
