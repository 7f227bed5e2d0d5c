// tcgen05 backward kernels of the CiM convolution (get_cim_output_signed.backward, lsq.py:244-386).
//
// Both GEMMs multiply a real-valued operand (grad_out masked by the STE clip bits) with an exact small
// integer operand (weight or activation digit planes).  fp32-level accuracy (1e-5 tolerance) on bf16
// tensor cores is obtained by splitting only the real operand into three bf16 terms
// (v = hi + mid + lo exactly covers the 24-bit fp32 mantissa); the integer operand is exact in bf16 and
// accumulation is fp32 in TMEM.  The masked operand is built on the fly in shared memory from grad_out
// and the 1-bit-per-partial-sum clip state the forward stored -- the 6-D tensors of the reference never
// exist.
//
//   dgrad  D[128 pixels x Nf crossbar rows] = sum_k sum_split A'_{k,split}[128 x Cout] * W_k[Nf x Cout]^T
//          A'_k[m,co] = go[m,co] * sum_j pass[m,i,k,j,co] * mask[k,j] * 2^(-abs*j)
//          -> gxu[b][f][l] (then col2im)
//   wgrad  D_i[128 crossbar rows x Cout] += sum_j sum_split X_j[128 x 128 pixels] * G'_{j,split}[Cout x 128 pixels]^T
//          G'_j[m,co] = go[m,co] * sum_k pass[m,i,k,j,co] * mask[k,j] * 2^(-wbs*k),  X_j = activation digit plane
//          accumulated over all pixel tiles of the CTA in TMEM, one partial [F x Cout] per CTA
//
// Warp roles (16 warps): 0-7 producers, 8-11 epilogue (one per TMEM lane quarter), 12 MMA issuer.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "cim_tc_layout.cuh"
#include "cim_v2.cuh"
#include "tc_ptx.cuh"

namespace cimq {

namespace {

using namespace ptx;

constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kEpilogueWarp0 = 8;
constexpr int kEpilogueWarps = 4;
constexpr int kMmaWarp = 12;
constexpr int kThreads = 512;
constexpr int kMaxStages = 4;
constexpr int kMaxPairs = 64;
constexpr size_t kSmemBudget = 227 * 1024 - 1024;
constexpr size_t kBarrierBytes = 3072;  // barriers, tmem slot, slice-weight table, pixel / row tables, LUT
// dgrad register split per warpgroup (setmaxnreg): 2 x 168 + 104 + 40 = 480 <= 512
constexpr int kDgRegsProducer = 168, kDgRegsEpilogue = 104, kDgRegsMma = 40;
// wgrad: 3 x 152 + 56 = 512
constexpr int kWgRegsProducer = 152, kWgRegsMma = 56;
constexpr int kWgProducerThreads = 384;  // wgrad: warps 0-11 build operands (8-11 also run the final epilogue)
constexpr int kNoRow = -2147483647 - 1;
#ifndef CIMQ_TIMERS
#define CIMQ_TIMERS 0
#endif
constexpr bool kTimers = CIMQ_TIMERS != 0;
#define CIMQ_TB() (dbg ? clock64() : 0ll)

struct BwdParams {
  Geo g;
  int Kc;           // Cout
  int Nf;           // dgrad N: padded crossbar rows
  int mtiles, stages;
  uint32_t a_bytes, b_bytes, stage_bytes, tmem_cols;
  int nxg, chunk0;  // wgrad: crossbars handled by this launch's blockIdx.y group
  // wgrad staged activation rows (same geometry as the forward producer, stride 1 only)
  int fastx, ow_log2, rpt, pitch_log2, col0;
  int pitch, slot_bytes;  // wgrad staging: bytes per staged row / per channel slot (rk rows + bank padding), see tune_row_layout
  int rk, prow;      // wgrad staging: rows per channel slot, staged row of (output row o, tap row ky) = o*prow + ky
  int xshared;       // wgrad staging: 1 = a tile is consecutive rows of ONE image and its output rows share input rows
                     // (prow == 1); 0 = every output row stages its own K rows (prow == K; also 1 for 1x1 kernels)
  int async_rows;    // wgrad staging: rows arrive by cp.async copies issued one chunk ahead; value = piece size (16 or 8), 0 = off
  uint32_t raw_bytes;
  int gfast;         // wgrad: every 8-pixel group is an aligned run of one image row (or past the end)
  int co0;           // first output channel of this launch (layers with Cout > 128 run as blocks of 128 channels)
  int sdiv;          // virtual 128-row chunks per crossbar: the ADC state of virtual chunk i is that of crossbar i / sdiv
  long long *debug;  // per-role cycle counters (builds with TIMERS=1 only)
  int cached;   // dgrad: 1 = producers keep grad_out of the tile and the state words of the chunk in registers
  uint32_t pw_off;  // dgrad: byte offset of the pass-weight table inside the raw region
  int dbg;   // development only (env CIMQ_V2_DBG, wgrad v2): 1 = no MMAs, 2 = no X stores, 4 = no G' stores, 8 = no X gather
  int fold;  // dgrad: 1 = the epilogue folds (col2im) straight into grad_x with fp32 reductions; 0 = writes gxu[b][f][l]
  int v2;    // 1 = `state` holds the v2 byte planes (cim_v2.cuh): D [NX][M][Cout] for dgrad, W for wgrad
  const uint8_t *state2, *state2w;
  uint32_t xt_off;     // wgrad: byte offset of the per-(chunk, crossbar row) index tables inside the dynamic shared memory
  const uint32_t *chmax;  // v2 wgrad: bit pattern of max |grad_out| per output channel (launch_go_scales)
  const float *go;
  const uint32_t *state;
  const uint8_t *xcodes;
  const uint8_t *wtb;  // dgrad B tiles (bf16)
  const float *s;
  const int8_t *mask;
  float *out;          // dgrad: gxuT [F][M]; wgrad: partial [ctas][F][Cout]
};

// two fp32 -> packed bf16x2 (first argument in the upper half)
__device__ __forceinline__ uint32_t cvt_bf16x2(float upper, float lower) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
  return d;
}
// (v0, v1) -> three packed bf16x2 words with hi + mid + lo == v exactly (24-bit mantissa, up to underflow);
// v0 lands in the low half (lower K index)
__device__ __forceinline__ void split3x2(float v0, float v1, uint32_t &hi, uint32_t &mid, uint32_t &lo) {
  hi = cvt_bf16x2(v1, v0);
  const float r0 = v0 - __uint_as_float(hi << 16), r1 = v1 - __uint_as_float(hi & 0xffff0000u);
  mid = cvt_bf16x2(r1, r0);
  const float q0 = r0 - __uint_as_float(mid << 16), q1 = r1 - __uint_as_float(mid & 0xffff0000u);
  lo = cvt_bf16x2(q1, q0);
}

struct Carve {
  uint8_t *stage_base;
  uint32_t full0, empty0, tfull0, tempty0;
  uint32_t *tmem_slot;
  float *wtab;    // [pairs] slice weights
  int4 *pixtab;   // [2][16] wgrad: {image, output row, first output column, fast-path flag} per 8-pixel group
  int *rowoff;    // [2][128] wgrad: global offset of staged row (output row, ky) or kNoRow
  float *lut;     // [NSA][16] wgrad: pass weight by clip-bit pattern of the weight slices
  uint8_t *raw;   // wgrad: 2 staging buffers of raw_bytes (after the stages)
};

__device__ __forceinline__ Carve carve_smem(uint8_t *smem_raw, int stages, uint32_t stage_bytes) {
  Carve c;
  c.stage_base = smem_raw;
  uint8_t *aux = smem_raw + (size_t)stages * stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(aux);
  c.full0 = smem_u32(bars);
  c.empty0 = c.full0 + 8 * kMaxStages;
  c.tfull0 = c.empty0 + 8 * kMaxStages;
  c.tempty0 = c.tfull0 + 16;
  c.tmem_slot = reinterpret_cast<uint32_t *>(aux + 112);
  c.wtab = reinterpret_cast<float *>(aux + 128);     // 64 floats
  c.pixtab = reinterpret_cast<int4 *>(aux + 512);    // 2 x 16 x 16 bytes
  c.rowoff = reinterpret_cast<int *>(aux + 1024);    // 2 x 128 ints
  c.lut = reinterpret_cast<float *>(aux + 2048);     // 8 x 16 floats
  c.raw = aux + kBarrierBytes;
  return c;
}

// Compile-time description of where the STE clip bits live in the ADC state words.
template <int NSW, int NSA, bool TERN>
struct ClipBits {
  static constexpr int PAIRS = NSW * NSA;
  static constexpr int SWORDS = ((TERN ? 3 : 1) * PAIRS + 31) / 32;  // state words per (crossbar, channel, pixel)
  static constexpr int CLIP0 = TERN ? 2 * PAIRS : 0;                 // first clip bit
  static constexpr int CW0 = CLIP0 >> 5;                             // first word holding clip bits
  static constexpr int CWN = ((CLIP0 + PAIRS - 1) >> 5) - CW0 + 1;   // words holding clip bits (1 or 2)
  static constexpr int CB = CLIP0 - 32 * CW0;                        // clip bit 0 inside word CW0
  static_assert(CWN <= 2, "clip bits span more than two state words");
};
// sum over t < N of (clip bit (start + t*stride) set ? 0 : w[t]); the clip bits of one (crossbar, channel,
// pixel) live in sw[0..CWN)
template <int N, int CWN>
__device__ __forceinline__ float pass_weight(const uint32_t (&sw)[CWN], int start, int stride, const float (&w)[N]) {
  float acc = 0.0f;
  if constexpr (CWN == 1) {
    const uint32_t win = sw[0] >> start;  // all clip bits are in one word: one shift, constant bit tests
#pragma unroll
    for (int t = 0; t < N; ++t) acc += ((win >> (t * stride)) & 1u) ? 0.0f : w[t];
  } else {
#pragma unroll
    for (int t = 0; t < N; ++t) {
      const int idx = start + t * stride;
      const uint32_t bit = (idx >= 32 ? (sw[1] >> (idx - 32)) : (sw[0] >> idx)) & 1u;
      acc += bit ? 0.0f : w[t];
    }
  }
  return acc;
}

// =====================================================================================================
// dgrad
// =====================================================================================================
template <int NSW, int NSA, bool TERN>
__global__ void __launch_bounds__(kThreads, 1) bwd_input_tc_kernel(const BwdParams P) {
  using CBits = ClipBits<NSW, NSA, TERN>;
  const Geo &g = P.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const Carve cv = carve_smem(smem_raw, P.stages, P.stage_bytes);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // (provably warp-uniform)
  const int Kc = P.Kc, Nf = P.Nf;
  const uint32_t sbo = (uint32_t)Kc * 16u;  // 8 rows x Kc bf16

  if (threadIdx.x == 0) {
    for (int sidx = 0; sidx < P.stages; ++sidx) {
      mbar_init(cv.full0 + 8 * sidx, kProducerThreads + 1);
      mbar_init(cv.empty0 + 8 * sidx, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(cv.tfull0 + 8 * b, 1);
      mbar_init(cv.tempty0 + 8 * b, kEpilogueWarps);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < g.pairs) {  // mask[k][j] * 2^(-abs*j)  (lsq.py:306, 373-374)
    const int j = threadIdx.x % g.NSA;
    cv.wtab[threadIdx.x] = (float)P.mask[threadIdx.x] * exp2f(-(float)(g.abs_ * j));
  }
  // pass weight of weight slice k by table lookup: index = the NSA clip bits of (k, j = 0..NSA-1), which sit NSW
  // apart in the state word (bit of (k, j) = CB + j*NSW + k), left in place (sparse index, PWN entries per slice)
  constexpr int PWBITS = (NSA - 1) * NSW + 1, PWN = 1 << (PWBITS <= 7 ? PWBITS : 0);
  constexpr bool kPwLut = PWBITS <= 7;
  constexpr uint32_t PWMASK = []() {
    uint32_t mk = 0;
    for (int j = 0; j < NSA; ++j)
      if (j * NSW < 32) mk |= 1u << (j * NSW);
    return mk;
  }();
  if (kPwLut && P.cached) {
    float *pwl = reinterpret_cast<float *>(cv.raw + P.pw_off);
    for (int t = threadIdx.x; t < NSW * PWN; t += kThreads) {
      const int k = t / PWN, pat = t % PWN;
      float acc = 0.0f;
      for (int j = 0; j < NSA; ++j)  // same terms, same order as pass_weight()
        acc += ((pat >> (j * NSW)) & 1) ? 0.0f : (float)P.mask[k * NSA + j] * exp2f(-(float)(g.abs_ * j));
      pwl[t] = acc;
    }
  }
  if (P.fold) {  // unfold row f = (ci, ky, kx) (nn.Unfold order, lsq.py:141) -> offset inside the image, tap index
    int *ftab = reinterpret_cast<int *>(cv.raw);
    for (int f = threadIdx.x; f < g.F; f += kThreads) {
      const int ci = f / g.KK, tap = f % g.KK, ky = tap / g.K, kx = tap % g.K;
      ftab[f] = (((ci * g.H + ky) * g.W + kx) << 7) | (kx << 5) | tap;
    }
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(cv.tmem_slot), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *cv.tmem_slot;
  const int rows_full = g.xbar < g.F ? g.xbar : g.F;

  if (warp < kProducerWarps) {
    reg_alloc<kDgRegsProducer>();  // warpgroups 0-1; taken from the epilogue / MMA warpgroups below
    // ------------------------------------------------------------------ producers
    // thread = (pixel row r, channel half h); per stage (crossbar i, weight slice k) it builds
    // A'[r, co] = go[r, co] * sum_j pass * wx[k][j] for its Kc/2 channels in groups of 8 (one 16-byte
    // store per bf16 term).  Loads of the next group are issued before the current one is processed.
    const int r = threadIdx.x & 127;
    const int h = threadIdx.x >> 7;
    const int cpt = Kc >> 1;          // channels per thread
    const int G = cpt >> 3;           // groups of 8 channels
    uint32_t it = 0;
    bool done = false;
    if constexpr (CBits::CWN == 1) {
      if (P.cached && !done) {
        // ---- register-resident operands (Cout <= 64): grad_out of the tile (reused by all NX*NSW stages) and the
        // state words of the chunk (reused by its NSW stages) live in registers; each group of eight is refilled
        // for the next chunk / tile right after its last use, so the loads fly during the remaining groups and the
        // following stage.  Rows past the last pixel read pixel 0: their A' rows only feed accumulator rows that
        // the epilogue never stores.
        const float *pwl = reinterpret_cast<const float *>(cv.raw + P.pw_off);
        const size_t sstride = (size_t)CBits::SWORDS * g.M;
        float gvr[32];
        uint32_t swr[32];
        const bool dbg = kTimers && P.debug != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
        long long d_wait = 0, d_prod = 0;
        auto tile_ptrs = [&](int mt, const float *&gop, const uint32_t *&stp) {
          const int64_t m = (int64_t)mt * kTcTileM + r;
          const bool live = m < g.M;
          const int b = live ? (int)(m / g.L) : 0, l = live ? (int)(m % g.L) : 0;
          gop = P.go + ((int64_t)b * g.Cout + P.co0 + h * cpt) * g.L + l;
          stp = P.state + (int64_t)CBits::CW0 * g.M + (live ? m : 0) + (size_t)(P.co0 + h * cpt) * sstride;
        };
        const float *gop = nullptr, *gop_n = nullptr;
        const uint32_t *stp = nullptr, *stp_n = nullptr;
        if ((int)blockIdx.x < P.mtiles) {
          tile_ptrs(blockIdx.x, gop, stp);
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c < cpt) {
              gvr[c] = __ldg(gop + (size_t)c * g.L);
              swr[c] = __ldg(stp + (size_t)c * sstride);
            }
        }
        for (int mt = blockIdx.x; mt < P.mtiles; mt += gridDim.x) {
          const int nmt = mt + gridDim.x;
          const bool more_tiles = nmt < P.mtiles;
          if (more_tiles) tile_ptrs(nmt, gop_n, stp_n);
          for (int i = 0; i < g.NX; ++i) {
            for (int k = 0; k < NSW; ++k, ++it) {
              const int sidx = it % P.stages;
              const uint32_t use = it / P.stages;
              const long long t0 = CIMQ_TB();
              mbar_wait(cv.empty0 + 8 * sidx, (use & 1) ^ 1);
              const long long t1 = CIMQ_TB();
              d_wait += t1 - t0;
              uint8_t *st_ptr = cv.stage_base + (size_t)sidx * P.stage_bytes;
              if (threadIdx.x == 0) {
                mbar_arrive_expect_tx(cv.full0 + 8 * sidx, P.b_bytes);
                bulk_copy_g2s(smem_u32(st_ptr + 3 * (size_t)P.a_bytes), P.wtb + (size_t)(i * NSW + k) * P.b_bytes,
                              P.b_bytes, cv.full0 + 8 * sidx);
              }
              float wx[NSA];
#pragma unroll
              for (int j = 0; j < NSA; ++j) wx[j] = cv.wtab[k * NSA + j];
              const uint32_t pwk = smem_u32(pwl + k * PWN);
              const bool last_k = k + 1 == NSW;
              // after the last slice of a chunk the state registers take the next chunk (or the next tile's first)
              const uint32_t *sp_next = nullptr;
              if (last_k) {
                if (i + 1 < g.NX) sp_next = stp + (size_t)((i + 1) / P.sdiv) * g.Cout * sstride;
                else if (more_tiles) sp_next = stp_n;
              }
              const float *gp_next = (last_k && i + 1 == g.NX && more_tiles) ? gop_n : nullptr;
#pragma unroll
              for (int cgi = 0; cgi < 4; ++cgi) {
                if (cgi < G) {
                  float v[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const uint32_t sw1[1] = {swr[8 * cgi + e]};
                    float pw;
                    if constexpr (kPwLut) pw = lds_const_f32(pwk + 4u * ((sw1[0] >> (CBits::CB + k)) & PWMASK));
                    else pw = pass_weight<NSA, 1>(sw1, CBits::CB + k, NSW, wx);
                    v[e] = gvr[8 * cgi + e] * pw;
                  }
                  if (sp_next != nullptr) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) swr[8 * cgi + e] = __ldg(sp_next + (size_t)(8 * cgi + e) * sstride);
                  }
                  if (gp_next != nullptr) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) gvr[8 * cgi + e] = __ldg(gp_next + (size_t)(8 * cgi + e) * g.L);
                  }
                  uint32_t hi[4], mid[4], lo[4];
#pragma unroll
                  for (int e2 = 0; e2 < 4; ++e2) split3x2(v[2 * e2], v[2 * e2 + 1], hi[e2], mid[e2], lo[e2]);
                  const uint32_t off = tc_tile_offset16(r, h * cpt + cgi * 8, kTcLBO, sbo);
                  *reinterpret_cast<uint4 *>(st_ptr + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                  *reinterpret_cast<uint4 *>(st_ptr + P.a_bytes + off) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
                  *reinterpret_cast<uint4 *>(st_ptr + 2 * (size_t)P.a_bytes + off) =
                      make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
              }
              fence_proxy_async();
              mbar_arrive(cv.full0 + 8 * sidx);
              d_prod += CIMQ_TB() - t1;
            }
          }
          gop = gop_n;
          stp = stp_n;
        }
        if (dbg) { P.debug[0] = d_wait; P.debug[1] = d_prod; }
        done = true;
      }
    }
    for (int mt = blockIdx.x; mt < P.mtiles && !done; mt += gridDim.x) {
      const int64_t m = (int64_t)mt * kTcTileM + r;
      const bool live = m < g.M;
      const int b = live ? (int)(m / g.L) : 0, l = live ? (int)(m % g.L) : 0;
      const float *gop = P.go + ((int64_t)b * g.Cout + P.co0 + h * cpt) * g.L + l;
      // rows past the last pixel read pixel 0 (valid addresses, no predicates on the loads) and are zeroed
      // through the slice weights below
      const uint32_t *stp = P.state + (int64_t)CBits::CW0 * g.M + (live ? m : 0);
      const float livef = live ? 1.0f : 0.0f;
      const size_t sstride = (size_t)CBits::SWORDS * g.M;  // state words between consecutive channels
      uint32_t sw_n[8][CBits::CWN];
      float gv_n[8];
      auto prefetch = [&](int i, int cg) {
        // one 64-bit multiply per group of eight channels, then pointer increments
        const uint32_t *sp = stp + (size_t)((i / P.sdiv) * g.Cout + P.co0 + h * cpt + cg) * sstride;
        const float *gp = gop + (size_t)cg * g.L;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
#pragma unroll
          for (int w = 0; w < CBits::CWN; ++w) sw_n[e][w] = __ldg(sp + (size_t)w * g.M);
          gv_n[e] = __ldg(gp);
          sp += sstride;
          gp += g.L;
        }
      };
      prefetch(0, 0);
      for (int i = 0; i < g.NX; ++i) {
        for (int k = 0; k < NSW; ++k, ++it) {
          const int sidx = it % P.stages;
          const uint32_t use = it / P.stages;
          mbar_wait(cv.empty0 + 8 * sidx, (use & 1) ^ 1);
          uint8_t *st_ptr = cv.stage_base + (size_t)sidx * P.stage_bytes;
          if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(cv.full0 + 8 * sidx, P.b_bytes);
            bulk_copy_g2s(smem_u32(st_ptr + 3 * (size_t)P.a_bytes), P.wtb + (size_t)(i * NSW + k) * P.b_bytes,
                          P.b_bytes, cv.full0 + 8 * sidx);
          }
          float wx[NSA];
#pragma unroll
          for (int j = 0; j < NSA; ++j) wx[j] = cv.wtab[k * NSA + j] * livef;
          const int start = CBits::CB + k;  // clip bit of (k, j) = CB + j*NSW + k
          for (int cgi = 0; cgi < G; ++cgi) {
            uint32_t sw[8][CBits::CWN];
            float gv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              gv[e] = gv_n[e];
#pragma unroll
              for (int w = 0; w < CBits::CWN; ++w) sw[e][w] = sw_n[e][w];
            }
            {  // next group in (i, k, cgi) order; state words do not depend on k (cache hits)
              int ni = i, nc = cgi + 1;
              if (nc == G) { nc = 0; if (k + 1 == NSW) ni = i + 1; }
              if (ni < g.NX) prefetch(ni, nc * 8);
            }
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              v[e] = gv[e] * pass_weight<NSA, CBits::CWN>(sw[e], start, NSW, wx);
            }
            uint32_t hi[4], mid[4], lo[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) split3x2(v[2 * e2], v[2 * e2 + 1], hi[e2], mid[e2], lo[e2]);
            const uint32_t off = tc_tile_offset16(r, h * cpt + cgi * 8, kTcLBO, sbo);
            *reinterpret_cast<uint4 *>(st_ptr + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(st_ptr + P.a_bytes + off) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
            *reinterpret_cast<uint4 *>(st_ptr + 2 * (size_t)P.a_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async();
          mbar_arrive(cv.full0 + 8 * sidx);
        }
      }
    }
  } else if (warp >= kMmaWarp) {
    reg_dealloc<kDgRegsMma>();  // warpgroup 3: the MMA issuer and three idle warps
    if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (whole warp walks the loops, lane 0 issues)
    {
      const uint32_t idesc = idesc_bf16_f32(kTcTileM, Nf);
      const int ksteps = Kc >> 4;
      uint32_t acc_it = 0;
      int m_sidx = 0;  // pipeline stage and its use count as running counters
      uint32_t m_use = 0;
      const bool dbg = kTimers && P.debug != nullptr && blockIdx.x == 0 && lane == 0;
      long long d_full = 0, d_tempty = 0;
      const long long t_begin = CIMQ_TB();
      for (int mt = blockIdx.x; mt < P.mtiles; mt += gridDim.x) {
        for (int i = 0; i < g.NX; ++i, ++acc_it) {
          const uint32_t buf = acc_it & 1, buse = acc_it >> 1;
          const long long ta = CIMQ_TB();
          mbar_wait<400>(cv.tempty0 + 8 * buf, (buse & 1) ^ 1);
          d_tempty += CIMQ_TB() - ta;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * Nf;
          for (int k = 0; k < NSW; ++k) {
            const int sidx = m_sidx;
            const uint32_t use = m_use;
            if (++m_sidx == P.stages) { m_sidx = 0; ++m_use; }
            const long long tb = CIMQ_TB();
            mbar_wait<400>(cv.full0 + 8 * sidx, use & 1);
            d_full += CIMQ_TB() - tb;
            tc_fence_after();
            const uint32_t a0 = smem_u32(cv.stage_base + (size_t)sidx * P.stage_bytes);
            const uint32_t b0 = a0 + 3 * P.a_bytes;
            if (lane == 0) {
              for (int sp = 0; sp < 3; ++sp)
                for (int ks = 0; ks < ksteps; ++ks) {
                  const uint64_t adesc = make_smem_desc(a0 + sp * P.a_bytes + ks * 2 * kTcLBO, kTcLBO, sbo);
                  const uint64_t bdesc = make_smem_desc(b0 + ks * 2 * kTcLBO, kTcLBO, sbo);
                  umma_f16(d_tmem, adesc, bdesc, idesc, (k | sp | ks) != 0 ? 1u : 0u);
                }
              umma_commit(cv.empty0 + 8 * sidx);
              if (k == NSW - 1) umma_commit(cv.tfull0 + 8 * buf);
            }
            __syncwarp();
          }
        }
      }
      if (dbg) { P.debug[4] = d_full; P.debug[5] = d_tempty; P.debug[6] = clock64() - t_begin; }
    }
    }
  } else if (warp >= kEpilogueWarp0 && warp < kEpilogueWarp0 + kEpilogueWarps) {
    reg_dealloc<kDgRegsEpilogue>();  // warpgroup 2
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    // w_sl * s_w (lsq.py:252), mean over act slices (lsq.py:376)
    const float scale = P.s[1] / (float)NSA;
    const int *ftab = reinterpret_cast<const int *>(cv.raw);  // fold: per unfold row {offset in the image << 7 | kx << 5 | tap}
    const bool dbg = kTimers && P.debug != nullptr && blockIdx.x == 0 && warp == kEpilogueWarp0 && lane == 0;
    long long d_tfull = 0, d_comp = 0;
    uint32_t acc_it = 0;
    for (int mt = blockIdx.x; mt < P.mtiles; mt += gridDim.x) {
      const int64_t m = (int64_t)mt * kTcTileM + r;
      const int eb = m < g.M ? (int)(m / g.L) : 0, el = m < g.M ? (int)(m % g.L) : 0;
      // fold: taps of this pixel that land inside the image, and the address of tap (0, 0) of channel 0
      uint32_t vm = 0;
      float *gxp = P.out;
      if (P.fold && m < g.M) {
        const int oy = el / g.OW, ox = el % g.OW;
        const int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
        for (int ky = 0; ky < g.K; ++ky)
          for (int kx = 0; kx < g.K; ++kx)
            if (iy0 + ky >= 0 && iy0 + ky < g.H && ix0 + kx >= 0 && ix0 + kx < g.W) vm |= 1u << (ky * g.K + kx);
        gxp = P.out + ((int64_t)eb * g.Cin * g.H + iy0) * g.W + ix0;
      }
      for (int i = 0; i < g.NX; ++i, ++acc_it) {
        const uint32_t buf = acc_it & 1, buse = acc_it >> 1;
        const int lo = i * g.xbar;
        const int rows = min(rows_full, g.F - lo);
        const long long te0 = CIMQ_TB();
        mbar_wait(cv.tfull0 + 8 * buf, buse & 1);
        const long long te1 = CIMQ_TB();
        d_tfull += te1 - te0;
        tc_fence_after();
        for (int c0 = 0; c0 < rows; c0 += 32) {
          int v[32];
          tmem_ld<32>(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * Nf + c0, v);
          tmem_ld_wait();
          if (P.fold) {
            // nn.Fold (lsq.py:378-383) fused: every unfolded gradient is added to its input pixel.  Lanes are
            // consecutive output pixels, so one warp-wide reduction covers consecutive addresses; grad_x
            // (zeroed by the caller) stays in L2 while it is accumulated.
            // The table entries are read first, all at once: the reductions below order the compiler's memory
            // operations, and a table read between two of them would expose its latency 32 times per batch.
            int te[32];
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) te[cc] = ftab[min(lo + c0 + cc, g.F - 1)];
#pragma unroll
            for (int cc = 0; cc < 32; ++cc)
              if (c0 + cc < rows) {
                const int e = te[cc];
                if ((vm >> (e & 31)) & 1u) atomicAdd(gxp + (e >> 7), __int_as_float(v[cc]) * scale);
              }
          } else if (m < g.M) {  // gxu[b][f][l]: image-major so that col2im's reads per output stay within one image
            float *dst = P.out + ((int64_t)eb * g.F + lo + c0) * g.L + el;
#pragma unroll
            for (int cc = 0; cc < 32; ++cc)
              if (c0 + cc < rows) {
                // later channel blocks (Cout > 128) add to what the first one wrote: same thread, launch order
                const float t = __int_as_float(v[cc]) * scale;
                dst[(int64_t)cc * g.L] = P.co0 > 0 ? dst[(int64_t)cc * g.L] + t : t;
              }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(cv.tempty0 + 8 * buf);
        d_comp += CIMQ_TB() - te1;
      }
    }
    if (dbg) { P.debug[8] = d_tfull; P.debug[9] = d_comp; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// =====================================================================================================
// wgrad
// =====================================================================================================
// The 8 activation codes (one per byte) of crossbar row (ci, ky, kx) at the 8 consecutive output pixels starting at
// m8, for geometries the staged-row path does not cover.  pt = {image, output row, first output column, kind} of
// the group: kind 1 = one aligned run of an image row, 0 = straddles rows / images / the end, -1 = past the end.
struct GatherGeo {  // the few geometry fields the gather needs, by value (a reference would pin Geo to the stack)
  int stride, pad, H, W, Cin, L, OW;
  int64_t M;
};
__device__ __noinline__ uint2 gather_codes_generic(GatherGeo g, const uint8_t *__restrict__ xcodes, int4 pt, int ci,
                                                   int ky, int kx, int64_t m8) {
  uint32_t c[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) c[e] = 0u;
  if (pt.w == 1) {
    const int iy = pt.y * g.stride - g.pad + ky;
    if (iy >= 0 && iy < g.H) {
      const uint8_t *row = xcodes + (((int64_t)pt.x * g.Cin + ci) * g.H + iy) * g.W;
      const int ix0 = pt.z * g.stride - g.pad + kx;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ix = ix0 + e * g.stride;
        if (ix >= 0 && ix < g.W) c[e] = __ldg(row + ix);
      }
    }
  } else if (pt.w == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int64_t m = m8 + e;
      if (m < g.M) {
        const int b = (int)(m / g.L), l = (int)(m % g.L), oy = l / g.OW, ox = l % g.OW;
        const int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) c[e] = xcodes[(((int64_t)b * g.Cin + ci) * g.H + iy) * g.W + ix];
      }
    }
  }
  return make_uint2(c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24), c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24));
}

__host__ __device__ inline size_t wg_table_bytes(int nxg) { return (size_t)nxg * (sizeof(int4) + 128 * sizeof(uint32_t)); }  // index tables
constexpr int kWgLBO = 144;  // padded K-stride of the G' tiles: producer lanes run along K (bank-conflict free)

// ---- wgrad X tile (im2col^T digit planes), v2 fast path: item counts and strides are compile-time per thread role so
// that every address is base + immediate and no guards remain in the unrolled loops.
// gather: the 8 activation codes (one per byte) of crossbar row (ci,ky,kx) at pixel groups pg0 + STEP*q, from the staged rows
template <int CNT, int STEP, int XI>
__device__ __forceinline__ void x_gather_fast(const uint8_t *src_row, int pg0, int ow_log2, int row_pitch_shift,
                                              bool frow, uint32_t (&xlo)[XI], uint32_t (&xhi)[XI]) {
  // src_row = staged byte of (output row 0, output column 0) for this crossbar row; output row r is r << row_pitch_shift
  // bytes further, output column c is c bytes further
#pragma unroll
  for (int q = 0; q < CNT; ++q) {
    const int p0 = (pg0 + STEP * q) * 8;
    uint32_t lo8 = 0u, hi8 = 0u;
    if (frow) {
      const uint8_t *src = src_row + ((p0 >> ow_log2) << row_pitch_shift) + (p0 & ((1 << ow_log2) - 1));
      const uint32_t sa = smem_u32(src);
      const uint32_t *al = reinterpret_cast<const uint32_t *>(src - (sa & 3u));
      const uint32_t w0 = al[0], w1 = al[1], w2 = al[2];
      const uint32_t bsh = (sa & 3u) * 8u;
      lo8 = __funnelshift_r(w0, w1, bsh);
      hi8 = __funnelshift_r(w1, w2, bsh);
    }
    xlo[q] = lo8;
    xhi[q] = hi8;
  }
}
// store: digit plane `sh` of the gathered codes as bf16 (value cmul << sh for a set bit), one 16-byte row segment per item
template <int CNT, int STEP, int XI>
__device__ __forceinline__ void x_store_fast(uint8_t *dst, const uint32_t (&xlo)[XI], const uint32_t (&xhi)[XI], int sh,
                                             uint32_t cmul) {
  const uint32_t bit = 0x01010101u << sh;
#pragma unroll
  for (int q = 0; q < CNT; ++q) {
    const uint32_t tl = xlo[q] & bit, th = xhi[q] & bit;  // bytes 0 or 2^sh; cmul = bf16 pattern >> sh
    *reinterpret_cast<uint4 *>(dst + q * STEP * kTcLBO) =
        make_uint4(__byte_perm(tl, 0u, 0x4140) * cmul, __byte_perm(tl, 0u, 0x4342) * cmul,
                   __byte_perm(th, 0u, 0x4140) * cmul, __byte_perm(th, 0u, 0x4342) * cmul);
  }
}

// V2: the pass counts come from plane W of the v2 state (cim_v2.cuh), one byte per (crossbar, pixel, channel), and the
// producer threads split differently: threads [0, 4*Kc) build the G' tiles (item = 4 channels x 8 pixels), the next 128
// or 256 threads the X tile -- see the V2 blocks below.
template <int NSW, int NSA, bool TERN, bool V2>
__global__ void __launch_bounds__(kThreads, 1) bwd_weight_tc_kernel(const BwdParams P) {
  using CBits = ClipBits<NSW, NSA, TERN>;
  constexpr bool kLut = NSW <= 4;  // pass weight of one activation slice by table lookup on its NSW clip bits
  const Geo &g = P.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const Carve cv = carve_smem(smem_raw, P.stages, P.stage_bytes);
  if (kTimers && P.debug != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) P.debug[10] = clock64();  // (absolute: entry)
  // (the warp index through a shuffle: the compiler then knows that the role branches below are warp-uniform and keeps
  // the MMA issuer's descriptors in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int Kc = P.Kc;                            // Cout = UMMA N
  const uint32_t a_sbo = 128u * 16u;              // X tile: 8 rows x 128 pixels bf16, LBO 128
  // G' tile: 16 k-groups of 144 bytes per 8 rows (+ 16 bytes in the v2 variant, whose producer lanes run along the
  // channels: eight consecutive channel quads then fall into eight different 16-byte bank groups)
  const uint32_t b_sbo = 16u * (uint32_t)kWgLBO + (V2 ? 16u : 0u);
  const int i_begin = blockIdx.y * P.nxg;
  const int i_end = min(g.NX, i_begin + P.nxg);

  if (threadIdx.x == 0) {
    for (int sidx = 0; sidx < P.stages; ++sidx) {
      mbar_init(cv.full0 + 8 * sidx, kWgProducerThreads / 32);
      mbar_init(cv.empty0 + 8 * sidx, 1);
    }
    mbar_init(cv.tfull0, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < g.pairs) {  // mask[k][j] * 2^(-wbs*k)  (lsq.py:306, 363-364)
    const int k = threadIdx.x / g.NSA;
    cv.wtab[threadIdx.x] = (float)P.mask[threadIdx.x] * exp2f(-(float)(g.wbs * k));
  }
  if (kLut && threadIdx.x >= 128 && threadIdx.x < 128 + NSA * 16) {
    const int j = (threadIdx.x - 128) >> 4, pat = (threadIdx.x - 128) & 15;
    float acc = 0.0f;
    for (int k = 0; k < NSW; ++k)
      if (!((pat >> k) & 1)) acc += (float)P.mask[k * NSA + j] * exp2f(-(float)(g.wbs * k));
    cv.lut[j * 16 + pat] = acc;
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(cv.tmem_slot), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *cv.tmem_slot;
  const int rows_full = g.xbar < g.F ? g.xbar : g.F;
  const bool has_work = blockIdx.x < P.mtiles;
  if (kTimers && P.debug != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) P.debug[11] = clock64();  // prologue done

  if (warp < 12) {
    // ------------------------------------------------------------------ producers (384 threads)
    reg_alloc<kWgRegsProducer>();  // warpgroups 0-2; the registers come from warpgroup 3 (MMA issuer + idle warps)
    const int tid = threadIdx.x;
    const bool dbg = kTimers && P.debug != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x == (P.dbg >> 8);  // (timed thread: CIMQ_V2_DBG_WG >> 8)
    long long d_stage = 0, d_wait = 0, d_x = 0, d_g = 0, d_bar = 0, d_conv = 0, d_tbar = 0;
    // X tile items (crossbar row fr, 8-pixel group pg = x_pg0 + x_step * q, q < XI) of this thread
    // v2: threads [0, n_g) build G' (one 4-channel x 8-pixel item = 32 elements); with n_g = 256 the G' threads also
    // take three X items each (pixel groups 10..15) and the other 128 threads ten (groups 0..9) -- measured with the
    // per-role instantiation of the tile loop below: 14/1 346 us, 12/2 336, 10/3 330, 8/4 345; with fewer G' threads
    // the 256 threads after them take eight X items each.
    constexpr int XI = V2 ? 14 : 6;
    const int n_g = V2 ? 4 * Kc : 0;
    int fr, x_pg0, x_step, x_cnt;
    if (!V2) { fr = tid & 127; x_pg0 = tid >> 7; x_step = 3; x_cnt = 6; }
    else if (n_g == 256) {
      if (tid >= 256) { fr = tid - 256; x_pg0 = 0; x_step = 1; x_cnt = 10; }
      else { fr = tid & 127; x_pg0 = 10 + 3 * (tid >> 7); x_step = 1; x_cnt = 3; }
    } else {
      const int xt = tid - n_g;
      fr = xt & 127; x_pg0 = xt >> 7; x_step = 2;
      x_cnt = (xt >= 0 && xt < 256) ? 8 : 0;
    }
    const bool aligned = (g.L % 8) == 0 && (g.OW % 8) == 0;
    const int pitch = P.pitch;
    const int slot_bytes = P.slot_bytes;
    const int HW = g.H * g.W;
    uint32_t it = 0, chunk_it = 0;
    int p_sidx = 0;
    uint32_t p_phase = 0;
    int tpar = 0;
    // ---- G' operands kept in registers (Cout <= 72: at most three (channel, 8-pixel group) items per thread, clip
    // bits in one state word): grad_out of the tile is reused by all chunks and planes, the state words of a chunk
    // by its NSA planes.  Each item is refilled for the next chunk / tile right after its last use, a whole stage
    // before it is needed again, so no load latency is exposed in the plane loop.
    constexpr bool kCacheOk = kLut && CBits::CWN == 1;
    const bool gcache = kCacheOk && Kc <= 72 && P.gfast;
    const int gpg = tid & 15, gco0 = tid >> 4;
    float gvc[3][8];
    uint32_t swc[3][8];
    // ---- index tables of this CTA's chunks (built once): xt_chunk[c] = {lo, rows, first channel, channels touched},
    // xt_row[c][fr] = valid << 31 | kx << 24 | ky << 16 | (channel - first channel)
    int4 *xt_chunk = reinterpret_cast<int4 *>(smem_raw + P.xt_off);
    uint32_t *xt_row = reinterpret_cast<uint32_t *>(smem_raw + P.xt_off + (size_t)P.nxg * sizeof(int4));
    for (int e = tid; e < (i_end - i_begin) * 128; e += kWgProducerThreads) {
      const int c = e >> 7, r = e & 127;
      const int lo_ = (i_begin + c) * g.xbar, rows_ = min(rows_full, g.F - lo_);
      const int c_lo_ = lo_ / g.KK;
      uint32_t v = 0u;
      if (r < rows_) {
        const int f = lo_ + r, ci_ = f / g.KK, tap = f % g.KK;
        v = 0x80000000u | ((uint32_t)(tap % g.K) << 24) | ((uint32_t)(tap / g.K) << 16) | (uint32_t)ci_;
      }
      xt_row[e] = v;
      if (r == 0) xt_chunk[c] = make_int4(lo_, rows_, c_lo_, (lo_ + rows_ - 1) / g.KK - c_lo_ + 1);
    }
    named_barrier_sync(1, kWgProducerThreads);
    const int ir_cpr = P.async_rows ? g.W / P.async_rows : 1;
    const int ir_rq = tid / ir_cpr, ir_c16 = tid - ir_rq * ir_cpr;
    const int ir_sl = ir_rq / (P.rk > 0 ? P.rk : 1), ir_row = ir_rq - ir_sl * (P.rk > 0 ? P.rk : 1);
    auto pix_group = [&](int mt_) {  // {image, output row, first output column, kind} of this thread's pixel group
      const int64_t m = (int64_t)mt_ * kTcTileM + gpg * 8;
      int4 e = make_int4(0, 0, 0, -1);
      if (m < g.M) {
        const int b = (int)(m / g.L), l = (int)(m % g.L), oy = l / g.OW, ox = l % g.OW;
        e = make_int4(b, oy, ox, (aligned && m + 7 < g.M && ox + 7 < g.OW) ? 1 : 0);
      }
      return e;
    };
    // The register-resident path is only taken when every 8-pixel group is one aligned run of an image row or lies
    // entirely past the last pixel (P.gfast, decided on the host): no per-element fallback is compiled into it.
    // Groups past the end read valid addresses (image 0 / the last valid group): their activation codes are zero
    // (rows outside the image), so whatever finite values they contribute are multiplied by zero.
    auto load_go = [&](const int4 &pt, int mt_, int co, float (&gv)[8]) {
      (void)mt_;
      const float4 *gp = reinterpret_cast<const float4 *>(P.go + ((int64_t)pt.x * g.Cout + P.co0 + co) * g.L +
                                                          pt.y * g.OW + pt.z);
      const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
      gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w;
      gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
    };
    auto load_state = [&](const int4 &pt, int mt_, int i, int co, uint32_t (&sw)[8]) {  // i: crossbar (state) index
      (void)pt;
      const int64_t mg = min((int64_t)mt_ * kTcTileM + gpg * 8, (int64_t)g.M - 8);
      const uint32_t *sp = P.state + ((int64_t)(i * g.Cout + P.co0 + co) * CBits::SWORDS + CBits::CW0) * g.M + mg;
      const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(sp)), s1 = __ldg(reinterpret_cast<const uint4 *>(sp) + 1);
      sw[0] = s0.x; sw[1] = s0.y; sw[2] = s0.z; sw[3] = s0.w;
      sw[4] = s1.x; sw[5] = s1.y; sw[6] = s1.z; sw[7] = s1.w;
    };
    // staged-row table of tile mt_: offset of the row's first byte source, or kNoRow.  Rows of a slot are
    // (output row, ky) pairs, or -- when a tile is consecutive rows of one image (P.prow == 1) -- the distinct
    // input rows oy0 - pad + r.  Async staging copies whole rows (offset of image column 0); the synchronous path
    // loads words starting at staged column 0 (image column -pad - col0).
    auto fill_rowoff = [&](int mt_, int *tab) {
      if (tid >= 128 && tid < 128 + P.rk) {
        const int rr = tid - 128;
        const int orow = P.xshared ? 0 : rr / g.K, ky = P.xshared ? rr : rr % g.K;
        const int64_t m_row = (int64_t)mt_ * kTcTileM + (orow << P.ow_log2);
        int off = kNoRow;
        if (m_row < g.M) {
          const int b = (int)(m_row / g.L), oy = (int)(m_row % g.L) >> P.ow_log2;
          const int iy = oy - g.pad + ky;  // stride 1
          if (iy >= 0 && iy < g.H) off = (b * g.Cin * g.H + iy) * g.W - (P.async_rows ? 0 : g.pad + P.col0);
        }
        tab[rr] = off;
      }
    };
    // async staging of the channels crossbar i_ touches: 16-byte pieces, zero-fill for rows outside the image
    auto issue_rows = [&](int i_, const int *tab, uint8_t *dstbuf) {
      const int4 ci_ = xt_chunk[i_ - i_begin];
      const int c_lo_ = ci_.z, nch_ = ci_.w;
      // pieces of 16 bytes, or of 8 for rows of 8 (mod 16) bytes (P.async_rows = piece size)
      const int ps = P.async_rows, cpr = ir_cpr, total = nch_ * P.rk * cpr;
      for (int q = tid; q < total && !(P.dbg & 16); q += kWgProducerThreads) {
        int rq, c16, sl, row;
        if (q == tid) { rq = ir_rq; c16 = ir_c16; sl = ir_sl; row = ir_row; }  // (the common case: one piece per thread)
        else { rq = q / cpr; c16 = q - rq * cpr; sl = rq / P.rk; row = rq - sl * P.rk; }
        const int off = tab[row];
        const bool ok = off != kNoRow;
        const uint8_t *src = P.xcodes + (ok ? (size_t)(c_lo_ + sl) * HW + off + ps * c16 : (size_t)0);
        const uint32_t dst = smem_u32(dstbuf + (size_t)sl * slot_bytes + (size_t)row * pitch + g.pad + P.col0 + ps * c16);
        if (ps == 16)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u)
                       : "memory");
        else
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(ok ? 8u : 0u)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (P.async_rows) {  // padding columns stay zero for the whole kernel
      for (uint32_t q = tid * 16u; q < 2u * P.raw_bytes; q += kWgProducerThreads * 16u)
        *reinterpret_cast<uint4 *>(cv.raw + q) = make_uint4(0u, 0u, 0u, 0u);
    }
    int4 gpt = make_int4(0, 0, 0, -1), gpt_n = gpt;
    if (gcache && (int)blockIdx.x < P.mtiles) {
      gpt = pix_group(blockIdx.x);
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (gco0 + 24 * q < Kc) {
          load_go(gpt, blockIdx.x, gco0 + 24 * q, gvc[q]);
          load_state(gpt, blockIdx.x, i_begin / P.sdiv, gco0 + 24 * q, swc[q]);
        }
    }
    // ---- v2 G' operands (threads [0, n_g)): item = (channel quad gq2, 8-pixel group gpg2).  grad_out of the tile
    // (4 channels x 8 pixels, pre-scaled by 2^100) is reused by all chunks and planes, the W bytes of a chunk
    // (one word = 4 channels per pixel) by its NSA planes; both are refilled right after their last use.  Groups
    // past the last pixel read the last valid group: their activation digits are zero (staged rows outside the image).
    const bool g_thread = V2 && tid < n_g;
    const int nq2 = Kc >> 2;
    const int gq2 = V2 ? tid % nq2 : 0, gpg2 = V2 ? tid / nq2 : 0;
    // grad_out of the item (4 channels x 8 pixels), scaled per channel to [2^12, 2^13) and split ONCE per tile into two
    // fp16 pieces (cim_v2.cuh); the W bytes of a chunk (one word = 4 channels per pixel) are re-paired per channel
    // ([pixel 2q | . | pixel 2q+1 | .]) once per chunk.  Raw grad_out / W words of the next tile / chunk are loaded a
    // stage ahead.  Groups past the last pixel read the last valid group: their activation digits are zero.
    constexpr int PB2 = v2::bwd_piece_bits(NSW);
    float gs2[4][8];
    uint32_t gp1[4][4], gp2[4][4], wsp[4][4], wnx2[8];
    float chs[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    auto v2_group_m = [&](int mt_) { return min((int64_t)mt_ * kTcTileM + gpg2 * 8, (int64_t)g.M - 8); };
    auto v2_load_go = [&](int mt_, int c) {
      if (P.dbg & 32) return;
      const int64_t m = v2_group_m(mt_);
      const int b = (int)(m / g.L), l = (int)(m % g.L);
      const float4 *gp = reinterpret_cast<const float4 *>(P.go + ((int64_t)b * g.Cout + P.co0 + 4 * gq2 + c) * g.L + l);
      const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
      gs2[c][0] = g0.x; gs2[c][1] = g0.y; gs2[c][2] = g0.z; gs2[c][3] = g0.w;
      gs2[c][4] = g1.x; gs2[c][5] = g1.y; gs2[c][6] = g1.z; gs2[c][7] = g1.w;
    };
    auto v2_load_w = [&](int mt_, int i_) {
      if (P.dbg & 32) return;
      const uint8_t *wp = P.state2w + ((int64_t)i_ * g.M + v2_group_m(mt_)) * g.Cout + P.co0 + 4 * gq2;
#pragma unroll
      for (int e = 0; e < 8; ++e) wnx2[e] = __ldg(reinterpret_cast<const uint32_t *>(wp + (int64_t)e * g.Cout));
    };
    if constexpr (V2) {
#pragma unroll
      for (int e = 0; e < 8; ++e) wnx2[e] = 0u;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) gs2[c][e] = 0.0f;
      if (g_thread) {
#pragma unroll
        for (int c = 0; c < 4; ++c) chs[c] = v2::bwd_scale_from_maxbits(__ldg(P.chmax + P.co0 + 4 * gq2 + c));
      }
      if (g_thread && (int)blockIdx.x < P.mtiles) {
#pragma unroll
        for (int c = 0; c < 4; ++c) v2_load_go(blockIdx.x, c);
        v2_load_w(blockIdx.x, i_begin);
      }
    }
    // The tile loop is instantiated per thread ROLE (v2 at 64 channels: 1 = G' thread with one X item, 2 = X thread with
    // fourteen; v2 at 16 / 32 channels: 3 = G' thread, 4 = X thread with eight items or none; v1: 0 = one for all): the G' registers (pieces, paired W bytes, raw
    // grad_out) and the X registers (gathered codes of 14 items) then never live in the same thread.
    auto run_tiles = [&](auto role_tag) {
      constexpr int ROLE = decltype(role_tag)::value;
      constexpr int XIr = ROLE == 1 ? 3 : (ROLE == 3 ? 1 : XI);
      const bool gthr = ROLE == 1 || ROLE == 3 || (ROLE == 0 && g_thread);
      // X items of the thread: compile-time where the role fixes them (the run-time branches cost 10 % at 64 channels)
      const int xc = ROLE == 1 ? 3 : (ROLE == 2 ? 10 : (ROLE == 3 ? 0 : x_cnt));
      for (int mt = blockIdx.x; mt < P.mtiles; mt += gridDim.x, tpar ^= 1) {
        const int64_t m0 = (int64_t)mt * kTcTileM;
        const long long tt0 = CIMQ_TB();
        if constexpr (V2) {
          if (gthr && !(P.dbg & 32)) {  // this tile's grad_out (loaded during the previous tile's last stage) -> scaled fp16 pieces
  #pragma unroll
            for (int c = 0; c < 4; ++c)
  #pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float v0 = gs2[c][2 * q] * chs[c], v1 = gs2[c][2 * q + 1] * chs[c];
                constexpr uint32_t kMask = 0xffffffffu << (24 - PB2), kRnd = 1u << (23 - PB2);
                const float a0 = __uint_as_float((__float_as_uint(v0) + kRnd) & kMask);
                const float a1 = __uint_as_float((__float_as_uint(v1) + kRnd) & kMask);
                const float r0 = v0 - a0, r1 = v1 - a1;
                const float b0 = __uint_as_float((__float_as_uint(r0) + kRnd) & kMask);
                const float b1 = __uint_as_float((__float_as_uint(r1) + kRnd) & kMask);
                const __half2 h1 = __floats2half2_rn(a0, a1), h2v = __floats2half2_rn(b0, b1);
                gp1[c][q] = *reinterpret_cast<const uint32_t *>(&h1);
                gp2[c][q] = *reinterpret_cast<const uint32_t *>(&h2v);
              }
          }
        }
        const int mt_n = mt + gridDim.x;
        if (gcache && mt_n < P.mtiles) gpt_n = pix_group(mt_n);
        // per 8-pixel group: image, output row, first output column; fast = one image row, fully valid, aligned
        if (tid < 16) {
          const int64_t m = m0 + tid * 8;
          int4 e = make_int4(0, 0, 0, -1);  // w = -1: entirely past the end
          if (m < g.M) {
            const int b = (int)(m / g.L), l = (int)(m % g.L), oy = l / g.OW, ox = l % g.OW;
            e = make_int4(b, oy, ox, (aligned && m + 7 < g.M && ox + 7 < g.OW) ? 1 : 0);
          }
          cv.pixtab[tpar * 16 + tid] = e;
        }
        int *rowoff = cv.rowoff + tpar * 128;
        // async staging fills the table of a tile one chunk before the tile starts (below); the first tile's here
        if (P.fastx && (!P.async_rows || mt == (int)blockIdx.x)) fill_rowoff(mt, rowoff);
        const long long tt1 = CIMQ_TB();
        named_barrier_sync(1, kWgProducerThreads);
        d_conv += tt1 - tt0;
        d_tbar += CIMQ_TB() - tt1;
        if (P.async_rows && mt == (int)blockIdx.x) {  // rows of the very first chunk
          issue_rows(i_begin, rowoff, cv.raw + (size_t)(chunk_it & 1) * P.raw_bytes);
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          named_barrier_sync(1, kWgProducerThreads);
        }
        const int4 *ptab = cv.pixtab + tpar * 16;
        for (int i = i_begin; i < i_end; ++i, ++chunk_it) {
          // (channel, tap) of this thread's crossbar row and the chunk's channel range: from the tables built at kernel
          // start -- five divisions by run-time values per chunk and thread were a quarter of the producers' instructions
          const int4 cinf = xt_chunk[i - i_begin];  // {lo, rows, c_lo, nch}
          const int lo = cinf.x, rows = cinf.y, c_lo = cinf.z;
          const uint32_t xinf = xt_row[(i - i_begin) * 128 + fr];
          const bool frow = (xinf >> 31) != 0;
          // (Measured: letting warps whose 32 rows all lie past the chunk's last row -- the 16- / 32-row last crossbar of
          // a 16- / 32-channel layer -- skip gather and stores changed nothing: 701 -> 700 us at 16 channels.)
          const int ci = (int)(xinf & 0xffffu), ky = (int)((xinf >> 16) & 0xffu), kx = (int)((xinf >> 24) & 0x7fu);
          const int st_next = (i + 1) / P.sdiv, st_first = i_begin / P.sdiv;  // state (crossbar) index of the next chunk
          uint8_t *raw = cv.raw + (size_t)(chunk_it & 1) * P.raw_bytes;
          const long long ts0 = CIMQ_TB();
          if (P.async_rows) {
            // the rows of this chunk are already in raw[chunk_it & 1]; start the next chunk's (or next tile's first)
            uint8_t *nbuf = cv.raw + (size_t)((chunk_it + 1) & 1) * P.raw_bytes;
            if (i + 1 < i_end) {
              issue_rows(i + 1, rowoff, nbuf);
            } else if (mt_n < P.mtiles) {
              fill_rowoff(mt_n, cv.rowoff + (tpar ^ 1) * 128);
              named_barrier_sync(1, kWgProducerThreads);
              issue_rows(i_begin, cv.rowoff + (tpar ^ 1) * 128, nbuf);
            }
          } else if (P.fastx) {
            // stage the input rows of the channels this crossbar touches (once per chunk, shared by all planes)
            const int nch = (lo + rows - 1) / g.KK - c_lo + 1;
            const int wpr_log2 = P.pitch_log2 - 2;
            const int xw = tid & ((1 << wpr_log2) - 1);
            const int rstep = kWgProducerThreads >> wpr_log2;
            const int rk = P.rk;
            const int total_rows = nch * rk;
            const int ix = 4 * xw - P.col0 - g.pad;
            const bool xok = ix >= 0 && ix < g.W;
            int row = tid >> wpr_log2, sl = 0;
            while (row >= rk) { row -= rk; ++sl; }
            for (int r0 = tid >> wpr_log2; r0 < total_rows; r0 += 4 * rstep) {
              uint32_t v[4];
              int dsto[4];
  #pragma unroll
              for (int u = 0; u < 4; ++u) {
                v[u] = 0u;
                dsto[u] = -1;
                if (r0 + u * rstep < total_rows) {
                  const int off = rowoff[row];
                  if (xok && off != kNoRow)
                    v[u] = __ldg(reinterpret_cast<const uint32_t *>(P.xcodes + (size_t)(c_lo + sl) * HW + off + 4 * xw));
                  dsto[u] = sl * slot_bytes + (row << P.pitch_log2) + 4 * xw;
                }
                row += rstep;
                while (row >= rk) { row -= rk; ++sl; }
              }
  #pragma unroll
              for (int u = 0; u < 4; ++u)
                if (dsto[u] >= 0) *reinterpret_cast<uint32_t *>(raw + dsto[u]) = v[u];
            }
            named_barrier_sync(1, kWgProducerThreads);
          }
          // ---- the activation codes of this thread's X items (row fr, 8-pixel group pg = tid/128 + 3q), gathered once
          // per chunk: the NSA digit planes below only shift and mask them
          uint32_t xlo[XIr], xhi[XIr];
          const bool xfast = V2 && P.fastx;  // (v2 layers have 1-bit digits)
          if (xfast && !(P.dbg & 8)) {
            const uint8_t *src_row = raw + (size_t)(ci - c_lo) * slot_bytes + (size_t)ky * pitch + kx + P.col0;
            // output row r of the tile is staged r * prow rows further down; prow is 1 or K (not a power of two in
            // general), so the row offset is a multiply
            const int rowb = P.prow * pitch;
            auto gather = [&](auto cnt, auto stp) {
  #pragma unroll
              for (int q = 0; q < decltype(cnt)::value; ++q) {
                const int p0 = (x_pg0 + decltype(stp)::value * q) * 8;
                uint32_t lo8 = 0u, hi8 = 0u;
                if (frow) {
                  const uint8_t *src = src_row + (p0 >> P.ow_log2) * rowb + (p0 & ((1 << P.ow_log2) - 1));
                  const uint32_t sa = smem_u32(src);
                  const uint32_t *al = reinterpret_cast<const uint32_t *>(src - (sa & 3u));
                  const uint32_t w0 = al[0], w1 = al[1], w2 = al[2];
                  const uint32_t bsh = (sa & 3u) * 8u;
                  lo8 = __funnelshift_r(w0, w1, bsh);
                  hi8 = __funnelshift_r(w1, w2, bsh);
                }
                xlo[q] = lo8;
                xhi[q] = hi8;
              }
            };
            if constexpr (ROLE == 1) gather(std::integral_constant<int, 3>{}, std::integral_constant<int, 1>{});
            else if constexpr (ROLE == 2) gather(std::integral_constant<int, 10>{}, std::integral_constant<int, 1>{});
            else if constexpr (ROLE == 3) { /* no X items */ }
            else if constexpr (ROLE == 4) {
              if (xc == 8) gather(std::integral_constant<int, 8>{}, std::integral_constant<int, 2>{});
            } else {
              if (xc == 14) gather(std::integral_constant<int, 14>{}, std::integral_constant<int, 1>{});
              else if (xc == 8) gather(std::integral_constant<int, 8>{}, std::integral_constant<int, 2>{});
              else if (xc == 1) gather(std::integral_constant<int, 1>{}, std::integral_constant<int, 1>{});
            }
          }
  #pragma unroll
          for (int q = 0; q < XIr; ++q) {
            if (xfast) break;
            const int pg = x_pg0 + x_step * q;
            xlo[q] = 0u;
            xhi[q] = 0u;
            if (pg < 16 && q < xc) {
            uint32_t lo8 = 0u, hi8 = 0u;
            if (frow) {
              if (P.fastx) {
                const int p0 = pg * 8;
                const uint8_t *src = raw + (size_t)(ci - c_lo) * slot_bytes +
                                     (((p0 >> P.ow_log2) * P.prow) + ky) * pitch +
                                     (p0 & ((1 << P.ow_log2) - 1)) + kx + P.col0;
                const uint32_t sa = smem_u32(src);
                const uint32_t *al = reinterpret_cast<const uint32_t *>(src - (sa & 3u));
                const uint32_t w0 = al[0], w1 = al[1], w2 = al[2];
                const uint32_t bsh = (sa & 3u) * 8u;
                lo8 = __funnelshift_r(w0, w1, bsh);
                hi8 = __funnelshift_r(w1, w2, bsh);
              } else {  // geometries without staged rows: out of line, it would be inlined six times here
                const uint2 c8 = gather_codes_generic(GatherGeo{g.stride, g.pad, g.H, g.W, g.Cin, g.L, g.OW, g.M}, P.xcodes,
                                                      ptab[pg], ci, ky, kx, m0 + pg * 8);
                lo8 = c8.x;
                hi8 = c8.y;
              }
            }
              xlo[q] = lo8;
              xhi[q] = hi8;
            }
          }
          d_stage += CIMQ_TB() - ts0;
          if constexpr (V2) {
  #pragma unroll
            for (int c = 0; c < 4; ++c)
  #pragma unroll
              for (int q = 0; q < 4; ++q) wsp[c][q] = __byte_perm(wnx2[2 * q], wnx2[2 * q + 1], c | ((4 + c) << 8));
          }
          for (int j = 0; j < NSA; ++j, ++it) {
            // (stage index and phase as running counters: `it % stages` is a division by a run-time value)
            const int sidx = p_sidx;
            const uint32_t use = p_phase;
            if (++p_sidx == P.stages) { p_sidx = 0; p_phase ^= 1u; }
            const long long tw0 = CIMQ_TB();
            mbar_wait(cv.empty0 + 8 * sidx, (use & 1) ^ 1);
            const long long tw1 = CIMQ_TB();
            d_wait += tw1 - tw0;
            uint8_t *st_ptr = cv.stage_base + (size_t)sidx * P.stage_bytes;
            // ---- X_j tile [128 crossbar rows x 128 pixels] (bf16 digits); item = (row fr, 8-pixel group pg)
            const int sh = g.abs_ * j;
            // v2: fp16 operands (the G' pieces are fp16); the digit carries the slice weight 2^j (mask[k][j] * 2^(-wbs*k)
            // = 2^j for 1-bit slices, lsq.py:306, 363-364), the pass COUNT of plane W multiplies grad_out
            const uint32_t one_bf = V2 ? 0x3C00u + ((uint32_t)j << 10) : 0x3F80u;
            if (xfast && !(P.dbg & 2)) {
              uint8_t *dst = st_ptr + tc_tile_offset16(fr, x_pg0 * 8, kTcLBO, a_sbo);
              const uint32_t cmul = one_bf >> sh;  // the 16-bit pattern of 1.0 divided by the bit weight of the digit
              if constexpr (ROLE == 1) x_store_fast<3, 1, XIr>(dst, xlo, xhi, sh, cmul);
              else if constexpr (ROLE == 2) x_store_fast<10, 1, XIr>(dst, xlo, xhi, sh, cmul);
              else if constexpr (ROLE == 3) { /* no X items */ }
              else if constexpr (ROLE == 4) {
                if (xc == 8) x_store_fast<8, 2, XIr>(dst, xlo, xhi, sh, cmul);
              } else {
                if (xc == 14) x_store_fast<14, 1, XIr>(dst, xlo, xhi, sh, cmul);
                else if (xc == 8) x_store_fast<8, 2, XIr>(dst, xlo, xhi, sh, cmul);
                else if (xc == 1) x_store_fast<1, 1, XIr>(dst, xlo, xhi, sh, cmul);
              }
            }
  #pragma unroll
            for (int q = 0; q < XIr; ++q) {
              if (xfast) break;
              const int pg = x_pg0 + x_step * q;
              if (pg >= 16 || q >= xc) continue;
              const uint32_t lo8 = xlo[q], hi8 = xhi[q];
              uint32_t d[4];
              if (g.amask == 1) {  // 1-bit digits: bf16(1) = 0x3F80; spread two bytes to 16-bit lanes, one multiply
                const uint32_t tl = (lo8 >> sh) & 0x01010101u, th = (hi8 >> sh) & 0x01010101u;
                d[0] = __byte_perm(tl, 0u, 0x4140) * one_bf;
                d[1] = __byte_perm(tl, 0u, 0x4342) * one_bf;
                d[2] = __byte_perm(th, 0u, 0x4140) * one_bf;
                d[3] = __byte_perm(th, 0u, 0x4342) * one_bf;
              } else {
                const uint32_t am = (uint32_t)g.amask;
  #pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  const uint32_t wsrc = e2 < 2 ? lo8 : hi8;
                  const uint32_t b0 = (wsrc >> (16 * (e2 & 1) + sh)) & am, b1 = (wsrc >> (16 * (e2 & 1) + 8 + sh)) & am;
                  d[e2] = cvt_bf16x2((float)b1, (float)b0);
                }
              }
              *reinterpret_cast<uint4 *>(st_ptr + tc_tile_offset16(fr, pg * 8, kTcLBO, a_sbo)) =
                  make_uint4(d[0], d[1], d[2], d[3]);
            }
            // ---- G'_j tiles (3 bf16 terms) [Cout x 128 pixels]; item = (channel co, 8-pixel group pg), lanes along pixels
            const long long tx1 = CIMQ_TB();
            d_x += tx1 - tw1;
            uint8_t *gb = st_ptr + P.a_bytes;
            if constexpr (V2) {
              // G'_j[co, m] = (go * 2^s_co as two fp16 pieces) * (pass count of activation slice j): exact products
              if (gthr) {
                const bool last_j = j + 1 == NSA;
                const bool next_chunk = i + 1 < i_end, next_tile = !next_chunk && mt_n < P.mtiles;
                if (last_j) {  // the W bytes of the next chunk (or of the next tile's first chunk) start their trip now
                  if (next_chunk) v2_load_w(mt, i + 1);
                  else if (next_tile) v2_load_w(mt_n, i_begin);
                }
                if (j == 0 && next_tile) {  // next tile's grad_out: three stages ahead of its split
                  // (measured: issuing these loads after the stage's arrival instead, one stage earlier for the W bytes,
                  // made the kernel 4 % slower)
  #pragma unroll
                  for (int c = 0; c < 4; ++c) v2_load_go(mt_n, c);
                }
                // count field j of both bytes -> fp16x2 (1024 + cnt * 4^j) -> cnt
                const uint32_t fmask = 0x00030003u << (2 * j);
                const __half2 sk = __float2half2_rn(1.0f / (float)(1 << (2 * j)));
                const __half2 ok = __float2half2_rn(-1024.0f / (float)(1 << (2 * j)));
  #pragma unroll
                for (int c = 0; c < 4; ++c) {
                  if (P.dbg & 4) break;
                  uint32_t a1[4], a2[4];
  #pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const uint32_t fw = (wsp[c][q] & fmask) | 0x64006400u;
                    const __half2 cnt = __hfma2(*reinterpret_cast<const __half2 *>(&fw), sk, ok);
                    const __half2 r1 = __hmul2(*reinterpret_cast<const __half2 *>(&gp1[c][q]), cnt);
                    const __half2 r2 = __hmul2(*reinterpret_cast<const __half2 *>(&gp2[c][q]), cnt);
                    a1[q] = *reinterpret_cast<const uint32_t *>(&r1);
                    a2[q] = *reinterpret_cast<const uint32_t *>(&r2);
                  }
                  const uint32_t off = tc_tile_offset16(4 * gq2 + c, gpg2 * 8, kWgLBO, b_sbo);
                  *reinterpret_cast<uint4 *>(gb + off) = make_uint4(a1[0], a1[1], a1[2], a1[3]);
                  *reinterpret_cast<uint4 *>(gb + P.b_bytes + off) = make_uint4(a2[0], a2[1], a2[2], a2[3]);
                }
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(cv.full0 + 8 * sidx);  // one arrival per warp (every thread fenced its own stores)
              d_g += CIMQ_TB() - tx1;
              continue;
            }
            float wv[NSW];
  #pragma unroll
            for (int k = 0; k < NSW; ++k) wv[k] = cv.wtab[k * NSA + j];
            const uint32_t lutj = smem_u32(cv.lut + j * 16);
            const int pg = tid & 15;
            if (gcache) {
              if constexpr (kCacheOk) {
                const bool last_j = j + 1 == NSA;
                const bool next_chunk = i + 1 < i_end, next_tile = !next_chunk && mt_n < P.mtiles;
  #pragma unroll
                for (int q = 0; q < 3; ++q) {
                  const int co = gco0 + 24 * q;
                  if (co < Kc) {
                    float v[8];
  #pragma unroll
                    for (int e = 0; e < 8; ++e)
                      v[e] = gvc[q][e] *
                             lds_const_f32(lutj + 4u * ((swc[q][e] >> (CBits::CB + j * NSW)) & ((1u << NSW) - 1u)));
                    if (last_j) {  // last use of this item's state words (and, at the last chunk, of its grad_out)
                      if (next_chunk) load_state(gpt, mt, st_next, co, swc[q]);
                      else if (next_tile) {
                        load_state(gpt_n, mt_n, st_first, co, swc[q]);
                        load_go(gpt_n, mt_n, co, gvc[q]);
                      }
                    }
                    uint32_t hi[4], mid[4], lo3[4];
  #pragma unroll
                    for (int e2 = 0; e2 < 4; ++e2) split3x2(v[2 * e2], v[2 * e2 + 1], hi[e2], mid[e2], lo3[e2]);
                    const uint32_t off = tc_tile_offset16(co, pg * 8, kWgLBO, b_sbo);
                    *reinterpret_cast<uint4 *>(gb + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(gb + P.b_bytes + off) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
                    *reinterpret_cast<uint4 *>(gb + 2 * (size_t)P.b_bytes + off) =
                        make_uint4(lo3[0], lo3[1], lo3[2], lo3[3]);
                  }
                }
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(cv.full0 + 8 * sidx);  // one arrival per warp (every thread fenced its own stores)
              d_g += CIMQ_TB() - tx1;
              continue;
            }
            const int4 pt = ptab[pg];
            const int64_t mg = m0 + pg * 8;
            // loads of the next channel are issued before the current one is processed (latency hiding)
            auto load_g = [&](int co, float (&gv)[8], uint32_t (&sw)[8][CBits::CWN]) {
              const uint32_t *sp = P.state + ((int64_t)((i / P.sdiv) * g.Cout + P.co0 + co) * CBits::SWORDS + CBits::CW0) * g.M + mg;
              if (pt.w == 1) {
                const float4 *gp = reinterpret_cast<const float4 *>(
                    P.go + ((int64_t)pt.x * g.Cout + P.co0 + co) * g.L + pt.y * g.OW + pt.z);
                const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
                gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w;
                gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
  #pragma unroll
                for (int w = 0; w < CBits::CWN; ++w) {
                  const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(sp + (int64_t)w * g.M));
                  const uint4 s1 = __ldg(reinterpret_cast<const uint4 *>(sp + (int64_t)w * g.M) + 1);
                  sw[0][w] = s0.x; sw[1][w] = s0.y; sw[2][w] = s0.z; sw[3][w] = s0.w;
                  sw[4][w] = s1.x; sw[5][w] = s1.y; sw[6][w] = s1.z; sw[7][w] = s1.w;
                }
              } else {
  #pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const int64_t m = mg + e;
                  gv[e] = 0.0f;
  #pragma unroll
                  for (int w = 0; w < CBits::CWN; ++w) sw[e][w] = 0xffffffffu;
                  if (m < g.M) {
                    const int b = (int)(m / g.L), l = (int)(m % g.L);
                    gv[e] = __ldg(&P.go[((int64_t)b * g.Cout + P.co0 + co) * g.L + l]);
  #pragma unroll
                    for (int w = 0; w < CBits::CWN; ++w) sw[e][w] = __ldg(sp + (int64_t)w * g.M + e);
                  }
                }
              }
            };
            float gv_n[8];
            uint32_t sw_n[8][CBits::CWN];
            const int co0 = tid >> 4;  // 0..23
            if (co0 < Kc) load_g(co0, gv_n, sw_n);
            for (int co = co0; co < Kc; co += 24) {
              float gv[8];
              uint32_t sw[8][CBits::CWN];
  #pragma unroll
              for (int e = 0; e < 8; ++e) {
                gv[e] = gv_n[e];
  #pragma unroll
                for (int w = 0; w < CBits::CWN; ++w) sw[e][w] = sw_n[e][w];
              }
              if (co + 24 < Kc) load_g(co + 24, gv_n, sw_n);
              float v[8];
  #pragma unroll
              for (int e = 0; e < 8; ++e) {
                if constexpr (kLut && CBits::CWN == 1) {
                  v[e] = gv[e] * lds_const_f32(lutj + 4u * ((sw[e][0] >> (CBits::CB + j * NSW)) & ((1u << NSW) - 1u)));
                } else {
                  v[e] = gv[e] * pass_weight<NSW, CBits::CWN>(sw[e], CBits::CB + j * NSW, 1, wv);
                }
              }
              uint32_t hi[4], mid[4], lo3[4];
  #pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) split3x2(v[2 * e2], v[2 * e2 + 1], hi[e2], mid[e2], lo3[e2]);
              const uint32_t off = tc_tile_offset16(co, pg * 8, kWgLBO, b_sbo);
              *reinterpret_cast<uint4 *>(gb + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4 *>(gb + P.b_bytes + off) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
              *reinterpret_cast<uint4 *>(gb + 2 * (size_t)P.b_bytes + off) = make_uint4(lo3[0], lo3[1], lo3[2], lo3[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(cv.full0 + 8 * sidx);  // one arrival per warp (every thread fenced its own stores)
            d_g += CIMQ_TB() - tx1;
          }
          if (P.async_rows) {  // my copies of the next chunk's rows have landed; publish them, retire this chunk's buffer
            const long long tb0 = CIMQ_TB();
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            named_barrier_sync(1, kWgProducerThreads);
            d_bar += CIMQ_TB() - tb0;
          }
        }
        gpt = gpt_n;
      }
    };
    if constexpr (V2) {
      if (n_g == 256) {
        if (g_thread) run_tiles(std::integral_constant<int, 1>{});
        else run_tiles(std::integral_constant<int, 2>{});
      } else {
        if (g_thread) run_tiles(std::integral_constant<int, 3>{});
        else run_tiles(std::integral_constant<int, 4>{});
      }
    } else {
      run_tiles(std::integral_constant<int, 0>{});
    }
    if (dbg) { P.debug[0] = d_wait; P.debug[1] = d_stage; P.debug[2] = d_x; P.debug[3] = d_g; P.debug[7] = d_bar; P.debug[16] = d_conv; P.debug[17] = d_tbar; }
  } else {
    reg_dealloc<kWgRegsMma>();
    // ------------------------------------------------------------------ MMA issuer
    if (warp == kMmaWarp && has_work) {  // the whole warp walks the loop (uniform control flow); lane 0 issues
      const uint32_t idesc = V2 ? idesc_f16_f32(128, Kc) : idesc_bf16_f32(128, Kc);
      constexpr int kPieces = V2 ? v2::kBwdPieces : 3;  // terms of the real-valued operand
      // (Measured: the two v2 piece tiles are contiguous and can be ONE B operand of N = 2 * Kc rows -- half the MMAs,
      // the X tile read once, the epilogue adds the two column halves; where tensor memory allows it (16 / 32
      // channels) that took 703 -> 698 us and 327 -> 323 us: the MMA count does not bound this kernel.  Not kept.)
      uint32_t it = 0;
      int m_sidx = 0;
      uint32_t m_phase = 0;
      bool first_tile = true;
      const bool dbg = kTimers && P.debug != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
      long long d_full = 0;
      const long long t_begin = CIMQ_TB();
      const uint32_t stage0 = smem_u32(cv.stage_base);
      for (int mt = blockIdx.x; mt < P.mtiles; mt += gridDim.x, first_tile = false) {
        for (int i = i_begin; i < i_end; ++i) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(i - i_begin) * Kc;
          for (int j = 0; j < NSA; ++j, ++it) {
            const int sidx = m_sidx;
            const uint32_t use = m_phase;
            if (++m_sidx == P.stages) { m_sidx = 0; m_phase ^= 1u; }
            const long long tf0 = CIMQ_TB();
            mbar_wait<CIMQ_MMA_SLEEP>(cv.full0 + 8 * sidx, use & 1);
            d_full += CIMQ_TB() - tf0;
            tc_fence_after();
            const uint32_t a0 = stage0 + (uint32_t)sidx * P.stage_bytes;
            // one descriptor per operand and stage; the k-steps and pieces only add to its address field (bits 0-13,
            // address >> 4: shared memory ends below 2^18, so the sum never carries out of the field)
            const uint64_t adesc0 = make_smem_desc(a0, kTcLBO, a_sbo);
            const uint64_t bdesc0 = make_smem_desc(a0 + P.a_bytes, kWgLBO, b_sbo);
            if (lane == 0) {
              if (!(P.dbg & 1)) {
#pragma unroll
                for (int sp = 0; sp < kPieces; ++sp)
#pragma unroll
                  for (int ks = 0; ks < 8; ++ks)  // 128 pixels = 8 x K16
                    umma_f16(d_tmem, adesc0 + (uint64_t)((ks * 2 * kTcLBO) >> 4),
                             bdesc0 + (uint64_t)((sp * P.b_bytes + ks * 2 * kWgLBO) >> 4), idesc,
                             (first_tile && j == 0 && sp == 0 && ks == 0) ? 0u : 1u);
              }
              umma_commit(cv.empty0 + 8 * sidx);
            }
            __syncwarp();
          }
        }
      }
      if (lane == 0) umma_commit(cv.tfull0);  // every accumulation of this CTA is complete
      if (dbg && lane == 0) { P.debug[4] = d_full; P.debug[6] = clock64() - t_begin; P.debug[12] = t_begin; P.debug[13] = clock64(); }
    }
  }
  if (warp >= kEpilogueWarp0 && warp < kEpilogueWarp0 + kEpilogueWarps) {
    // ------------------------------------------------------------------ epilogue (once, at the end)
    const int quarter = warp & 3;
    const int frow = quarter * 32 + lane;  // TMEM lane = crossbar row
    const float scale = P.s[0] / (float)NSW;  // x_sl * s_a (lsq.py:295), mean over weight slices (lsq.py:366)
    float *part = P.out + (int64_t)blockIdx.x * g.F * g.Cout;
    if (has_work) {
      mbar_wait(cv.tfull0, 0);
      tc_fence_after();
    }
    if (kTimers && P.debug != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && warp == kEpilogueWarp0 && lane == 0) P.debug[14] = clock64();
    for (int i = i_begin; i < i_end; ++i) {
      const int lo = i * g.xbar;
      const int rows = min(rows_full, g.F - lo);
      for (int c0 = 0; c0 < Kc; c0 += 16) {
        int v[16];
        if (has_work) {
          tmem_ld<16>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i - i_begin) * Kc + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int cc = 0; cc < 16; ++cc) v[cc] = 0;
        }
        if (frow < rows) {
          float4 *dst = reinterpret_cast<float4 *>(part + (int64_t)(lo + frow) * g.Cout + P.co0 + c0);
          // v2: undo the per-channel power-of-two scale of grad_out
          float cs[16];
#pragma unroll
          for (int cc = 0; cc < 16; ++cc)
            cs[cc] = V2 ? scale * v2::bwd_scale_inverse(v2::bwd_scale_from_maxbits(__ldg(P.chmax + P.co0 + c0 + cc))) : scale;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            dst[q4] = make_float4(__int_as_float(v[4 * q4]) * cs[4 * q4], __int_as_float(v[4 * q4 + 1]) * cs[4 * q4 + 1],
                                  __int_as_float(v[4 * q4 + 2]) * cs[4 * q4 + 2], __int_as_float(v[4 * q4 + 3]) * cs[4 * q4 + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kTimers && P.debug != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) P.debug[15] = clock64();  // all roles done
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// Sum of the per-CTA partials [part][F][Cout] -> grad_w [Cout][F] (weight layout, lsq.py:369).  Block = 32
// consecutive elements of the partial layout (coalesced 128-byte reads) x 8 slices of the part range; the slices
// are combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256) bwd_weight_tc_finish_kernel(Geo g, int nparts,
                                                                   const float *__restrict__ partial,
                                                                   float *__restrict__ gw) {
  __shared__ float red[8][32];
  const int64_t n = (int64_t)g.Cout * g.F;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * 32 + lane;  // e = f * Cout + co
  float v = 0.0f;
  if (e < n) {
    float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;  // four loads in flight, fixed summation order
    int p = slice;
    for (; p + 24 < nparts; p += 32) {
      const float a0 = __ldg(partial + (int64_t)p * n + e), a1 = __ldg(partial + (int64_t)(p + 8) * n + e);
      const float a2 = __ldg(partial + (int64_t)(p + 16) * n + e), a3 = __ldg(partial + (int64_t)(p + 24) * n + e);
      v0 += a0; v1 += a1; v2 += a2; v3 += a3;
    }
    for (; p < nparts; p += 8) v0 += __ldg(partial + (int64_t)p * n + e);
    v = (v0 + v1) + (v2 + v3);
  }
  red[slice][lane] = v;
  __syncthreads();
  if (slice == 0 && e < n) {
    float t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][lane];
    const int f = (int)(e / g.Cout), co = (int)(e % g.Cout);
    gw[(int64_t)co * g.F + f] = t;
  }
}

// bf16 weight digit tiles for dgrad: tile (channel block cb, i, k) = [Nf rows (crossbar row) x Kc channels]
// K-major no-swizzle, Kc = min(Cout, 128)
// v2s != 0: digits scaled by 2^(-k - 8*(co & 1)) (exact in bf16), the inverse of what the v2 dgrad producers leave on
// the masked operand (pass-count field position of the D byte) beyond the slice weight 2^k.
__global__ void weight_tiles_bwd_kernel(Geo g, int Nf, int Kc, int v2s, const int8_t *__restrict__ wcodes,
                                        uint16_t *__restrict__ tiles) {
  const int64_t tile_elems = (int64_t)Nf * Kc;
  const int64_t tiles_per_block = (int64_t)g.NX * g.NSW;
  const int64_t n = (int64_t)(g.Cout / Kc) * tiles_per_block * tile_elems;
  const uint32_t sbo = (uint32_t)Kc * 16u;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tile = idx / tile_elems;
    const int within = (int)(idx % tile_elems);
    const int fr = within / Kc, cl = within % Kc;
    const int cb = (int)(tile / tiles_per_block), ik = (int)(tile % tiles_per_block);
    const int i = ik / g.NSW, k = ik % g.NSW;
    const int co = cb * Kc + cl;
    const int f = i * g.xbar + fr;
    const int hi = min((i + 1) * g.xbar, g.F);
    int digit = 0;
    if (f < hi) {
      const int code = wcodes[(int64_t)co * g.F + f];
      const int mag = code < 0 ? -code : code;
      digit = (mag >> (g.wbs * k)) & g.wmask;
      if (code < 0) digit = -digit;
    }
    float dv = (float)digit;
    if (v2s) dv = ldexpf(dv, -k - 8 * (co & 1));
    tiles[tile * tile_elems + tc_tile_offset16(fr, cl, kTcLBO, sbo) / 2] = __bfloat16_as_ushort(__float2bfloat16_rn(dv));
  }
}

// output channels per launch: layers with more run as blocks (v2: 64, one producer thread holds at most 32 channels)
inline int bwd_channel_block(const Geo &g, bool v2 = false) {
  const int cap = v2 ? 64 : 128;
  return g.Cout > cap ? cap : g.Cout;
}

inline int wgrad_chunks_per_group(const Geo &g, bool v2 = false) {
  int n = 512 / bwd_channel_block(g, v2);
  return n < g.NX ? n : g.NX;
}

}  // namespace

inline bool bwd_slices_supported(const Geo &g) {
  return (g.NSW == g.NSA) && (g.NSW == 2 || g.NSW == 3 || g.NSW == 4 || g.NSW == 8);
}

#define CIMQ_BWD_DISPATCH(KERNEL, ...)                                                          \
  do {                                                                                          \
    const bool tern = g.adc_mode != CIMQ_ADC_MULTIBIT;                                           \
    if (g.NSW == 2 && tern) { KERNEL(2, 2, true, __VA_ARGS__); }                                 \
    else if (g.NSW == 2) { KERNEL(2, 2, false, __VA_ARGS__); }                                   \
    else if (g.NSW == 3 && tern) { KERNEL(3, 3, true, __VA_ARGS__); }                            \
    else if (g.NSW == 3) { KERNEL(3, 3, false, __VA_ARGS__); }                                   \
    else if (g.NSW == 4 && tern) { KERNEL(4, 4, true, __VA_ARGS__); }                            \
    else if (g.NSW == 4) { KERNEL(4, 4, false, __VA_ARGS__); }                                   \
    else if (g.NSW == 8 && tern) { KERNEL(8, 8, true, __VA_ARGS__); }                            \
    else { KERNEL(8, 8, false, __VA_ARGS__); }                                                   \
  } while (0)

template <int W, int A, bool T>
static int launch_wgrad_instance(const BwdParams &P, dim3 grid, size_t smem, cudaStream_t st) {
  if constexpr (W <= v2::kMaxNS) {
    if (P.v2) {
      CIMQ_CUDA_OK(cudaFuncSetAttribute(bwd_weight_tc_kernel<W, A, T, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bwd_weight_tc_kernel<W, A, T, true><<<grid, kThreads, smem, st>>>(P);
      return 0;
    }
  }
  CIMQ_CUDA_OK(cudaFuncSetAttribute(bwd_weight_tc_kernel<W, A, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  bwd_weight_tc_kernel<W, A, T, false><<<grid, kThreads, smem, st>>>(P);
  return 0;
}

// Crossbars deeper than 128 rows are processed as 128-row "virtual" chunks: neither gradient couples the rows of
// a crossbar (dgrad: the rows are the N dimension; wgrad: the rows are the M dimension), only the pass mask --
// the ADC state -- is shared, and virtual chunk i reads the state of crossbar i / sdiv.  Needs xbar % 128 == 0.
static Geo bwd_virtual_geo(const Geo &g, int *sdiv) {
  Geo v = g;
  *sdiv = 1;
  const int rows = g.xbar < g.F ? g.xbar : g.F;
  if (rows > 128 && g.xbar % 128 == 0) {
    *sdiv = g.xbar / 128;
    v.xbar = 128;
    v.NX = (g.F + 127) / 128;
  }
  return v;
}

bool tc_backward_supported(const Geo &g) {
  if (!bwd_slices_supported(g)) return false;
  if (g.Cout % 16 != 0 || (g.Cout > 128 && (g.Cout % 128 != 0 || g.Cout > 512))) return false;
  if (g.pairs > kMaxPairs) return false;
  const int rows = g.xbar < g.F ? g.xbar : g.F;
  // wgrad M tile / dgrad N tile is one (virtual) chunk of <= 128 rows
  if (rows > 128 && g.xbar % 128 != 0) return false;
  if ((int64_t)g.B * g.Cin * g.H * g.W >= (1ll << 31)) return false;
  return true;
}

bool v2_backward_supported(const Geo &g) {
  if (!tc_backward_supported(g) || !v2::supported(g)) return false;
  if (g.Cout > 64 && g.Cout % 64 != 0) return false;
  // 8-pixel groups of the wgrad operands never straddle images or the end
  if (g.L % 8 != 0 || g.M % 8 != 0) return false;
  return true;
}

int64_t wtiles_bwd_bytes(const Geo &g0) {
  if (!tc_backward_supported(g0)) return 0;
  int sdiv;
  const Geo g = bwd_virtual_geo(g0, &sdiv);
  return (int64_t)g.NX * g.NSW * tc_nf(g) * g.Cout * 2;
}

int launch_weight_tiles_bwd(const Geo &g0, const int8_t *wcodes, void *tiles, cudaStream_t st) {
  int sdiv;
  const Geo g = bwd_virtual_geo(g0, &sdiv);
  const int64_t n = (int64_t)g.NX * g.NSW * tc_nf(g) * g.Cout;
  weight_tiles_bwd_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(g, tc_nf(g), bwd_channel_block(g), 0, wcodes,
                                                                 reinterpret_cast<uint16_t *>(tiles));
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

static int wgrad_ctas(const Geo &g, bool v2, int *nxg_out, int *groups_out) {
  const int mtiles = (g.M + kTcTileM - 1) / kTcTileM;
  const int nxg = wgrad_chunks_per_group(g, v2);
  const int groups = (g.NX + nxg - 1) / nxg;
  int ctas = 148 / groups;
  if (ctas > mtiles) ctas = mtiles;
  if (ctas < 1) ctas = 1;
  if (nxg_out) *nxg_out = nxg;
  if (groups_out) *groups_out = groups;
  return ctas;
}

int64_t bwd_tc_partial_bytes(const Geo &g0) {
  int sdiv;
  const Geo g = bwd_virtual_geo(g0, &sdiv);
  int ctas = wgrad_ctas(g, false, nullptr, nullptr);
  const int c2 = wgrad_ctas(g, true, nullptr, nullptr);
  if (c2 > ctas) ctas = c2;
  return (int64_t)ctas * g.F * g.Cout * 4;
}

bool bwd_input_tc_can_fold(const Geo &g) {
  // the fold table packs {offset, kx, tap}: 5 bits of tap (K <= 5), offsets below 2^24, table in shared memory
  return g.K <= 5 && (int64_t)g.Cin * g.H * g.W < (1 << 24) && g.F <= 4096;
}

int launch_bwd_input_tc(const Geo &g0, const float *go, const uint32_t *state, const void *wtb, const float *s,
                        const int8_t *mask, float *out, int fold, bool v2, cudaStream_t st) {
  CIMQ_REQUIRE(!v2, "dgrad: the v2 state is handled by launch_bwd_input_v2");
  BwdParams P;
  memset(&P, 0, sizeof(P));
  const Geo g = bwd_virtual_geo(g0, &P.sdiv);
  P.g = g;
  P.v2 = v2 ? 1 : 0;
  P.state2 = reinterpret_cast<const uint8_t *>(state);
  P.Kc = bwd_channel_block(g, v2);
  P.Nf = tc_nf(g);
  P.mtiles = (g.M + kTcTileM - 1) / kTcTileM;
  P.a_bytes = (uint32_t)(kTcTileM * P.Kc * 2);
  P.b_bytes = (uint32_t)(P.Nf * P.Kc * 2);
  P.stage_bytes = 3 * P.a_bytes + P.b_bytes;
  P.fold = fold;
  P.debug = g_tc_debug;
  P.raw_bytes = fold ? (uint32_t)((g.F * 4 + 15) & ~15) : 0u;
  // register-resident operands: 32 channels per producer thread at most, clip bits in one state word
  P.cached = P.Kc <= 64 ? 1 : 0;
  P.pw_off = P.raw_bytes;
  if (P.cached && (g.NSA - 1) * g.NSW + 1 <= 7) P.raw_bytes += (uint32_t)(g.NSW * (1 << ((g.NSA - 1) * g.NSW + 1)) * 4);
  int stages = (int)((kSmemBudget - kBarrierBytes - P.raw_bytes) / P.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  CIMQ_REQUIRE(stages >= 1, "dgrad tile does not fit shared memory");
  P.stages = stages;
  uint32_t cols = 32;
  while (cols < 2u * P.Nf) cols <<= 1;
  P.tmem_cols = cols;
  P.go = go; P.state = state; P.wtb = reinterpret_cast<const uint8_t *>(wtb); P.s = s; P.mask = mask; P.out = out;
  const size_t smem = (size_t)stages * P.stage_bytes + kBarrierBytes + P.raw_bytes + 1024;
  const int grid = P.mtiles < 148 ? P.mtiles : 148;
#define CIMQ_LAUNCH_DGRAD(W, A, T, ...)                                                                         \
  do {                                                                                                          \
    CIMQ_CUDA_OK(cudaFuncSetAttribute(bwd_input_tc_kernel<W, A, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                              \
    bwd_input_tc_kernel<W, A, T><<<grid, kThreads, smem, st>>>(P);                                              \
  } while (0)
  for (P.co0 = 0; P.co0 < g.Cout; P.co0 += P.Kc) {  // one launch per block of <= 128 output channels
    P.wtb = reinterpret_cast<const uint8_t *>(wtb) + (size_t)(P.co0 / P.Kc) * g.NX * g.NSW * P.b_bytes;
    CIMQ_BWD_DISPATCH(CIMQ_LAUNCH_DGRAD, 0);
  }
#undef CIMQ_LAUNCH_DGRAD
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

// Bank layout of the staged activation rows (async staging).  A producer warp gathers 32 consecutive crossbar rows
// = (channel, ky, kx) triples; its lanes read row ky of channel slot (channel - first channel), byte kx + col0 + pixel
// offset.  With a power-of-two row pitch and slot = rk * pitch (64 B, 384 B at 32x32 images) every channel and every
// second ky fell on the same banks: 8.4 wavefronts per load instruction, 16 M conflict wavefronts of 62 M at the
// microbench layer (ncu l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld).  Pitch and slot padding (16-byte steps:
// the cp.async pieces stay aligned) are picked to minimise the worst distinct-words-per-bank count over the warps of
// all crossbars.
static void tune_row_layout(const Geo &g, int rk, int needp, int col0, int *pitch_out, int *slot_out) {
  const int rows_full = g.xbar < g.F ? g.xbar : g.F;
  const int base = (needp + 15) & ~15;
  long best = -1;
  for (int pitch = base; pitch <= base + 48; pitch += 16)
    for (int pad = 0; pad < 128; pad += 16) {
      const int slot = rk * pitch + pad;
      long cost = 0;
      for (int i = 0; i < g.NX; ++i) {
        const int lo = i * g.xbar, rows = rows_full < g.F - lo ? rows_full : g.F - lo, c_lo = lo / g.KK;
        for (int w0 = 0; w0 < rows; w0 += 32)
          for (int word = 0; word < 3; ++word) {  // the three aligned words of an 8-byte window
            int seen[32][32], nseen[32];
            for (int b = 0; b < 32; ++b) nseen[b] = 0;
            int worst = 0;
            for (int r = w0; r < w0 + 32 && r < rows; ++r) {
              const int f = lo + r, ci = f / g.KK - c_lo, tap = f % g.KK, ky = tap / g.K, kx = tap % g.K;
              const int wa = ((ci * slot + ky * pitch + kx + col0) >> 2) + word, b = wa & 31;
              bool dup = false;
              for (int q = 0; q < nseen[b]; ++q) dup = dup || seen[b][q] == wa;
              if (!dup) seen[b][nseen[b]++] = wa;
              if (nseen[b] > worst) worst = nseen[b];
            }
            cost += worst;
          }
      }
      cost = cost * 4096 + slot;  // ties: the smaller buffer
      if (best < 0 || cost < best) { best = cost; *pitch_out = pitch; *slot_out = slot; }
    }
}

int launch_bwd_weight_tc(const Geo &g0, const float *go, const uint8_t *xcodes, const uint32_t *state,
                         const float *s, const int8_t *mask, float *partial, float *gw, bool v2, const void *scales,
                         cudaStream_t st) {
  BwdParams P;
  memset(&P, 0, sizeof(P));
  const Geo g = bwd_virtual_geo(g0, &P.sdiv);
  P.g = g;
  P.v2 = v2 ? 1 : 0;
  P.state2 = reinterpret_cast<const uint8_t *>(state);
  P.state2w = P.state2 + (v2 ? v2::plane_bytes(g0) : 0);
  P.Kc = bwd_channel_block(g, v2);
  P.mtiles = (g.M + kTcTileM - 1) / kTcTileM;
  P.a_bytes = 128u * 128u * 2u;
  P.debug = g_tc_debug;
  P.b_bytes = (uint32_t)(P.Kc / 8) * (16u * (uint32_t)kWgLBO + (v2 ? 16u : 0u));  // 8-row groups x (16 k-groups x 144 B [+ 16])
  P.stage_bytes = P.a_bytes + (v2 ? v2::kBwdPieces : 3) * P.b_bytes;
  CIMQ_REQUIRE(!v2 || scales != nullptr, "wgrad (v2): the grad_out scales are missing");
  if (v2) P.chmax = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(scales) + (((int64_t)g0.M * 4 + 255) & ~(int64_t)255));
  int groups = 1;
  const int ctas = wgrad_ctas(g, v2, &P.nxg, &groups);
  const size_t tbytes = wg_table_bytes(P.nxg);
  // staged activation rows (stride 1, output width a power of two between 8 and 128, 4-byte aligned rows)
  P.fastx = 0; P.raw_bytes = 0;
  if (g.stride == 1 && g.W % 4 == 0 && g.OW >= 8 && g.OW <= kTcTileM && (g.OW & (g.OW - 1)) == 0 &&
      g.pad < g.K) {
    int owl = 0;
    while ((1 << owl) < g.OW) ++owl;
    // rows by 16-byte cp.async copies: 16-byte aligned rows on both sides (W % 16 == 0, image column 0 staged at
    // byte 16); otherwise 4-byte loads with image column -pad on a word boundary
    const bool async_rows = g.W % 8 == 0 && g.pad <= 16 && (reinterpret_cast<uintptr_t>(xcodes) & 15u) == 0;
    const int col0 = async_rows ? 16 - g.pad : (4 - g.pad % 4) % 4;
    const int needp = (g.OW - 1) + g.K + col0 + 4;  // + 4: the unaligned 8-byte window reads one word further
    int pl = 2;
    while ((1 << pl) < needp) ++pl;
    const int rows = g.xbar < g.F ? g.xbar : g.F;
    const int nch = (rows + g.KK - 2) / g.KK + 1;  // channels one crossbar can touch
    // a tile that is 128 consecutive pixels of one image reads consecutive input rows: output rows share them
    const int rpt = kTcTileM / g.OW;
    const bool shared_rows = g.L % kTcTileM == 0;
    const int rk = shared_rows ? rpt - 1 + g.K : rpt * g.K;
    int pitch = 1 << pl, slot = rk * pitch;  // synchronous staging maps threads to the words of a power-of-two row
    if (async_rows) tune_row_layout(g, rk, needp, col0, &pitch, &slot);
    const size_t raw = ((size_t)nch * slot + 8 + 15) & ~(size_t)15;
    if (pl <= 9 && rk <= 128 && 2 * raw + 2 * P.stage_bytes + kBarrierBytes + tbytes + 16 <= kSmemBudget) {
      P.fastx = 1; P.ow_log2 = owl; P.rpt = rpt; P.pitch_log2 = pl; P.col0 = col0; P.pitch = pitch; P.slot_bytes = slot;
      P.rk = rk; P.prow = shared_rows ? 1 : g.K; P.xshared = shared_rows ? 1 : 0; P.async_rows = async_rows ? (g.W % 16 == 0 ? 16 : 8) : 0;
      P.raw_bytes = (uint32_t)raw;
    }
  }
  int stages = (int)((kSmemBudget - kBarrierBytes - 2 * (size_t)P.raw_bytes - tbytes - 16) / P.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  CIMQ_REQUIRE(stages >= 1, "wgrad tile does not fit shared memory");
  P.stages = stages;
  P.gfast = (g.L % 8 == 0 && g.OW % 8 == 0 && g.M % 8 == 0 && g.M >= 8 &&
             (reinterpret_cast<uintptr_t>(go) & 15u) == 0 && (reinterpret_cast<uintptr_t>(state) & 15u) == 0) ? 1 : 0;
  uint32_t cols = 32;
  while (cols < (uint32_t)(P.nxg * P.Kc)) cols <<= 1;
  P.tmem_cols = cols;
  P.go = go; P.state = state; P.xcodes = xcodes; P.s = s; P.mask = mask; P.out = partial;
  P.xt_off = (uint32_t)(((size_t)stages * P.stage_bytes + kBarrierBytes + 2 * (size_t)P.raw_bytes + 15) & ~(size_t)15);
  const size_t smem = (size_t)P.xt_off + tbytes + 1024;
  dim3 grid(ctas, groups);
  if (const char *e = getenv("CIMQ_V2_DBG")) P.dbg = v2 ? atoi(e) : 0;
  if (const char *e = getenv("CIMQ_V2_DBG_WG")) P.dbg = v2 ? atoi(e) : 0;  // (this kernel only)
#define CIMQ_LAUNCH_WGRAD(W, A, T, ...)                                        \
  do {                                                                         \
    if (launch_wgrad_instance<W, A, T>(P, grid, smem, st)) return 1;            \
  } while (0)
  for (P.co0 = 0; P.co0 < g.Cout; P.co0 += P.Kc) {  // one launch per block of <= 128 output channels
    CIMQ_BWD_DISPATCH(CIMQ_LAUNCH_WGRAD, 0);
  }
#undef CIMQ_LAUNCH_WGRAD
  CIMQ_CUDA_OK(cudaGetLastError());
  const int64_t n = (int64_t)g.Cout * g.F;
  bwd_weight_tc_finish_kernel<<<(int)((n + 31) / 32), 256, 0, st>>>(g, ctas, partial, gw);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cimq
