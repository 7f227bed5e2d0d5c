// Second-generation tcgen05 forward kernel of the CiM convolution (get_cim_output_signed.forward, lsq.py:92-237)
// for 1-bit slices: ~4 CUDA-core instructions per partial sum instead of ~16.
//
//   producers (warps 0-7)    the v1 producer (cim_conv_tc_kernel.cuh): im2col digit planes of 128 pixels per crossbar
//                            chunk, written as e4m3 bytes (2.0 for a set bit), weight digits +-0.5.  Two groups of
//                            four warps build alternate stages (a producer thread is latency bound: one warp per
//                            scheduler assembling its pixel's row byte by byte).
//   MMA issuers (16, 18)      GEMM1: tcgen05.mma.kind::f8f6f4 with **fp16 accumulators** -- partial sums
//                            p[128 pixels x NSW*CT] of one activation digit plane, exact integers (|p| <= 128;
//                            the reference stores them as fp16 too, lsq.py:169), two TMEM buffers.
//                            GEMM2 (after the epilogue of that plane): tcgen05.mma.kind::f16 with the **A operand
//                            in TMEM** -- out[128 x CT] += code[128 x (k,c)] * diag(n(i,k,j,c) * 2^(k+j)), one
//                            N=16,K=16 instruction per (weight slice, 16-channel group), fp32 accumulator that lives
//                            in TMEM for the whole tile: the shift-and-add of lsq.py:233 costs no CUDA-core work.
//   epilogue (warps 8-15)    thread = pixel.  tcgen05.ld.pack::16b delivers two fp16 partial sums per register;
//                            |p| (ALU pipe), code = sat(|p| - (tp-1)) with p's sign OR-ed in (FMA + ALU pipe),
//                            clip = sat(|p| - (tg-1)); the ternary codes go straight back to tensor memory, in
//                            place, as GEMM2's A operand (tcgen05.st); the backward's inputs are accumulated as
//                            small integers in packed-half registers (see cim_v2.cuh) and leave as bytes.
//   constants (warp 17)      one bulk copy per chunk: thresholds + B2 slabs -> shared memory (double buffered).
// Measured alternatives (B200, microbench layer; training / inference forward, baseline 334 / 221 us): 16 epilogue warps
// with 16 channels per thread and two producer groups (88 / 72 registers): 338 / 230; 16 epilogue warps and ONE producer
// group (96 / 80 registers): 328 / 251.  The epilogue warps are the busy role (ncu: ~45 % of the stall samples, almost no
// waiting) while the producers wait 40 % of their time, but halving the work per epilogue thread does not shorten the
// per-plane chain tcgen05.ld -> quantise -> tcgen05.st -> fence -> arrive.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "cim_conv_tc_kernel.cuh"
#include "cim_v2.cuh"

namespace cimq {
namespace v2 {

using namespace ptx;
using tcfwd::kProducerThreads;
using tcfwd::kProducerWarps;

constexpr int kProducerGroups = 2;
constexpr int kEpiWarp0 = kProducerGroups * kProducerWarps;  // 8
constexpr int kMmaWarp = kEpiWarp0 + 8, kConstWarp = kMmaWarp + 1, kMma2Warp = kMmaWarp + 2;
constexpr int kThreads = (kMmaWarp + 4) * 32;                // 640: five warpgroups
// Registers per warpgroup (setmaxnreg).  The pool setmaxnreg.inc draws from is what the CTA was LAUNCHED with
// (640 threads x 96 registers, the __launch_bounds__ cap), not the whole register file: the split must satisfy
// 2 x producers + 2 x epilogue + MMA <= 5 x 96 = 480, or the last warpgroup to ask waits forever.
constexpr int kRegsProducer = 96, kRegsMma = 40, kRegsEpilogue = 120;
static_assert(2 * kRegsProducer + 2 * kRegsEpilogue + kRegsMma <= 5 * 96, "setmaxnreg split exceeds the launch allocation");
constexpr int kMaxStages = 4;
constexpr size_t kAuxBytes = 4096;
constexpr size_t kSmemBudget = 227 * 1024 - 1024;

struct V2Params {
  tcfwd::TcParams tc;  // geometry + producer plan (tc.wtiles: e4m3 weight tiles)
  int CT, nct, EW, CH, N1, G;
  uint32_t block_bytes, b2_off;
  uint32_t d2_col;     // first TMEM column of the output accumulators
  int d2_bufs;
  int ngroups;         // producer groups in use (2 needs >= 2 pipeline stages, see make_plan)
  int dbg;             // development only (env CIMQ_V2_DBG): 1 = skip GEMM2 MMAs, 2 = skip GEMM1 MMAs, 4 = skip the
                       // epilogue arithmetic, 8 = producers skip the row assembly -- results are then garbage
  const uint8_t *consts;  // [nct][NX] constants blocks
  const float *oscale;    // {o0, o1}: out = (acc * o0) * o1
  float *out;
  uint8_t *state;         // v2 state planes or NULL (inference)
  float mb_qn, mb_qp, mb_nchi, mb_pclo;  // multi-bit ADC: clamp range, -(chi-1), clo+1
};

struct Bars {
  uint32_t full0, empty0, tfull0, tempty0, a2full0, cfull0, cempty0, d2full0, d2empty0;
};

__device__ __forceinline__ __half2 h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t u32(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }

// One weight slice of one activation plane for this thread's CH channels: partial sums in (packed fp16), codes out
// (packed fp16, in place), state fields updated.  R = CH / 2 registers.
template <int R, bool MB, bool WS>
__device__ __forceinline__ void quantise_slice(uint32_t (&p)[R], const uint4 *thr_tp, const uint4 *thr_tg,
                                               __half2 (&dfld)[R], __half2 (&wfld)[R], __half2 (&cfld)[R],
                                               const __half2 n4k, const __half2 n4j, const __half2 p4k,
                                               const __half2 qn2, const __half2 qp2, const __half2 nchi2,
                                               const __half2 pclo2) {
  const __half2 one2 = __floats2half2_rn(1.0f, 1.0f);
  if constexpr (MB) {
    const __half2 neg2 = __floats2half2_rn(-1.0f, -1.0f);
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const __half2 pv = h2(p[q]);
      const __half2 code = __hmin2(__hmax2(pv, qn2), qp2);  // clamp(p, Qn, Qp), lsq.py:228
      if constexpr (WS) {
        // STE mask off where p >= chi or p <= clo (lsq.py:310-313)
        const __half2 clip = __hadd2(__hfma2_sat(pv, one2, nchi2), __hfma2_sat(pv, neg2, pclo2));
        dfld[q] = __hfma2(clip, n4k, dfld[q]);
        wfld[q] = __hfma2(clip, n4j, wfld[q]);
      }
      p[q] = u32(code);
    }
  } else {
#pragma unroll
    for (int u = 0; u < R / 4; ++u) {
      const uint4 t4 = thr_tp[u];
      const uint32_t tpw[4] = {t4.x, t4.y, t4.z, t4.w};
      uint32_t tgw[4] = {0u, 0u, 0u, 0u};
      if constexpr (WS) {
        const uint4 g4 = thr_tg[u];
        tgw[0] = g4.x; tgw[1] = g4.y; tgw[2] = g4.z; tgw[3] = g4.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int q = 4 * u + e;
        const __half2 a = __habs2(h2(p[q]));
        const __half2 cab = __hfma2_sat(a, one2, h2(tpw[e]));        // 1 where |p| >= tp   (lsq.py:201 / 224)
        const uint32_t code = u32(cab) | (p[q] & 0x80008000u);      // sign of p: ternary / binary code
        if constexpr (WS) {
          const __half2 clip = __hfma2_sat(a, one2, h2(tgw[e]));    // 1 where |p| >= tg   (STE clip, lsq.py:310)
          dfld[q] = __hfma2(clip, n4k, dfld[q]);
          wfld[q] = __hfma2(clip, n4j, wfld[q]);
          cfld[q] = __hfma2(h2(code), p4k, cfld[q]);
        }
        p[q] = code;
      }
    }
  }
}

// low bytes of the four fp16 values 1024 + n (n < 256) held in two registers -> one word of four bytes
__device__ __forceinline__ uint32_t field_bytes(__half2 a, __half2 b) { return __byte_perm(u32(a), u32(b), 0x6420); }

template <int R>
__device__ __forceinline__ void store_fields(uint8_t *dst, const __half2 (&f)[R]) {
  // R registers = 2R channels = 2R bytes
#pragma unroll
  for (int v = 0; v < R / 8; ++v) {
    uint4 w;
    w.x = field_bytes(f[8 * v + 0], f[8 * v + 1]);
    w.y = field_bytes(f[8 * v + 2], f[8 * v + 3]);
    w.z = field_bytes(f[8 * v + 4], f[8 * v + 5]);
    w.w = field_bytes(f[8 * v + 6], f[8 * v + 7]);
    reinterpret_cast<uint4 *>(dst)[v] = w;
  }
  if constexpr (R % 8 == 4) {  // CH = 8: 8 bytes
    uint2 w;
    w.x = field_bytes(f[R - 4], f[R - 3]);
    w.y = field_bytes(f[R - 2], f[R - 1]);
    *reinterpret_cast<uint2 *>(dst + 2 * (R - 4)) = w;
  }
}

// NS: digit planes per operand; CH: channels per epilogue thread; EW: epilogue warpgroups; MB: multi-bit ADC;
// WS: write the ADC state (training)
template <int NS, int CH, int EW, bool MB, bool WS>
__global__ void __launch_bounds__(kThreads, 1) conv_v2_kernel(const V2Params P) {
  constexpr int CT = CH * EW;
  constexpr int N1 = NS * CT;      // GEMM1 N
  constexpr int R = CH / 2;        // packed registers per weight slice
  constexpr int G = CT / 16;       // GEMM2 channel groups
  constexpr int kEpiWarps = 4 * EW;
  const tcfwd::TcParams &T = P.tc;
  const Geo &g = T.g;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: stages | raw x2 | constants x2 | barriers, tmem slot, row table
  tcfwd::Smem sm;
  sm.stage_base = smem_raw;
  const size_t raw_off = (size_t)T.stages * T.stage_bytes;
  sm.raw = smem_raw + raw_off;
  const size_t c_off = (raw_off + 2 * (size_t)P.ngroups * T.raw_bytes + 127) & ~(size_t)127;  // smem_raw: 1024-byte aligned
  uint8_t *cbuf = smem_raw + c_off;
  uint8_t *pp = cbuf + 2 * (size_t)P.block_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(pp);
  Bars B;
  B.full0 = smem_u32(bars);
  B.empty0 = B.full0 + 8 * kMaxStages;
  B.tfull0 = B.empty0 + 8 * kMaxStages;
  B.tempty0 = B.tfull0 + 16;
  B.a2full0 = B.tempty0 + 16;
  B.cfull0 = B.a2full0 + 16;
  B.cempty0 = B.cfull0 + 16;
  B.d2full0 = B.cempty0 + 16;
  B.d2empty0 = B.d2full0 + 16;  // ends at byte 64 + 7*16 = 176
  sm.full0 = B.full0;
  sm.empty0 = B.empty0;
  sm.tfull0 = B.tfull0;
  sm.tempty0 = B.tempty0;
  sm.tmem_slot = reinterpret_cast<uint32_t *>(pp + 192);
  sm.rowoff = reinterpret_cast<int *>(pp + 256);  // per producer group 2 x 128 ints
  sm.ttab = nullptr;
  ChunkLayout *cltab = reinterpret_cast<ChunkLayout *>(pp + 256 + kProducerGroups * 2 * 128 * 4);  // [NX] (<= 64 chunks)
  sm.cltab = g.NX <= (int)((kAuxBytes - 256 - kProducerGroups * 1024) / sizeof(ChunkLayout)) ? cltab : nullptr;
  if (sm.cltab != nullptr && (int)threadIdx.x < g.NX) cltab[threadIdx.x] = chunk_layout(g, threadIdx.x);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // (provably warp-uniform)
  if (threadIdx.x == 0) {
    for (int s = 0; s < T.stages; ++s) {
      mbar_init(B.full0 + 8 * s, kProducerWarps + 1);  // one arrival per producer warp of the group + the weight-tile copy
      mbar_init(B.empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(B.tfull0 + 8 * b, 1);
      mbar_init(B.tempty0 + 8 * b, 1);
      mbar_init(B.a2full0 + 8 * b, kEpiWarps);
      mbar_init(B.cfull0 + 8 * b, 1);
      mbar_init(B.cempty0 + 8 * b, 1 + kEpiWarps);
      mbar_init(B.d2full0 + 8 * b, 1);
      mbar_init(B.d2empty0 + 8 * b, kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(sm.tmem_slot), T.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_slot;
  const int ntiles = T.mtiles * T.nct;
  const int rows_full = g.xbar < g.F ? g.xbar : g.F;

  if (warp < kEpiWarp0) {
    // =========================== producers ===========================
    reg_dealloc<kRegsProducer>();
    const int gidx = warp / kProducerWarps;
    if (gidx < P.ngroups) {
      tcfwd::Smem smg = sm;  // the group's staging buffers and row tables
      smg.raw = sm.raw + (size_t)gidx * 2 * T.raw_bytes;
      smg.rowoff = sm.rowoff + gidx * 256;
      tcfwd::producer_loop<NS, 1, false>(T, smg, ntiles, threadIdx.x & (kProducerThreads - 1), gidx, P.ngroups);
    }
  } else if (warp >= kMmaWarp) {
    reg_dealloc<kRegsMma>();
    if (warp == kMmaWarp) {
      // =========================== GEMM1 issuer ===========================
      // (the whole warp walks the loops -- uniform control flow keeps descriptors and counters in uniform registers, a
      // single-lane branch made the compiler wrap every MMA in an elect / broadcast loop -- and lane 0 issues)
      // partial sums of one activation digit plane per TMEM buffer; runs ahead of the epilogue by one plane.  Its
      // commits track only its own MMAs: a stage is released as soon as the GEMM1s that read it are done.
      const uint32_t idesc1 = idesc_e4m3_f16(kTcTileM, N1);
      const uint32_t sbo = 8u * (uint32_t)T.Kp;
      uint32_t pl = 0;
      int m_sidx = 0;       // pipeline stage and its use count as running counters (no division by a run-time value)
      uint32_t m_use = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int i = 0; i < g.NX; ++i) {
          const int sidx = m_sidx;
          const uint32_t use = m_use;
          if (++m_sidx == T.stages) { m_sidx = 0; ++m_use; }
          const int rows = min(rows_full, g.F - i * g.xbar);
          const int ksteps = (rows + 31) >> 5;
          mbar_wait(B.full0 + 8 * sidx, use & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sm.stage_base + (size_t)sidx * T.stage_bytes);
          const uint32_t b0 = a0 + NS * T.a_bytes;
          for (int j = 0; j < NS; ++j, ++pl) {
            const uint32_t buf = pl & 1;
            mbar_wait(B.tempty0 + 8 * buf, ((pl >> 1) & 1) ^ 1);  // GEMM2 of the plane two back has read its codes
            tc_fence_after();
            const uint32_t d1 = tmem_base + buf * N1;
            if (lane == 0) {
              for (int ks = 0; ks < ksteps && !(P.dbg & 2); ++ks) {
                const uint64_t adesc = make_smem_desc(a0 + j * T.a_bytes + ks * 2 * kTcLBO, kTcLBO, sbo);
                const uint64_t bdesc = make_smem_desc(b0 + ks * 2 * kTcLBO, kTcLBO, sbo);
                umma_f8(d1, adesc, bdesc, idesc1, ks > 0 ? 1u : 0u);
              }
              umma_commit(B.tfull0 + 8 * buf);
              if (j == NS - 1) umma_commit(B.empty0 + 8 * sidx);  // stage consumed -> producers
            }
            __syncwarp();
          }
        }
      }
    } else if (warp == kMma2Warp) {
      // =========================== GEMM2 issuer (whole warp, lane 0 issues) ===========================
      // shift-and-add of one plane as soon as its codes are in tensor memory: a thread of its own, so that waiting
      // for the epilogue never delays the next GEMM1
      const uint32_t idesc2 = idesc_f16_f32(kTcTileM, 16);
      const uint32_t cb_addr = smem_u32(cbuf);
      uint32_t pl = 0, chunk_it = 0, tile_it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
        const uint32_t tb = P.d2_bufs == 2 ? (tile_it & 1) : 0u;
        const uint32_t tuse = P.d2_bufs == 2 ? (tile_it >> 1) : tile_it;
        mbar_wait(B.d2empty0 + 8 * tb, (tuse & 1) ^ 1);  // the tile's output accumulator has been drained
        const uint32_t d2 = tmem_base + P.d2_col + tb * CT;
        for (int i = 0; i < g.NX; ++i, ++chunk_it) {
          const uint32_t cpar = chunk_it & 1;
          mbar_wait(B.cfull0 + 8 * cpar, (chunk_it >> 1) & 1);  // B2 slabs of this chunk
          for (int j = 0; j < NS; ++j, ++pl) {
            const uint32_t buf = pl & 1;
            mbar_wait(B.a2full0 + 8 * buf, (pl >> 1) & 1);      // codes written by the epilogue
            tc_fence_after();
            const uint32_t a2 = tmem_base + buf * N1;
            const uint32_t b2 = cb_addr + cpar * P.block_bytes + P.b2_off + (uint32_t)(j * NS * G) * kSlabBytes;
            if (lane == 0) {
#pragma unroll
              for (int k = 0; k < NS; ++k)
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                  if (P.dbg & 1) continue;
                  const int h = (16 * gi) / CH;
                  const uint32_t acol = a2 + k * CT + h * CH + (16 * gi - h * CH) / 2;
                  const uint64_t bdesc = make_smem_desc(b2 + (uint32_t)(k * G + gi) * kSlabBytes, kTcLBO, 256u);
                  umma_f16_ts(d2 + 16 * gi, acol, bdesc, idesc2, (i | j | k) != 0 ? 1u : 0u);
                }
              umma_commit(B.tempty0 + 8 * buf);                      // partial-sum buffer free again
              if (j == NS - 1) umma_commit(B.cempty0 + 8 * cpar);    // constants of the chunk no longer read by the MMA
              if (i == g.NX - 1 && j == NS - 1) umma_commit(B.d2full0 + 8 * tb);
            }
            __syncwarp();
          }
        }
      }
    } else if (warp == kConstWarp && lane == 0) {
      // =========================== constants loader ===========================
      uint32_t chunk_it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int ct = tile % T.nct;
        for (int i = 0; i < g.NX; ++i, ++chunk_it) {
          const uint32_t cpar = chunk_it & 1;
          mbar_wait(B.cempty0 + 8 * cpar, ((chunk_it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(B.cfull0 + 8 * cpar, P.block_bytes);
          bulk_copy_g2s(smem_u32(cbuf + (size_t)cpar * P.block_bytes),
                        P.consts + (size_t)(ct * g.NX + i) * P.block_bytes, P.block_bytes, B.cfull0 + 8 * cpar);
        }
      }
    }
  } else {
    // =========================== epilogue ===========================
    reg_alloc<kRegsEpilogue>();
    const int e = warp - kEpiWarp0;
    const int quarter = warp & 3;
    const int half = e >> 2;
    if (half < EW) {
      const int r = quarter * 32 + lane;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const float o0 = P.oscale[0], o1 = P.oscale[1];
      const __half2 qn2 = __float2half2_rn(P.mb_qn), qp2 = __float2half2_rn(P.mb_qp);
      const __half2 nchi2 = __float2half2_rn(P.mb_nchi), pclo2 = __float2half2_rn(P.mb_pclo);
      const int64_t plane = plane_bytes(g);
      uint32_t pl = 0, chunk_it = 0, tile_it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
        const int mt = tile / T.nct, ct = tile % T.nct;
        const int m = mt * kTcTileM + r;
        const int c_first = ct * CT + half * CH;
        for (int i = 0; i < g.NX; ++i, ++chunk_it) {
          const uint32_t cpar = chunk_it & 1;
          // this chunk's constants block: thresholds [half][pair][tp | tg][CH] fp16 (none for the multi-bit ADC, whose
          // epilogue still takes part in the hand-shake that recycles the buffer)
          mbar_wait(B.cfull0 + 8 * cpar, (chunk_it >> 1) & 1);
          const uint4 *thr =
              reinterpret_cast<const uint4 *>(cbuf + (size_t)cpar * P.block_bytes) + half * (NS * NS * 2 * CH / 8);
          // fields start at 1024 (fp16 integers 1024..2047 have their value in the low mantissa bits) + "nothing clipped"
          constexpr float kF0 = 1024.0f + NS * (NS == 3 ? 21.0f : 5.0f);
          __half2 dfld[R], wfld[R];
#pragma unroll
          for (int q = 0; q < R; ++q) dfld[q] = wfld[q] = __float2half2_rn(kF0);
          // j is a REAL loop: the body (NS unrolled weight slices) is ~6 KB of code; unrolled over j as well the
          // epilogue alone outgrew the instruction cache and a third of its stall samples were instruction fetches
#pragma unroll 1
          for (int j = 0; j < NS; ++j, ++pl) {
            const uint32_t buf = pl & 1;
            mbar_wait(B.tfull0 + 8 * buf, (pl >> 1) & 1);
            tc_fence_after();
            __half2 cfld[R];
#pragma unroll
            for (int q = 0; q < R; ++q) cfld[q] = __float2half2_rn(1024.0f + (NS == 3 ? 21.0f : 5.0f));
            const __half2 n4j = __float2half2_rn(-(float)(1 << (2 * j)));
            const uint32_t tcol = lane_base + buf * N1 + half * CH;
            const uint4 *thr_j = thr + (j * 2) * (CH / 8);
            // the next weight slice's partial sums are fetched while this one is quantised
            uint32_t pa[R], pb[R];
            tmem_ld_pack16<R>(tcol, pa);
#pragma unroll
            for (int k = 0; k < NS; ++k) {
              uint32_t(&p)[R] = (k & 1) ? pb : pa;
              uint32_t(&pn)[R] = (k & 1) ? pa : pb;
              tmem_ld_wait();
              if (k + 1 < NS) tmem_ld_pack16<R>(tcol + (k + 1) * CT, pn);
              const __half2 n4k = __float2half2_rn(-(float)(1 << (2 * k))), p4k = __float2half2_rn((float)(1 << (2 * k)));
              const uint4 *tp = MB ? nullptr : thr_j + (k * NS * 2) * (CH / 8);
              if (!(P.dbg & 4))
                quantise_slice<R, MB, WS>(p, tp, MB ? nullptr : tp + CH / 8, dfld, wfld, cfld, n4k, n4j, p4k, qn2, qp2,
                                          nchi2, pclo2);
              tmem_st<R>(tcol + k * CT, p);  // codes: GEMM2's A operand, in place (two per 32-bit column)
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(B.a2full0 + 8 * buf);
            if constexpr (WS && !MB) {
              if (m < g.M)
                store_fields<R>(P.state + 2 * plane + (((int64_t)i * NS + j) * g.M + m) * g.Cout + c_first, cfld);
            }
          }
          if constexpr (WS) {
            if (m < g.M) {
              uint8_t *dp = P.state + ((int64_t)i * g.M + m) * g.Cout + c_first;
              store_fields<R>(dp, dfld);
              store_fields<R>(dp + plane, wfld);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(B.cempty0 + 8 * cpar);  // done with this chunk's thresholds
        }
        // ---- the tile's output: accumulated by GEMM2 in tensor memory
        const uint32_t tb = P.d2_bufs == 2 ? (tile_it & 1) : 0u;
        const uint32_t tuse = P.d2_bufs == 2 ? (tile_it >> 1) : tile_it;
        mbar_wait(B.d2full0 + 8 * tb, tuse & 1);
        tc_fence_after();
        const int b = m < g.M ? m / g.L : 0, l = m < g.M ? m % g.L : 0;
        float *op = P.out + ((size_t)b * g.Cout + c_first) * g.L + l;
        constexpr int VW = CH < 16 ? CH : 16;
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += VW) {
          int v[VW];
          tmem_ld<VW>(lane_base + P.d2_col + tb * CT + half * CH + c0, v);
          tmem_ld_wait();
          if (m < g.M) {
#pragma unroll
            for (int cc = 0; cc < VW; ++cc)
              op[(size_t)(c0 + cc) * g.L] = __fmul_rn(__fmul_rn(__int_as_float(v[cc]), o0), o1);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.d2empty0 + 8 * tb);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, T.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static bool make_plan(const Geo &g, V2Params &P, size_t &smem) {
  if (!supported(g)) return false;
  const ConstLayout cl = const_layout(g);
  tcfwd::TcParams &T = P.tc;
  T.g = g;
  T.Kp = tc_kp(g);
  T.CT = cl.CT;
  T.nct = g.Cout / cl.CT;
  T.mtiles = (g.M + kTcTileM - 1) / kTcTileM;
  T.a_bytes = (uint32_t)(kTcTileM * T.Kp);
  T.b_bytes = (uint32_t)(g.NSW * cl.CT * T.Kp);
  T.stage_bytes = g.NSA * T.a_bytes + T.b_bytes;
  T.ttab_bytes = 0;
  P.CT = cl.CT; P.nct = T.nct; P.EW = cl.EW; P.CH = cl.CH; P.N1 = g.NSW * cl.CT; P.G = cl.G;
  P.block_bytes = cl.block_bytes; P.b2_off = cl.b2_off;
  P.d2_col = 2u * P.N1;
  P.d2_bufs = (2 * P.N1 + 2 * cl.CT <= 512) ? 2 : 1;
  uint32_t need = 2u * P.N1 + (uint32_t)P.d2_bufs * cl.CT, cols = 32;
  while (cols < need) cols <<= 1;
  if (cols > 512) return false;
  T.tmem_cols = cols;
  tcfwd::plan_producer(g, T);
  if (g.K != 3) { T.fast = 0; T.raw_bytes = 0; }  // only the 3x3 staged producer is compiled into this kernel
  // Two producer groups build alternate stages.  Every stage slot must then belong to one group or be revisited only
  // after the other group's use of it was consumed (a group that runs two uses ahead on a slot would pass the parity
  // wait of the mbarrier spuriously): true for >= 2 stages, not for 1 -- fall back to a single group then, which
  // also frees two staging buffers.
  // preference: staged producer with two groups, staged with one, generic gather with two, generic with one
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (attempt == 1 && !T.fast) break;  // there was no staged plan to drop
    const bool fast = attempt == 0 && T.fast;
    const size_t raw = fast ? T.raw_bytes : 0;
    for (int groups = kProducerGroups; groups >= 1; --groups) {
      const size_t fixed = 2 * (size_t)groups * raw + 128 + 2 * (size_t)P.block_bytes + kAuxBytes;
      if (fixed + (size_t)(groups > 1 ? 2 : 1) * T.stage_bytes > kSmemBudget) continue;
      int stages = (int)((kSmemBudget - fixed) / T.stage_bytes);
      if (stages > kMaxStages) stages = kMaxStages;
      if (stages > g.NX + 1) stages = g.NX + 1;
      if (groups > 1 && stages < 2) continue;
      if (!fast) { T.fast = 0; T.raw_bytes = 0; }
      T.stages = stages;
      P.ngroups = groups;
      smem = (size_t)stages * T.stage_bytes + fixed + 1024;
      return true;
    }
  }
  return false;
}

template <int NS, int CH, int EW>
static int launch_instance(const V2Params &P, size_t smem, int grid, bool mb, bool ws, cudaStream_t st) {
#define CIMQ_V2_LAUNCH(MB_, WS_)                                                                          \
  do {                                                                                                    \
    CIMQ_CUDA_OK(cudaFuncSetAttribute(conv_v2_kernel<NS, CH, EW, MB_, WS_>,                               \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    conv_v2_kernel<NS, CH, EW, MB_, WS_><<<grid, kThreads, smem, st>>>(P);                                \
  } while (0)
  if (mb && ws) CIMQ_V2_LAUNCH(true, true);
  else if (mb) CIMQ_V2_LAUNCH(true, false);
  else if (ws) CIMQ_V2_LAUNCH(false, true);
  else CIMQ_V2_LAUNCH(false, false);
#undef CIMQ_V2_LAUNCH
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace v2

bool v2_forward_supported(const Geo &g) {
  v2::V2Params P;
  size_t smem;
  return v2::make_plan(g, P, smem);
}

int launch_conv_v2_forward(const Geo &g, const uint8_t *xcodes, const void *wtiles, const void *table, float *out,
                           uint8_t *state, cudaStream_t st) {
  using namespace v2;
  V2Params P;
  memset(&P, 0, sizeof(P));
  size_t smem = 0;
  CIMQ_REQUIRE(make_plan(g, P, smem), "layer not covered by the v2 forward kernel");
  CIMQ_REQUIRE(xcodes && wtiles && table && out, "conv_forward (v2): NULL argument");
  tcfwd::TcParams &T = P.tc;
  const WtLayout wl = wt_layout(g);
  CIMQ_REQUIRE(wl.fwd8_bytes > 0, "conv_forward (v2): weight tiles lack the e4m3 section");
  T.xcodes = xcodes;
  if (T.tma_rows && (reinterpret_cast<uintptr_t>(xcodes) & 15u) != 0) {
    T.tma_rows = 0;
    T.prefetch = (T.raw_bytes / 4 + kProducerThreads - 1) / kProducerThreads <= (uint32_t)tcfwd::kPrefetchWords;
  }
  T.wtiles = reinterpret_cast<const uint8_t *>(wtiles) + wl.fwd8_off;
  T.lut = reinterpret_cast<const int2 *>(reinterpret_cast<const uint8_t *>(wtiles) + wl.lut_off);
  T.table = nullptr; T.ttab = nullptr; T.s = nullptr; T.out = nullptr; T.state = nullptr;
  T.debug = nullptr;
  const uint8_t *sec = reinterpret_cast<const uint8_t *>(table) + table_v2_offset(g);
  P.oscale = reinterpret_cast<const float *>(sec);
  P.consts = sec + 256;
  P.out = out;
  P.state = state;
  const bool mb = g.adc_mode == CIMQ_ADC_MULTIBIT;
  if (mb) {
    P.mb_qn = (float)g.qn; P.mb_qp = (float)g.qp;
    P.mb_nchi = -(float)(g.clip_hi - 1); P.mb_pclo = (float)(g.clip_lo + 1);
  }
  const int ntiles = T.mtiles * T.nct;
  const int grid = ntiles < 148 ? ntiles : 148;
  const bool ws = state != nullptr;
  if (const char *e = getenv("CIMQ_V2_DBG")) {
    P.dbg = atoi(e);
    if (P.dbg & 16) P.ngroups = 1;
    T.debug = (P.dbg & 8) ? reinterpret_cast<long long *>(1) : nullptr;
  }
  if (g.NSW == 3) {
    if (P.CH == 32) return launch_instance<3, 32, 2>(P, smem, grid, mb, ws, st);
    if (P.CH == 16 && P.EW == 2) return launch_instance<3, 16, 2>(P, smem, grid, mb, ws, st);
    return launch_instance<3, 16, 1>(P, smem, grid, mb, ws, st);
  }
  if (P.CH == 32) return launch_instance<2, 32, 2>(P, smem, grid, mb, ws, st);
  if (P.CH == 16 && P.EW == 2) return launch_instance<2, 16, 2>(P, smem, grid, mb, ws, st);
  return launch_instance<2, 16, 1>(P, smem, grid, mb, ws, st);
}

}  // namespace cimq
