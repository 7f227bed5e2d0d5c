// Second-generation tcgen05 dgrad of the CiM convolution (get_cim_output_signed.backward, lsq.py:244-386; the
// grad-input half: lsq.py:298-313, 371-383) for layers whose forward stored the v2 byte-plane state (cim_v2.cuh), plus
// the grad_out scale pre-pass both v2 backward kernels share.
//
//   grad_xunf[m, f] = s_w / NSA * sum_{k, co} go[m, co] * cnt_k[i(f), m, co] * 2^k * w_k[f, co]      (then nn.Fold)
//   cnt_k = number of activation slices j whose partial sum (i, k, j) was NOT clipped: plane D of the v2 state.
//
// GEMM per (128-pixel tile, crossbar chunk i, weight slice k):  D[128 x Nf] += A'_k[128 x Cout] * B_k[Nf x Cout]^T
//   A'_k[m, co] = (go[m, co] * 2^s_m as two fp16 pieces) * cnt_k        exact products, see cim_v2.cuh
//   B_k[n, co]  = 2^k * w_k[f(n), co] in fp16, columns in dgrad_col_f order (complete kernel rows first)
//
// Warp roles (20 warps):
//   0-7   producers: thread = (pixel, channel half).  grad_out of the tile is split into pieces ONCE (it is reused by
//         all NX*NSW stages); a stage costs one LOP3 + one HFMA2 (count as fp16x2 out of the D byte) + two HMUL2 per
//         channel pair and one 16-byte store per 8 channels and piece.
//   8-15  epilogue: two warps per TMEM lane quarter.  For 3x3 / stride 1 / pad 1 layers nn.Fold is done in registers:
//         the three taps of a kernel row land on horizontally adjacent pixels = adjacent lanes, so a tuple is combined
//         with two shuffles and leaves as ONE fp32 reduction per pixel (lanes = consecutive addresses; grad_x stays in
//         L2 while it is accumulated).  Other geometries: one predicated reduction per unfolded element.
//   16    MMA issuer (tcgen05.mma.kind::f16, fp32 accumulators, up to three TMEM buffers); 17: bulk-copy loader of the D bytes.
// Measured dead ends (B200, microbench layer, kept out of the code): 16-byte vector reductions after a 4x4 lane transpose
// (same time), staging the folded sums in shared memory and reducing through cp.reduce.async.bulk (+25 %), and a vertical
// fold in shared memory that halves the number of reductions (+35 %: the extra barrier and one pipeline stage less cost
// more than the reductions saved).
#include <stdlib.h>
#include <string.h>

#include "cim_tc_layout.cuh"
#include "cim_v2.cuh"
#include "tc_ptx.cuh"

namespace cimq {
namespace v2 {
namespace {

using namespace ptx;

constexpr int kDgProducerThreads = 256;
constexpr int kDgEpiWarp0 = 8, kDgEpiWarps = 8;
constexpr int kDgMmaWarp = 16;
constexpr int kDgThreads = 640;
// setmaxnreg split: the pool is what the CTA was launched with (640 x 96)
constexpr int kDgRegsProducer = 128, kDgRegsEpilogue = 88, kDgRegsMma = 40;
static_assert(2 * kDgRegsProducer + 2 * kDgRegsEpilogue + kDgRegsMma <= 5 * 96, "setmaxnreg split exceeds the launch allocation");
constexpr int kDgMaxStages = 4;
constexpr size_t kDgSmemBudget = 227 * 1024 - 1024;
constexpr int kTupStride = 48;  // tuple table entries per chunk: two runs of 24 (even / odd tuples, one per epilogue warp of a quarter)

struct DgParams {
  Geo g;
  int Kc, Nf, mtiles, stages;
  int nbuf;  // accumulator buffers in tensor memory (2..4)
  uint32_t a_bytes, b_bytes, stage_bytes, tmem_cols;
  uint32_t aux_off, ftab_off, tup_off;  // byte offsets inside the dynamic shared memory
  int co0;
  int fold;  // 1: fold into grad_x with fp32 reductions; 0: write gxu[b][f][l] (deterministic path, then col2im)
  int fast;  // fold in registers (3x3, stride 1, pad 1, output width 8 / 16 / 32)
  int dbulk; // the D bytes of a (tile, chunk) are one contiguous run (Cout == Kc): a loader thread brings them to shared
             // memory by bulk copy, so that no global load of the producers is in flight at their proxy fences
  uint32_t dbuf_off;
  int dbg;   // development only (env CIMQ_V2_DBG): 1 = no MMAs, 2 = epilogue drains nothing, 4 = producers store nothing, 8 = no weight-tile copies, 16 = no grad_out loads / split
  const uint8_t *stateD;
  const float *go, *rowscale, *s;
  const uint8_t *wtb;
  float *out;
};

__device__ __forceinline__ __half2 h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t u32(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }

// (v0, v1) -> two fp16x2 words with p1 + p2 == v rounded to 2*PB significant bits (v0 in the low half)
template <int PB>
__device__ __forceinline__ void split_pieces(float v0, float v1, uint32_t &p1, uint32_t &p2) {
  constexpr uint32_t kMask = 0xffffffffu << (24 - PB), kRnd = 1u << (23 - PB);
  const float a0 = __uint_as_float((__float_as_uint(v0) + kRnd) & kMask);
  const float a1 = __uint_as_float((__float_as_uint(v1) + kRnd) & kMask);
  const float r0 = v0 - a0, r1 = v1 - a1;
  const float b0 = __uint_as_float((__float_as_uint(r0) + kRnd) & kMask);
  const float b1 = __uint_as_float((__float_as_uint(r1) + kRnd) & kMask);
  p1 = u32(__floats2half2_rn(a0, a1));
  p2 = u32(__floats2half2_rn(b0, b1));
}

__device__ __forceinline__ void red_add_pred(float *addr, float v, uint32_t pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "@p red.global.add.f32 [%0], %1;\n\t}"
      ::"l"(addr), "f"(v), "r"(pred));  // (no "memory" clobber: nothing in this kernel reads grad_x back)
}

// ---------------------------------------------------------------------------------------------------------------
// grad_out scales: rowscale[m] = 2^s with max_co |go[m, co]| * 2^s in [2^12, 2^13); chmax[co] = bit pattern of
// max_m |go[m, co]| (atomicMax; zeroed by the caller).  One pass over grad_out, thread = pixel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kScaleThreads = 128;  // thread = four consecutive pixels (16-byte loads); small blocks fill the SMs on small layers
__global__ void __launch_bounds__(kScaleThreads) go_scales_kernel(Geo g, const float *__restrict__ go, float *__restrict__ rowscale,
                                                                  uint32_t *__restrict__ chmax) {
  extern __shared__ uint32_t sm_ch[];  // [Cout]
  for (int c = threadIdx.x; c < g.Cout; c += kScaleThreads) sm_ch[c] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // channel maxima stay in registers: lane l of a warp keeps the running maximum of channels c = l (mod 32)
  constexpr int kCB = 16;  // 32-channel blocks (Cout <= 512)
  uint32_t chreg[kCB];
#pragma unroll
  for (int cb = 0; cb < kCB; ++cb) chreg[cb] = 0u;
  const int m4s = g.M >> 2;  // pixel quads (L % 4 == 0: a quad lies inside one image)
  for (int q0 = blockIdx.x * kScaleThreads; q0 < m4s; q0 += gridDim.x * kScaleThreads) {
    const int q = q0 + threadIdx.x;
    const bool live = q < m4s;
    const int m = live ? 4 * q : 0;
    const int b = m / g.L, l = m % g.L;
    const float4 *gp = reinterpret_cast<const float4 *>(go + (int64_t)b * g.Cout * g.L + l);
    const int64_t cstride = g.L >> 2;  // float4 per channel row
    uint32_t rm0 = 0u, rm1 = 0u, rm2 = 0u, rm3 = 0u;
#pragma unroll
    for (int cb = 0; cb < kCB; ++cb) {
      if (cb * 32 < g.Cout) {
#pragma unroll
        for (int c8 = 0; c8 < 32; c8 += 16) {
          if (cb * 32 + c8 < g.Cout) {  // Cout % 16 == 0
            float4 v[16];  // sixteen 16-byte loads in flight per thread: the kernel is latency bound
#pragma unroll
            for (int e = 0; e < 16; ++e)
              v[e] = live ? __ldg(gp + (int64_t)(cb * 32 + c8 + e) * cstride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const uint32_t a0 = __float_as_uint(v[e].x) & 0x7fffffffu, a1 = __float_as_uint(v[e].y) & 0x7fffffffu;
              const uint32_t a2 = __float_as_uint(v[e].z) & 0x7fffffffu, a3 = __float_as_uint(v[e].w) & 0x7fffffffu;
              rm0 = max(rm0, a0); rm1 = max(rm1, a1); rm2 = max(rm2, a2); rm3 = max(rm3, a3);
              const uint32_t w = __reduce_max_sync(0xffffffffu, max(max(a0, a1), max(a2, a3)));
              if (lane == c8 + e) chreg[cb] = max(chreg[cb], w);
            }
          }
        }
      }
    }
    if (live)
      *reinterpret_cast<float4 *>(rowscale + m) = make_float4(bwd_scale_from_maxbits(rm0), bwd_scale_from_maxbits(rm1),
                                                              bwd_scale_from_maxbits(rm2), bwd_scale_from_maxbits(rm3));
  }
#pragma unroll
  for (int cb = 0; cb < kCB; ++cb)
    if (cb * 32 + lane < g.Cout && chreg[cb] != 0u) atomicMax(&sm_ch[cb * 32 + lane], chreg[cb]);
  __syncthreads();
  for (int c = threadIdx.x; c < g.Cout; c += kScaleThreads)
    if (sm_ch[c] != 0u) atomicMax(&chmax[c], sm_ch[c]);
}

// Small layers (fewer than ~300 blocks of 512 pixels): the four warps of a block share 32 pixel quads and split the
// CHANNELS, four times as many blocks and a quarter of the loads per thread -- the 8x8 / 16x16 layers of a ResNet at
// batch 256 were 32 / 128 blocks of the kernel above, 15 us of load latency each.  E = loads in flight per batch.
template <int E>
__global__ void __launch_bounds__(kScaleThreads) go_scales_split_kernel(Geo g, const float *__restrict__ go,
                                                                        float *__restrict__ rowscale,
                                                                        uint32_t *__restrict__ chmax) {
  __shared__ uint4 rm_s[4][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int cpw = g.Cout >> 2;  // channels per warp (a multiple of E)
  constexpr int kCB = 4;        // 32-channel blocks per warp (Cout <= 512)
  uint32_t chreg[kCB];
#pragma unroll
  for (int cb = 0; cb < kCB; ++cb) chreg[cb] = 0u;
  const int m4s = g.M >> 2;
  for (int q0 = blockIdx.x * 32; q0 < m4s; q0 += gridDim.x * 32) {
    const int q = q0 + lane;
    const bool live = q < m4s;
    const int m = live ? 4 * q : 0;
    const int b = m / g.L, l = m % g.L;
    const float4 *gp = reinterpret_cast<const float4 *>(go + ((int64_t)b * g.Cout + w * cpw) * g.L + l);
    const int64_t cstride = g.L >> 2;
    uint32_t rm0 = 0u, rm1 = 0u, rm2 = 0u, rm3 = 0u;
#pragma unroll
    for (int cb = 0; cb < kCB; ++cb) {
      if (cb * 32 < cpw) {
#pragma unroll
        for (int c8 = 0; c8 < 32; c8 += E) {
          if (cb * 32 + c8 < cpw) {
            float4 v[E];
#pragma unroll
            for (int e = 0; e < E; ++e)
              v[e] = live ? __ldg(gp + (int64_t)(cb * 32 + c8 + e) * cstride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int e = 0; e < E; ++e) {
              const uint32_t a0 = __float_as_uint(v[e].x) & 0x7fffffffu, a1 = __float_as_uint(v[e].y) & 0x7fffffffu;
              const uint32_t a2 = __float_as_uint(v[e].z) & 0x7fffffffu, a3 = __float_as_uint(v[e].w) & 0x7fffffffu;
              rm0 = max(rm0, a0); rm1 = max(rm1, a1); rm2 = max(rm2, a2); rm3 = max(rm3, a3);
              const uint32_t wm = __reduce_max_sync(0xffffffffu, max(max(a0, a1), max(a2, a3)));
              if (lane == c8 + e) chreg[cb] = max(chreg[cb], wm);
            }
          }
        }
      }
    }
    rm_s[w][lane] = make_uint4(rm0, rm1, rm2, rm3);
    __syncthreads();
    if (w == 0 && live) {
      uint4 a = rm_s[0][lane];
#pragma unroll
      for (int ww = 1; ww < 4; ++ww) {
        const uint4 o = rm_s[ww][lane];
        a.x = max(a.x, o.x); a.y = max(a.y, o.y); a.z = max(a.z, o.z); a.w = max(a.w, o.w);
      }
      *reinterpret_cast<float4 *>(rowscale + m) = make_float4(bwd_scale_from_maxbits(a.x), bwd_scale_from_maxbits(a.y),
                                                              bwd_scale_from_maxbits(a.z), bwd_scale_from_maxbits(a.w));
    }
    __syncthreads();
  }
#pragma unroll
  for (int cb = 0; cb < kCB; ++cb)
    if (cb * 32 + lane < cpw && chreg[cb] != 0u) atomicMax(&chmax[w * cpw + cb * 32 + lane], chreg[cb]);
}

// ---------------------------------------------------------------------------------------------------------------
// dgrad
// ---------------------------------------------------------------------------------------------------------------
struct DgSmem {
  uint8_t *stage_base;
  uint32_t full0, empty0, tfull0, tempty0, dfull0, dempty0;
  uint32_t *tmem_slot;
  uint8_t *dbuf;  // [2][128 * Kc] D bytes of the current / next chunk (dbulk)
  int *ftab;  // [F]: fold entry of unfold row f: offset inside the image << 7 | kx << 5 | tap
  int *tup;   // [NX][2][24]: complete tuple tt = 2v + eh of chunk i at [i][eh][v]: ((ci * H + ky) * W) << 2 | ky
  DgradCols *dcols;  // [NX] column layout of every chunk (computed once: three run-time divisions each)
};

// NS: digit planes per operand; CPT: channels per producer thread (Kc / 2)
template <int NS, int CPT>
__global__ void __launch_bounds__(kDgThreads, 1) bwd_input_v2_kernel(const DgParams P) {
  constexpr int PB = bwd_piece_bits(NS);
  constexpr int NPAIR = CPT / 2;
  constexpr int NW = CPT / 4;  // D words per thread and chunk
  const Geo &g = P.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  DgSmem sm;
  sm.stage_base = smem_raw;
  {
    uint8_t *aux = smem_raw + P.aux_off;
    sm.full0 = smem_u32(aux);
    sm.empty0 = sm.full0 + 8 * kDgMaxStages;
    sm.tfull0 = sm.empty0 + 8 * kDgMaxStages;
    sm.tempty0 = sm.tfull0 + 32;
    sm.dfull0 = sm.tempty0 + 32;
    sm.dempty0 = sm.dfull0 + 16;  // ends at byte 160
    sm.tmem_slot = reinterpret_cast<uint32_t *>(aux + 160);
    sm.dbuf = smem_raw + P.dbuf_off;
    sm.ftab = reinterpret_cast<int *>(smem_raw + P.ftab_off);
    sm.tup = reinterpret_cast<int *>(smem_raw + P.tup_off);
    sm.dcols = reinterpret_cast<DgradCols *>(smem_raw + P.tup_off + (size_t)g.NX * kTupStride * 4);
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // (provably warp-uniform)
  const int Kc = P.Kc, Nf = P.Nf;
  const uint32_t sbo = (uint32_t)Kc * 16u;  // 8 rows x Kc fp16

  if (threadIdx.x == 0) {
    for (int sidx = 0; sidx < P.stages; ++sidx) {
      mbar_init(sm.full0 + 8 * sidx, kDgProducerThreads / 32 + 1);  // one arrival per producer warp + the weight-tile copy
      mbar_init(sm.empty0 + 8 * sidx, 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(sm.tfull0 + 8 * b, 1);
      mbar_init(sm.tempty0 + 8 * b, kDgEpiWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sm.dfull0 + 8 * b, 1);
      mbar_init(sm.dempty0 + 8 * b, kDgProducerThreads / 32);
    }
    fence_barrier_init();
  }
  for (int f = threadIdx.x; f < g.F; f += kDgThreads) {  // nn.Unfold row order (lsq.py:141)
    const int ci = f / g.KK, tap = f % g.KK, ky = tap / g.K, kx = tap % g.K;
    sm.ftab[f] = (((ci * g.H + ky) * g.W + kx) << 7) | (kx << 5) | tap;
  }
  for (int e = threadIdx.x; e < g.NX * kTupStride; e += kDgThreads) {
    const int i = e / kTupStride, eh = (e % kTupStride) / 24, v = e % 24;
    const DgradCols dc = dgrad_cols(g, i);
    const int tt = 2 * v + eh;
    int val = 3;  // "no such tuple": kernel row 3 never passes the row-validity test
    if (tt < dc.ntup) {
      const int t = dc.t0 + tt, ci = t / g.K, ky = t % g.K;
      val = (((ci * g.H + ky) * g.W) << 2) | ky;
    }
    sm.tup[e] = val;
  }
  if ((int)threadIdx.x < g.NX) sm.dcols[threadIdx.x] = dgrad_cols(g, threadIdx.x);
  if (warp == kDgMmaWarp) tmem_alloc(smem_u32(sm.tmem_slot), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_slot;

  if (warp < kDgEpiWarp0) {
    // =========================== producers ===========================
    reg_alloc<kDgRegsProducer>();
    const int r = threadIdx.x & 127, h = threadIdx.x >> 7;
    uint32_t p1[NPAIR], p2[NPAIR], sp[NPAIR], dnx[NW];
    float gnx[CPT], rs_n = 1.0f;
    // rows past the last pixel read pixel 0: their A' rows only feed accumulator rows the epilogue never stores
    auto tile_ptrs = [&](int mt, const float *&gop, const uint8_t *&dp, const float *&rsp) {
      const int64_t m = (int64_t)mt * kTcTileM + r;
      const bool live = m < g.M;
      const int b = live ? (int)(m / g.L) : 0, l = live ? (int)(m % g.L) : 0;
      gop = P.go + ((int64_t)b * g.Cout + P.co0 + h * CPT) * g.L + l;
      dp = P.stateD + (live ? m : 0) * (int64_t)g.Cout + P.co0 + h * CPT;
      rsp = P.rowscale + (live ? m : 0);
    };
    auto load_go = [&](const float *gop, const float *rsp) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) gnx[c] = __ldg(gop + (size_t)c * g.L);
      rs_n = __ldg(rsp);
    };
    auto load_d = [&](const uint8_t *dp, int i) {
      const uint8_t *wp = dp + (int64_t)i * g.M * g.Cout;
      if constexpr (NW == 8) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(wp)), b4 = __ldg(reinterpret_cast<const uint4 *>(wp) + 1);
        dnx[0] = a.x; dnx[1] = a.y; dnx[2] = a.z; dnx[3] = a.w; dnx[4] = b4.x; dnx[5] = b4.y; dnx[6] = b4.z; dnx[7] = b4.w;
      } else if constexpr (NW == 4) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(wp));
        dnx[0] = a.x; dnx[1] = a.y; dnx[2] = a.z; dnx[3] = a.w;
      } else {
        const uint2 a = __ldg(reinterpret_cast<const uint2 *>(wp));
        dnx[0] = a.x; dnx[1] = a.y;
      }
    };
    const float *gop = nullptr, *gop_n = nullptr, *rsp = nullptr, *rsp_n = nullptr;
    const uint8_t *dp = nullptr, *dp_n = nullptr;
#pragma unroll
    for (int w = 0; w < NW; ++w) dnx[w] = 0u;
#pragma unroll
    for (int c = 0; c < CPT; ++c) gnx[c] = 0.0f;
    if ((int)blockIdx.x < P.mtiles) {
      tile_ptrs(blockIdx.x, gop, dp, rsp);
      load_go(gop, rsp);
      if (!P.dbulk) load_d(dp, 0);
    }
    uint32_t chunk_it = 0;
    int p_sidx = 0;       // pipeline stage and its phase as running counters
    uint32_t p_phase = 0;
    for (int mt = blockIdx.x; mt < P.mtiles; mt += gridDim.x) {
      const int nmt = mt + gridDim.x;
      const bool more_tiles = nmt < P.mtiles;
      if (more_tiles) tile_ptrs(nmt, gop_n, dp_n, rsp_n);
      // grad_out of the tile -> scaled fp16 pieces (once per tile)
      if (!(P.dbg & 16)) {
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) split_pieces<PB>(gnx[2 * q] * rs_n, gnx[2 * q + 1] * rs_n, p1[q], p2[q]);
      }
      for (int i = 0; i < g.NX; ++i, ++chunk_it) {
        if (P.dbulk) {  // this chunk's D bytes were brought to shared memory by the loader thread
          const uint32_t db = chunk_it & 1;
          mbar_wait(sm.dfull0 + 8 * db, (chunk_it >> 1) & 1);
          const uint8_t *src = sm.dbuf + (size_t)db * (kTcTileM * Kc) + r * Kc + h * CPT;
          if constexpr (NW == 8) {
            const uint4 a = *reinterpret_cast<const uint4 *>(src), b4 = *reinterpret_cast<const uint4 *>(src + 16);
            dnx[0] = a.x; dnx[1] = a.y; dnx[2] = a.z; dnx[3] = a.w; dnx[4] = b4.x; dnx[5] = b4.y; dnx[6] = b4.z; dnx[7] = b4.w;
          } else if constexpr (NW == 4) {
            const uint4 a = *reinterpret_cast<const uint4 *>(src);
            dnx[0] = a.x; dnx[1] = a.y; dnx[2] = a.z; dnx[3] = a.w;
          } else {
            const uint2 a = *reinterpret_cast<const uint2 *>(src);
            dnx[0] = a.x; dnx[1] = a.y;
          }
        }
        // D bytes of two adjacent channels spread to the two halves of a register: [b(c), 0, b(c+1), 0]
#pragma unroll
        for (int q = 0; q < NPAIR; ++q) sp[q] = __byte_perm(dnx[q >> 1], 0u, (q & 1) ? 0x4342 : 0x4140);
        if (P.dbulk) {
          __syncwarp();
          if (lane == 0) mbar_arrive(sm.dempty0 + 8 * (chunk_it & 1));
        }
#pragma unroll
        for (int k = 0; k < NS; ++k) {
          const int sidx = p_sidx;
          const uint32_t use = p_phase;
          if (++p_sidx == P.stages) { p_sidx = 0; p_phase ^= 1u; }
          if (k == NS - 1 && !P.dbulk) {  // the D bytes of the next chunk (or of the next tile's first chunk) start their trip
            if (i + 1 < g.NX) load_d(dp, i + 1);
            else if (more_tiles) load_d(dp_n, 0);
          }
          if (k == 0 && i + 1 == g.NX && more_tiles && !(P.dbg & 16)) load_go(gop_n, rsp_n);  // next tile's grad_out: a chunk ahead
          mbar_wait(sm.empty0 + 8 * sidx, (use & 1) ^ 1);
          uint8_t *st_ptr = sm.stage_base + (size_t)sidx * P.stage_bytes;
          if (threadIdx.x == 0) {
            if (P.dbg & 8) mbar_arrive(sm.full0 + 8 * sidx);
            else {
              mbar_arrive_expect_tx(sm.full0 + 8 * sidx, P.b_bytes);
              bulk_copy_g2s(smem_u32(st_ptr + kBwdPieces * (size_t)P.a_bytes), P.wtb + (size_t)(i * NS + k) * P.b_bytes,
                            P.b_bytes, sm.full0 + 8 * sidx);
            }
          }
          // count field k of both bytes -> fp16x2 (1024 + cnt * 4^k) -> cnt
          const uint32_t fmask = 0x00030003u << (2 * k);
          const __half2 sk = __float2half2_rn(1.0f / (float)(1 << (2 * k)));
          const __half2 ok = __float2half2_rn(-1024.0f / (float)(1 << (2 * k)));
#pragma unroll
          for (int cg8 = 0; cg8 < CPT / 8; ++cg8) {
            if (P.dbg & 4) break;
            uint32_t a1[4], a2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int q = 4 * cg8 + e;
              const __half2 cnt = __hfma2(h2((sp[q] & fmask) | 0x64006400u), sk, ok);
              a1[e] = u32(__hmul2(h2(p1[q]), cnt));
              a2[e] = u32(__hmul2(h2(p2[q]), cnt));
            }
            const uint32_t off = tc_tile_offset16(r, h * CPT + cg8 * 8, kTcLBO, sbo);
            *reinterpret_cast<uint4 *>(st_ptr + off) = make_uint4(a1[0], a1[1], a1[2], a1[3]);
            *reinterpret_cast<uint4 *>(st_ptr + P.a_bytes + off) = make_uint4(a2[0], a2[1], a2[2], a2[3]);
          }
          fence_proxy_async();  // every thread publishes its own stores to the async proxy, then one arrival per warp
          __syncwarp();
          if (lane == 0) mbar_arrive(sm.full0 + 8 * sidx);
        }
      }
      gop = gop_n; dp = dp_n; rsp = rsp_n;
    }
  } else if (warp >= kDgMmaWarp) {
    reg_dealloc<kDgRegsMma>();
    if (warp == kDgMmaWarp) {
      // =========================== MMA issuer (whole warp walks the loops, lane 0 issues) ===========================
      const uint32_t idesc = idesc_f16_f32(kTcTileM, Nf);
      const int ksteps = Kc >> 4;
      uint32_t a_buf = 0, a_use = 0;  // accumulator buffer and how often it has been used
      int m_sidx = 0;
      uint32_t m_phase = 0;
      for (int mt = blockIdx.x; mt < P.mtiles; mt += gridDim.x) {
        for (int i = 0; i < g.NX; ++i) {
          const uint32_t buf = a_buf, buse = a_use;
          if (++a_buf == (uint32_t)P.nbuf) { a_buf = 0; ++a_use; }
          mbar_wait<CIMQ_MMA_SLEEP>(sm.tempty0 + 8 * buf, (buse & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * Nf;
          for (int k = 0; k < NS; ++k) {
            const int sidx = m_sidx;
            const uint32_t use = m_phase;
            if (++m_sidx == P.stages) { m_sidx = 0; m_phase ^= 1u; }
            mbar_wait<CIMQ_MMA_SLEEP>(sm.full0 + 8 * sidx, use & 1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sm.stage_base + (size_t)sidx * P.stage_bytes);
            const uint32_t b0 = a0 + kBwdPieces * P.a_bytes;
            if (lane == 0) {
#pragma unroll
              for (int pc = 0; pc < kBwdPieces; ++pc)
                for (int ks = 0; ks < ksteps && !(P.dbg & 1); ++ks) {
                  const uint64_t adesc = make_smem_desc(a0 + pc * P.a_bytes + ks * 2 * kTcLBO, kTcLBO, sbo);
                  const uint64_t bdesc = make_smem_desc(b0 + ks * 2 * kTcLBO, kTcLBO, sbo);
                  umma_f16(d_tmem, adesc, bdesc, idesc, (k | pc | ks) != 0 ? 1u : 0u);
                }
              umma_commit(sm.empty0 + 8 * sidx);
              if (k == NS - 1) umma_commit(sm.tfull0 + 8 * buf);
            }
            __syncwarp();
          }
        }
      }
    }
    else if (warp == kDgMmaWarp + 1 && lane == 0 && P.dbulk) {
      // =========================== D-byte loader ===========================
      // (Measured dead end: the producers' 16-byte reads of these rows are 4-way bank conflicted at Kc = 64 (row pitch
      // 64 bytes; ncu: 16 wavefronts per LDS.128).  One bulk copy PER ROW into rows of Kc + 16 bytes removes the
      // conflicts and costs far more than it saves: 128 small copies per chunk, dgrad 235 -> 350 us.)
      uint32_t chunk_it = 0;
      for (int mt = blockIdx.x; mt < P.mtiles; mt += gridDim.x) {
        const int64_t m0 = (int64_t)mt * kTcTileM;
        const uint32_t bytes = (uint32_t)(min((int64_t)kTcTileM, g.M - m0) * Kc);
        for (int i = 0; i < g.NX; ++i, ++chunk_it) {
          const uint32_t db = chunk_it & 1;
          mbar_wait<100>(sm.dempty0 + 8 * db, ((chunk_it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(sm.dfull0 + 8 * db, bytes);
          bulk_copy_g2s(smem_u32(sm.dbuf + (size_t)db * (kTcTileM * Kc)), P.stateD + ((int64_t)i * g.M + m0) * g.Cout,
                        bytes, sm.dfull0 + 8 * db);
        }
      }
    }
  } else {
    // =========================== epilogue ===========================
    reg_dealloc<kDgRegsEpilogue>();
    const int e = warp - kDgEpiWarp0;
    const int quarter = warp & 3, eh = e >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sw_ns = P.s[1] / (float)NS;  // w_sl * s_w (lsq.py:252), mean over the activation slices (lsq.py:376)
    // fast fold: lanes are consecutive pixels of image rows; neighbours inside the row take part in the shuffles
    const int oxl = lane & (g.OW - 1);
    const float ml = (P.fast && oxl >= 1) ? 1.0f : 0.0f, mr = (P.fast && oxl + 1 < g.OW) ? 1.0f : 0.0f;
    uint32_t a_buf = 0, a_use = 0;
    for (int mt = blockIdx.x; mt < P.mtiles; mt += gridDim.x) {
      const int64_t m = (int64_t)mt * kTcTileM + r;
      const bool live = m < g.M;
      const int eb = live ? (int)(m / g.L) : 0, el = live ? (int)(m % g.L) : 0;
      const int oy = el / g.OW, ox = el % g.OW;
      const float scale = live ? sw_ns * bwd_scale_inverse(__ldg(P.rowscale + m)) : 0.0f;
      // scales of the horizontally adjacent pixels (zero where the neighbour is in another image row)
      const float sc_r = __shfl_down_sync(0xffffffffu, scale, 1) * mr, sc_l = __shfl_up_sync(0xffffffffu, scale, 1) * ml;
      // fold: taps of this pixel that land inside the image, and the address of tap (0, 0) of channel 0
      const int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
      uint32_t vm = 0, vy = 0;
      if (P.fold && live) {
        for (int ky = 0; ky < g.K; ++ky) {
          if (iy0 + ky >= 0 && iy0 + ky < g.H) vy |= 1u << ky;
          for (int kx = 0; kx < g.K; ++kx)
            if (iy0 + ky >= 0 && iy0 + ky < g.H && ix0 + kx >= 0 && ix0 + kx < g.W) vm |= 1u << (ky * g.K + kx);
        }
      }
      float *gxp = P.out + ((int64_t)eb * g.Cin * g.H + iy0) * g.W + ix0;  // generic: + (ci*H + ky)*W + kx
      float *gxc = gxp + g.pad;                                             // fast: + (ci*H + ky)*W, column ox
      for (int i = 0; i < g.NX; ++i) {
        const uint32_t buf = a_buf, buse = a_use;
        if (++a_buf == (uint32_t)P.nbuf) { a_buf = 0; ++a_use; }
        const DgradCols dc = sm.dcols[i];
        const int ntup = (P.dbg & 2) ? 0 : (P.fast ? dc.ntup : 0);
        const int *tup = sm.tup + i * kTupStride + eh * 24;
        mbar_wait(sm.tfull0 + 8 * buf, buse & 1);
        tc_fence_after();
        const uint32_t tcol = lane_base + buf * Nf;
        // ---- complete kernel rows: both warps of the quarter read the batch, each folds every other tuple.  Written
        // in phases (table entries, tensor-memory loads, all shuffles, all reductions) so that the latencies overlap:
        // shuffles and reductions are ordered among themselves and would otherwise alternate.
        for (int t0 = 0; t0 < ntup; t0 += 16) {
          // warp eh reads from column 3 * (t0 + eh) on: its tuples t0 + eh + 2u then sit at the static local columns 6u
          // (reads may run past the accumulator into allocated, unused tensor memory: see tmem_cols on the host side)
          int va[16], vb[16], vc[16];
          const int c0 = 3 * (t0 + eh), cend = 3 * ntup;
          tmem_ld<16>(tcol + c0, va);
          if (c0 + 16 < cend) tmem_ld<16>(tcol + c0 + 16, vb);
          if (c0 + 32 < cend) tmem_ld<16>(tcol + c0 + 32, vc);
          const int4 e0 = *reinterpret_cast<const int4 *>(tup + (t0 >> 1)), e1 = *reinterpret_cast<const int4 *>(tup + (t0 >> 1) + 4);
          const int ent[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
          tmem_ld_wait();
          float sum[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            auto col = [&](int c) { return __int_as_float(c < 16 ? va[c & 15] : (c < 32 ? vb[c & 15] : vc[c & 15])); };
            // tap kx of pixel ox lands on column ox - 1 + kx: column ox collects tap 0 of its right neighbour, its own
            // tap 1 and tap 2 of its left neighbour, each with the power-of-two scale of the row it comes from
            const float fr = __shfl_down_sync(0xffffffffu, col(6 * u), 1), fl = __shfl_up_sync(0xffffffffu, col(6 * u + 2), 1);
            sum[u] = fmaf(fr, sc_r, fmaf(fl, sc_l, col(6 * u + 1) * scale));
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) red_add_pred(gxc + (ent[u] >> 2), sum[u], (vy >> (ent[u] & 3)) & 1u);
        }
        // ---- everything else, one unfolded element at a time: columns [3 * ntup, rows), split between the two warps
        const int c_begin = 3 * ntup;
        for (int c0 = c_begin + 16 * eh; c0 < dc.rows && !(P.dbg & 2); c0 += 32) {
          int v[16];
          tmem_ld<16>(tcol + c0, v);  // (fast fold: c0 is not a multiple of 16 and the read may run into the slack columns)
          tmem_ld_wait();
#pragma unroll
          for (int cc = 0; cc < 16; ++cc) {
            const int col = c0 + cc;
            if (col < dc.rows) {  // warp-uniform
              const float val = __int_as_float(v[cc]);
              const int f = dgrad_col_f(g, dc, col);
              if (P.fold) {
                const int ent = sm.ftab[f];
                red_add_pred(gxp + (ent >> 7), val * scale, (vm >> (ent & 31)) & 1u);
              } else if (live) {  // gxu[b][f][l]; later channel blocks add to what the first one wrote (same thread)
                float *dst = P.out + ((int64_t)eb * g.F + f) * g.L + el;
                const float t = val * scale;
                *dst = P.co0 > 0 ? *dst + t : t;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.tempty0 + 8 * buf);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kDgMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// fp16 weight digit tiles for the v2 dgrad: tile (channel block cb, i, k) = [Nf columns x Kc channels] K-major
// no-swizzle; column n holds unfold row dgrad_col_f(i, n); value 2^k * digit_k (the slice weight of lsq.py:306 with
// the activation-slice factor 2^(-abs*j) cancelled, see the header).
__global__ void weight_tiles_dgrad2_kernel(Geo g, int Nf, int Kc, const int8_t *__restrict__ wcodes,
                                           uint16_t *__restrict__ tiles) {
  const int64_t tile_elems = (int64_t)Nf * Kc;
  const int64_t tiles_per_block = (int64_t)g.NX * g.NSW;
  const int64_t n = (int64_t)(g.Cout / Kc) * tiles_per_block * tile_elems;
  const uint32_t sbo = (uint32_t)Kc * 16u;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tile = idx / tile_elems;
    const int within = (int)(idx % tile_elems);
    const int col = within / Kc, cl = within % Kc;
    const int cb = (int)(tile / tiles_per_block), ik = (int)(tile % tiles_per_block);
    const int i = ik / g.NSW, k = ik % g.NSW;
    const int co = cb * Kc + cl;
    const int f = dgrad2_tile_row(g, i, col);
    int digit = 0;
    if (f >= 0) {
      const int code = wcodes[(int64_t)co * g.F + f];
      const int mag = code < 0 ? -code : code;
      digit = (mag >> (g.wbs * k)) & g.wmask;
      if (code < 0) digit = -digit;
    }
    tiles[tile * tile_elems + tc_tile_offset16(col, cl, kTcLBO, sbo) / 2] =
        __half_as_ushort(__float2half_rn((float)(digit * (1 << k))));
  }
}

}  // namespace

}  // namespace v2

int64_t bwd_v2_scales_bytes(const Geo &g) { return (((int64_t)g.M * 4 + 255) & ~(int64_t)255) + (((int64_t)g.Cout * 4 + 255) & ~(int64_t)255); }

// rowscale = scales, chmax = scales + align256(M * 4)
int launch_go_scales(const Geo &g, const float *go, void *scales, cudaStream_t st) {
  float *rowscale = reinterpret_cast<float *>(scales);
  uint32_t *chmax = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(scales) + (((int64_t)g.M * 4 + 255) & ~(int64_t)255));
  CIMQ_CUDA_OK(cudaMemsetAsync(chmax, 0, (size_t)g.Cout * 4, st));
  CIMQ_REQUIRE(g.L % 4 == 0 && (reinterpret_cast<uintptr_t>(go) & 15u) == 0 && (reinterpret_cast<uintptr_t>(scales) & 15u) == 0,
               "go_scales: grad_out rows must be 16-byte aligned");
  int blocks = (g.M / 4 + v2::kScaleThreads - 1) / v2::kScaleThreads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 148 * 2 && g.Cout % 16 == 0 && g.Cout <= 512) {  // small layer: split the channels over the warps
    int sb = (g.M / 4 + 31) / 32;
    if (sb > 148 * 4) sb = 148 * 4;
    if (g.Cout % 32 == 0) v2::go_scales_split_kernel<8><<<sb, v2::kScaleThreads, 0, st>>>(g, go, rowscale, chmax);
    else v2::go_scales_split_kernel<4><<<sb, v2::kScaleThreads, 0, st>>>(g, go, rowscale, chmax);
    CIMQ_CUDA_OK(cudaGetLastError());
    return 0;
  }
  v2::go_scales_kernel<<<blocks, v2::kScaleThreads, (size_t)g.Cout * 4, st>>>(g, go, rowscale, chmax);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_weight_tiles_bwd2(const Geo &g, const int8_t *wcodes, void *tiles, cudaStream_t st) {
  const int Kc = g.Cout > 64 ? 64 : g.Cout;
  const int64_t n = (int64_t)g.NX * g.NSW * tc_nf(g) * g.Cout;
  v2::weight_tiles_dgrad2_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(g, tc_nf(g), Kc, wcodes,
                                                                        reinterpret_cast<uint16_t *>(tiles));
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_bwd_input_v2(const Geo &g, const float *go, const uint8_t *state, const void *wtb, const float *s,
                        const void *scales, float *out, int fold, cudaStream_t st) {
  using namespace v2;
  DgParams P;
  memset(&P, 0, sizeof(P));
  P.g = g;
  P.Kc = g.Cout > 64 ? 64 : g.Cout;
  P.Nf = tc_nf(g);
  P.mtiles = (g.M + kTcTileM - 1) / kTcTileM;
  P.a_bytes = (uint32_t)(kTcTileM * P.Kc * 2);
  P.b_bytes = (uint32_t)(P.Nf * P.Kc * 2);
  P.stage_bytes = kBwdPieces * P.a_bytes + P.b_bytes;
  P.fold = fold;
  P.fast = (fold && g.K == 3 && g.stride == 1 && g.pad == 1 && (g.OW == 8 || g.OW == 16 || g.OW == 32) &&
            (g.L % kTcTileM == 0 || 2 * g.L == kTcTileM)) ? 1 : 0;  // a tile is rows of one image, or two whole images
  const size_t ftab_bytes = ((size_t)g.F * 4 + 15) & ~(size_t)15;
  const size_t tup_bytes = (size_t)g.NX * (kTupStride * 4 + sizeof(v2::DgradCols));
  P.dbulk = (g.Cout == P.Kc && (reinterpret_cast<uintptr_t>(state) & 15u) == 0) ? 1 : 0;
  const size_t dbuf_bytes = P.dbulk ? 2 * (size_t)kTcTileM * P.Kc : 0;
  const size_t fixed = 256 + ftab_bytes + tup_bytes + 16 + dbuf_bytes;
  CIMQ_REQUIRE(fixed + P.stage_bytes <= kDgSmemBudget, "dgrad (v2): tile does not fit shared memory");
  int stages = (int)((kDgSmemBudget - fixed) / P.stage_bytes);
  if (stages > kDgMaxStages) stages = kDgMaxStages;
  P.stages = stages;
  P.aux_off = (uint32_t)((size_t)stages * P.stage_bytes);
  P.ftab_off = P.aux_off + 256;
  P.tup_off = P.ftab_off + (uint32_t)ftab_bytes;
  P.dbuf_off = P.tup_off + (uint32_t)((tup_bytes + 15) & ~(size_t)15);
  // as many accumulator buffers as fit (the epilogue's drain then hides behind the next chunks' MMAs); the fast fold
  // reads up to 18 columns past the last one
  P.nbuf = (512 - (P.fast ? 32 : 0)) / P.Nf;
  if (P.nbuf > 4) P.nbuf = 4;
  uint32_t cols = 32;
  while (cols < (uint32_t)(P.nbuf * P.Nf) + (P.fast ? 32u : 0u)) cols <<= 1;
  P.tmem_cols = cols;
  P.go = go; P.stateD = state; P.s = s; P.out = out;
  P.rowscale = reinterpret_cast<const float *>(scales);
  const size_t smem = (size_t)stages * P.stage_bytes + fixed + 1024;
  const int grid = P.mtiles < 148 ? P.mtiles : 148;
  if (const char *e = getenv("CIMQ_V2_DBG")) P.dbg = atoi(e);
#define CIMQ_LAUNCH_DG2(NS_, CPT_)                                                                                  \
  do {                                                                                                              \
    CIMQ_CUDA_OK(cudaFuncSetAttribute(bwd_input_v2_kernel<NS_, CPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                      (int)smem));                                                                  \
    bwd_input_v2_kernel<NS_, CPT_><<<grid, kDgThreads, smem, st>>>(P);                                              \
  } while (0)
  for (P.co0 = 0; P.co0 < g.Cout; P.co0 += P.Kc) {  // one launch per block of <= 64 output channels
    P.wtb = reinterpret_cast<const uint8_t *>(wtb) + (size_t)(P.co0 / P.Kc) * g.NX * g.NSW * P.b_bytes;
    if (g.NSW == 3) {
      if (P.Kc == 64) CIMQ_LAUNCH_DG2(3, 32);
      else if (P.Kc == 32) CIMQ_LAUNCH_DG2(3, 16);
      else CIMQ_LAUNCH_DG2(3, 8);
    } else {
      if (P.Kc == 64) CIMQ_LAUNCH_DG2(2, 32);
      else if (P.Kc == 32) CIMQ_LAUNCH_DG2(2, 16);
      else CIMQ_LAUNCH_DG2(2, 8);
    }
  }
#undef CIMQ_LAUNCH_DG2
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cimq
