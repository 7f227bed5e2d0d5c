// tcgen05 (5th-gen tensor core) forward kernel of the CiM convolution for sm_100a -- kernel template.
// Reference semantics: get_cim_output_signed.forward, lsq.py:92-237.
//
// Implicit GEMM, one persistent CTA per SM, warp-specialised (16 warps, registers rebalanced with
// setmaxnreg):
//
//   warps 0-3   producers   per (pixel tile, crossbar chunk): stage the input rows the tile needs in shared
//                           memory (uint8 codes, coalesced 4-byte loads, zero padding), then every thread
//                           assembles the im2col row of ITS output pixel -- whole input channels with
//                           compile-time byte positions, see cim_tc_layout.cuh for the K order -- splits it
//                           into NSA digit planes and stores them as the K-major A operands (16-byte
//                           stores, then fence.proxy.async).  One lane bulk-copies the pre-tiled weight digit
//                           planes of the chunk (cp.async.bulk -> mbarrier complete_tx).
//   warp  12    MMA issuer  (warps 13-15 idle) one thread: for every activation digit plane j one
//                           accumulation group D[128 x NSW*CT] (int32, TMEM) = A_j[128 x K] * B[K x NSW*CT]
//                           with tcgen05.mma.kind::i8 (K = 32 per instruction); TMEM is double buffered so
//                           the tensor core runs ahead of the epilogue by one digit plane.
//   warps 4-11  epilogue    tcgen05.ld the partial sums (thread = output pixel, registers = channels),
//                           quantise each to its ADC code with integer compares against per-(crossbar,
//                           slice pair, channel) thresholds held in shared memory, accumulate
//                           code*alpha*2^shift in fp32 registers across slice pairs and crossbar chunks,
//                           record code/clip bits for the backward.  The partial-sum tensor never exists in
//                           memory.
//
// Pipelines: shared-memory stages (full/empty mbarriers, producers <-> MMA) and two TMEM accumulator
// buffers (tmem_full/tmem_empty mbarriers, MMA <-> epilogue).
#pragma once

#include "cim_tc_layout.cuh"
#include "tc_ptx.cuh"

namespace cimq {
namespace tcfwd {

using namespace ptx;

constexpr int kProducerWarps = 4;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kEpilogueWarps = 8;
constexpr int kMmaWarp = kProducerWarps + kEpilogueWarps;
// 16 warps = 4 warpgroups: registers are allocated per warpgroup, so the fourth warpgroup (MMA issuer
// + three idle warps) costs nothing extra and setmaxnreg can move its registers to the epilogue.
constexpr int kThreads = 512;
constexpr int kRegsProducer = 104, kRegsMma = 40, kRegsEpilogue = 184;  // 4*32*(104+40+2*184) = 65536
constexpr int kMaxStages = 4;
constexpr int kNoRow = -2147483647 - 1;  // staged row outside the image / past the last pixel
constexpr int kMaxSlots = 16;  // staged input channels per chunk: <= 14 complete (K=3, 128 rows) + head + tail

struct TcParams {
  Geo g;
  int Kp, CT, nct, mtiles, stages;
  uint32_t a_bytes, b_bytes, stage_bytes, tmem_cols;
  // fast producer (input rows staged in shared memory)
  int fast;          // 1: staged producer, 0: generic per-element gather
  int owt;           // output pixels per staged row segment (= OW, divides 128)
  int rpt;           // output rows per tile (128 / OW)
  int rk;            // staged input rows per channel slot: (rpt-1)*stride+K when the rows of a tile are
                     // consecutive rows of one image (shared between output rows), else rpt*K
  int prow;          // staged row of output row `orow`, tap row 0: orow * prow
  int shared_rows;   // 1: a tile reads consecutive input rows of one image (rk, prow as above)
  uint32_t rk_magic; // ceil(2^32 / rk): q / rk = umulhi(q, rk_magic) for q < 2^16
  int tma_rows;      // 1: staged rows arrive by 16-byte cp.async copies; needs W % 16 == 0 and a 16-byte aligned image
  int prefetch;      // 1: the rows of the next stage are loaded into registers while this one is assembled
  int pitch_log2;    // staged row pitch in bytes (power of two)
  int col0;          // staged column of input column -pad (alignment shift)
  uint32_t raw_bytes;    // one staging buffer
  uint32_t ttab_bytes;   // one epilogue table slice per warpgroup
  const uint8_t *xcodes;
  const uint8_t *wtiles;   // [nct*NX] tiles of b_bytes
  const int2 *lut;         // per operand position: {element offset of the tap, tap index}
  const int4 *table;       // AoS table (generic paths)
  const uint32_t *ttab;    // tiled SoA table [ct][i][pair][tp|tg|amp][CT]
  const float *s;
  float *out;
  uint32_t *state;
  long long *debug;  // optional (tools/prof_fwd.py --timers): per-role cycle counters of CTA 0
};
// Per-role cycle counters (tools/prof_fwd.py --timers) exist only in builds made with `make TIMERS=1`;
// the shipped library pays no registers for them.
#ifndef CIMQ_TIMERS
#define CIMQ_TIMERS 0
#endif
constexpr bool kTimers = CIMQ_TIMERS != 0;
#define CIMQ_T0() (dbg ? clock64() : 0ll)

// shared-memory carve-up (all offsets from the dynamic smem base)
struct Smem {
  uint8_t *stage_base;   // stages * stage_bytes
  uint8_t *raw;          // 2 * raw_bytes
  uint32_t *ttab;        // 2 warpgroups * 2 buffers * ttab_bytes
  int *rowoff;           // 2 * 128 ints: global offset of staged row (orow, ky) or kNoRow
  uint32_t full0, empty0, tfull0, tempty0;
  uint32_t *tmem_slot;
  const ChunkLayout *cltab = nullptr;  // [NX] chunk layouts precomputed in shared memory (v2 kernel), or NULL
};
constexpr size_t kAuxBytes = 2048;  // barriers, tmem slot, rowoff table

__device__ __forceinline__ Smem carve(uint8_t *base, const TcParams &P) {
  Smem s;
  s.stage_base = base;
  uint8_t *p = base + (size_t)P.stages * P.stage_bytes;
  s.raw = p;
  p += 2 * (size_t)P.raw_bytes;
  s.ttab = reinterpret_cast<uint32_t *>(p);
  p += 4 * (size_t)P.ttab_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(p);
  s.full0 = smem_u32(bars);
  s.empty0 = s.full0 + 8 * kMaxStages;
  s.tfull0 = s.empty0 + 8 * kMaxStages;
  s.tempty0 = s.tfull0 + 16;
  s.tmem_slot = reinterpret_cast<uint32_t *>(p + 112);
  s.rowoff = reinterpret_cast<int *>(p + 128);
  return s;
}

// insert the low byte of `b` at byte position POS of w
template <int POS>
__device__ __forceinline__ uint32_t put_byte(uint32_t w, uint32_t b) {
  constexpr uint32_t sel = POS == 0 ? 0x3214u : POS == 1 ? 0x3240u : POS == 2 ? 0x3410u : 0x4210u;
  return __byte_perm(w, b, sel);
}

// ---------------------------------------------------------------------------------------------------
// fast producer: one stage = digit planes of 128 im2col rows for crossbar chunk i
// ---------------------------------------------------------------------------------------------------
constexpr int kPrefetchWords = 16;  // staged words a producer thread may hold in registers for the next stage

// which input rows / columns a producer thread stages (fixed for the kernel)
struct StageThread {
  int xw;        // word column inside a staged row
  int row0;      // first staged row (slot-major row list) of this thread
  int rstep;     // rows advanced per pass of all producer threads
  bool xok;      // the word lies inside the image horizontally
};
__device__ __forceinline__ StageThread stage_thread(const TcParams &P, int tid) {
  StageThread t;
  const int wpr_log2 = P.pitch_log2 - 2;  // words per staged row (<= 128)
  t.xw = tid & ((1 << wpr_log2) - 1);
  t.row0 = tid >> wpr_log2;
  t.rstep = kProducerThreads >> wpr_log2;
  const int ix = 4 * t.xw - P.col0 - P.g.pad;  // input column of this thread's word
  t.xok = ix >= 0 && ix < P.g.W;
  return t;
}

// staged rows of a tile: rowoff[staged row] = offset of staged column 0 of that input row in channel 0 of its
// image, or kNoRow (outside the image / past the last pixel)
__device__ __forceinline__ bool stage_set_rowoff(const TcParams &P, int mt, int *rowoff, int tid) {
  const Geo &g = P.g;
  const int r = tid;
  if (r < P.rk) {
    int off = kNoRow;
    if (P.shared_rows) {  // consecutive output rows of one image: input rows oy0*stride - pad + r
      const int m0 = mt * kTcTileM;
      if (m0 < g.M) {
        const int b = m0 / g.L, oy0 = (m0 % g.L) / g.OW;
        const int iy = oy0 * g.stride - g.pad + r;
        if (iy >= 0 && iy < g.H) off = (b * g.Cin * g.H + iy) * g.W - (P.tma_rows ? 0 : g.pad + P.col0);
      }
    } else {
      const int orow = r / g.K, ky = r % g.K;
      const int m_row = mt * kTcTileM + orow * P.owt;
      if (m_row < g.M) {
        const int b = m_row / g.L, oy = (m_row % g.L) / g.OW;
        const int iy = oy * g.stride - g.pad + ky;
        if (iy >= 0 && iy < g.H) off = (b * g.Cin * g.H + iy) * g.W - (P.tma_rows ? 0 : g.pad + P.col0);
      }
    }
    rowoff[r] = off;
    return off != kNoRow;
  }
  return false;
}

// Load up to NW words of chunk i's staged rows, starting at pass `pass0` of the thread's row walk.  Slot s < nfull
// holds channel cf0+s, then the head channel cf0-1, then the tail channel cf0+nfull; slots are contiguous
// (slot_bytes = rk * pitch), so word q of the slot-major row list lives at raw32[q].
template <int NW>
__device__ __forceinline__ void stage_load(const TcParams &P, const StageThread &t, const ChunkLayout &cl,
                                           const int *rowoff, int pass0, uint32_t (&v)[NW]) {
  const Geo &g = P.g;
  const int nslots = cl.nfull + (cl.nhead > 0 ? 1 : 0) + (cl.ntail > 0 ? 1 : 0);
  const int total_rows = nslots * P.rk, HW = g.H * g.W;
  int rl = t.row0 + pass0 * t.rstep;
  int sl = rl / P.rk, row = rl - sl * P.rk;
#pragma unroll
  for (int u = 0; u < NW; ++u) {
    v[u] = 0u;
    if (rl < total_rows) {
      int ch = cl.cf0 + sl;
      if (sl >= cl.nfull) ch = (sl == cl.nfull && cl.nhead > 0) ? cl.cf0 - 1 : cl.cf0 + cl.nfull;
      const int off = rowoff[row];
      if (t.xok && off != kNoRow)
        v[u] = __ldg(reinterpret_cast<const uint32_t *>(P.xcodes + (size_t)ch * HW + off + 4 * t.xw));
    }
    rl += t.rstep;
    row += t.rstep;
    while (row >= P.rk) { row -= P.rk; ++sl; }
  }
}
template <int NW>
__device__ __forceinline__ void stage_store(const TcParams &P, const StageThread &t, const ChunkLayout &cl,
                                            int pass0, const uint32_t (&v)[NW], uint8_t *raw, int tid) {
  const int nslots = cl.nfull + (cl.nhead > 0 ? 1 : 0) + (cl.ntail > 0 ? 1 : 0);
  const int total_words = (nslots * P.rk) << (P.pitch_log2 - 2);
  uint32_t *raw32 = reinterpret_cast<uint32_t *>(raw);
#pragma unroll
  for (int u = 0; u < NW; ++u) {
    const int q = tid + (pass0 + u) * kProducerThreads;
    if (q < total_words) raw32[q] = v[u];
  }
}

// Staged rows by asynchronous 16-byte copies (cp.async): row q of the slot-major list (slot = q / rk) is one input
// row of W bytes, copied to column pad+col0 (= 16, so both sides are 16-byte aligned) of its staged row; rows
// outside the image are zero-filled by the same instruction (source size 0).  The padding columns are zeroed once
// at kernel start and never written again.  No registers are held while the rows are in flight; the issuing
// thread waits for its own copies (cp.async.wait_group) before the barrier that publishes the buffer.
__device__ __forceinline__ void stage_issue_async(const TcParams &P, const ChunkLayout &cl, const int *rowoff,
                                                  uint8_t *raw, int tid) {
  const Geo &g = P.g;
  const int nslots = cl.nfull + (cl.nhead > 0 ? 1 : 0) + (cl.ntail > 0 ? 1 : 0);
  const int cpr = g.W >> 4;  // 16-byte pieces per row
  const int total = nslots * P.rk * cpr, HW = g.H * g.W;
  const int ch_head = cl.nhead > 0 ? cl.cf0 - 1 : cl.cf0 + cl.nfull, ch_tail = cl.cf0 + cl.nfull;
  const int cshift = (cpr & (cpr - 1)) == 0 ? 31 - __clz(cpr) : -1;  // (row widths of 16 * 2^n bytes: a shift)
  for (int q = tid; q < total; q += kProducerThreads) {
    const int rq = cshift >= 0 ? q >> cshift : q / cpr, c16 = q - rq * cpr;
    const int sl = (int)__umulhi((uint32_t)rq, P.rk_magic), row = rq - sl * P.rk;
    const int off = rowoff[row];
    const bool ok = off != kNoRow;
    const int ch = sl < cl.nfull ? cl.cf0 + sl : (sl == cl.nfull ? ch_head : ch_tail);
    const uint8_t *src = P.xcodes + (ok ? (size_t)ch * HW + off + 16 * c16 : (size_t)0);
    const uint32_t dst = smem_u32(raw + ((size_t)rq << P.pitch_log2) + g.pad + P.col0 + 16 * c16);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// fast producer, assembly: digit planes of this thread's im2col row of chunk `cl` from the staged rows
// ---------------------------------------------------------------------------------------------------
// ENC: how a digit becomes an operand byte.  0 = the integer digit (kind::i8); 1 = e4m3 2.0 (0x40) for a set bit of a
// 1-bit digit plane (kind::f8f6f4 of the v2 kernel; the weight tiles then hold +-0.5 so that products are +-1).
template <int ENC>
__device__ __forceinline__ uint32_t digit_plane_word(uint32_t w, int sh, uint32_t amask4) {
  if constexpr (ENC == 0) return (w >> sh) & amask4;
  else return ((w >> sh) & 0x01010101u) << 6;
}

template <int NSA, int KT, int ENC = 0>
__device__ __forceinline__ void produce_fast(const TcParams &P, const ChunkLayout &cl, uint8_t *st_ptr,
                                             const uint8_t *raw, int r, int pix_base) {
  constexpr int KK = KT * KT;
  constexpr int NFMAX = 126 / KK < kMaxSlots - 2 ? 126 / KK : kMaxSlots - 2;  // complete channels in 128 bytes
  const Geo &g = P.g;
  const int pitch = 1 << P.pitch_log2;
  const int slot_bytes = P.rk * pitch;
  // ---- 2. assemble this thread's im2col row: complete channels at compile-time byte positions
  uint32_t w[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) w[t] = 0u;
  const uint8_t *pb = raw + pix_base;
#pragma unroll
  for (int n = 0; n < NFMAX; ++n) {
    if (n < cl.nfull) {
      const uint8_t *cb = pb + (size_t)n * slot_bytes;
#pragma unroll
      for (int ky = 0; ky < KT; ++ky)
#pragma unroll
        for (int kx = 0; kx < KT; ++kx) {
          const int pos = n * KK + ky * KT + kx;  // compile time after unrolling
          const uint32_t b = cb[ky * pitch + kx];
          switch (pos & 3) {
            case 0: w[pos >> 2] = put_byte<0>(w[pos >> 2], b); break;
            case 1: w[pos >> 2] = put_byte<1>(w[pos >> 2], b); break;
            case 2: w[pos >> 2] = put_byte<2>(w[pos >> 2], b); break;
            default: w[pos >> 2] = put_byte<3>(w[pos >> 2], b); break;
          }
        }
    }
  }
  // ---- 2b. (3x3 kernels) taps of the two channels cut by the chunk edges: gather them into a 16-byte block
  //          (head taps, then tail taps) and OR it into w[] at byte nfull*9 -- a switch on the (uniform) number of
  //          complete channels makes the word positions compile-time
  bool partial_done = false;
  if constexpr (KT == 3) {
    unsigned long long ph = 0ull, pt = 0ull;
    const uint8_t *cbh = pb + (size_t)cl.nfull * slot_bytes;
    const uint8_t *cbt = pb + (size_t)(cl.nfull + (cl.nhead > 0 ? 1 : 0)) * slot_bytes;
#pragma unroll
    for (int e = 0; e < KK - 1; ++e) {
      if (e < cl.nhead) {
        const int tap = cl.head_tap0 + e;
        ph |= (unsigned long long)cbh[(tap / KT) * pitch + tap % KT] << (8 * e);
      }
      if (e < cl.ntail) pt |= (unsigned long long)cbt[(e / KT) * pitch + e % KT] << (8 * e);
    }
    const int hs = 8 * cl.nhead;  // 0..64
    const unsigned long long lo64 = ph | (hs < 64 ? pt << hs : 0ull);
    const unsigned long long hi64 = hs == 0 ? 0ull : (hs < 64 ? pt >> (64 - hs) : pt);
    const uint32_t p0 = (uint32_t)lo64, p1 = (uint32_t)(lo64 >> 32), p2 = (uint32_t)hi64, p3 = (uint32_t)(hi64 >> 32);
#define CIMQ_MERGE_CASE(N)                                                                    \
  case N: {                                                                                   \
    constexpr int START = N * 9, IDX = START >> 2, SH = (START & 3) * 8;                       \
    if constexpr (SH == 0) {                                                                  \
      w[IDX] |= p0;                                                                           \
      if constexpr (IDX + 1 < 32) w[IDX + 1] |= p1;                                           \
      if constexpr (IDX + 2 < 32) w[IDX + 2] |= p2;                                           \
      if constexpr (IDX + 3 < 32) w[IDX + 3] |= p3;                                           \
    } else {                                                                                  \
      w[IDX] |= p0 << SH;                                                                     \
      if constexpr (IDX + 1 < 32) w[IDX + 1] |= __funnelshift_l(p0, p1, SH);                  \
      if constexpr (IDX + 2 < 32) w[IDX + 2] |= __funnelshift_l(p1, p2, SH);                  \
      if constexpr (IDX + 3 < 32) w[IDX + 3] |= __funnelshift_l(p2, p3, SH);                  \
      if constexpr (IDX + 4 < 32) w[IDX + 4] |= p3 >> (32 - SH);                              \
    }                                                                                         \
  } break;
    switch (cl.nfull) {
      CIMQ_MERGE_CASE(0) CIMQ_MERGE_CASE(1) CIMQ_MERGE_CASE(2) CIMQ_MERGE_CASE(3) CIMQ_MERGE_CASE(4)
      CIMQ_MERGE_CASE(5) CIMQ_MERGE_CASE(6) CIMQ_MERGE_CASE(7) CIMQ_MERGE_CASE(8) CIMQ_MERGE_CASE(9)
      CIMQ_MERGE_CASE(10) CIMQ_MERGE_CASE(11) CIMQ_MERGE_CASE(12) CIMQ_MERGE_CASE(13) CIMQ_MERGE_CASE(14)
      default: break;
    }
#undef CIMQ_MERGE_CASE
    partial_done = true;
  }
  // ---- 3. digit planes (LSB first, slicing_act lsq.py:466-480) -> A operands, 16 bytes per store
  const uint32_t amask4 = (uint32_t)g.amask * 0x01010101u;
  const int ngroups = ((cl.rows + 31) & ~31) >> 4;
#pragma unroll
  for (int gi = 0; gi < 8; ++gi) {
    if (gi < ngroups) {
      const uint32_t off = tc_tile_offset(r, gi * 16, P.Kp);
#pragma unroll
      for (int j = 0; j < NSA; ++j) {
        const int sh = g.abs_ * j;
        *reinterpret_cast<uint4 *>(st_ptr + (size_t)j * P.a_bytes + off) =
            make_uint4(digit_plane_word<ENC>(w[4 * gi], sh, amask4), digit_plane_word<ENC>(w[4 * gi + 1], sh, amask4),
                       digit_plane_word<ENC>(w[4 * gi + 2], sh, amask4), digit_plane_word<ENC>(w[4 * gi + 3], sh, amask4));
      }
    }
  }
  // ---- 4. (other kernel sizes) taps of the channels cut by the chunk edges: byte stores after the vector stores
  if (!partial_done) {
    int pos = cl.nfull * KK;
    for (int part = 0; part < 2; ++part) {
      const int cnt = part == 0 ? cl.nhead : cl.ntail;
      if (cnt == 0) continue;
      const int sl = cl.nfull + ((part == 1 && cl.nhead > 0) ? 1 : 0);
      const int tap0 = part == 0 ? cl.head_tap0 : 0;
      const uint8_t *cb = pb + (size_t)sl * slot_bytes;
      for (int e = 0; e < cnt; ++e, ++pos) {
        const int tap = tap0 + e, ky = tap / KT, kx = tap % KT;
        const uint32_t b = cb[ky * pitch + kx];
        const uint32_t off = tc_tile_offset(r, pos, P.Kp);
#pragma unroll
        for (int j = 0; j < NSA; ++j)
          st_ptr[(size_t)j * P.a_bytes + off] = (uint8_t)digit_plane_word<ENC>(b, g.abs_ * j, (uint32_t)g.amask);
      }
    }
  }
}

// generic producer: per-element gather through the position LUT (any geometry)
template <int NSA, int ENC = 0>
__device__ __forceinline__ void produce_generic(const TcParams &P, int i, uint8_t *st_ptr, int r, int base,
                                                uint32_t vm) {
  const Geo &g = P.g;
  const uint32_t amask4 = (uint32_t)g.amask * 0x01010101u;
  const int lo = i * g.xbar;
  const int rows = min(g.xbar, g.F - lo);
  const int ngroups = ((rows + 31) & ~31) >> 4;
  for (int gi = 0; gi < ngroups; ++gi) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int kk = gi * 16 + e;
      uint32_t code = 0;
      if (kk < rows) {
        const int2 lt = __ldg(&P.lut[lo + kk]);
        if ((vm >> lt.y) & 1u) code = P.xcodes[base + lt.x];
      }
      w[e >> 2] |= code << (8 * (e & 3));
    }
    const uint32_t off = tc_tile_offset(r, gi * 16, P.Kp);
#pragma unroll
    for (int j = 0; j < NSA; ++j) {
      const int sh = g.abs_ * j;
      *reinterpret_cast<uint4 *>(st_ptr + (size_t)j * P.a_bytes + off) =
          make_uint4(digit_plane_word<ENC>(w[0], sh, amask4), digit_plane_word<ENC>(w[1], sh, amask4),
                     digit_plane_word<ENC>(w[2], sh, amask4), digit_plane_word<ENC>(w[3], sh, amask4));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// the producer role (4 warps): per (pixel tile, crossbar chunk) one pipeline stage = NSA digit planes of 128 im2col
// rows + the chunk's weight tile (bulk copy).  Shared by the v1 (kind::i8) and v2 (kind::f8f6f4) kernels.
// ---------------------------------------------------------------------------------------------------
// K5: also compile the 5x5 instance of the staged producer (the v2 kernel leaves 5x5 layers to the generic gather:
// every instance is ~45 KB of code, and the producers' working set should stay inside the instruction cache)
template <int NSA, int ENC, bool K5 = true>
__device__ __forceinline__ void producer_loop(const TcParams &P, const Smem &sm, int ntiles, int tid = threadIdx.x,
                                              int gidx = 0, int ngroups = 1) {
  // A producer GROUP is 128 threads (tid = index inside the group) that build whole pipeline stages; with
  // ngroups > 1 (v2 kernel) group gidx builds stages gidx, gidx + ngroups, ... of the CTA's (tile, chunk) sequence,
  // with its own staging buffers (sm.raw, sm.rowoff point at the group's) and named barrier (1 + gidx).
  const Geo &g = P.g;
  const int r = tid;  // tile row = output pixel
  const uint32_t bar = 1u + (uint32_t)gidx;
  const bool dbg = kTimers && P.debug != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long d_wait = 0, d_prod = 0, d_tile = 0;
  uint32_t it = (uint32_t)gidx, lit = 0;  // global / group-local stage counters
  // stage slot / use count of stage `it` as running counters, chunk layouts from a table where the kernel provides one:
  // divisions by run-time values (it % stages, the four of chunk_layout) were a sizeable part of a producer stage
  int p_sidx = gidx % P.stages;
  uint32_t p_use = (uint32_t)(gidx / P.stages);
  auto next_slot = [&]() {
    p_sidx += ngroups;
    while (p_sidx >= P.stages) { p_sidx -= P.stages; ++p_use; }
  };
  auto layout_of = [&](int i_) { return sm.cltab != nullptr ? sm.cltab[i_] : chunk_layout(g, i_); };
  // (tile, i) of stage `it`: advance by n stages
  auto advance = [&](int &tile_, int &i_, int n) {
    i_ += n;
    while (i_ >= g.NX) { i_ -= g.NX; tile_ += gridDim.x; }
  };
  if (P.fast) {
    // ---- staged producer.  Stage `it` is (tile, chunk i); its input rows sit in raw[lit & 1].
    const StageThread stt = stage_thread(P, tid);
    const int pix_base = (((r / P.owt) * P.prow) << P.pitch_log2) + (r % P.owt) * g.stride + P.col0;
    int tile = blockIdx.x, i = 0, tpar = 0;
    advance(tile, i, gidx);
    uint32_t v[kPrefetchWords];
    if (tile < ntiles) {
      stage_set_rowoff(P, tile / P.nct, sm.rowoff, tid);
      named_barrier_sync(bar, kProducerThreads);
      if (P.tma_rows) {
        // padding columns are zero for the whole kernel: clear both buffers once, then only rows are written
        for (uint32_t q = tid * 16u; q < 2u * P.raw_bytes; q += kProducerThreads * 16u)
          *reinterpret_cast<uint4 *>(sm.raw + q) = make_uint4(0u, 0u, 0u, 0u);
        named_barrier_sync(bar, kProducerThreads);
        stage_issue_async(P, layout_of(i), sm.rowoff, sm.raw, tid);
        stage_async_wait();
        named_barrier_sync(bar, kProducerThreads);
      }
      if (P.prefetch) {
        const ChunkLayout cl0 = layout_of(i);
        stage_load<kPrefetchWords>(P, stt, cl0, sm.rowoff, 0, v);
        stage_store<kPrefetchWords>(P, stt, cl0, 0, v, sm.raw, tid);
        named_barrier_sync(bar, kProducerThreads);
      }
    }
    while (tile < ntiles) {
      const int ct = P.nct == 1 ? 0 : tile % P.nct;
      const ChunkLayout cl = layout_of(i);
      uint8_t *raw = sm.raw + (size_t)(lit & 1) * P.raw_bytes;
      // the group's next stage
      int ni = i, ntile = tile;
      advance(ntile, ni, ngroups);
      const int ntpar = ntile != tile ? tpar ^ 1 : tpar;
      const bool more = ntile < ntiles;
      ChunkLayout ncl = cl;
      if (more) {
        ncl = layout_of(ni);
        if (ntile != tile) {  // new tile: publish its row table first (its buffer was last read a whole tile ago)
          stage_set_rowoff(P, ntile / P.nct, sm.rowoff + ntpar * 128, tid);
          named_barrier_sync(bar, kProducerThreads);
        }
        // its input rows start their trip from L2 / HBM now: by bulk copy straight into the other staging buffer,
        // or into registers, while this stage is built
        if (P.tma_rows)
          stage_issue_async(P, ncl, sm.rowoff + ntpar * 128, sm.raw + (size_t)((lit + 1) & 1) * P.raw_bytes, tid);
        else if (P.prefetch) stage_load<kPrefetchWords>(P, stt, ncl, sm.rowoff + ntpar * 128, 0, v);
      }
      if (!P.prefetch && !P.tma_rows) {  // too many rows for the registers: stage this chunk now, eight loads in flight
        const int nslots = cl.nfull + (cl.nhead > 0 ? 1 : 0) + (cl.ntail > 0 ? 1 : 0);
        const int passes = ((nslots * P.rk) + stt.rstep - 1) / stt.rstep;
        for (int p0 = 0; p0 < passes; p0 += 8) {
          uint32_t v8[8];
          stage_load<8>(P, stt, cl, sm.rowoff + tpar * 128, p0, v8);
          stage_store<8>(P, stt, cl, p0, v8, raw, tid);
        }
        named_barrier_sync(bar, kProducerThreads);
      }
      const int sidx = p_sidx;
      const uint32_t use = p_use;
      next_slot();
      long long t0 = CIMQ_T0();
      mbar_wait(sm.empty0 + 8 * sidx, (use & 1) ^ 1);
      long long t1 = CIMQ_T0();
      d_wait += t1 - t0;
      uint8_t *st_ptr = sm.stage_base + (size_t)sidx * P.stage_bytes;
      if (tid == 0) {
        mbar_arrive_expect_tx(sm.full0 + 8 * sidx, P.b_bytes);
        bulk_copy_g2s(smem_u32(st_ptr + (size_t)NSA * P.a_bytes), P.wtiles + (size_t)(ct * g.NX + i) * P.b_bytes,
                      P.b_bytes, sm.full0 + 8 * sidx);
      }
      if (ENC == 1 && P.debug == reinterpret_cast<long long *>(1)) { /* development: skip the row assembly */ }
      else if (!K5 || g.K == 3) produce_fast<NSA, 3, ENC>(P, cl, st_ptr, raw, r, pix_base);
      else if constexpr (K5) produce_fast<NSA, 5, ENC>(P, cl, st_ptr, raw, r, pix_base);
      fence_proxy_async();  // every thread publishes its own stores to the async proxy; one arrival per warp (every
      __syncwarp();         // arrival wakes the warps sleeping on ANY mbarrier of the CTA: 128 per stage made them spin)
      if ((tid & 31) == 0) mbar_arrive(sm.full0 + 8 * sidx);
      if (P.prefetch || P.tma_rows) {
        if (P.prefetch && more)
          stage_store<kPrefetchWords>(P, stt, ncl, 0, v, sm.raw + (size_t)((lit + 1) & 1) * P.raw_bytes, tid);
        if (P.tma_rows) stage_async_wait();  // my copies of the next stage's rows have landed
        // next stage's rows visible to all producers; everyone is done reading this stage's rows
        named_barrier_sync(bar, kProducerThreads);
      }
      d_prod += CIMQ_T0() - t1;
      i = ni; tile = ntile; tpar = ntpar; it += ngroups; ++lit;
    }
  } else {
    int tile = blockIdx.x, i = 0;
    advance(tile, i, gidx);
    int cur_tile = -1, base = 0;
    uint32_t vm = 0;  // bit t set = tap t of this pixel is inside the image
    while (tile < ntiles) {
      const int mt = tile / P.nct, ct = tile % P.nct;
      if (tile != cur_tile) {
        cur_tile = tile;
        const int m = mt * kTcTileM + r;
        base = 0;
        vm = 0;
        if (m < g.M) {
          const int b = m / g.L, l = m % g.L, oy = l / g.OW, ox = l % g.OW;
          const int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
          base = (b * g.Cin * g.H + iy0) * g.W + ix0;
          for (int ky = 0; ky < g.K; ++ky)
            for (int kx = 0; kx < g.K; ++kx)
              if (iy0 + ky >= 0 && iy0 + ky < g.H && ix0 + kx >= 0 && ix0 + kx < g.W) vm |= 1u << (ky * g.K + kx);
        }
      }
      const int sidx = p_sidx;
      const uint32_t use = p_use;
      next_slot();
      mbar_wait(sm.empty0 + 8 * sidx, (use & 1) ^ 1);
      uint8_t *st_ptr = sm.stage_base + (size_t)sidx * P.stage_bytes;
      if (tid == 0) {
        mbar_arrive_expect_tx(sm.full0 + 8 * sidx, P.b_bytes);
        bulk_copy_g2s(smem_u32(st_ptr + (size_t)NSA * P.a_bytes), P.wtiles + (size_t)(ct * g.NX + i) * P.b_bytes,
                      P.b_bytes, sm.full0 + 8 * sidx);
      }
      produce_generic<NSA, ENC>(P, i, st_ptr, r, base, vm);
      fence_proxy_async();
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(sm.full0 + 8 * sidx);
      advance(tile, i, ngroups);
      it += ngroups;
    }
  }
  if (dbg) { P.debug[0] = d_wait; P.debug[1] = d_prod; P.debug[2] = d_tile; }
}

// ---------------------------------------------------------------------------------------------------
// the kernel.  NSW/NSA: weight / activation digit planes; CH: output channels per epilogue thread (CT/2);
// MB: multi-bit ADC (clamp) instead of the binary / ternary threshold ADC; WS: also write the ADC state the
// backward pass reads (training) -- a compile-time switch, so the inference kernel carries none of it
// ---------------------------------------------------------------------------------------------------
template <int NSW, int NSA, int CH, bool MB, bool WS>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const TcParams P) {
  constexpr int CT = 2 * CH;
  constexpr int NROWS = NSW * CT;  // UMMA N
  constexpr int PAIRS = NSW * NSA;
  constexpr int SWORDS_MAX = (3 * PAIRS + 31) / 32;
  const Geo &g = P.g;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const Smem sm = carve(smem_raw, P);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // (provably warp-uniform)

  if (threadIdx.x == 0) {
    for (int sidx = 0; sidx < P.stages; ++sidx) {
      mbar_init(sm.full0 + 8 * sidx, kProducerWarps + 1);
      mbar_init(sm.empty0 + 8 * sidx, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sm.tfull0 + 8 * b, 1);
      mbar_init(sm.tempty0 + 8 * b, kEpilogueWarps);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(sm.tmem_slot), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_slot;

  const int ntiles = P.mtiles * P.nct;
  const int rows_full = g.xbar < g.F ? g.xbar : g.F;

  if (warp < kProducerWarps) {
    // =========================== producers ===========================
    reg_dealloc<kRegsProducer>();
    producer_loop<NSA, 0>(P, sm, ntiles);
  } else if (warp >= kMmaWarp) {
    // =========================== MMA issuer ===========================
    reg_dealloc<kRegsMma>();
    if (warp == kMmaWarp) {
      // (the whole warp walks the loops: uniform control flow keeps descriptors and counters in uniform registers -- a
      // single-lane branch made the compiler wrap every MMA in an elect / broadcast loop; lane 0 issues)
      const bool dbg = kTimers && P.debug != nullptr && blockIdx.x == 0 && lane == 0;
      long long d_full = 0, d_tempty = 0, t_begin = CIMQ_T0();
      const uint32_t idesc = idesc_i8_u8s8(kTcTileM, NROWS);
      const uint32_t sbo = 8u * (uint32_t)P.Kp;
      uint32_t acc_it = 0;
      int m_sidx = 0;  // pipeline stage and its use count as running counters (no division by a run-time value)
      uint32_t m_use = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int i = 0; i < g.NX; ++i) {
          const int sidx = m_sidx;
          const uint32_t use = m_use;
          if (++m_sidx == P.stages) { m_sidx = 0; ++m_use; }
          const int rows = min(rows_full, g.F - i * g.xbar);
          const int ksteps = (rows + 31) >> 5;
          long long t0 = CIMQ_T0();
          mbar_wait<400>(sm.full0 + 8 * sidx, use & 1);
          d_full += CIMQ_T0() - t0;
          tc_fence_after();
          const uint32_t a0 = smem_u32(sm.stage_base + (size_t)sidx * P.stage_bytes);
          const uint32_t b0 = a0 + NSA * P.a_bytes;
          for (int j = 0; j < NSA; ++j, ++acc_it) {
            const uint32_t buf = acc_it & 1, buse = acc_it >> 1;
            long long t2 = CIMQ_T0();
            mbar_wait<400>(sm.tempty0 + 8 * buf, (buse & 1) ^ 1);
            d_tempty += CIMQ_T0() - t2;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * NROWS;
            if (lane == 0) {
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t adesc = make_smem_desc(a0 + j * P.a_bytes + ks * 2 * kTcLBO, kTcLBO, sbo);
                const uint64_t bdesc = make_smem_desc(b0 + ks * 2 * kTcLBO, kTcLBO, sbo);
                umma_i8(d_tmem, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
              }
              umma_commit(sm.tfull0 + 8 * buf);  // accumulator of digit plane j complete -> epilogue
              if (j == NSA - 1) umma_commit(sm.empty0 + 8 * sidx);  // all MMAs reading this stage complete -> producers
            }
            __syncwarp();
          }
        }
      }
      if (dbg) { P.debug[4] = d_full; P.debug[5] = d_tempty; P.debug[6] = clock64() - t_begin; }
    }
  } else {
    // =========================== epilogue ===========================
    reg_alloc<kRegsEpilogue>();
    const int e = warp - kProducerWarps;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = e >> 2;       // which half of the channel tile (= epilogue warpgroup)
    const int wgt = threadIdx.x - (kProducerWarps + 4 * half) * 32;  // thread index inside the warpgroup
    const int r = quarter * 32 + lane;
    const float sa = P.s[0], sw = P.s[1];
    const int swords = g.state_words;
    constexpr bool want_state = WS;
    const bool dbg = kTimers && P.debug != nullptr && blockIdx.x == 0 && warp == kProducerWarps && lane == 0;
    long long d_tfull = 0, d_comp = 0, d_tab = 0, d_st = 0;
    // per-warpgroup table slice [pair][tp|tg|amp][CH], double buffered by chunk parity
    constexpr int SLICE_WORDS = PAIRS * 3 * CH;
    uint32_t *tbuf = sm.ttab + (size_t)half * 2 * (P.ttab_bytes / 4);
    uint32_t acc_it = 0, chunk_it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int mt = tile / P.nct, ct = tile % P.nct;
      const int m = mt * kTcTileM + r;
      const int c_first = ct * CT + half * CH;
      float acc[CH];
#pragma unroll
      for (int cc = 0; cc < CH; ++cc) acc[cc] = 0.0f;
      for (int i = 0; i < g.NX; ++i, ++chunk_it) {
        // stage this chunk's thresholds / amplitudes for our channels (the previous use of this buffer was
        // two chunks ago; the barrier below orders it)
        uint32_t *tb = tbuf + (size_t)(chunk_it & 1) * (P.ttab_bytes / 4);
        long long ta = CIMQ_T0();
        {
          const uint32_t *src = P.ttab + ((size_t)(ct * g.NX + i) * PAIRS * 3) * CT + half * CH;
          for (int idx = wgt; idx < SLICE_WORDS / 4; idx += 128) {
            const int row = idx / (CH / 4), c4 = idx % (CH / 4);  // row = pair*3 + array
            *reinterpret_cast<uint4 *>(tb + row * CH + 4 * c4) =
                __ldg(reinterpret_cast<const uint4 *>(src + (size_t)row * CT + 4 * c4));
          }
        }
        named_barrier_sync(2 + half, 128);
        d_tab += CIMQ_T0() - ta;
        // ADC state, accumulated on the FMA pipe: every 32-bit state word is two fp32 accumulators (bits
        // [0,HB) and [HB,32)) that start at 2^23, so their low mantissa bits are an exact integer, and
        // receive decision * 2^bit.  State bit of (type t, act slice j, weight slice k) = t*PAIRS + j*NSW + k:
        // the k loop below is a real loop (small code: the instruction cache matters here), so a group of NSW
        // consecutive bits must stay inside one accumulator -- HB is a multiple of NSW.
        constexpr int HB = (16 / NSW) * NSW;
        float stf[CH][2 * SWORDS_MAX];
#pragma unroll
        for (int cc = 0; cc < CH; ++cc)
#pragma unroll
          for (int w = 0; w < 2 * SWORDS_MAX; ++w) stf[cc][w] = 8388608.0f;
#pragma unroll
        for (int j = 0; j < NSA; ++j, ++acc_it) {
          const uint32_t buf = acc_it & 1, buse = acc_it >> 1;
          long long tw = CIMQ_T0();
          mbar_wait(sm.tfull0 + 8 * buf, buse & 1);
          long long tc0 = CIMQ_T0();
          d_tfull += tc0 - tw;
          tc_fence_after();
          // first bit of this activation slice in the +1 / -1 / clip groups, its accumulator and weight there
          constexpr int kTypes = MB ? 1 : 3;
          int aidx[3];
          float abase[3];
#pragma unroll
          for (int t = 0; t < kTypes; ++t) {
            const int b0 = t * PAIRS + j * NSW;
            const bool hi = (b0 & 31) >= HB;
            aidx[t] = 2 * (b0 >> 5) + (hi ? 1 : 0);
            abase[t] = (float)(1u << ((b0 & 31) - (hi ? HB : 0)));
          }
          float wk = 1.0f;  // 2^k
          const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * NROWS + half * CH;
          // The accumulator is read VW columns at a time; as soon as a group is converted to fp32 its registers
          // are free and the next group (same weight slice, or the next one) is fetched while this one is quantised.
          constexpr int VW = CH < 16 ? CH : 16, NG = CH / VW;
          int v[VW];
          tmem_ld<VW>(tcol, v);
#pragma unroll 1
          for (int k = 0; k < NSW; ++k, wk *= 2.0f) {
            const uint32_t *trow = tb + (k * NSA + j) * 3 * CH;
            const float kPos = abase[0] * wk, kNeg = abase[MB ? 0 : 1] * wk, kClp = abase[MB ? 0 : 2] * wk;
#pragma unroll
            for (int gq = 0; gq < NG; ++gq) {
              tmem_ld_wait();
              // int32 partial sums -> fp32 without the conversion unit (exact for |p| < 2^22)
              float pfv[VW];
#pragma unroll
              for (int cc = 0; cc < VW; ++cc) {
                if constexpr (MB) pfv[cc] = psum_as_stored(v[cc]);
                else pfv[cc] = __int_as_float(v[cc] + 0x4B400000) - 12582912.0f;
              }
              if (gq + 1 < NG) tmem_ld<VW>(tcol + k * CT + (gq + 1) * VW, v);
              else if (k + 1 < NSW) tmem_ld<VW>(tcol + (k + 1) * CT, v);
#pragma unroll
              for (int c4 = 0; c4 < VW / 4; ++c4) {
                const int cb = gq * VW + 4 * c4;
                const float4 amp4 = *reinterpret_cast<const float4 *>(trow + 2 * CH + cb);
                const float ampv[4] = {amp4.x, amp4.y, amp4.z, amp4.w};
                if constexpr (MB) {
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int cc = cb + u;
                    const float ph = pfv[4 * c4 + u];
                    const float cf = fminf(fmaxf(ph, (float)g.qn), (float)g.qp);
                    acc[cc] += __fmul_rn(__fmul_rn(cf, sw), sa) * ampv[u];
                    if (ph >= (float)g.clip_hi || ph <= (float)g.clip_lo) stf[cc][aidx[0]] += kPos;  // lsq.py:310-311
                  }
                } else {
                  const float4 tp4 = *reinterpret_cast<const float4 *>(trow + cb);
                  const float tpv[4] = {tp4.x, tp4.y, tp4.z, tp4.w};
                  float tgv[4] = {0.f, 0.f, 0.f, 0.f};
                  if (want_state) {
                    const float4 tg4 = *reinterpret_cast<const float4 *>(trow + CH + cb);
                    tgv[0] = tg4.x; tgv[1] = tg4.y; tgv[2] = tg4.z; tgv[3] = tg4.w;
                  }
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const int cc = cb + u;
                    const float pf = pfv[4 * c4 + u];
                    const float tpos = __saturatef(pf - tpv[u]);    // 1 if p >= tp
                    const float tneg = __saturatef(-pf - tpv[u]);   // 1 if p <= -tp
                    acc[cc] = fmaf(tpos - tneg, ampv[u], acc[cc]);  // ternary / sign code times alpha*2^shift
                    if (want_state) {
                      const float tclp = __saturatef(fabsf(pf) - tgv[u]);  // 1 if |p| >= tg (STE clip, lsq.py:310)
                      stf[cc][aidx[0]] = fmaf(tpos, kPos, stf[cc][aidx[0]]);
                      stf[cc][aidx[1]] = fmaf(tneg, kNeg, stf[cc][aidx[1]]);
                      stf[cc][aidx[2]] = fmaf(tclp, kClp, stf[cc][aidx[2]]);
                    }
                  }
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sm.tempty0 + 8 * buf);  // this warp has drained the accumulator
          d_comp += CIMQ_T0() - tc0;
        }
        long long ts = CIMQ_T0();
        uint32_t stw[CH][SWORDS_MAX];
#pragma unroll
        for (int cc = 0; cc < CH; ++cc)
#pragma unroll
          for (int w = 0; w < SWORDS_MAX; ++w)
            stw[cc][w] = (__float_as_uint(stf[cc][2 * w]) & 0x7fffffu) |
                         ((__float_as_uint(stf[cc][2 * w + 1]) & 0x7fffffu) << HB);
        if (want_state && m < g.M) {
          uint32_t *sp = P.state + ((size_t)i * g.Cout + c_first) * swords * g.M + m;  // [i][c][word][m]
#pragma unroll
          for (int cc = 0; cc < CH; ++cc)
#pragma unroll
            for (int w = 0; w < SWORDS_MAX; ++w)
              if (w < swords) {
                *sp = stw[cc][w];
                sp += g.M;
              }
        }
        d_st += CIMQ_T0() - ts;
      }
      if (m < g.M) {
        const int b = m / g.L, l = m % g.L;
#pragma unroll
        for (int cc = 0; cc < CH; ++cc) P.out[((size_t)b * g.Cout + c_first + cc) * g.L + l] = acc[cc];
      }
    }
    if (dbg) { P.debug[8] = d_tfull; P.debug[9] = d_comp; P.debug[10] = d_tab; P.debug[11] = d_st; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

template <int NSW, int NSA, int CH>
int launch_instance(const TcParams &P, size_t smem, int grid, cudaStream_t st) {
  const bool mb = P.g.adc_mode == CIMQ_ADC_MULTIBIT, ws = P.state != nullptr;
#define CIMQ_TC_LAUNCH(MB_, WS_)                                                                      \
  do {                                                                                                \
    CIMQ_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<NSW, NSA, CH, MB_, WS_>,                         \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    conv_tc_kernel<NSW, NSA, CH, MB_, WS_><<<grid, kThreads, smem, st>>>(P);                          \
  } while (0)
  if (mb && ws) CIMQ_TC_LAUNCH(true, true);
  else if (mb) CIMQ_TC_LAUNCH(true, false);
  else if (ws) CIMQ_TC_LAUNCH(false, true);
  else CIMQ_TC_LAUNCH(false, false);
#undef CIMQ_TC_LAUNCH
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

void plan_producer(const Geo &g, TcParams &P);  // cim_conv_tc.cu

// one translation unit per slice count (parallel compilation)
int launch_ns2(const TcParams &P, size_t smem, int grid, int ch, cudaStream_t st);
int launch_ns3(const TcParams &P, size_t smem, int grid, int ch, cudaStream_t st);
int launch_ns4(const TcParams &P, size_t smem, int grid, int ch, cudaStream_t st);
int launch_ns8(const TcParams &P, size_t smem, int grid, int ch, cudaStream_t st);

}  // namespace tcfwd
}  // namespace cimq
