// Shared-memory operand layout of the tcgen05 forward kernel, shared between the weight
// preparation kernel (which writes weight digit tiles to global memory already in this order, so
// one cp.async.bulk drops them into shared memory) and the conv kernel.
//
// Both MMA operands are K-major with NO swizzle ("interleaved" canonical layout): the tile is a
// grid of 8-row x 16-byte core matrices, each 128 contiguous bytes.
//   byte offset of (row r, k-byte kk) = (r >> 3) * SBO + (kk >> 4) * LBO + (r & 7) * 16 + (kk & 15)
//   LBO (core matrices adjacent in K)   = 128
//   SBO (core matrices adjacent in M/N) = 8 * Kp        (Kp = padded K bytes of the tile)
// so a tile of R rows is exactly R*Kp bytes for any Kp that is a multiple of 32 (one
// tcgen05.mma.kind::i8 consumes K = 32 bytes = two core matrices along K).
#pragma once

#include "cimq_common.cuh"

namespace cimq {

constexpr int kTcTileM = 128;        // output pixels per CTA tile (UMMA M)
constexpr int kTcLBO = 128;          // bytes between K-adjacent core matrices
constexpr int kTcMaxKp = 256;        // largest padded crossbar depth handled by the tc kernel

__host__ __device__ inline int tc_kp(const Geo &g) {
  int rows = g.xbar < g.F ? g.xbar : g.F;
  return (rows + 31) & ~31;
}
__host__ __device__ inline uint32_t tc_tile_offset(int r, int kk, int Kp) {
  return (uint32_t)((r >> 3) * (8 * Kp) + (kk >> 4) * kTcLBO + (r & 7) * 16 + (kk & 15));
}


// ---- K order inside a crossbar chunk ------------------------------------------------------------------
// A chunk is the SET of unfold rows f in [lo, hi) (the partition is the reference's, lsq.py:172-185); inside
// a chunk the order is free (integer sum), so the tensor-core operands use: first every input channel that
// lies completely inside the chunk (K*K taps each, channel-major), then the taps of the channel cut by
// the chunk's lower edge ("head"), then those cut by its upper edge ("tail").  The producer can then
// assemble whole channels with compile-time byte positions.
struct ChunkLayout {
  int lo, rows;
  int cf0;        // first complete channel
  int nfull;      // number of complete channels
  int nhead;      // rows of channel cf0-1 (taps head_tap0 .. KK-1)
  int head_tap0;
  int ntail;      // rows of channel cf0+nfull (taps 0 .. ntail-1)
};
__host__ __device__ inline ChunkLayout chunk_layout(const Geo &g, int i) {
  ChunkLayout c;
  c.lo = i * g.xbar;
  const int hi = (c.lo + g.xbar < g.F) ? c.lo + g.xbar : g.F;
  c.rows = hi - c.lo;
  c.cf0 = (c.lo + g.KK - 1) / g.KK;
  const int cf1 = hi / g.KK;
  if (cf1 >= c.cf0) {
    c.nfull = cf1 - c.cf0;
    c.nhead = c.cf0 * g.KK - c.lo;
    c.ntail = hi - cf1 * g.KK;
  } else {  // the chunk lies strictly inside one channel
    c.nfull = 0;
    c.nhead = c.rows;
    c.ntail = 0;
  }
  c.head_tap0 = c.lo - (c.cf0 - 1) * g.KK;
  return c;
}
// unfold row f stored at operand position `pos` (0 <= pos < rows) of chunk i
__host__ __device__ inline int chunk_row_at(const Geo &g, const ChunkLayout &c, int pos) {
  if (pos < c.nfull * g.KK) return (c.cf0 + pos / g.KK) * g.KK + pos % g.KK;
  pos -= c.nfull * g.KK;
  if (pos < c.nhead) return c.lo + pos;
  return (c.cf0 + c.nfull) * g.KK + (pos - c.nhead);
}

// ---- backward (bf16) tiles --------------------------------------------------------------------------
// 16-bit operands use the same K-major no-swizzle scheme with 8-element (16-byte) core-matrix rows:
//   byte offset of (row r, k-element kk) = (r >> 3) * SBO + (kk >> 3) * LBO + (r & 7) * 16 + (kk & 7) * 2
__host__ __device__ inline uint32_t tc_tile_offset16(int r, int kk, int lbo, int sbo) {
  return (uint32_t)((r >> 3) * sbo + (kk >> 3) * lbo + (r & 7) * 16 + (kk & 7) * 2);
}
// dgrad GEMM N dimension: crossbar rows padded to a multiple of 16
__host__ __device__ inline int tc_nf(const Geo &g) {
  int rows = g.xbar < g.F ? g.xbar : g.F;
  return (rows + 15) & ~15;
}

}  // namespace cimq
