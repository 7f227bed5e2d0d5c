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
