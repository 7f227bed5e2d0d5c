// "v2" data formats shared by the second-generation kernels (forward cim_conv_v2.cu, backward cim_bwd_tc.cu /
// cim_bwd.cu, preparation cim_prep.cu).
//
// Why a second format.  The first-generation forward quantised every partial sum with ~16 fp32 / int32
// instructions and stored 3 bits per partial sum in 32-bit words that every backward kernel re-read in full
// (335 MB, three times).  v2 keeps partial sums as fp16 in tensor memory (integers up to 2048 are exact; the
// reference itself stores them as fp16, lsq.py:169), quantises two at a time with packed-half instructions, lets
// the tensor core do the shift-and-add, and writes the backward's inputs as the small integers the backward
// actually needs:
//
// ADC state v2 (uint8, channel fastest -- "NHWC": every consumer thread reads its channels with 16-byte loads)
//   plane D  [NX][M][Cout]       byte = sum_k 4^k * #{j : partial sum (i,k,j) NOT clipped}   (dgrad:  sum_j pass)
//   plane W  [NX][M][Cout]       byte = sum_j 4^j * #{k : partial sum (i,k,j) NOT clipped}   (wgrad:  sum_k pass)
//   plane C  [NX][NSA][M][Cout]  byte = sum_k 4^k * (code(i,k,j) + 1), code in {-1,0,1}      (grad_alpha; absent
//                                                                                             for the multi-bit ADC)
// With 1-bit slices the STE weights mask[k][j]*2^(-abs*j) = 2^k and mask[k][j]*2^(-wbs*k) = 2^j do not depend on
// the slice that is summed over, so the backward needs only these counts (0..NS, two bits each for NS <= 3).
//
// Constants block, one per (channel tile ct, crossbar i), copied to shared memory by one bulk copy per chunk:
//   thresholds  fp16 [EW halves][pair q = k*NSA + j][2: -(tp-1), -(tg-1)][CH]     (ternary / binary ADC)
//   B2 slabs    fp16 [j][k][g = CT/16] 16x16 K-major no-swizzle, diagonal: n(i,k,j,c) * mask[k][j] where
//               alpha_q = n * alpha_scale (the 8-bit alpha quantiser makes n an integer, lsq.py:566-571), or
//               mask[k][j] alone for the multi-bit ADC.  out = alpha_scale * sum code * n * mask  (exact integers
//               in the fp32 accumulator; one rounding instead of the reference's ~45).
#pragma once

#include "cim_tc_layout.cuh"

namespace cimq {
namespace v2 {

constexpr int kMaxNS = 3;          // digit planes per operand (two-bit count fields)
constexpr int kSlabBytes = 512;    // one 16x16 fp16 B2 slab
constexpr int kThrClamp = 4096;    // |threshold - 1| is clamped here: no partial sum (<= 2048) reaches it

// channels per CTA tile (multiple of 16, <= 64)
__host__ __device__ inline int channel_tile(const Geo &g) {
  if (g.Cout % 64 == 0) return 64;
  if (g.Cout % 32 == 0) return 32;
  if (g.Cout % 16 == 0) return 16;
  return 0;
}
// epilogue warpgroups sharing a tile's channels (each thread then owns CH = CT / EW channels of its pixel)
__host__ __device__ inline int epilogue_groups(int CT) { return CT >= 32 ? 2 : 1; }

// Does the v2 path cover this layer?  (1-bit slices, 2 or 3 digit planes each, crossbar <= 128 rows, partial sums
// exact in fp16, int8 mask without wrap-around.)
__host__ __device__ inline bool supported(const Geo &g) {
  if (g.abs_ != 1 || g.wbs != 1) return false;
  if (g.NSW != g.NSA || g.NSW < 2 || g.NSW > kMaxNS) return false;
  if (channel_tile(g) == 0) return false;
  if (tc_kp(g) > 128) return false;
  if (g.K > 5) return false;
  if ((int64_t)g.B * g.Cin * g.H * g.W >= (1ll << 31)) return false;
  if ((int64_t)g.M * g.Cout >= (1ll << 31)) return false;
  return true;
}

struct ConstLayout {
  int CT, EW, CH, G;
  uint32_t thr_bytes;   // all halves
  uint32_t b2_off, b2_bytes;
  uint32_t block_bytes; // multiple of 128
};
__host__ __device__ inline ConstLayout const_layout(const Geo &g) {
  ConstLayout c;
  c.CT = channel_tile(g);
  c.EW = epilogue_groups(c.CT);
  c.CH = c.CT / c.EW;
  c.G = c.CT / 16;
  c.thr_bytes = g.adc_mode == CIMQ_ADC_MULTIBIT ? 0u : (uint32_t)(g.pairs * 2 * c.CT * 2);
  c.b2_off = (c.thr_bytes + 127u) & ~127u;
  c.b2_bytes = (uint32_t)(g.pairs * c.G * kSlabBytes);
  c.block_bytes = (c.b2_off + c.b2_bytes + 127u) & ~127u;
  return c;
}
// v2 section of the ADC table buffer: header (256 bytes: {o0, o1, status...}) + [nct][NX] blocks
__host__ __device__ inline int64_t const_section_bytes(const Geo &g) {
  if (!supported(g)) return 0;
  const ConstLayout c = const_layout(g);
  return 256 + (int64_t)(g.Cout / c.CT) * g.NX * c.block_bytes;
}

// ---- state planes ---------------------------------------------------------------------------------------------
__host__ __device__ inline int64_t plane_bytes(const Geo &g) { return (int64_t)g.NX * g.M * g.Cout; }
__host__ __device__ inline int64_t state_bytes(const Geo &g) {
  if (!supported(g)) return 0;
  return plane_bytes(g) * (g.adc_mode == CIMQ_ADC_MULTIBIT ? 2 : 2 + g.NSA);
}
__host__ __device__ inline int64_t state_d_off(const Geo &) { return 0; }
__host__ __device__ inline int64_t state_w_off(const Geo &g) { return plane_bytes(g); }
__host__ __device__ inline int64_t state_c_off(const Geo &g, int i, int j) {
  return 2 * plane_bytes(g) + ((int64_t)i * g.NSA + j) * g.M * g.Cout;
}


// ---- backward operands (cim_bwd_v2.cu, the v2 branch of cim_bwd_tc.cu) ------------------------------------------
// Both backward GEMMs multiply grad_out * (pass count) by exact small integers.  grad_out is scaled by a power of two
// per pixel row (dgrad) / per output channel (wgrad) so that its largest magnitude lies in [2^12, 2^13), and split
// ONCE per tile into kBwdPieces fp16 pieces of PB significant bits each (round to nearest; PB = 9 for three digit
// planes, 11 for two): piece * count (count <= NS) then has at most 11 significant bits, i.e. the per-stage operand
// piece * count is exact in fp16 and costs one packed multiply per two elements.  The only error of the whole GEMM
// input is the initial rounding of grad_out to 2*PB bits: relative 2^-19 (PB = 9) or 2^-23 (PB = 11) per element.
constexpr int kBwdPieces = 2;
constexpr int kBwdScaleExp = 12;  // scaled row / channel maximum in [2^12, 2^13)
__host__ __device__ constexpr int bwd_piece_bits(int NS) { return NS >= 3 ? 9 : 11; }
// power-of-two scale for a row / channel whose largest |grad_out| has the fp32 bit pattern `maxbits`
__device__ __forceinline__ float bwd_scale_from_maxbits(uint32_t maxbits) {
  int s = kBwdScaleExp + 127 - (int)((maxbits >> 23) & 0xffu);
  s = s < -100 ? -100 : (s > 100 ? 100 : s);
  return __uint_as_float((uint32_t)(s + 127) << 23);
}
// 1 / scale for a power of two in [2^-126, 2^126]
__device__ __forceinline__ float bwd_scale_inverse(float scale) { return __uint_as_float(0x7f000000u - __float_as_uint(scale)); }

// dgrad column (GEMM N) order inside crossbar chunk i: the complete K-tuples -- the K taps (ci, ky, 0..K-1) of one
// kernel row, consecutive unfold rows -- come first, then the taps cut off by the chunk's lower and upper edge.  The
// epilogue folds a complete tuple with two warp shuffles and ONE reduction per pixel.
struct DgradCols {
  int lo, rows;   // first unfold row, rows of the chunk
  int t0, ntup;   // first complete tuple (f / K), number of complete tuples
  int nhead;      // leftover rows before the first complete tuple
};
__host__ __device__ inline DgradCols dgrad_cols(const Geo &g, int i) {
  DgradCols c;
  c.lo = i * g.xbar;
  const int hi = (c.lo + g.xbar < g.F) ? c.lo + g.xbar : g.F;
  c.rows = hi - c.lo;
  c.t0 = (c.lo + g.K - 1) / g.K;
  const int t1 = hi / g.K;
  c.ntup = t1 > c.t0 ? t1 - c.t0 : 0;
  c.nhead = c.ntup > 0 ? c.t0 * g.K - c.lo : 0;
  return c;
}
// unfold row f at column `pos` (0 <= pos < rows) of the chunk
__host__ __device__ inline int dgrad_col_f(const Geo &g, const DgradCols &c, int pos) {
  const int mid = c.ntup * g.K;
  if (pos < mid) return c.t0 * g.K + pos;
  pos -= mid;
  if (pos < c.nhead) return c.lo + pos;
  return (c.ntup > 0 ? (c.t0 + c.ntup) * g.K : c.lo) + (pos - c.nhead);
}
// unfold row held by column n of the v2 dgrad weight tile of chunk i (-1: padding column)
__host__ __device__ inline int dgrad2_tile_row(const Geo &g, int i, int n) {
  const DgradCols dc = dgrad_cols(g, i);
  return n < dc.rows ? dgrad_col_f(g, dc, n) : -1;
}

}  // namespace v2
}  // namespace cimq
