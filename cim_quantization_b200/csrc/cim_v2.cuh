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

}  // namespace v2
}  // namespace cimq
