// grad_alpha_q of the CiM convolution (get_cim_output_signed.backward, lsq.py:315-334) from plane C of the v2 state
// (cim_v2.cuh):  grad_alpha_q[i,k,j,co] = mask[k][j] / sqrt(numel) * sum_m code(i,k,j,m,co) * go[m,co].
//
// Plane C holds code + 1 as 2-bit fields, byte = sum_k 4^k * (code_k + 1), layout [NX][NSA][M][Cout]:
//   sum_m code * go = sum_m (code + 1) * go - sum_m go.
// A field is multiplied in WITHOUT an integer-to-float conversion: the masked bits, read as an fp32 subnormal, are
// field * 2^(pos - 149) exactly, and go is pre-scaled by 2^100 so that the product is a normal number; the power-of-two
// factors are undone once at the end.  2 instructions (mask, FMA) per partial sum, which is what bounds this kernel
// (755 M partial sums at the microbench layer): it streams 319 MB but is issue-bound, not HBM-bound.
//
// The first version of this kernel read the planes straight from global memory with a thread = (channel quad, 8 pixels)
// mapping and spent as many instructions on addresses and on re-scaling grad_out for each of the NX*NSA planes as on
// the sums (17 per state byte).  Here a block stages tiles of P pixels -- every plane of a crossbar group plus the
// grad_out tile -- in shared memory with cp.async (double buffered); a thread owns ONE channel and every 256 / Cout-th
// plane of it: it reads grad_out once per pixel and plane subset and keeps its accumulators in registers for the
// whole kernel (no reduction inside the block), 7 instructions per state byte (byte load, 3 masks, 3 FMAs).
#include "cim_v2.cuh"
#include "cimq_common.cuh"

namespace cimq {
namespace v2 {
namespace {

constexpr int kAlphaThreads = 256;
constexpr int kAlphaMaxNXG = 5;  // crossbars per block

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

struct AlphaParams {
  Geo g;
  int P;        // pixels per tile (8 or 16; L % P == 0)
  int ntiles;   // M / P
  int phases;   // threads per channel = 256 / Cout
  int gstride;  // floats per channel row of the staged grad_out tile (P + 1: conflict-free for lanes = channels)
  uint32_t c_bytes, buf_bytes;  // staged plane bytes per tile (all planes of the group), bytes per buffer
  const float *go;
  const uint8_t *cplanes;  // plane C
  float *partial;          // [gridDim.x][table entries]
};

// NS: digit planes per operand; PPT: planes (i, j) per thread.  A block handles the `npl` planes of a group of crossbars
// (blockIdx.y); thread = (channel c, phase ph) owns planes ph, ph + phases, ... of its channel at EVERY pixel of the tile,
// so no two threads share an accumulator and the block needs no final reduction.
template <int NS, int PPT>
__device__ __forceinline__ void alpha_body(const AlphaParams &P, int i0, int npl, uint8_t *smem) {
  const Geo &g = P.g;
  const int tid = threadIdx.x;
  const int c = tid % g.Cout, ph = tid / g.Cout;
  const int plane_px = P.P * g.Cout;  // bytes of one plane of a tile
  float acc[PPT][NS], gsum = 0.0f;
#pragma unroll
  for (int q = 0; q < PPT; ++q)
#pragma unroll
    for (int k = 0; k < NS; ++k) acc[q][k] = 0.0f;

  // ---- copy plan: P and Cout are powers of two, chunk / element indices are shifts and masks of the thread index
  const int cpp_log2 = 31 - __clz(plane_px >> 4);  // 16-byte chunks per plane
  const int p_log2 = 31 - __clz(P.P);
  const int nchunks = npl << cpp_log2, nel = g.Cout << p_log2;
  auto issue = [&](int tile, int b) {
    const int64_t m0 = (int64_t)tile * P.P;
    const uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(smem + (size_t)b * P.buf_bytes);
    // planes: npl runs of P * Cout contiguous bytes
    const uint8_t *cbase = P.cplanes + ((int64_t)i0 * NS * g.M + m0) * g.Cout;
    for (int q = tid; q < nchunks; q += kAlphaThreads) {
      const int pl = q >> cpp_log2, ch16 = (q & ((1 << cpp_log2) - 1)) << 4;
      cp_async16(sbuf + pl * plane_px + ch16, cbase + (int64_t)pl * g.M * g.Cout + ch16);  // plane (i, j) = i0*NS + pl
    }
    // grad_out: Cout runs of P contiguous floats (one image: L % P == 0)
    const int bimg = (int)(m0 / g.L), l0 = (int)(m0 % g.L);
    const float *gbase = P.go + (int64_t)bimg * g.Cout * g.L + l0;
    const uint32_t sgs = sbuf + P.c_bytes;
    for (int q = tid; q < nel; q += kAlphaThreads) {
      const int cc = q >> p_log2, p = q & (P.P - 1);
      cp_async4(sgs + (uint32_t)(cc * P.gstride + p) * 4u, gbase + (int64_t)cc * g.L + p);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int tile = blockIdx.x, b = 0;
  if (tile < P.ntiles) issue(tile, 0);
  for (; tile < P.ntiles; tile += gridDim.x, b ^= 1) {
    const int nxt = tile + gridDim.x;
    if (nxt < P.ntiles) {
      issue(nxt, b ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint8_t *buf = smem + (size_t)b * P.buf_bytes;
    const float *gs = reinterpret_cast<const float *>(buf + P.c_bytes) + c * P.gstride;
    // one running pointer per owned plane (planes past the group's end alias the first: their sums are never written)
    const uint8_t *bq[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) bq[q] = buf + (ph + q * P.phases < npl ? ph + q * P.phases : 0) * plane_px + c;
#pragma unroll 4
    for (int p = 0; p < P.P; ++p) {
      const float gv = gs[p];
      gsum += gv;
      const float gsc = gv * 1.2676506002282294e30f;  // 2^100
#pragma unroll
      for (int q = 0; q < PPT; ++q) {
        const uint32_t w = *bq[q];
        bq[q] += g.Cout;
#pragma unroll
        for (int k = 0; k < NS; ++k) acc[q][k] = fmaf(gsc, __uint_as_float(w & (3u << (2 * k))), acc[q][k]);
      }
    }
    __syncthreads();  // the buffer is refilled by the copies issued at the top of the next iteration
  }

  // ---- undo the scaling, subtract sum go (every thread has summed all pixels of its channel)
  const int64_t n = table_entries(g);
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int pl = ph + q * P.phases;
    if (pl < npl) {
      const int i = i0 + pl / NS, j = pl % NS;
#pragma unroll
      for (int k = 0; k < NS; ++k)
        P.partial[(int64_t)blockIdx.x * n + ((int64_t)i * g.pairs + k * g.NSA + j) * g.Cout + c] =
            acc[q][k] * exp2f((float)(49 - 2 * k)) - gsum;
    }
  }
}

template <int NS>
__global__ void __launch_bounds__(kAlphaThreads, 4) bwd_alpha_v3_kernel(const AlphaParams P) {
  extern __shared__ __align__(16) uint8_t smem_alpha[];
  const int i0 = blockIdx.y * kAlphaMaxNXG;
  const int npl = min(kAlphaMaxNXG, P.g.NX - i0) * NS;
  const int ppt = (npl + P.phases - 1) / P.phases;
  if (ppt <= 1) alpha_body<NS, 1>(P, i0, npl, smem_alpha);
  else if (ppt <= 2) alpha_body<NS, 2>(P, i0, npl, smem_alpha);
  else if (ppt <= 4) alpha_body<NS, 4>(P, i0, npl, smem_alpha);
  else if (ppt <= 8) alpha_body<NS, 8>(P, i0, npl, smem_alpha);
  else alpha_body<NS, kAlphaMaxNXG * NS>(P, i0, npl, smem_alpha);
}


// ---- compile-time variant (Cout = 16 / 32 / 64 / 128): thread = (channel c, pixel phase sub) owns EVERY plane of its
// channel at the pixels sub, sub + PH, ... of the tile.  The mapping above gives a thread whole planes: at 16 channels
// (6 planes, 16 threads per channel) 10 of 16 threads had nothing to do, at 32 channels (9 planes, 8 threads) one thread
// of eight had two planes -- those layers ran at a quarter / a half of the 64-channel rate per partial sum.  Here the
// work is 4 pixels x all planes for every thread, grad_out is read once per pixel for all planes, a tile is always
// 1024 bytes per plane (PX = 1024 / Cout pixels) and every shared-memory address is base + immediate (no pointer
// updates): 6 instructions per state byte (byte load, 2 masks, 3 FMAs on prefix sums of the fields).  The PH sums of a channel are added through
// shared memory once, at the end of the block.
template <int NS, int NPL, int COUT>
__device__ __forceinline__ void alpha_body_ct(const AlphaParams &P, int i0, uint8_t *smem) {
  constexpr int PH = kAlphaThreads / COUT, PX = 1024 / COUT, PT = PX / PH, GS = PX + 1;
  constexpr uint32_t CB = NPL * 1024u, BUF = (CB + COUT * GS * 4u + 15u) & ~15u;
  static_assert(PT == 4, "four pixels per thread and tile");
  const Geo &g = P.g;
  const int tid = threadIdx.x;
  const int c = tid % COUT, sub = tid / COUT;
  float acc[NPL][NS], gsum = 0.0f;
#pragma unroll
  for (int q = 0; q < NPL; ++q)
#pragma unroll
    for (int k = 0; k < NS; ++k) acc[q][k] = 0.0f;

  auto issue = [&](int tile, int b) {
    const int64_t m0 = (int64_t)tile * PX;
    const uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(smem + (size_t)b * BUF);
    const uint8_t *cbase = P.cplanes + ((int64_t)i0 * NS * g.M + m0) * COUT;
#pragma unroll
    for (int q0 = 0; q0 < NPL * 64; q0 += kAlphaThreads) {  // 64 16-byte chunks per plane
      const int q = q0 + tid;
      if (NPL * 64 % kAlphaThreads == 0 || q < NPL * 64) {
        const int pl = q >> 6, ch16 = (q & 63) << 4;
        cp_async16(sbuf + pl * 1024 + ch16, cbase + (int64_t)pl * g.M * COUT + ch16);
      }
    }
    const int bimg = (int)(m0 / g.L), l0 = (int)(m0 % g.L);
    const float *gbase = P.go + (int64_t)bimg * COUT * g.L + l0;
#pragma unroll
    for (int q0 = 0; q0 < 1024; q0 += kAlphaThreads) {  // Cout runs of PX floats
      const int q = q0 + tid, cc = q / PX, p = q % PX;
      cp_async4(sbuf + CB + (uint32_t)(cc * GS + p) * 4u, gbase + (int64_t)cc * g.L + p);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int tile = blockIdx.x, b = 0;
  if (tile < P.ntiles) issue(tile, 0);
  for (; tile < P.ntiles; tile += gridDim.x, b ^= 1) {
    const int nxt = tile + gridDim.x;
    if (nxt < P.ntiles) {
      issue(nxt, b ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint8_t *bp = smem + (size_t)b * BUF + sub * COUT + c;
    const float *gs = reinterpret_cast<const float *>(smem + (size_t)b * BUF + CB) + c * GS + sub;
#pragma unroll
    for (int t = 0; t < PT; ++t) {
      const float gv = gs[t * PH];
      gsum += gv;
      const float gsc = gv * 1.2676506002282294e30f;  // 2^100
#pragma unroll
      for (int q = 0; q < NPL; ++q) {
        const uint32_t w = bp[q * 1024 + t * PH * COUT];
        // PREFIX sums: acc[k] accumulates the fields 0..k together (the low 2k+2 bits; the whole byte for the last one,
        // which needs no mask) -- one instruction less per byte; the fields are separated once, after the loop
#pragma unroll
        for (int k = 0; k < NS; ++k)
          acc[q][k] = fmaf(gsc, __uint_as_float(k == NS - 1 ? w : (w & ((4u << (2 * k)) - 1u))), acc[q][k]);
      }
    }
    __syncthreads();  // the buffer is refilled by the copies issued at the top of the next iteration
  }
#pragma unroll
  for (int q = 0; q < NPL; ++q)
#pragma unroll
    for (int k = NS - 1; k > 0; --k) acc[q][k] -= acc[q][k - 1];  // field k * 4^k alone

  // ---- add the PH pixel phases of each channel (fixed order), undo the scaling, subtract sum go
  constexpr int NV = NPL * NS;
  float *red = reinterpret_cast<float *>(smem);  // [NV + 1][PH][COUT]
#pragma unroll
  for (int q = 0; q < NPL; ++q)
#pragma unroll
    for (int k = 0; k < NS; ++k) red[((q * NS + k) * PH + sub) * COUT + c] = acc[q][k];
  red[(NV * PH + sub) * COUT + c] = gsum;
  __syncthreads();
  float gt = 0.0f;
#pragma unroll
  for (int s = 0; s < PH; ++s) gt += red[(NV * PH + s) * COUT + c];
  const int64_t n = table_entries(g);
  for (int v = sub; v < NV; v += PH) {
    float a = 0.0f;
#pragma unroll
    for (int s = 0; s < PH; ++s) a += red[(v * PH + s) * COUT + c];
    const int pl = v / NS, k = v % NS;
    const int i = i0 + pl / NS, j = pl % NS;
    P.partial[(int64_t)blockIdx.x * n + ((int64_t)i * g.pairs + k * g.NSA + j) * COUT + c] =
        a * exp2f((float)(49 - 2 * k)) - gt;
  }
}

template <int NS, int COUT>
__global__ void __launch_bounds__(kAlphaThreads, 4) bwd_alpha_ct_kernel(const AlphaParams P) {
  extern __shared__ __align__(16) uint8_t smem_alpha[];
  const int i0 = blockIdx.y * kAlphaMaxNXG;
  switch (min(kAlphaMaxNXG, P.g.NX - i0)) {
    case 1: alpha_body_ct<NS, 1 * NS, COUT>(P, i0, smem_alpha); break;
    case 2: alpha_body_ct<NS, 2 * NS, COUT>(P, i0, smem_alpha); break;
    case 3: alpha_body_ct<NS, 3 * NS, COUT>(P, i0, smem_alpha); break;
    case 4: alpha_body_ct<NS, 4 * NS, COUT>(P, i0, smem_alpha); break;
    default: alpha_body_ct<NS, 5 * NS, COUT>(P, i0, smem_alpha); break;
  }
}

template <int NS, int COUT>
int launch_alpha_ct(const AlphaParams &P, dim3 grid, int nxg, cudaStream_t st) {
  constexpr int PX = 1024 / COUT;
  const size_t buf = ((size_t)nxg * NS * 1024 + (size_t)COUT * (PX + 1) * 4 + 15) & ~(size_t)15;
  const size_t red = ((size_t)nxg * NS * NS + 1) * kAlphaThreads * 4;
  const size_t smem = 2 * buf > red ? 2 * buf : red;
  CIMQ_CUDA_OK(cudaFuncSetAttribute(bwd_alpha_ct_kernel<NS, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bwd_alpha_ct_kernel<NS, COUT><<<grid, kAlphaThreads, smem, st>>>(P);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace
}  // namespace v2

// Does the staged kernel cover this layer?  (channels divide the block, tiles stay inside one image)
bool alpha_v3_supported(const Geo &g) {
  if (!v2::supported(g) || g.adc_mode == CIMQ_ADC_MULTIBIT) return false;
  if (g.Cout > 256 || 256 % g.Cout != 0) return false;
  return g.L % 8 == 0;
}

// number of per-block partials the launch below writes (the finish kernel sums them): blocks along x
// the compile-time variant covers Cout = 16 / 32 / 64 / 128 when its tile (1024 / Cout pixels) divides an image
static bool alpha_ct_geo(const Geo &g) {
  return (g.Cout == 16 || g.Cout == 32 || g.Cout == 64 || g.Cout == 128) && g.L % (1024 / g.Cout) == 0 && g.NSW == g.NSA;
}

int alpha_v3_blocks(const Geo &g) {
  const int P = alpha_ct_geo(g) ? 1024 / g.Cout : (g.L % 16 == 0 ? 16 : 8);
  const int ntiles = g.M / P;
  const int groups = (g.NX + v2::kAlphaMaxNXG - 1) / v2::kAlphaMaxNXG;
  int bx = 148 * 4 / groups;
  if (bx > ntiles / 2) bx = ntiles / 2;  // at least two tiles per block (double buffering); small layers are latency
  if (bx < 1) bx = 1;                    // bound and want many blocks, although every block writes one partial per entry
  return bx;
}

int launch_alpha_v3(const Geo &g, const float *go, const uint8_t *cplanes, float *partial, cudaStream_t st) {
  using namespace v2;
  AlphaParams P;
  P.g = g;
  int nxg = g.NX < kAlphaMaxNXG ? g.NX : kAlphaMaxNXG;
  P.P = g.L % 16 == 0 ? 16 : 8;  // measured at the microbench layer: 8 -> 199 us, 16 -> 177 us, 32 -> 212 us
  // keep two buffers of a block under ~100 KB
  while (P.P > 8 && 2 * ((size_t)nxg * g.NSA * P.P * g.Cout + (size_t)g.Cout * (P.P + 1) * 4) > 100 * 1024) P.P >>= 1;
  P.ntiles = g.M / P.P;
  P.phases = kAlphaThreads / g.Cout;
  P.gstride = P.P + 1;
  P.c_bytes = (uint32_t)(nxg * g.NSA * P.P * g.Cout);
  P.buf_bytes = (P.c_bytes + (uint32_t)(g.Cout * P.gstride * 4) + 15u) & ~15u;
  P.go = go; P.cplanes = cplanes; P.partial = partial;
  if (alpha_ct_geo(g)) {
    CIMQ_REQUIRE((reinterpret_cast<uintptr_t>(cplanes) & 15u) == 0, "alpha-grad (v2): state planes must be 16-byte aligned");
    P.P = 1024 / g.Cout;
    P.ntiles = g.M / P.P;
    const dim3 grid(alpha_v3_blocks(g), (g.NX + kAlphaMaxNXG - 1) / kAlphaMaxNXG);
#define CIMQ_ALPHA_CT(NS_)                                                        \
    switch (g.Cout) {                                                             \
      case 16: return launch_alpha_ct<NS_, 16>(P, grid, nxg, st);                 \
      case 32: return launch_alpha_ct<NS_, 32>(P, grid, nxg, st);                 \
      case 64: return launch_alpha_ct<NS_, 64>(P, grid, nxg, st);                 \
      default: return launch_alpha_ct<NS_, 128>(P, grid, nxg, st);                \
    }
    if (g.NSW == 3) { CIMQ_ALPHA_CT(3) } else if (g.NSW == 2) { CIMQ_ALPHA_CT(2) }
#undef CIMQ_ALPHA_CT
  }
  const size_t smem = 2 * (size_t)P.buf_bytes;
  CIMQ_REQUIRE(smem <= 200 * 1024, "alpha-grad (v2): tile does not fit shared memory");
  CIMQ_REQUIRE((reinterpret_cast<uintptr_t>(cplanes) & 15u) == 0, "alpha-grad (v2): state planes must be 16-byte aligned");
  const int groups = (g.NX + kAlphaMaxNXG - 1) / kAlphaMaxNXG;
  dim3 grid(alpha_v3_blocks(g), groups);
  if (g.NSW == 3) {
    CIMQ_CUDA_OK(cudaFuncSetAttribute(bwd_alpha_v3_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_alpha_v3_kernel<3><<<grid, kAlphaThreads, smem, st>>>(P);
  } else {
    CIMQ_CUDA_OK(cudaFuncSetAttribute(bwd_alpha_v3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_alpha_v3_kernel<2><<<grid, kAlphaThreads, smem, st>>>(P);
  }
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cimq
