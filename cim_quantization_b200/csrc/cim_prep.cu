// Per-step preparation kernels (tiny): ADC decision table and weight digit planes.
#include <cuda_fp16.h>

#include "cim_tc_layout.cuh"
#include "cim_v2.cuh"

namespace cimq {

namespace {

constexpr int kNever = 0x3fffffff;  // threshold that no partial sum reaches

// Scaled partial sum seen by the ADC: fp16 storage, * s_w, * s_a, / alpha_q -- one IEEE op at a
// time (lsq.py:169, 195, 223 / 257-264).
__device__ __forceinline__ float adc_input(int p, float sw, float sa, float aq) {
  float v = __fmul_rn(__fmul_rn(psum_as_stored(p), sw), sa);
  return __fdiv_rn(v, aq);
}

// smallest p in [1, 65536] with pred(p), assuming pred is monotone; kNever if none.
template <class Pred>
__device__ __forceinline__ int first_true(Pred pred) {
  int lo = 1, hi = 65537;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (pred(mid)) hi = mid; else lo = mid + 1;
  }
  return lo > 65536 ? kNever : lo;
}

// table entry e = ((i*NSW + k)*NSA + j)*Cout + c  ->  int4 {tp, tg, amp bits, 0}
__global__ void adc_table_kernel(Geo g, int CT, const float *__restrict__ s, const float *__restrict__ alpha_q,
                                 const int8_t *__restrict__ mask, int4 *__restrict__ table,
                                 uint32_t *__restrict__ tiled, int32_t *__restrict__ status) {
  const int64_t n = table_entries(g);
  const float sa = s[0], sw = s[1];
  bool bad = !(sa > 0.0f) || !(sw > 0.0f) || isinf(sa) || isinf(sw);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int q = (int)((e / g.Cout) % g.pairs);
    float mk = (float)mask[q];  // mask is [NSW, NSA] row-major: index k*NSA + j == q
    int tp = kNever, tg = kNever;
    float amp = mk;
    if (g.adc_mode != CIMQ_ADC_MULTIBIT) {
      float aq = alpha_q[e];
      if (!(aq > 0.0f) || isinf(aq)) bad = true;
      amp = __fmul_rn(aq, mk);  // code * alpha_q * binary_mask with code = +-1 (lsq.py:225, 233)
      tg = first_true([&](int p) { return adc_input(p, sw, sa, aq) >= 1.00001f; });  // lsq.py:310
      if (g.adc_mode == CIMQ_ADC_TERNARY)
        tp = first_true([&](int p) { return rintf(adc_input(p, sw, sa, aq)) >= 1.0f; });  // lsq.py:224
      else
        tp = 1;  // sign(p)
    }
    table[e] = make_int4(tp, tg, __float_as_int(amp), 0);
    if (CT > 0) {  // tiled copy for the tcgen05 epilogue: [ct][i][pair][tp | tg | amp][CT]
      const int c = (int)(e % g.Cout), i = (int)(e / ((int64_t)g.pairs * g.Cout));
      const int ct = c / CT, cl = c % CT;
      uint32_t *row = tiled + ((((int64_t)ct * g.NX + i) * g.pairs + q) * 3) * CT + cl;
      // thresholds as fp32 "threshold - 1" so that sat(|p| - row) is the 0/1 decision (exact: |p| < 2^24)
      row[0] = __float_as_uint(tp == kNever ? 3.0e38f : (float)(tp - 1));
      row[CT] = __float_as_uint(tg == kNever ? 3.0e38f : (float)(tg - 1));
      row[2 * CT] = __float_as_uint(amp);
    }
  }
  if (bad && status != nullptr) atomicOr(status, 1);
}

// v2 constants blocks (cim_v2.cuh): thresholds as negated fp16 "threshold - 1", B2 slabs with the integer amplitudes.
// One CUDA block per (channel tile, crossbar); reads the AoS table written by adc_table_kernel just before.
__global__ void __launch_bounds__(256) v2_consts_kernel(Geo g, v2::ConstLayout cl, const float *__restrict__ s,
                                                        const float *__restrict__ alpha_q,
                                                        const float *__restrict__ alpha_scale,
                                                        const int8_t *__restrict__ mask,
                                                        const int4 *__restrict__ table, uint8_t *__restrict__ section,
                                                        int32_t *__restrict__ status) {
  const int ct = blockIdx.x / g.NX, i = blockIdx.x % g.NX;
  uint8_t *blk = section + 256 + (size_t)blockIdx.x * cl.block_bytes;
  const bool mb = g.adc_mode == CIMQ_ADC_MULTIBIT;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float *hdr = reinterpret_cast<float *>(section);
    if (mb) { hdr[0] = s[1]; hdr[1] = s[0]; }            // (code * s_w) * s_a, lsq.py:228-230
    else { hdr[0] = alpha_scale ? alpha_scale[0] : 0.0f; hdr[1] = 1.0f; }
  }
  for (uint32_t o = threadIdx.x * 16u; o < cl.block_bytes; o += 256u * 16u)
    *reinterpret_cast<uint4 *>(blk + o) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  bool bad = false;
  const float scale = (!mb && alpha_scale) ? alpha_scale[0] : 0.0f;
  __half *thr = reinterpret_cast<__half *>(blk);
  __half *b2 = reinterpret_cast<__half *>(blk + cl.b2_off);
  for (int idx = threadIdx.x; idx < g.pairs * cl.CT; idx += 256) {
    const int q = idx / cl.CT, cl_c = idx % cl.CT;  // q = k*NSA + j
    const int k = q / g.NSA, j = q % g.NSA;
    const int c = ct * cl.CT + cl_c;
    const int64_t e = ((int64_t)i * g.pairs + q) * g.Cout + c;
    float amp = (float)mask[q];
    if (!mb) {
      const int4 t = table[e];
      const int h = cl_c / cl.CH, ch = cl_c % cl.CH;
      const int tpm = min(t.x - 1, v2::kThrClamp), tgm = min(t.y - 1, v2::kThrClamp);
      __half *row = thr + ((size_t)(h * g.pairs + q) * 2) * cl.CH + ch;
      row[0] = __float2half_rn((float)(-tpm));  // integer negation: threshold 0 stays +0.0
      row[cl.CH] = __float2half_rn((float)(-tgm));
      // alpha_q = n * scale with integer n (lsq.py:566-571): the tensor core accumulates code * n * mask exactly
      const float aq = alpha_q[e];
      const float n = rintf(__fdiv_rn(aq, scale));
      if (!(scale > 0.0f) || __fmul_rn(n, scale) != aq || !(n >= 0.0f) || n >= 2048.0f) bad = true;
      amp = n * amp;
    }
    if (fabsf(amp) > 60000.0f) bad = true;
    // slab (j, k, g16): row = output channel within the group, K element = the same channel (diagonal)
    const int g16 = cl_c / 16, cc = cl_c % 16;
    __half *slab = b2 + (size_t)((j * g.NSW + k) * cl.G + g16) * (v2::kSlabBytes / 2);
    slab[tc_tile_offset16(cc, cc, kTcLBO, 256) / 2] = __float2half_rn(amp);
  }
  if (bad && status != nullptr) atomicOr(status, 2);
}

__global__ void weight_digits_kernel(Geo g, const int8_t *__restrict__ wcodes, float *__restrict__ wdigits) {
  const int64_t n = (int64_t)g.Cout * g.F;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int code = wcodes[idx];
    int mag = code < 0 ? -code : code, sgn = code < 0 ? -1 : 1;
    for (int k = 0; k < g.NSW; ++k)  // sign-magnitude digits, slicing_weights_signed lsq.py:438-464
      wdigits[(int64_t)k * n + idx] = (float)(sgn * ((mag >> (g.wbs * k)) & g.wmask));
  }
}

// int8 digit tiles for the tcgen05 kernel: tile (ct, i) = NSW*CT rows x Kp bytes, row = k*CT + c_local.
// enc 0: the digit as int8; enc 1: e4m3 +-0.5 (0x30 / 0xB0) for the +-1 digits of 1-bit slices (v2 forward, whose
// activation planes hold 2.0 for a set bit, so that products are +-1).
__global__ void weight_tiles_kernel(Geo g, int CT, int Kp, int enc, const int8_t *__restrict__ wcodes,
                                    int8_t *__restrict__ wtiles) {
  const int nct = g.Cout / CT;
  const int rows = g.NSW * CT;
  const int64_t tile_bytes = (int64_t)rows * Kp;
  const int64_t n = (int64_t)nct * g.NX * tile_bytes;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t tile = idx / tile_bytes;
    int within = (int)(idx % tile_bytes);
    int r = within / Kp, kk = within % Kp;  // logical (row, k-byte); scattered to its layout slot
    int ct = (int)(tile / g.NX), i = (int)(tile % g.NX);
    int k = r / CT, c = ct * CT + r % CT;
    const ChunkLayout cl = chunk_layout(g, i);
    int digit = 0;
    if (kk < cl.rows) {
      const int f = chunk_row_at(g, cl, kk);  // K order inside the chunk: see cim_tc_layout.cuh
      int code = wcodes[(int64_t)c * g.F + f];
      int mag = code < 0 ? -code : code;
      digit = (mag >> (g.wbs * k)) & g.wmask;
      if (code < 0) digit = -digit;
    }
    if (enc == 1) digit = digit > 0 ? 0x30 : (digit < 0 ? (int)(int8_t)0xB0 : 0);
    wtiles[tile * tile_bytes + tc_tile_offset(r, kk, Kp)] = (int8_t)digit;
  }
}

// ---- alpha_cim range quantiser (lsq.py:566-571) ---------------------------------------------------------
// alpha_q = clamp(round_pass(alpha / scale), qn, qp) * scale,  scale = (max(alpha) - min(alpha)) / (qp - qn)
// One block: the tensor has NX*pairs*Cout elements (a few thousand).  aux = {scale, max, min, #max, #min}.
constexpr int kAqThreads = 1024;

__device__ __forceinline__ float block_reduce_f(float v, bool is_max, float *sm) {
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : fminf(v, t);
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  v = sm[threadIdx.x & 31];
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : fminf(v, t);
  }
  __syncthreads();
  return v;
}
__device__ __forceinline__ double block_reduce_d(double v, double *sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  v = sm[threadIdx.x & 31];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(kAqThreads) alpha_quant_fwd_kernel(const float *__restrict__ alpha, int n, float qn,
                                                                     float qp, float *__restrict__ aq,
                                                                     float *__restrict__ aux) {
  __shared__ float smf[32];
  __shared__ double smd[32];
  float mx = -INFINITY, mn = INFINITY;
  for (int i = threadIdx.x; i < n; i += kAqThreads) { mx = fmaxf(mx, alpha[i]); mn = fminf(mn, alpha[i]); }
  mx = block_reduce_f(mx, true, smf);
  mn = block_reduce_f(mn, false, smf);
  const float scale = __fdiv_rn(__fsub_rn(mx, mn), __fsub_rn(qp, qn));
  double cmx = 0.0, cmn = 0.0;
  for (int i = threadIdx.x; i < n; i += kAqThreads) {
    const float a = alpha[i];
    const float r = rintf(__fdiv_rn(a, scale));  // round_pass value is exactly the rounded value
    aq[i] = __fmul_rn(fminf(fmaxf(r, qn), qp), scale);
    cmx += a == mx ? 1.0 : 0.0;
    cmn += a == mn ? 1.0 : 0.0;
  }
  cmx = block_reduce_d(cmx, smd);
  cmn = block_reduce_d(cmn, smd);
  if (threadIdx.x == 0) { aux[0] = scale; aux[1] = mx; aux[2] = mn; aux[3] = (float)cmx; aux[4] = (float)cmn; }
}

// autograd of the above w.r.t. alpha: STE through round, inclusive clamp mask, scale through max/min with
// ties sharing the gradient evenly (torch's full-reduction max/min backward)
__global__ void __launch_bounds__(kAqThreads) alpha_quant_bwd_kernel(const float *__restrict__ alpha,
                                                                     const float *__restrict__ gaq, int n, float qn,
                                                                     float qp, const float *__restrict__ aux,
                                                                     float *__restrict__ galpha) {
  __shared__ double smd[32];
  const float scale = aux[0], mx = aux[1], mn = aux[2], cmx = aux[3], cmn = aux[4];
  double gs = 0.0;
  for (int i = threadIdx.x; i < n; i += kAqThreads) {
    const float a = alpha[i], g = gaq[i];
    const float r = rintf(__fdiv_rn(a, scale));
    const bool inside = r >= qn && r <= qp;
    const float c = fminf(fmaxf(r, qn), qp);
    const float gt = inside ? g * scale : 0.0f;
    gs += (double)g * c - (double)gt * ((double)a / ((double)scale * scale));
    galpha[i] = gt / scale;
  }
  gs = block_reduce_d(gs, smd);
  const float grange = (float)(gs / (double)(qp - qn));
  for (int i = threadIdx.x; i < n; i += kAqThreads) {
    const float a = alpha[i];
    float v = galpha[i];
    if (a == mx) v += grange / cmx;
    if (a == mn) v -= grange / cmn;
    galpha[i] = v;
  }
}

}  // namespace

int launch_alpha_quant(const float *alpha, int64_t n, int qn, int qp, const float *gaq, float *out, float *aux,
                       cudaStream_t st) {
  CIMQ_REQUIRE(n > 0 && n < (1ll << 30), "alpha_quant: bad element count");
  if (gaq == nullptr)
    alpha_quant_fwd_kernel<<<1, kAqThreads, 0, st>>>(alpha, (int)n, (float)qn, (float)qp, out, aux);
  else
    alpha_quant_bwd_kernel<<<1, kAqThreads, 0, st>>>(alpha, gaq, (int)n, (float)qn, (float)qp, aux, out);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int64_t table_v2_offset(const Geo &g) { return (table_tiled_offset(g) + table_entries(g) * 12 + 255) & ~(int64_t)255; }
int64_t table_total_bytes(const Geo &g) { return table_v2_offset(g) + v2::const_section_bytes(g); }

int launch_adc_table(const Geo &g, const float *s, const float *alpha_q, const float *alpha_scale, const int8_t *mask,
                     void *table, int32_t *status, cudaStream_t st) {
  CIMQ_REQUIRE(g.adc_mode == CIMQ_ADC_MULTIBIT || alpha_q != nullptr, "adc_table: alpha_q is NULL");
  int64_t n = table_entries(g);
  int blocks = (int)((n + 127) / 128);
  adc_table_kernel<<<blocks, 128, 0, st>>>(
      g, tc_channel_tile_for(g), s, alpha_q, mask, reinterpret_cast<int4 *>(table),
      reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(table) + table_tiled_offset(g)), status);
  CIMQ_CUDA_OK(cudaGetLastError());
  // v2 constants: needs the integer amplitudes, i.e. the scale of the alpha quantiser (ternary / binary ADC)
  if (v2::supported(g) && (g.adc_mode == CIMQ_ADC_MULTIBIT || alpha_scale != nullptr)) {
    const v2::ConstLayout cl = v2::const_layout(g);
    v2_consts_kernel<<<(g.Cout / cl.CT) * g.NX, 256, 0, st>>>(
        g, cl, s, alpha_q, alpha_scale, mask, reinterpret_cast<const int4 *>(table),
        reinterpret_cast<uint8_t *>(table) + table_v2_offset(g), status);
    CIMQ_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

WtLayout wt_layout(const Geo &g) {
  auto align = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
  WtLayout w;
  const bool fwd = tc_forward_supported(g);
  w.fwd_off = 0;
  w.fwd_bytes = 0;
  if (fwd) {
    const int CT = tc_channel_tile_for(g);
    w.fwd_bytes = (int64_t)(g.Cout / CT) * g.NX * g.NSW * CT * tc_kp(g);
  }
  w.lut_off = align(w.fwd_bytes);
  w.lut_bytes = fwd ? (int64_t)g.F * 8 : 0;
  w.bwd_off = align(w.lut_off + w.lut_bytes);
  w.bwd_bytes = wtiles_bwd_bytes(g);
  w.fwd8_off = align(w.bwd_off + w.bwd_bytes);
  w.fwd8_bytes = 0;
  if (fwd && v2_forward_supported(g))
    w.fwd8_bytes = (int64_t)g.NX * g.NSW * g.Cout * tc_kp(g);  // (Cout/CT) * NX tiles of NSW*CT rows x Kp
  w.bwd2_off = align(w.fwd8_off + w.fwd8_bytes);
  w.bwd2_bytes = (w.fwd8_bytes > 0 && v2_backward_supported(g)) ? w.bwd_bytes : 0;
  w.total = align(w.bwd2_off + w.bwd2_bytes);
  if (w.fwd_bytes == 0 && w.bwd_bytes == 0) w.total = 0;
  return w;
}

int64_t wtiles_bytes(const Geo &g) { return wt_layout(g).total; }

int launch_weight_prepare(const Geo &g, const int8_t *wcodes, float *wdigits, void *wtiles, cudaStream_t st) {
  if (wdigits != nullptr) {
    int64_t n = (int64_t)g.Cout * g.F;
    weight_digits_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(g, wcodes, wdigits);
    CIMQ_CUDA_OK(cudaGetLastError());
  }
  const WtLayout wl = wt_layout(g);
  if (wtiles != nullptr && wl.fwd_bytes > 0) {
    const int CT = tc_channel_tile_for(g);
    weight_tiles_kernel<<<(int)((wl.fwd_bytes + 255) / 256), 256, 0, st>>>(
        g, CT, tc_kp(g), 0, wcodes, reinterpret_cast<int8_t *>(wtiles) + wl.fwd_off);
    CIMQ_CUDA_OK(cudaGetLastError());
    if (wl.fwd8_bytes > 0) {
      weight_tiles_kernel<<<(int)((wl.fwd8_bytes + 255) / 256), 256, 0, st>>>(
          g, v2::channel_tile(g), tc_kp(g), 1, wcodes, reinterpret_cast<int8_t *>(wtiles) + wl.fwd8_off);
      CIMQ_CUDA_OK(cudaGetLastError());
    }
    if (launch_im2col_lut(g, reinterpret_cast<uint8_t *>(wtiles) + wl.lut_off, st)) return 1;
  }
  if (wtiles != nullptr && wl.bwd_bytes > 0) {
    if (launch_weight_tiles_bwd(g, wcodes, reinterpret_cast<uint8_t *>(wtiles) + wl.bwd_off, st)) return 1;
  }
  if (wtiles != nullptr && wl.bwd2_bytes > 0) {
    if (launch_weight_tiles_bwd2(g, wcodes, reinterpret_cast<uint8_t *>(wtiles) + wl.bwd2_off, st)) return 1;
  }
  return 0;
}

}  // namespace cimq
