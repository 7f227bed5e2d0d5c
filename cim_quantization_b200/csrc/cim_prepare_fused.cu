// One launch per layer per step for everything that depends only on the parameters (not on the batch): the two LSQ
// step sizes (lsq.py:548, 554), the weight codes (lsq.py:555), the nbits_alpha quantiser of alpha_cim (lsq.py:566-571),
// the ADC decision table, the v2 constants blocks and the weight digit tiles of the v2 kernels.  Replaces
// cimq_step_sizes + cimq_lsq_quantize(w) + cimq_alpha_quantize + cimq_adc_table2 + cimq_weight_prepare (ten launches
// of a few microseconds each -- a fixed cost per layer that a 19-layer model pays 19 times per step and that does not
// shrink with the per-GPU batch).  Results are bit-identical to the separate entry points (tests/test_gpu_v2.py).
//
// No grid-wide dependency: every block recomputes the scalars (step sizes; max / min of alpha_cim, a few thousand
// values) and every output element is a pure function of those scalars and of one weight / one alpha value.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "cim_tc_layout.cuh"
#include "cim_v2.cuh"

namespace cimq {

namespace {

constexpr int kThreads = 512;
constexpr int kNever = 0x3fffffff;

struct PrepArgs {
  Geo g;
  v2::ConstLayout cl;
  int Kp, Nf, Kc2;          // forward K padding, dgrad N padding, dgrad channel block
  int n_alpha;              // NX * pairs * Cout (0 for the multi-bit ADC)
  float ga, gw, aq_qn, aq_qp;
  const float *weight, *alpha_act, *alpha_weight, *alpha_cim;
  const int8_t *mask;
  float *s;                 // [2] = {s_a, s_w}
  int8_t *wcodes;           // [Cout, F]
  float *alpha_q, *aux;     // [n_alpha], [8]
  int4 *table;              // AoS table {tp, tg, amp, 0}
  uint8_t *v2sec;           // v2 section of the table buffer (header + constants blocks)
  uint8_t *fwd8;            // e4m3 forward tiles
  uint16_t *bwd2;           // fp16 dgrad tiles (v2)
  int2 *lut;                // im2col LUT
  int32_t *status;
  int64_t seg[7];           // prefix offsets of the work segments
};

__device__ __forceinline__ float adc_in(int p, float sw, float sa, float aq) {
  return __fdiv_rn(__fmul_rn(__fmul_rn(psum_as_stored(p), sw), sa), aq);
}
template <class Pred>
__device__ __forceinline__ int first_true_p(Pred pred) {
  int lo = 1, hi = 65537;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (pred(mid)) hi = mid; else lo = mid + 1;
  }
  return lo > 65536 ? kNever : lo;
}
__device__ __forceinline__ float block_reduce_max(float v, float *sm) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  v = (threadIdx.x & 31) < kThreads / 32 ? sm[threadIdx.x & 31] : -INFINITY;
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  return v;
}
// weight code rint(clamp(w / s_w, qn, qp)) with the IEEE division (== lsq_code of lsq_quant.cu, which only
// approximates the quotient where the approximation cannot change the result)
__device__ __forceinline__ int weight_code(float w, float sw, float qn, float qp) {
  return (int)rintf(fminf(fmaxf(__fdiv_rn(w, sw), qn), qp));
}
__device__ __forceinline__ int weight_digit(const Geo &g, int code, int k) {
  const int mag = code < 0 ? -code : code;
  const int d = (mag >> (g.wbs * k)) & g.wmask;
  return code < 0 ? -d : d;
}

__global__ void __launch_bounds__(kThreads) layer_prepare_kernel(const PrepArgs A) {
  __shared__ float smf[32];
  const Geo &g = A.g;
  const bool mb = g.adc_mode == CIMQ_ADC_MULTIBIT;
  // ---- scalars, recomputed by every block
  const float a = A.alpha_act[0], ag = __fmul_rn(a, A.ga);
  const float sa = __fadd_rn(__fsub_rn(a, ag), ag);       // grad_scale value, lsq.py:23-26
  const float w_ = A.alpha_weight[0], wg = __fmul_rn(w_, A.gw);
  const float sw = __fadd_rn(__fsub_rn(w_, wg), wg);
  float mx = -INFINITY, mn = INFINITY, scale = 0.0f;
  if (!mb) {
    for (int i = threadIdx.x; i < A.n_alpha; i += kThreads) {
      const float v = A.alpha_cim[i];
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
    mx = block_reduce_max(mx, smf);
    mn = -block_reduce_max(-mn, smf);
    scale = __fdiv_rn(__fsub_rn(mx, mn), __fsub_rn(A.aq_qp, A.aq_qn));  // lsq.py:567
  }
  const float qn_w = -(float)(1 << (g.wbits - 1)), qp_w = (float)((1 << (g.wbits - 1)) - 1);
  bool bad = !(sa > 0.0f) || !(sw > 0.0f) || isinf(sa) || isinf(sw);
  bool bad2 = false;
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      A.s[0] = sa;
      A.s[1] = sw;
      float *hdr = reinterpret_cast<float *>(A.v2sec);
      if (mb) { hdr[0] = sw; hdr[1] = sa; } else { hdr[0] = scale; hdr[1] = 1.0f; }
    }
    if (!mb) {  // aux = {scale, max, min, #max, #min} for the backward of the alpha quantiser
      float cmx = 0.0f, cmn = 0.0f;
      for (int i = threadIdx.x; i < A.n_alpha; i += kThreads) {
        const float v = A.alpha_cim[i];
        cmx += v == mx ? 1.0f : 0.0f;
        cmn += v == mn ? 1.0f : 0.0f;
      }
      // counts are small integers: summing them with the max-reduction of partial prefix sums would be wrong, so
      // reduce with shuffles + shared memory adds
      for (int o = 16; o > 0; o >>= 1) { cmx += __shfl_xor_sync(0xffffffffu, cmx, o); cmn += __shfl_xor_sync(0xffffffffu, cmn, o); }
      __shared__ float c1[32], c2[32];
      if ((threadIdx.x & 31) == 0) { c1[threadIdx.x >> 5] = cmx; c2[threadIdx.x >> 5] = cmn; }
      __syncthreads();
      if (threadIdx.x == 0) {
        float t1 = 0.0f, t2 = 0.0f;
        for (int w8 = 0; w8 < kThreads / 32; ++w8) { t1 += c1[w8]; t2 += c2[w8]; }
        A.aux[0] = scale; A.aux[1] = mx; A.aux[2] = mn; A.aux[3] = t1; A.aux[4] = t2;
      }
    }
  }
  const v2::ConstLayout &cl = A.cl;
  const int nct = g.Cout / cl.CT;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < A.seg[6]; idx += (int64_t)gridDim.x * kThreads) {
    if (idx < A.seg[0]) {
      // ---- one (crossbar, slice pair, channel): alpha_q, thresholds (AoS table + v2 constants block)
      const int64_t e = idx;
      const int c = (int)(e % g.Cout), q = (int)((e / g.Cout) % g.pairs), i = (int)(e / ((int64_t)g.pairs * g.Cout));
      const float mk = (float)A.mask[q];
      int tp = kNever, tg = kNever;
      float amp = mk;
      if (!mb) {
        const float al = A.alpha_cim[e];
        const float r = rintf(__fdiv_rn(al, scale));                      // round_pass value
        const float n = fminf(fmaxf(r, A.aq_qn), A.aq_qp);
        const float aq = __fmul_rn(n, scale);
        A.alpha_q[e] = aq;
        if (!(aq > 0.0f) || isinf(aq)) bad = true;
        if (!(scale > 0.0f) || n >= 2048.0f) bad2 = true;
        amp = __fmul_rn(aq, mk);
        tg = first_true_p([&](int p) { return adc_in(p, sw, sa, aq) >= 1.00001f; });   // lsq.py:310
        tp = g.adc_mode == CIMQ_ADC_TERNARY ? first_true_p([&](int p) { return rintf(adc_in(p, sw, sa, aq)) >= 1.0f; })
                                           : 1;                                        // lsq.py:224 / sign(p)
        const int ct = c / cl.CT, cl_c = c % cl.CT, h = cl_c / cl.CH, ch = cl_c % cl.CH;
        __half *thr = reinterpret_cast<__half *>(A.v2sec + 256 + (size_t)(ct * g.NX + i) * cl.block_bytes);
        __half *row = thr + ((size_t)(h * g.pairs + q) * 2) * cl.CH + ch;
        row[0] = __float2half_rn((float)(-min(tp - 1, v2::kThrClamp)));  // integer negation: threshold 0 stays +0.0
        row[cl.CH] = __float2half_rn((float)(-min(tg - 1, v2::kThrClamp)));
      }
      A.table[e] = make_int4(tp, tg, __float_as_int(amp), 0);
    } else if (idx < A.seg[1]) {
      // ---- one element of a B2 slab: n(i,k,j,c) * mask[k][j] on the diagonal, zero elsewhere
      const int64_t t = idx - A.seg[0];
      const int el = (int)(t & 255);
      int64_t sl = t >> 8;                       // slab index: ((ct*NX + i)*pairs + (j*NSW + k))*G + g16
      const int g16 = (int)(sl % cl.G); sl /= cl.G;
      const int jk = (int)(sl % g.pairs); sl /= g.pairs;
      const int i = (int)(sl % g.NX), ct = (int)(sl / g.NX);
      const int j = jk / g.NSW, k = jk % g.NSW;
      // element el of the slab in storage order -> (row r, k-element kk): offset = (r>>3)*256 + (kk>>3)*128 + (r&7)*16 + (kk&7)*2
      const int r = ((el >> 7) << 3) | ((el >> 3) & 7), kk = (((el >> 6) & 1) << 3) | (el & 7);
      float v = 0.0f;
      if (r == kk) {
        const int c = ct * cl.CT + g16 * 16 + r, q = k * g.NSA + j;
        v = (float)A.mask[q];
        if (!mb) {
          const float al = A.alpha_cim[((int64_t)i * g.pairs + q) * g.Cout + c];
          const float n = fminf(fmaxf(rintf(__fdiv_rn(al, scale)), A.aq_qn), A.aq_qp);
          v *= n;
        }
        if (fabsf(v) > 60000.0f) bad2 = true;
      }
      __half *b2 = reinterpret_cast<__half *>(A.v2sec + 256 + (size_t)(ct * g.NX + i) * cl.block_bytes + cl.b2_off);
      b2[(size_t)(jk * cl.G + g16) * 256 + el] = __float2half_rn(v);
    } else if (idx < A.seg[2]) {
      const int64_t t = idx - A.seg[1];
      A.wcodes[t] = (int8_t)weight_code(A.weight[t], sw, qn_w, qp_w);
    } else if (idx < A.seg[3]) {
      // ---- e4m3 forward tile byte: tile (ct, i) = NSW*CT rows x Kp bytes in tc_tile_offset order
      const int64_t t = idx - A.seg[2];
      const int64_t tile_bytes = (int64_t)g.NSW * cl.CT * A.Kp;
      const int64_t tile = t / tile_bytes;
      const int within = (int)(t % tile_bytes);
      const int r = within / A.Kp, kk = within % A.Kp;
      const int ct = (int)(tile / g.NX), i = (int)(tile % g.NX);
      const int k = r / cl.CT, c = ct * cl.CT + r % cl.CT;
      const ChunkLayout lay = chunk_layout(g, i);
      int digit = 0;
      if (kk < lay.rows) {
        const int f = chunk_row_at(g, lay, kk);
        digit = weight_digit(g, weight_code(A.weight[(int64_t)c * g.F + f], sw, qn_w, qp_w), k);
      }
      A.fwd8[tile * tile_bytes + tc_tile_offset(r, kk, A.Kp)] = digit > 0 ? 0x30 : (digit < 0 ? 0xB0 : 0);
    } else if (idx < A.seg[4]) {
      // ---- fp16 dgrad tile element: tile (channel block, i, k) = [Nf columns x Kc2 channels], column n = unfold row
      // dgrad_col_f(i, n), value 2^k * digit (same bytes as weight_tiles_dgrad2_kernel, cim_bwd_v2.cu)
      const int64_t t = idx - A.seg[3];
      const int64_t tile_elems = (int64_t)A.Nf * A.Kc2;
      const int64_t tile = t / tile_elems;
      const int within = (int)(t % tile_elems);
      const int col = within / A.Kc2, cl_c = within % A.Kc2;
      const int64_t per_block = (int64_t)g.NX * g.NSW;
      const int cb = (int)(tile / per_block), ik = (int)(tile % per_block);
      const int i = ik / g.NSW, k = ik % g.NSW;
      const int co = cb * A.Kc2 + cl_c;
      const int f = v2::dgrad2_tile_row(g, i, col);
      int digit = 0;
      if (f >= 0) digit = weight_digit(g, weight_code(A.weight[(int64_t)co * g.F + f], sw, qn_w, qp_w), k);
      A.bwd2[tile * tile_elems + tc_tile_offset16(col, cl_c, kTcLBO, (uint32_t)A.Kc2 * 16u) / 2] =
          __half_as_ushort(__float2half_rn((float)(digit * (1 << k))));
    } else if (idx >= A.seg[5]) {
      // ---- padding of the constants blocks (between thresholds and slabs, after the slabs): zero, so that the block is
      // byte-identical to what cimq_adc_table2 writes
      const int64_t t = idx - A.seg[5];
      const uint32_t pad0 = (cl.b2_off - cl.thr_bytes) / 2, pad1 = (cl.block_bytes - cl.b2_off - cl.b2_bytes) / 2;
      const int64_t blk = t / (pad0 + pad1);
      const uint32_t e = (uint32_t)(t % (pad0 + pad1));
      uint16_t *bp = reinterpret_cast<uint16_t *>(A.v2sec + 256 + (size_t)blk * cl.block_bytes);
      bp[e < pad0 ? cl.thr_bytes / 2 + e : (cl.b2_off + cl.b2_bytes) / 2 + (e - pad0)] = 0;
    } else {
      // ---- im2col LUT of the generic forward producer
      const int p = (int)(idx - A.seg[4]);
      const int i = p / g.xbar;
      const ChunkLayout lay = chunk_layout(g, i);
      const int f = chunk_row_at(g, lay, p - lay.lo);
      const int ci = f / g.KK, tap = f % g.KK, ky = tap / g.K, kx = tap % g.K;
      A.lut[p] = make_int2((ci * g.H + ky) * g.W + kx, tap);
    }
    (void)nct;
  }
  if (A.status != nullptr) {
    if (bad) atomicOr(A.status, 1);
    if (bad2) atomicOr(A.status, 2);
  }
}

}  // namespace

int launch_layer_prepare(const Geo &g, const float *weight, const float *alpha_act, const float *alpha_weight, float ga,
                         float gw, const float *alpha_cim, int aq_qn, int aq_qp, const int8_t *mask, float *s,
                         int8_t *wcodes, float *alpha_q, float *aux, void *table, void *wtiles, int32_t *status,
                         cudaStream_t st) {
  CIMQ_REQUIRE(v2_forward_supported(g) && v2_backward_supported(g),
               "layer_prepare: layer not covered by the v2 kernels (cimq_info_t.tc_v2)");
  const bool mb = g.adc_mode == CIMQ_ADC_MULTIBIT;
  CIMQ_REQUIRE(weight && alpha_act && alpha_weight && mask && s && wcodes && table && wtiles, "layer_prepare: NULL argument");
  CIMQ_REQUIRE(mb || (alpha_cim && alpha_q && aux), "layer_prepare: alpha_cim / alpha_q / aux is NULL");
  PrepArgs A;
  A.g = g;
  A.cl = v2::const_layout(g);
  A.Kp = tc_kp(g);
  A.Nf = tc_nf(g);
  A.Kc2 = g.Cout > 64 ? 64 : g.Cout;
  A.n_alpha = mb ? 0 : (int)table_entries(g);
  A.ga = ga; A.gw = gw; A.aq_qn = (float)aq_qn; A.aq_qp = (float)aq_qp;
  A.weight = weight; A.alpha_act = alpha_act; A.alpha_weight = alpha_weight; A.alpha_cim = alpha_cim; A.mask = mask;
  A.s = s; A.wcodes = wcodes; A.alpha_q = alpha_q; A.aux = aux;
  A.table = reinterpret_cast<int4 *>(table);
  A.v2sec = reinterpret_cast<uint8_t *>(table) + table_v2_offset(g);
  const WtLayout wl = wt_layout(g);
  CIMQ_REQUIRE(wl.fwd8_bytes > 0 && wl.bwd2_bytes > 0, "layer_prepare: weight tile buffer lacks the v2 sections");
  A.fwd8 = reinterpret_cast<uint8_t *>(wtiles) + wl.fwd8_off;
  A.bwd2 = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(wtiles) + wl.bwd2_off);
  A.lut = reinterpret_cast<int2 *>(reinterpret_cast<uint8_t *>(wtiles) + wl.lut_off);
  A.status = status;
  const int nct = g.Cout / A.cl.CT;
  A.seg[0] = table_entries(g);
  A.seg[1] = A.seg[0] + (int64_t)nct * g.NX * g.pairs * A.cl.G * 256;
  A.seg[2] = A.seg[1] + (int64_t)g.Cout * g.F;
  A.seg[3] = A.seg[2] + wl.fwd8_bytes;
  A.seg[4] = A.seg[3] + wl.bwd2_bytes / 2;
  A.seg[5] = A.seg[4] + g.F;
  A.seg[6] = A.seg[5] + (int64_t)nct * g.NX * ((A.cl.block_bytes - A.cl.thr_bytes - A.cl.b2_bytes) / 2);
  int blocks = (int)((A.seg[6] + kThreads * 4 - 1) / (kThreads * 4));
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  layer_prepare_kernel<<<blocks, kThreads, 0, st>>>(A);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cimq
