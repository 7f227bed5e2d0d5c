// tcgen05 (5th-gen tensor core) forward kernel of the CiM convolution for sm_100a.
// Reference semantics: get_cim_output_signed.forward, lsq.py:92-237.
//
// Implicit GEMM, one persistent CTA per SM, warp-specialised:
//
//   warps 0-3   producers   gather the im2col rows of 128 output pixels for one crossbar chunk
//                           (uint8 codes, nn.Unfold K order), split them into NSA digit planes and
//                           store them as the K-major A operands in shared memory (generic proxy,
//                           then fence.proxy.async); one lane bulk-copies the pre-tiled weight
//                           digit planes of the chunk (cp.async.bulk -> mbarrier complete_tx).
//   warp  12    MMA issuer  (warps 13-15 idle) one thread: for every activation digit plane j one accumulation group
//                           D[128 x NSW*CT] (int32, TMEM) = A_j[128 x K] * B[K x NSW*CT] with
//                           tcgen05.mma.kind::i8 (K = 32 per instruction); TMEM is double buffered
//                           so the tensor core runs ahead of the epilogue by one digit plane.
//   warps 4-11  epilogue    tcgen05.ld the partial sums (thread = output pixel, registers =
//                           channels), quantise each to its ADC code with two integer compares
//                           against the per-(crossbar,slice pair,channel) thresholds, accumulate
//                           code*alpha*2^shift in fp32 registers across slice pairs and crossbar
//                           chunks, record code/clip bits for the backward.  The partial-sum tensor
//                           never exists in memory.
//
// Pipelines: shared-memory stages (full/empty mbarriers, producers <-> MMA) and two TMEM
// accumulator buffers (tmem_full/tmem_empty mbarriers, MMA <-> epilogue).
#include "cim_tc_layout.cuh"
#include "tc_ptx.cuh"

namespace cimq {

namespace {

using namespace ptx;

constexpr int kProducerWarps = 4;
constexpr int kEpilogueWarps = 8;
constexpr int kMmaWarp = kProducerWarps + kEpilogueWarps;
// 16 warps = 4 warpgroups: registers are allocated per warpgroup, so the fourth warpgroup (MMA issuer
// + three idle warps) costs nothing extra and setmaxnreg can move its registers to the epilogue.
constexpr int kThreads = 512;
constexpr int kRegsProducer = 72, kRegsMma = 40, kRegsEpilogue = 200;  // 4*32*(72+40+2*200) = 65536
constexpr int kMaxStages = 4;

struct TcParams {
  Geo g;
  int Kp, CT, nct, mtiles, stages;
  uint32_t a_bytes, b_bytes, stage_bytes, tmem_cols;
  const uint8_t *xcodes;
  const uint8_t *wtiles;   // [nct*NX] tiles of b_bytes, then the im2col LUT
  const int2 *lut;         // per crossbar row f: {element offset of the tap, tap index}
  const int4 *table;
  const float *s;
  float *out;
  uint32_t *state;
};

// ---- the kernel ------------------------------------------------------------------------------------
// NSW/NSA: weight / activation digit planes; CH: output channels per epilogue thread (= CT/2).
template <int NSW, int NSA, int CH>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const TcParams P) {
  constexpr int CT = 2 * CH;
  constexpr int NROWS = NSW * CT;  // UMMA N
  constexpr int PAIRS = NSW * NSA;
  const Geo &g = P.g;
  constexpr int SB = 3;  // state bits per pair for binary/ternary
  constexpr int SWORDS_MAX = (SB * PAIRS + 31) / 32;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: stages, then barriers
  uint8_t *stage_base = smem_raw;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)P.stages * P.stage_bytes);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kMaxStages;
  const uint32_t tfull0 = empty0 + 8 * kMaxStages, tempty0 = tfull0 + 16;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int sidx = 0; sidx < P.stages; ++sidx) {
      mbar_init(full0 + 8 * sidx, kProducerWarps * 32 + 1);
      mbar_init(empty0 + 8 * sidx, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, kEpilogueWarps);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ntiles = P.mtiles * P.nct;
  const int rows_full = g.xbar < g.F ? g.xbar : g.F;

  if (warp < kProducerWarps) {
    // =========================== producers ===========================
    reg_dealloc<kRegsProducer>();
    const int r = threadIdx.x;  // tile row = output pixel
    const uint32_t amask4 = (uint32_t)g.amask * 0x01010101u;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int mt = tile / P.nct, ct = tile % P.nct;
      const int m = mt * kTcTileM + r;
      int base = 0;
      uint32_t vm = 0;  // bit t set: tap t of this pixel is inside the image
      if (m < g.M) {
        int b = m / g.L, l = m % g.L, oy = l / g.OW, ox = l % g.OW;
        int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
        base = (b * g.Cin * g.H + iy0) * g.W + ix0;
        for (int ky = 0; ky < g.K; ++ky)
          for (int kx = 0; kx < g.K; ++kx)
            if (iy0 + ky >= 0 && iy0 + ky < g.H && ix0 + kx >= 0 && ix0 + kx < g.W) vm |= 1u << (ky * g.K + kx);
      }
      for (int i = 0; i < g.NX; ++i, ++it) {
        const int sidx = it % P.stages;
        const uint32_t use = it / P.stages;
        mbar_wait(empty0 + 8 * sidx, (use & 1) ^ 1);
        uint8_t *st_ptr = stage_base + (size_t)sidx * P.stage_bytes;
        if (threadIdx.x == 0) {
          mbar_arrive_expect_tx(full0 + 8 * sidx, P.b_bytes);
          bulk_copy_g2s(smem_u32(st_ptr + (size_t)NSA * P.a_bytes),
                        P.wtiles + (size_t)(ct * g.NX + i) * P.b_bytes, P.b_bytes, full0 + 8 * sidx);
        }
        const int lo = i * g.xbar;
        const int rows = min(rows_full, g.F - lo);
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
          for (int j = 0; j < NSA; ++j) {  // LSB-first digit planes (slicing_act, lsq.py:466-480)
            const int sh = g.abs_ * j;
            uint4 d = make_uint4((w[0] >> sh) & amask4, (w[1] >> sh) & amask4, (w[2] >> sh) & amask4,
                                 (w[3] >> sh) & amask4);
            *reinterpret_cast<uint4 *>(st_ptr + (size_t)j * P.a_bytes + off) = d;
          }
        }
        fence_proxy_async();
        mbar_arrive(full0 + 8 * sidx);
      }
    }
  } else if (warp >= kMmaWarp) {
    // =========================== MMA issuer ===========================
    reg_dealloc<kRegsMma>();
    if (warp == kMmaWarp && lane == 0) {
      // instruction descriptor: D = S32, A = U8, B = S8, both K-major, N = NROWS, M = 128
      const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(NROWS >> 3) << 17) |
                             ((uint32_t)(kTcTileM >> 4) << 24);
      const uint32_t sbo = 8u * (uint32_t)P.Kp;
      uint32_t it = 0, acc_it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int i = 0; i < g.NX; ++i, ++it) {
          const int sidx = it % P.stages;
          const uint32_t use = it / P.stages;
          const int rows = min(rows_full, g.F - i * g.xbar);
          const int ksteps = (rows + 31) >> 5;
          mbar_wait(full0 + 8 * sidx, use & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(stage_base + (size_t)sidx * P.stage_bytes);
          const uint32_t b0 = a0 + NSA * P.a_bytes;
          for (int j = 0; j < NSA; ++j, ++acc_it) {
            const uint32_t buf = acc_it & 1, buse = acc_it >> 1;
            mbar_wait(tempty0 + 8 * buf, (buse & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * NROWS;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t adesc = make_smem_desc(a0 + j * P.a_bytes + ks * 2 * kTcLBO, kTcLBO, sbo);
              const uint64_t bdesc = make_smem_desc(b0 + ks * 2 * kTcLBO, kTcLBO, sbo);
              umma_i8(d_tmem, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(tfull0 + 8 * buf);  // accumulator of digit plane j complete -> epilogue
          }
          umma_commit(empty0 + 8 * sidx);  // all MMAs reading this stage complete -> producers
        }
      }
    }
  } else {
    // =========================== epilogue ===========================
    reg_alloc<kRegsEpilogue>();
    const int e = warp - kProducerWarps;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = e >> 2;       // which half of the channel tile
    const int r = quarter * 32 + lane;
    const float sa = P.s[0], sw = P.s[1];
    const int swords = g.state_words;
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int mt = tile / P.nct, ct = tile % P.nct;
      const int m = mt * kTcTileM + r;
      const int c_first = ct * CT + half * CH;
      float acc[CH];
#pragma unroll
      for (int cc = 0; cc < CH; ++cc) acc[cc] = 0.0f;
      for (int i = 0; i < g.NX; ++i) {
        uint32_t stw[CH][SWORDS_MAX];
#pragma unroll
        for (int cc = 0; cc < CH; ++cc)
#pragma unroll
          for (int w = 0; w < SWORDS_MAX; ++w) stw[cc][w] = 0u;
#pragma unroll
        for (int j = 0; j < NSA; ++j, ++acc_it) {
          const uint32_t buf = acc_it & 1, buse = acc_it >> 1;
          mbar_wait(tfull0 + 8 * buf, buse & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < NSW; ++k) {
            int v[CH];
            tmem_ld<CH>(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * NROWS + k * CT + half * CH, v);
            const int q = k * NSA + j;
            const int4 *te = P.table + ((size_t)i * PAIRS + q) * g.Cout + c_first;
            tmem_ld_wait();
            if (g.adc_mode == CIMQ_ADC_MULTIBIT) {
#pragma unroll
              for (int cc = 0; cc < CH; ++cc) {
                const float amp = __int_as_float(__ldg(&te[cc]).z);
                const float ph = psum_as_stored(v[cc]);
                const float cf = fminf(fmaxf(ph, (float)g.qn), (float)g.qp);
                acc[cc] += __fmul_rn(__fmul_rn(cf, sw), sa) * amp;
                if (ph > (float)g.qp || ph < (float)g.qn) stw[cc][q >> 5] |= 1u << (q & 31);
              }
            } else {
#pragma unroll
              for (int cc = 0; cc < CH; ++cc) {
                const int4 t = __ldg(&te[cc]);
                const int p = v[cc];
                const float amp = __int_as_float(t.z);
                const bool pos = p >= t.x, neg = p <= -t.x;
                acc[cc] += pos ? amp : (neg ? -amp : 0.0f);
                if (P.state != nullptr) {
                  const bool clip = (p >= t.y) || (p <= -t.y);
                  if (pos) stw[cc][q >> 5] |= 1u << (q & 31);
                  if (neg) stw[cc][(PAIRS + q) >> 5] |= 1u << ((PAIRS + q) & 31);
                  if (clip) stw[cc][(2 * PAIRS + q) >> 5] |= 1u << ((2 * PAIRS + q) & 31);
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * buf);  // this warp has drained the accumulator
        }
        if (P.state != nullptr && m < g.M) {
#pragma unroll
          for (int cc = 0; cc < CH; ++cc)
#pragma unroll
            for (int w = 0; w < SWORDS_MAX; ++w)
              if (w < swords)
                P.state[(((size_t)i * g.Cout + c_first + cc) * swords + w) * g.M + m] = stw[cc][w];
        }
      }
      if (m < g.M) {
        const int b = m / g.L, l = m % g.L;
#pragma unroll
        for (int cc = 0; cc < CH; ++cc) P.out[((size_t)b * g.Cout + c_first + cc) * g.L + l] = acc[cc];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

__global__ void im2col_lut_kernel(Geo g, int2 *__restrict__ lut) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < g.F; f += gridDim.x * blockDim.x) {
    int ci = f / g.KK, tap = f % g.KK, ky = tap / g.K, kx = tap % g.K;  // nn.Unfold order (lsq.py:141)
    lut[f] = make_int2((ci * g.H + ky) * g.W + kx, tap);
  }
}

template <int NSW, int NSA, int CH>
int launch_instance(const TcParams &P, size_t smem, int grid, cudaStream_t st) {
  CIMQ_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<NSW, NSA, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  conv_tc_kernel<NSW, NSA, CH><<<grid, kThreads, smem, st>>>(P);
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

inline bool slices_supported(int nsw, int nsa) {
  return (nsw == 2 && nsa == 2) || (nsw == 3 && nsa == 3) || (nsw == 4 && nsa == 4) || (nsw == 8 && nsa == 8);
}

constexpr size_t kSmemBudget = 227 * 1024 - 1024;  // dynamic shared memory per CTA minus alignment slack
constexpr size_t kBarrierBytes = 256;

}  // namespace

int tc_channel_tile_for(const Geo &g) {
  // channels per CTA: multiple of 16, at most 64, NSW*CT <= 256 (UMMA N) and 2*NSW*CT <= 512 (TMEM)
  for (int ct = 64; ct >= 16; ct >>= 1)
    if (g.Cout % ct == 0 && g.NSW * ct <= 256) return ct;
  return 0;
}

bool tc_forward_supported(const Geo &g) {
  if (!slices_supported(g.NSW, g.NSA)) return false;
  if (g.K > 5) return false;                         // tap validity mask is 32 bits
  if (tc_channel_tile_for(g) == 0) return false;
  if (tc_kp(g) > kTcMaxKp) return false;
  if ((int64_t)g.B * g.Cin * g.H * g.W >= (1ll << 31)) return false;
  if (3 * g.pairs > 32 * 6) return false;
  int CT = tc_channel_tile_for(g), Kp = tc_kp(g);
  size_t stage = (size_t)g.NSA * kTcTileM * Kp + (size_t)g.NSW * CT * Kp;
  return stage + kBarrierBytes <= kSmemBudget;
}

int launch_im2col_lut(const Geo &g, void *lut, cudaStream_t st) {
  im2col_lut_kernel<<<(g.F + 127) / 128, 128, 0, st>>>(g, reinterpret_cast<int2 *>(lut));
  CIMQ_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_conv_tc_forward(const Geo &g, const uint8_t *xcodes, const void *wtiles, const void *table,
                           const float *s, const int8_t *, float *out, uint32_t *state, cudaStream_t st) {
  CIMQ_REQUIRE(tc_forward_supported(g), "layer not covered by the tcgen05 kernel");
  CIMQ_REQUIRE(xcodes && wtiles && table && s && out, "conv_forward: NULL argument");
  TcParams P;
  P.g = g;
  P.Kp = tc_kp(g);
  P.CT = tc_channel_tile_for(g);
  P.nct = g.Cout / P.CT;
  P.mtiles = (g.M + kTcTileM - 1) / kTcTileM;
  P.a_bytes = (uint32_t)(kTcTileM * P.Kp);
  P.b_bytes = (uint32_t)(g.NSW * P.CT * P.Kp);
  P.stage_bytes = g.NSA * P.a_bytes + P.b_bytes;
  int stages = (int)((kSmemBudget - kBarrierBytes) / P.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > g.NX + 1) stages = g.NX + 1;
  P.stages = stages;
  uint32_t need = 2u * g.NSW * P.CT, cols = 32;
  while (cols < need) cols <<= 1;
  P.tmem_cols = cols;
  P.xcodes = xcodes;
  const WtLayout wl = wt_layout(g);
  P.wtiles = reinterpret_cast<const uint8_t *>(wtiles) + wl.fwd_off;
  P.lut = reinterpret_cast<const int2 *>(reinterpret_cast<const uint8_t *>(wtiles) + wl.lut_off);
  P.table = reinterpret_cast<const int4 *>(table);
  P.s = s;
  P.out = out;
  P.state = state;
  size_t smem = (size_t)P.stages * P.stage_bytes + kBarrierBytes + 1024;
  int ntiles = P.mtiles * P.nct;
  int grid = ntiles < 148 ? ntiles : 148;
  const int ch = P.CT / 2;
#define CIMQ_TC_CASE(W, A, C) \
  if (g.NSW == W && g.NSA == A && ch == C) return launch_instance<W, A, C>(P, smem, grid, st)
  CIMQ_TC_CASE(3, 3, 32);
  CIMQ_TC_CASE(3, 3, 16);
  CIMQ_TC_CASE(3, 3, 8);
  CIMQ_TC_CASE(4, 4, 32);
  CIMQ_TC_CASE(4, 4, 16);
  CIMQ_TC_CASE(4, 4, 8);
  CIMQ_TC_CASE(2, 2, 32);
  CIMQ_TC_CASE(2, 2, 16);
  CIMQ_TC_CASE(2, 2, 8);
  CIMQ_TC_CASE(8, 8, 16);
  CIMQ_TC_CASE(8, 8, 8);
#undef CIMQ_TC_CASE
  CIMQ_REQUIRE(false, "no tcgen05 kernel instance for NSW=%d NSA=%d CT=%d", g.NSW, g.NSA, P.CT);
}

}  // namespace cimq
