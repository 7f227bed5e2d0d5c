// Host side of the tcgen05 forward kernel: coverage test, shared-memory plan, launch.
#include "cim_conv_tc_kernel.cuh"

namespace cimq {

namespace tcfwd {

// Staged ("fast") producer geometry: which input rows a 128-pixel tile needs per crossbar chunk and how they are
// laid out in the staging buffers.  Needs P.Kp; leaves P.fast = 0 when the layer does not qualify (the generic
// per-element gather is used then).  Shared with the v2 kernel (cim_conv_v2.cu).
void plan_producer(const Geo &g, TcParams &P) {
  P.fast = 0; P.owt = 0; P.rpt = 0; P.rk = 0; P.rk_magic = 0; P.tma_rows = 0; P.prow = 0; P.shared_rows = 0; P.prefetch = 0; P.pitch_log2 = 0; P.col0 = 0; P.raw_bytes = 0;
  if ((g.K == 3 || g.K == 5) && g.W % 4 == 0 && g.OW <= kTcTileM && kTcTileM % g.OW == 0 && P.Kp <= 128 &&
      (kTcTileM / g.OW) * g.K <= 128 && g.pad < g.K) {
    // input rows by 16-byte asynchronous copies need 16-byte aligned rows on both sides: W % 16 == 0 and column 0 of the image
    // staged at byte 16 (col0 = 16 - pad); otherwise 4-byte loads with column -pad on a word boundary
    const bool tma = g.W % 16 == 0 && g.pad <= 16;
    const int col0 = tma ? 16 - g.pad : (4 - g.pad % 4) % 4;
    const int needp = (g.OW - 1) * g.stride + g.K + col0;
    int pl = 2;
    while ((1 << pl) < needp) ++pl;
    int nslots = 0;
    for (int i = 0; i < g.NX; ++i) {
      const ChunkLayout cl = chunk_layout(g, i);
      const int n = cl.nfull + (cl.nhead > 0) + (cl.ntail > 0);
      nslots = n > nslots ? n : nslots;
    }
    // a tile that is 128 consecutive pixels of one image reads consecutive input rows: output rows share them
    const int rpt = kTcTileM / g.OW;
    const bool shared_rows = g.L % kTcTileM == 0;
    const int rk = shared_rows ? (rpt - 1) * g.stride + g.K : rpt * g.K;
    const size_t raw = ((size_t)nslots * rk * (1u << pl) + 15) & ~(size_t)15;
    if (nslots <= kMaxSlots && raw <= 48 * 1024 && pl <= 9 && rk <= 128) {
      P.fast = 1; P.owt = g.OW; P.rpt = rpt; P.pitch_log2 = pl; P.col0 = col0;
      P.rk = rk; P.prow = shared_rows ? g.stride : g.K; P.shared_rows = shared_rows ? 1 : 0;
      P.rk_magic = (uint32_t)(((1ull << 32) + rk - 1) / rk);
      P.tma_rows = tma && rk >= 2 ? 1 : 0;
      P.prefetch = !P.tma_rows && (raw / 4 + kProducerThreads - 1) / kProducerThreads <= (size_t)kPrefetchWords;
      P.raw_bytes = (uint32_t)raw;
    }
  }
}

}  // namespace tcfwd

namespace {

using namespace tcfwd;

constexpr size_t kSmemBudget = 227 * 1024 - 1024;  // dynamic shared memory per CTA minus alignment slack

__global__ void im2col_lut_kernel(Geo g, int2 *__restrict__ lut) {
  // lut[lo + pos] describes the unfold row stored at operand position `pos` of its chunk
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g.F; p += gridDim.x * blockDim.x) {
    const int i = p / g.xbar;
    const ChunkLayout cl = chunk_layout(g, i);
    const int f = chunk_row_at(g, cl, p - cl.lo);
    const int ci = f / g.KK, tap = f % g.KK, ky = tap / g.K, kx = tap % g.K;  // nn.Unfold order (lsq.py:141)
    lut[p] = make_int2((ci * g.H + ky) * g.W + kx, tap);
  }
}

inline bool slices_supported(int nsw, int nsa) {
  return (nsw == 2 && nsa == 2) || (nsw == 3 && nsa == 3) || (nsw == 4 && nsa == 4) || (nsw == 8 && nsa == 8);
}

// Fill the launch plan; returns false if the layer does not fit.
bool make_plan(const Geo &g, TcParams &P, size_t &smem) {
  P.g = g;
  P.Kp = tc_kp(g);
  P.CT = tc_channel_tile_for(g);
  if (P.CT == 0) return false;
  P.nct = g.Cout / P.CT;
  P.mtiles = (g.M + kTcTileM - 1) / kTcTileM;
  P.a_bytes = (uint32_t)(kTcTileM * P.Kp);
  P.b_bytes = (uint32_t)(g.NSW * P.CT * P.Kp);
  P.stage_bytes = g.NSA * P.a_bytes + P.b_bytes;
  uint32_t need = 2u * g.NSW * P.CT, cols = 32;
  while (cols < need) cols <<= 1;
  P.tmem_cols = cols;
  P.ttab_bytes = (uint32_t)((g.pairs * 3 * (P.CT / 2) * 4 + 15) & ~15);
  plan_producer(g, P);
  for (int attempt = 0; attempt < 2; ++attempt) {
    const size_t fixed = 2 * (size_t)P.raw_bytes + 4 * (size_t)P.ttab_bytes + kAuxBytes;
    if (fixed + P.stage_bytes <= kSmemBudget) {
      int stages = (int)((kSmemBudget - fixed) / P.stage_bytes);
      if (stages > kMaxStages) stages = kMaxStages;
      if (stages > g.NX + 1) stages = g.NX + 1;
      P.stages = stages;
      smem = (size_t)stages * P.stage_bytes + fixed + 1024;
      return true;
    }
    P.fast = 0;  // retry without the staging buffers
    P.raw_bytes = 0;
  }
  return false;
}

}  // namespace

long long *g_tc_debug = nullptr;  // set by cimq_debug_set_timers (development aid, not part of the product path)

int tc_channel_tile_for(const Geo &g) {
  // channels per CTA: multiple of 16, at most 64, NSW*CT <= 256 (UMMA N) and 2*NSW*CT <= 512 (TMEM); an
  // epilogue thread keeps CT/2 * (accumulator + partial sum + 2 floats per state word) in registers
  const int swords = (3 * g.pairs + 31) / 32;
  for (int ct = 64; ct >= 16; ct >>= 1)
    if (g.Cout % ct == 0 && g.NSW * ct <= 256 && (ct / 2) * (2 * swords + 2) <= 140) return ct;
  return 0;
}

bool tc_forward_supported(const Geo &g) {
  if (!slices_supported(g.NSW, g.NSA)) return false;
  if (g.K > 5) return false;  // tap validity mask of the generic producer is 32 bits
  if (tc_kp(g) > kTcMaxKp) return false;
  if ((int64_t)g.B * g.Cin * g.H * g.W >= (1ll << 31)) return false;
  if (g.state_words > 6) return false;
  TcParams P;
  size_t smem;
  return make_plan(g, P, smem);
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
  size_t smem = 0;
  CIMQ_REQUIRE(make_plan(g, P, smem), "tcgen05 forward plan failed");
  const WtLayout wl = wt_layout(g);
  P.xcodes = xcodes;
  if (P.tma_rows && (reinterpret_cast<uintptr_t>(xcodes) & 15u) != 0) {  // cp.async needs a 16-byte aligned image
    P.tma_rows = 0;  // same staging layout, rows by 4-byte loads instead
    P.prefetch = (P.raw_bytes / 4 + kProducerThreads - 1) / kProducerThreads <= (uint32_t)kPrefetchWords;
  }
  P.wtiles = reinterpret_cast<const uint8_t *>(wtiles) + wl.fwd_off;
  P.lut = reinterpret_cast<const int2 *>(reinterpret_cast<const uint8_t *>(wtiles) + wl.lut_off);
  P.table = reinterpret_cast<const int4 *>(table);
  P.ttab = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(table) + table_tiled_offset(g));
  P.s = s;
  P.out = out;
  P.state = state;
  P.debug = g_tc_debug;
  const int ntiles = P.mtiles * P.nct;
  const int grid = ntiles < 148 ? ntiles : 148;
  const int ch = P.CT / 2;
  if (g.NSW == 2) return launch_ns2(P, smem, grid, ch, st);
  if (g.NSW == 3) return launch_ns3(P, smem, grid, ch, st);
  if (g.NSW == 4) return launch_ns4(P, smem, grid, ch, st);
  return launch_ns8(P, smem, grid, ch, st);
}

}  // namespace cimq
