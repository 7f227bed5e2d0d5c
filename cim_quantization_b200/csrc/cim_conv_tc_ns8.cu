// tcgen05 forward kernel instances for NSW = NSA = 8 (one translation unit per slice count).
#include "cim_conv_tc_kernel.cuh"

namespace cimq {
namespace tcfwd {

int launch_ns8(const TcParams &P, size_t smem, int grid, int ch, cudaStream_t st) {
  if (ch == 8) return launch_instance<8, 8, 8>(P, smem, grid, st);
  CIMQ_REQUIRE(false, "no tcgen05 forward instance for 8 slices, %d channels per thread", ch);
}

}  // namespace tcfwd
}  // namespace cimq
