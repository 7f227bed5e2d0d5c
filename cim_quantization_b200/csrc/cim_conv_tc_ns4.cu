// tcgen05 forward kernel instances for NSW = NSA = 4 (one translation unit per slice count).
#include "cim_conv_tc_kernel.cuh"

namespace cimq {
namespace tcfwd {

int launch_ns4(const TcParams &P, size_t smem, int grid, int ch, cudaStream_t st) {
  if (ch == 16) return launch_instance<4, 4, 16>(P, smem, grid, st);
  if (ch == 8) return launch_instance<4, 4, 8>(P, smem, grid, st);
  CIMQ_REQUIRE(false, "no tcgen05 forward instance for 4 slices, %d channels per thread", ch);
}

}  // namespace tcfwd
}  // namespace cimq
