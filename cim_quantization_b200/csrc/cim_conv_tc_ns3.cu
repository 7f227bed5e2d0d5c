// tcgen05 forward kernel instances for NSW = NSA = 3 (one translation unit per slice count).
#include "cim_conv_tc_kernel.cuh"

namespace cimq {
namespace tcfwd {

int launch_ns3(const TcParams &P, size_t smem, int grid, int ch, cudaStream_t st) {
  if (ch == 32) return launch_instance<3, 3, 32>(P, smem, grid, st);
  if (ch == 16) return launch_instance<3, 3, 16>(P, smem, grid, st);
  if (ch == 8) return launch_instance<3, 3, 8>(P, smem, grid, st);
  CIMQ_REQUIRE(false, "no tcgen05 forward instance for 3 slices, %d channels per thread", ch);
}

}  // namespace tcfwd
}  // namespace cimq
