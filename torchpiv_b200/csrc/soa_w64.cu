// Pair-packed PIV pass kernels for 64 px interrogation windows (see piv_soa.cuh).
#define PIVB200_SOA_IMPL
#include "soa_launch.cuh"

namespace pivb200 {
int launch_soa_w64(int loader, const CUtensorMap& ta, const CUtensorMap& tb, const PassParams& p, cudaStream_t stream) {
    return launch_soa_w<64>(loader, ta, tb, p, stream);
}
}  // namespace pivb200
