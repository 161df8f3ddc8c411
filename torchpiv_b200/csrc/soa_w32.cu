// Pair-packed PIV pass kernels for 32 px interrogation windows (see piv_soa.cuh).
#define PIVB200_SOA_IMPL
#include "soa_launch.cuh"

namespace pivb200 {
int launch_soa_w32(int loader, const CUtensorMap& ta, const CUtensorMap& tb, const PassParams& p, cudaStream_t stream) {
    return launch_soa_w<32>(loader, ta, tb, p, stream);
}
}  // namespace pivb200
