// Fused PIV pass kernels for 16 px interrogation windows (see piv_fused.cuh).
#define PIVB200_FUSED_IMPL
#include "fused_launch.cuh"

namespace pivb200 {
int launch_fused_w16(int loader, int sink, const CUtensorMap& ta, const CUtensorMap& tb,
                     const PassParams& p, cudaStream_t stream) {
    return launch_w<16>(loader, sink, ta, tb, p, stream);
}
}  // namespace pivb200
