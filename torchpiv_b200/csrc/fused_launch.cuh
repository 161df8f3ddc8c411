// Host-side launcher of the fused pass kernels; one translation unit per window size
// (fused_w64.cu / fused_w32.cu / fused_w16.cu) so they compile in parallel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdlib>

#include "piv_params.h"

namespace pivb200 {

// returns a cudaError_t (0 = ok); -1 if the (loader, sink) combination is not built
int launch_fused_w64(int loader, int sink, const CUtensorMap& ta, const CUtensorMap& tb,
                     const PassParams& p, cudaStream_t stream);
int launch_fused_w32(int loader, int sink, const CUtensorMap& ta, const CUtensorMap& tb,
                     const PassParams& p, cudaStream_t stream);
int launch_fused_w16(int loader, int sink, const CUtensorMap& ta, const CUtensorMap& tb,
                     const PassParams& p, cudaStream_t stream);

void count_launch();

}  // namespace pivb200

#ifdef PIVB200_FUSED_IMPL
#include "piv_fused.cuh"

namespace pivb200 {

template <int W, int LOADER, int SINK>
static int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const PassParams& p_in,
                      cudaStream_t stream) {
    PassParams p = p_in;
    {
        // lock-step barriers (see piv_fused.cuh); PIVB200_SYNC_MASK overrides for experiments
        static const int env_mask = [] { const char* e = getenv("PIVB200_SYNC_MASK"); return e ? atoi(e) : -1; }();
        // measured best on B200: one barrier per job (before the inverse column step) inside groups of four
        // warps (one per scheduler): a group shares its instruction fetches, different groups overlap their
        // FP-bound and shared-memory-bound phases
        p.sync_mask = env_mask >= 0 ? env_mask : 16;
        static const int env_group = [] { const char* e = getenv("PIVB200_SYNC_GROUP"); return e ? atoi(e) : 1; }();
        static const int env_skew = [] { const char* e = getenv("PIVB200_SKEW_NS"); return e ? atoi(e) : 0; }();
        p.sync_group = env_group;
        p.skew_ns = env_skew;
    }
    using S = Smem<W, LOADER>;
    auto kern = piv_fused_kernel<W, LOADER, SINK>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::CTA_BYTES);
    if (err != cudaSuccess) return static_cast<int>(err);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // persistent: one CTA of NWARPS warps per SM, each warp strides over the jobs
    const long long njobs = (p.n_total + Geo<W>::NW - 1) / Geo<W>::NW;
    // PIVB200_NWARPS caps the warps per CTA (occupancy experiments)
    static const int env_warps = [] { const char* e = getenv("PIVB200_NWARPS"); return e ? atoi(e) : 0; }();
    int nwarps = (env_warps > 0 && env_warps < S::NWARPS) ? env_warps : S::NWARPS;
    if (p.sync_group) nwarps = (nwarps & ~3) < 4 ? 4 : (nwarps & ~3);       // whole groups of four warps, at least one
    if (p.sync_group >= 4 && nwarps % p.sync_group != 0) p.sync_group = 1;     // groups of N warps need N | nwarps
    long long grid = (njobs + nwarps - 1) / nwarps;
    if (grid > sms) grid = sms;
    kern<<<static_cast<unsigned>(grid), nwarps * 32, S::CTA_BYTES, stream>>>(ta, tb, p);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

template <int W>
static int launch_w(int loader, int sink, const CUtensorMap& ta, const CUtensorMap& tb,
                    const PassParams& p, cudaStream_t stream) {
    if (sink == SK_DISP && loader == LD_FRAME_INT) return launch_one<W, LD_FRAME_INT, SK_DISP>(ta, tb, p, stream);
    if constexpr (W == 64) {
        if (sink == SK_DISP && loader == LD_FRAME_TC) return launch_one<W, LD_FRAME_TC, SK_DISP>(ta, tb, p, stream);
    }
    if (sink == SK_DISP && loader == LD_FRAME_ALN) return launch_one<W, LD_FRAME_ALN, SK_DISP>(ta, tb, p, stream);
    if (sink == SK_WIN && loader == LD_FRAME_ALN) return launch_one<W, LD_FRAME_ALN, SK_WIN>(ta, tb, p, stream);
    if (sink == SK_DISP && loader == LD_FRAME_CWS) return launch_one<W, LD_FRAME_CWS, SK_DISP>(ta, tb, p, stream);
    if (sink == SK_WIN && loader == LD_FRAME_INT) return launch_one<W, LD_FRAME_INT, SK_WIN>(ta, tb, p, stream);
    if (sink == SK_WIN && loader == LD_FRAME_CWS) return launch_one<W, LD_FRAME_CWS, SK_WIN>(ta, tb, p, stream);
    if (sink == SK_CORR && loader == LD_EXPL_F32) return launch_one<W, LD_EXPL_F32, SK_CORR>(ta, tb, p, stream);
    if (sink == SK_CORR && loader == LD_EXPL_U8) return launch_one<W, LD_EXPL_U8, SK_CORR>(ta, tb, p, stream);
    return -1;
}

}  // namespace pivb200
#endif
