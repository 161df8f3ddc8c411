// Host-side launcher of the pair-packed pass kernels (piv_soa.cuh); one translation unit per window size.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdlib>

#include "piv_params.h"

namespace pivb200 {

// displacement sink only; returns a cudaError_t (0 = ok), -1 if the loader is not built
int launch_soa_w32(int loader, const CUtensorMap& ta, const CUtensorMap& tb, const PassParams& p, cudaStream_t stream);
int launch_soa_w64(int loader, const CUtensorMap& ta, const CUtensorMap& tb, const PassParams& p, cudaStream_t stream);
int launch_soa_w16(int loader, const CUtensorMap& ta, const CUtensorMap& tb, const PassParams& p, cudaStream_t stream);

void count_launch();

}  // namespace pivb200

#ifdef PIVB200_SOA_IMPL
#include "piv_soa.cuh"

namespace pivb200 {

template <int W, int LOADER>
static int launch_soa_one(const CUtensorMap& ta, const CUtensorMap& tb, const PassParams& p_in, cudaStream_t stream) {
    PassParams p = p_in;
    // lock-step barriers (piv_soa.cuh): PIVB200_SOA_SYNC = mask of phase boundaries, PIVB200_SOA_GROUP = warps per group
    // (0 = the warps of one scheduler).  Measured best on B200 (profiles/r02h_lockstep_sweep.txt): ONE meeting point per job.
    // The 64 px kernels (72 KB of per-job code, 3 warps per scheduler; meeting before the product) and the 32 px CWS
    // kernel (meeting before frame b's row transform: bit 6) want the warps of a SCHEDULER in step -- they share its
    // instruction fetches: 8 % for the 64 px pass, 4 % for the CWS pass --, the other kernels groups of four consecutive
    // warps (one per scheduler) meeting before the product, whose FP32-bound and shared-memory-bound phases overlap
    // inside a scheduler.  Meeting at every boundary, or CTA-wide, is slower everywhere.
    static const int env_sync = [] { const char* e = getenv("PIVB200_SOA_SYNC"); return e ? atoi(e) : -1; }();
    static const int env_group = [] { const char* e = getenv("PIVB200_SOA_GROUP"); return e ? atoi(e) : -1; }();
    constexpr bool per_scheduler = (W == 64) || (W == 32 && LOADER == LD_FRAME_CWS);
    p.sync_mask = env_sync >= 0 ? env_sync : ((W == 32 && LOADER == LD_FRAME_CWS) ? 64 : 4);
    p.sync_group = env_group >= 0 ? env_group : (per_scheduler ? 0 : 4);
    using S = SmemS<W, LOADER>;
    auto kern = piv_soa_kernel<W, LOADER>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::CTA_BYTES);
    if (err != cudaSuccess) return static_cast<int>(err);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long njobs = (p.n_total + GeoS<W>::NW - 1) / GeoS<W>::NW;
    static const int env_warps = [] { const char* e = getenv("PIVB200_NWARPS"); return e ? atoi(e) : 0; }();
    int nwarps = (env_warps > 0 && env_warps < S::NWARPS) ? env_warps : S::NWARPS;
    nwarps &= ~3;                                   // whole groups of four warps (TMEM lane quarters, barrier groups)
    if (nwarps < 4) nwarps = 4;
    if (p.sync_group < 0 || (p.sync_group > 0 && nwarps % p.sync_group != 0)) p.sync_mask = 0;
    long long grid = (njobs + nwarps - 1) / nwarps;
    if (grid > sms) grid = sms;
    kern<<<static_cast<unsigned>(grid), nwarps * 32, S::CTA_BYTES, stream>>>(ta, tb, p);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

template <int W>
static int launch_soa_w(int loader, const CUtensorMap& ta, const CUtensorMap& tb, const PassParams& p, cudaStream_t stream) {
    if (loader == LD_FRAME_INT) return launch_soa_one<W, LD_FRAME_INT>(ta, tb, p, stream);
    if (loader == LD_FRAME_ALN) return launch_soa_one<W, LD_FRAME_ALN>(ta, tb, p, stream);
    if (loader == LD_FRAME_CWS) return launch_soa_one<W, LD_FRAME_CWS>(ta, tb, p, stream);
    return -1;
}

}  // namespace pivb200
#endif
