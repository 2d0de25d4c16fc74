// gauss_wring.cu -- launch side of the EXPERIMENT kernel of gauss_wring.cuh (SM_GAUSS_KERNEL=wring, radius 5-8, diffusion-only
// and u8-flag passes).  A translation unit of its own: its instantiations compile beside gauss.cu.
#include "../../include/slime_b200.h"
#include "kernels.cuh"
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>

#define SM_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return sm_fail(err__ == cudaErrorMemoryAllocation ? SM_ERR_OOM : SM_ERR_CUDA,     \
                           "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                           cudaGetErrorString(err__));                                         \
    } while (0)

#define SM_TRY(expr)                    \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != SM_OK) return rc__; \
    } while (0)

// EXPERIMENT (SM_GAUSS_KERNEL=wring, radius 5-8): the rows kernel with its column-tap state in a private shared-memory ring
template <int R, int CM, bool SURF>
static int launch_gauss_wring(sm_engine* e, const smk::GsArgs& a0, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    auto kern = smk::k_gauss_wring<R, CM, SURF, 2>;
    const size_t smem = smk::gw_smem_bytes<R>();
    int per_sm = 0;
    SM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, smk::kGwNT, smem));
    if (per_sm < 1) per_sm = 1;
    smk::GsArgs a = a0;
    const uint64_t gx = (e->W + smk::gr_cta_cols<R>() - 1) / smk::gr_cta_cols<R>();
    const uint64_t cap = (uint64_t)e->num_sms * per_sm;
    uint64_t chunk = (uint64_t)e->gauss_chunk;
    if (chunk == 0) {
        const double want_chunks = (double)e->rows / 256.0;
        uint64_t waves = (uint64_t)llround((double)gx * want_chunks / (double)cap);
        if (waves < 1) waves = 1;
        uint64_t n_chunks = waves * cap / gx;
        if (n_chunks < 1) n_chunks = 1;
        chunk = (e->rows + n_chunks - 1) / n_chunks;
        if (chunk < 32) chunk = 32;
    }
    a.chunk_rows = (int)chunk;
    dim3 grid((unsigned)gx, (unsigned)((e->rows + chunk - 1) / chunk));
    kern<<<grid, smk::kGwNT, smem, e->stream>>>(a, tc, gc);
    SM_CUDA(cudaGetLastError());
    return SM_OK;
}


int sm_gauss_wring_dispatch(sm_engine* e, int R, bool flags, bool surf, const smk::GsArgs& a, const smd::TrailConsts& tc,
                            const smk::GaussConsts& gc)
{
    auto go = [&](auto r_tag) -> int {
        constexpr int RR = decltype(r_tag)::value;
        if (!flags) return launch_gauss_wring<RR, smk::GS_NONE, false>(e, a, tc, gc);
        return surf ? launch_gauss_wring<RR, smk::GS_FLAGS, true>(e, a, tc, gc) : launch_gauss_wring<RR, smk::GS_FLAGS, false>(e, a, tc, gc);
    };
    using std::integral_constant;
    switch (R) {
    case 5: return go(integral_constant<int, 5>{});
    case 6: return go(integral_constant<int, 6>{});
    case 7: return go(integral_constant<int, 7>{});
    case 8: return go(integral_constant<int, 8>{});
    default: return sm_fail(SM_ERR_BAD_ARG, "the private-ring kernel is built for radius 5-8 (got %d)", R);
    }
}
