// gauss.cu -- launch side of the Gaussian extension (no reference semantics): kernel selection per radius / map, grid
// sizing, the tile and two-pass fallbacks.  A translation unit of its own so that its ~190 kernel instantiations
// (gauss_rows.cuh, gauss_stream.cuh, k_gauss_fused) compile beside engine.cu instead of after it.
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

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// Variants that lost their comparison in round 1 and are no longer built (profiles/README.md has the tables): the rows kernel at
// radius 6-8 (167+ registers) and with packing level 1, FFMA2 forms of the streaming and tile kernels, the private-ring kernel.
constexpr int kRowsMaxBuiltR = 5;

// The streaming Gaussian kernel (gauss_stream.cuh) applies: it also merges u8 deposit flags and keeps the sampler's
// block-linear copy in step, so a Gaussian full step runs the same agent kernel as the box-blur step.
bool sm_engine::gauss_stream_ok() const
{
    return gauss_stream && !gauss_two_pass && W % 4 == 0 && W >= (uint32_t)smk::kGsMinW && rows >= (uint32_t)smk::kGsMinRows;
}

// The register-streaming kernel (gauss_rows.cuh) applies: small radius, W % 4 == 0.  Same capabilities as the streaming
// kernel (u8 flags, sampler copy, strips with ghost rows).
bool sm_engine::gauss_rows_ok() const
{
    const int R = (int)lroundf(params.blur_radius);
    return gauss_rows && !gauss_two_pass && R >= 1 && R <= gauss_rows_max_r && R <= kRowsMaxBuiltR && W % 4 == 0 &&
           W >= (uint32_t)smk::kGrMinW && rows >= (uint32_t)smk::kGrMinRows;
}

template <int R, int CM, bool SURF, int PK>
static int launch_gauss_rows_pk(sm_engine* e, const smk::GsArgs& a0, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    auto kern = smk::k_gauss_rows<R, CM, SURF, PK>;
    int per_sm = 0;
    SM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, smk::kGrNT, 0));
    if (per_sm < 1) per_sm = 1;
    // chunk height: about 64 / 128 / 256 rows (2R rows of pipeline fill per chunk), nudged so that the grid is a whole
    // number of waves of num_sms x resident CTAs
    smk::GsArgs a = a0;
    const uint64_t gx = (e->W + smk::gr_cta_cols<R>() - 1) / smk::gr_cta_cols<R>();
    const uint64_t cap = (uint64_t)e->num_sms * per_sm;
    uint64_t chunk = (uint64_t)e->tuning.gauss_chunk_rows;
    if (chunk == 0) {
        const double want_chunks = (double)e->rows / (R <= 2 ? 64.0 : R <= 4 ? 128.0 : 256.0);
        uint64_t waves = (uint64_t)llround((double)gx * want_chunks / (double)cap);
        if (waves < 1) waves = 1;
        uint64_t n_chunks = waves * cap / gx;
        if (n_chunks < 1) n_chunks = 1;
        chunk = (e->rows + n_chunks - 1) / n_chunks;
        if (chunk < 16) chunk = 16;
    }
    a.chunk_rows = (int)chunk;
    dim3 grid((unsigned)gx, (unsigned)((e->rows + chunk - 1) / chunk));
    kern<<<grid, smk::kGrNT, 0, e->stream>>>(a, tc, gc);
    SM_CUDA(cudaGetLastError());
    return SM_OK;
}

template <int R, int CM, bool SURF>
static int launch_gauss_rows(sm_engine* e, const smk::GsArgs& a, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    // FFMA2 taps (level 2: column taps + the aligned half of the row taps): 3-6 % faster at radius 3-4, 1-3 % at 1-2; at
    // radius 5 the extra registers cost a resident CTA (0.60 against 0.65 of the HBM peak) and the scalar form is used
    if constexpr (R <= 4) {
        const bool packed = e->gauss_rows_packing != 1;              // 0 auto (packed), 1 scalar, 2 packed
        if (packed) return launch_gauss_rows_pk<R, CM, SURF, 2>(e, a, tc, gc);
        return launch_gauss_rows_pk<R, CM, SURF, 0>(e, a, tc, gc);
    } else if constexpr (R == 5) {
        return launch_gauss_rows_pk<R, CM, SURF, 0>(e, a, tc, gc);
    } else {
        return sm_fail(SM_ERR_STATE, "k_gauss_rows is not built for radius %d", R);   // unreachable: gauss_rows_ok()
    }
}

template <int R, int CM, bool SURF, bool PK>
static int launch_gauss_stream_pk(sm_engine* e, const smk::GsArgs& a0, const smd::TrailConsts& tc, const smk::GaussConsts& gc);

template <int R, int CM, bool SURF>
static int launch_gauss_stream(sm_engine* e, const smk::GsArgs& a, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    // (an FFMA2 form of the taps removed 16 % of the instructions and changed nothing: this kernel waits on its barriers
    // and shared-memory round trips, not on issue slots)
    return launch_gauss_stream_pk<R, CM, SURF, false>(e, a, tc, gc);
}

template <int R, int CM, bool SURF, bool PK>
static int launch_gauss_stream_pk(sm_engine* e, const smk::GsArgs& a0, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    auto kern = smk::k_gauss_stream<R, CM, SURF, PK>;
    const size_t smem = smk::gs_smem_bytes<R>();
    SM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, smk::kGsNT, smem));
    if (per_sm < 1) per_sm = 1;
    // chunk height: about 256 rows (row halo 2R/chunk), nudged so that the grid is a whole number of waves
    // of num_sms x resident CTAs
    smk::GsArgs a = a0;
    const uint64_t gx = (e->W + smk::kGsTX - 1) / smk::kGsTX;
    const uint64_t cap = (uint64_t)e->num_sms * per_sm;
    uint64_t chunk = (uint64_t)e->tuning.gauss_chunk_rows;
    if (chunk == 0) {
        const double want_chunks = (double)e->rows / 256.0;
        uint64_t waves = (uint64_t)llround((double)gx * want_chunks / (double)cap);
        if (waves < 1) waves = 1;
        uint64_t n_chunks = waves * cap / gx;
        if (n_chunks < 1) n_chunks = 1;
        chunk = (e->rows + n_chunks - 1) / n_chunks;
        if (chunk < 32) chunk = 32;
    }
    chunk = (chunk + 7) / 8 * 8;
    a.chunk_rows = (int)chunk;
    dim3 grid((unsigned)gx, (unsigned)((e->rows + chunk - 1) / chunk));
    kern<<<grid, smk::kGsNT, smem, e->stream>>>(a, tc, gc);
    SM_CUDA(cudaGetLastError());
    return SM_OK;
}

// Everything that can make a Gaussian pass fail, checked BEFORE anything is launched (sm_set_params, and the top of
// sm_step / sm_diffuse_only): a bad radius or sigma must not leave a step half done (agents moved, deposits unmerged).
int sm_engine::check_gauss(bool has_counts) const
{
    if (!(cfg.flags & SM_FLAG_GAUSSIAN_BLUR)) return SM_OK;
    const float br = params.blur_radius;
    if (!(br >= 0.5f && br < 8.5f)) return sm_fail(SM_ERR_BAD_ARG, "gaussian blur radius must round to 1..8 (got %g)", br);
    if (!(params.blur_sigma > 0.0f) || !(params.blur_sigma < 1.0e6f)) return sm_fail(SM_ERR_BAD_ARG, "gaussian blur sigma must be > 0 and finite");
    const int R = (int)lroundf(br);
    if (world != 1) {
        // strips: diffusion-only passes (BASELINE config 5 at 2/4/8 GPUs); the R rows of the neighbours a pass reads are the
        // ghost rows sm_diffuse_only exchanges after every pass
        // full steps on strips (round 2): the pass also reads R deposit rows of each neighbour, pulled over NVLink after barrier 1
        // (exchange.cu: p2p_after_agents); peer-store exchange only
        if (has_counts && comm_ready && !p2p)
            return sm_fail(SM_ERR_STATE, "SM_FLAG_GAUSSIAN_BLUR: full steps on strips need the peer-store exchange (sm_tuning.exchange = 0)");
        if (!gauss_fast_ok())
            return sm_fail(SM_ERR_STATE, "SM_FLAG_GAUSSIAN_BLUR on strips needs the streaming kernels (W %% 4 == 0, W >= %d, >= %d rows per strip)",
                           smk::kGsMinW, smk::kGsMinRows);
        if ((uint32_t)R > ghost) return sm_fail(SM_ERR_STATE, "strip has %u ghost rows, the blur needs %d", ghost, R);
    }
    return SM_OK;
}

int sm_engine::launch_gauss(bool has_counts, const TrailPass& p)
{
    const smk::TrailGeom& g = p.g;
    const smd::TrailConsts& tc = p.tc;
    SM_TRY(check_gauss(has_counts));
    const int R = (int)lroundf(params.blur_radius);
    smk::GaussConsts gc{};
    gc.R = R;
    {   // weights: exp(-d^2 / 2 sigma^2) in f64, normalised in f64, rounded to f32 (oracle: so_gauss_weights)
        double tmp[17], s = 0.0;
        for (int d = -R; d <= R; ++d) {
            tmp[d + R] = exp(-(double)(d * d) / (2.0 * (double)params.blur_sigma * (double)params.blur_sigma));
            s += tmp[d + R];
        }
        for (int d = 0; d <= 2 * R; ++d) gc.set(d, (float)(tmp[d] / s));
    }
    const float* tin0 = trail_ptr(cur);
    float* tout0 = trail_ptr(1 - cur);
    if (gauss_fast_ok()) {
        // streaming single pass (gauss_rows.cuh for small radii, gauss_stream.cuh otherwise): counts or flags merged,
        // sampler copy written in a full step
        const bool use_rows = gauss_rows_ok();
        smk::GsArgs a{};
        a.tin = p.tin; a.cin = p.cm == smk::CM_NONE ? nullptr : p.cin; a.czero = p.cm == smk::CM_NONE ? nullptr : p.czero; a.tout = p.tout;
        a.W = (int)W; a.H = (int)rows; a.wrap_y = g.wrap_y;
        a.surf = (unsigned long long)g.surf; a.surf_row0 = g.surf_row0;
        const bool surf = g.surf != 0;
        auto go_rows = [&](auto r_tag) -> int {
            constexpr int RR = decltype(r_tag)::value;
            if (p.cm == smk::CM_NONE) return launch_gauss_rows<RR, smk::GS_NONE, false>(this, a, tc, gc);
            if (p.cm == smk::CM_COUNTS) return surf ? launch_gauss_rows<RR, smk::GS_COUNTS, true>(this, a, tc, gc)
                                                    : launch_gauss_rows<RR, smk::GS_COUNTS, false>(this, a, tc, gc);
            return surf ? launch_gauss_rows<RR, smk::GS_FLAGS, true>(this, a, tc, gc)
                        : launch_gauss_rows<RR, smk::GS_FLAGS, false>(this, a, tc, gc);
        };
        auto go = [&](auto r_tag) -> int {
            constexpr int RR = decltype(r_tag)::value;
            if (p.cm == smk::CM_NONE) return launch_gauss_stream<RR, smk::GS_NONE, false>(this, a, tc, gc);
            if (p.cm == smk::CM_COUNTS) return surf ? launch_gauss_stream<RR, smk::GS_COUNTS, true>(this, a, tc, gc)
                                                    : launch_gauss_stream<RR, smk::GS_COUNTS, false>(this, a, tc, gc);
            return surf ? launch_gauss_stream<RR, smk::GS_FLAGS, true>(this, a, tc, gc)
                        : launch_gauss_stream<RR, smk::GS_FLAGS, false>(this, a, tc, gc);
        };
        using std::integral_constant;
        if (use_rows) {
            switch (R) {
            case 1: SM_TRY(go_rows(integral_constant<int, 1>{})); break;
            case 2: SM_TRY(go_rows(integral_constant<int, 2>{})); break;
            case 3: SM_TRY(go_rows(integral_constant<int, 3>{})); break;
            case 4: SM_TRY(go_rows(integral_constant<int, 4>{})); break;
            case 5: SM_TRY(go_rows(integral_constant<int, 5>{})); break;
            case 6: SM_TRY(go_rows(integral_constant<int, 6>{})); break;
            case 7: SM_TRY(go_rows(integral_constant<int, 7>{})); break;
            default: SM_TRY(go_rows(integral_constant<int, 8>{})); break;
            }
            timing.kernel_launches += 1;
            return SM_OK;
        }
        switch (R) {
        case 1: SM_TRY(go(integral_constant<int, 1>{})); break;
        case 2: SM_TRY(go(integral_constant<int, 2>{})); break;
        case 3: SM_TRY(go(integral_constant<int, 3>{})); break;
        case 4: SM_TRY(go(integral_constant<int, 4>{})); break;
        case 5: SM_TRY(go(integral_constant<int, 5>{})); break;
        case 6: SM_TRY(go(integral_constant<int, 6>{})); break;
        case 7: SM_TRY(go(integral_constant<int, 7>{})); break;
        default: SM_TRY(go(integral_constant<int, 8>{})); break;
        }
        timing.kernel_launches += 1;
        return SM_OK;
    }
    if (W % 4 == 0 && W >= 160 && rows >= 64 && !gauss_two_pass) {
        // fused single pass: tiles with halos staged in shared memory (k_gauss_fused)
        dim3 grid(blocks_for(W, smk::kGaussTX), blocks_for(rows, smk::kGaussTY));
        auto go = [&](auto r_tag) -> int {
            constexpr int RR = decltype(r_tag)::value;
            const size_t smem = smk::gauss_smem_bytes<RR>();
            if (has_counts) {
                SM_CUDA(cudaFuncSetAttribute(smk::k_gauss_fused<RR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                smk::k_gauss_fused<RR, true><<<grid, 256, smem, stream>>>(tin0, counts_ptr(ccur), counts_ptr(1 - ccur), tout0, g, tc, gc);
            } else {
                SM_CUDA(cudaFuncSetAttribute(smk::k_gauss_fused<RR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                smk::k_gauss_fused<RR, false><<<grid, 256, smem, stream>>>(tin0, nullptr, nullptr, tout0, g, tc, gc);
            }
            return SM_OK;
        };
        using std::integral_constant;
        switch (R) {
        case 1: SM_TRY(go(integral_constant<int, 1>{})); break;
        case 2: SM_TRY(go(integral_constant<int, 2>{})); break;
        case 3: SM_TRY(go(integral_constant<int, 3>{})); break;
        case 4: SM_TRY(go(integral_constant<int, 4>{})); break;
        case 5: SM_TRY(go(integral_constant<int, 5>{})); break;
        case 6: SM_TRY(go(integral_constant<int, 6>{})); break;
        case 7: SM_TRY(go(integral_constant<int, 7>{})); break;
        default: SM_TRY(go(integral_constant<int, 8>{})); break;
        }
        timing.kernel_launches += 1;
        return SM_OK;
    }
    const size_t cells = (size_t)rows * W;
    if (!gauss_dec) SM_CUDA(cudaMalloc(&gauss_dec, cells * sizeof(float)));
    if (!gauss_hb) SM_CUDA(cudaMalloc(&gauss_hb, cells * sizeof(float)));
    const float* tin = trail_ptr(cur);
    float* tout = trail_ptr(1 - cur);
    const uint32_t* cin = counts_ptr(ccur);
    uint32_t* czero = counts_ptr(1 - ccur);
    const unsigned bs = 256;
    const size_t smem = (bs + 2 * R) * sizeof(float);
    for (uint32_t y0 = 0; y0 < rows; y0 += 32768) {
        uint32_t ny = std::min<uint32_t>(32768, rows - y0);
        dim3 grid(blocks_for(W, bs), ny);
        if (has_counts) smk::k_gauss_h<true><<<grid, bs, smem, stream>>>(tin, cin, czero, gauss_dec, gauss_hb, g, tc, gc, (int64_t)y0);
        else smk::k_gauss_h<false><<<grid, bs, smem, stream>>>(tin, nullptr, nullptr, gauss_dec, gauss_hb, g, tc, gc, (int64_t)y0);
        timing.kernel_launches += 1;
    }
    for (uint32_t y0 = 0; y0 < rows; y0 += 32768) {
        uint32_t ny = std::min<uint32_t>(32768, rows - y0);
        dim3 grid(blocks_for(W, bs), ny);
        smk::k_gauss_v<<<grid, bs, 0, stream>>>(gauss_dec, gauss_hb, tout, czero, has_counts ? 1 : 0, g, tc, gc, (int64_t)y0);
        timing.kernel_launches += 1;
    }
    return SM_OK;
}

