// hostcheck.cpp -- TEST-ONLY host instantiation of the engine's __host__ __device__
// arithmetic (device_math.cuh, agent_core.cuh, trail_core.cuh) so that the statement
// sequences the CUDA kernels execute can be compared with the oracle on a machine
// without a GPU.  Never linked into libslime_b200.so and never used as a fallback.
#include <cstdint>
#include <cstddef>
#include "../../slime_mold_b200/csrc/agent_core.cuh"
#include "../../slime_mold_b200/csrc/trail_core.cuh"
#include "../../slime_mold_b200/csrc/gauss_stream.cuh"
#include "../../slime_mold_b200/csrc/gauss_rows.cuh"
#include <pthread.h>
#include <cstring>
#include <thread>
#include <vector>

struct HostLd {
    float operator()(const float* p) const { return *p; }
};

struct hc_params {   // SimSizeUniform, /root/reference/src/main.rs:29-46
    uint32_t width, height;
    float decay_factor, agent_jitter, agent_speed_min, agent_speed_max, agent_turn_speed;
    float agent_sensor_angle, agent_sensor_distance, diffusion_rate, pheromone_deposition_amount;
    float blur_radius, blur_sigma;
    uint32_t pad;
};

static smd::AgentConsts make_consts(const hc_params* p)
{
    smd::AgentConsts c{};
    c.W = p->width; c.H = p->height;
    c.Wf = (float)c.W; c.Hf = (float)c.H;
    c.rcpW = 1.0f / c.Wf; c.rcpH = 1.0f / c.Hf;
    c.neg_zero = -0.0f;
    c.xmax = c.Wf - 2.0f; c.ymax = c.Hf - 2.0f;
    c.speed_min = p->agent_speed_min; c.speed_max = p->agent_speed_max;
    c.turn_speed = p->agent_turn_speed;
    c.sensor_angle = p->agent_sensor_angle; c.sensor_distance = p->agent_sensor_distance;
    c.jitter = p->agent_jitter;
    c.row_base = 0;
    c.rows_local = (int32_t)c.H;
    c.ghost = 0;
    c.fold_hi = (int32_t)c.H;
    c.fold_lo = 0;
    return c;
}

extern "C" {

void hc_sincos_array(const float* x, float* s, float* c, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) smd::sincos(x[i], s[i], c[i]);
}
void hc_fmod_array(const float* a, const float* b, float* r, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) r[i] = smd::fmod_exact(a[i], b[i], 1.0f / b[i]);
}
void hc_div9_array(const float* a, float* r, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) r[i] = smd::div9(a[i]);
}
void hc_hash01_array(const int32_t* idx, const float* x, const float* y, float* r, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) r[i] = smd::hash01(idx[i], x[i], y[i]);
}
void hc_init_agents(float* xyas, uint64_t first, uint64_t n, uint32_t W, uint32_t H, float smin, float smax, uint64_t seed)
{
    for (uint64_t i = 0; i < n; ++i)
        smd::agent_init(seed, first + i, (float)W, (float)H, smin, smax, xyas[4 * i], xyas[4 * i + 1], xyas[4 * i + 2], xyas[4 * i + 3]);
}
void hc_agents_phase_split(float* agents, const uint32_t* ids, uint64_t n, const float* trail, uint32_t* counts,
                           const hc_params* p)
{
    smd::AgentConsts c = make_consts(p);
    for (uint64_t i = 0; i < n; ++i) {
        int32_t cx, cy;
        int32_t idx = ids ? (int32_t)ids[i] : (int32_t)i;
        smd::FetchLinear<int64_t, HostLd> fetch{trail, (int64_t)c.W, (int64_t)c.row_base, HostLd()};
        smd::agent_update(agents[4 * i], agents[4 * i + 1], agents[4 * i + 2], agents[4 * i + 3], idx, c, fetch, cx, cy);
        if (cx >= 0) counts[(size_t)cy * p->width + cx] += 1u;
    }
}
void hc_trail_pass(const float* in, uint32_t* counts, float* out, const hc_params* p)
{
    smd::TrailConsts tc{};
    tc.dep = p->pheromone_deposition_amount;
    volatile float d = p->decay_factor * 0.001f;
    tc.decay_sub = d;
    tc.rate = fminf(fmaxf(p->diffusion_rate, 0.0f), 1.0f);
    volatile float om = 1.0f - tc.rate;
    tc.one_minus_rate = om;
    const int64_t W = p->width, H = p->height;
    for (int64_t y = 0; y < H; ++y)
        for (int64_t x = 0; x < W; ++x) {
            float v[9];
            int j = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int64_t off = ((y + dy + H) % H) * W + (x + dx + W) % W;
                    float t = in[off];
                    if (counts) t = smd::merge_deposit(t, counts[off], tc.dep);
                    v[j++] = smd::decay_cell(t, tc.decay_sub);
                }
            out[y * W + x] = smd::box9_mix(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], tc);
        }
    if (counts) for (int64_t i = 0; i < W * H; ++i) counts[i] = 0;
}

}  // extern "C"

// ---- CTA emulation of the streaming Gaussian kernel (gauss_stream.cuh) -------------------------------------------
// The kernel body is written against a context; here the context is a host thread per CUDA thread and a pthread
// barrier per CTA, so the index logic, the ring bookkeeping and the statement order of the very source the GPU runs
// can be compared with the oracle on a machine without a GPU.
static bool g_stream_packed = false;
extern "C" void hc_gauss_stream_set_packed(int on) { g_stream_packed = on != 0; }

struct HostGsCtx {
    int t, bxv, byv;
    pthread_barrier_t* bar;
    int surf_w;
    int tid() const { return t; }
    int bx() const { return bxv; }
    int by() const { return byv; }
    void sync() const { pthread_barrier_wait(bar); }
    void ld4(smk::F4& v, const float* p, bool valid) const { if (valid) std::memcpy(&v, p, 16); }
    void ldu4(smk::U4& v, const uint32_t* p, bool valid) const { if (valid) std::memcpy(&v, p, 16); }
    void ldu1(uint32_t& v, const uint32_t* p, bool valid) const { if (valid) std::memcpy(&v, p, 4); }
    void surf_write(smk::F4 v, unsigned long long surf, int x, int y) const
    {
        std::memcpy(reinterpret_cast<float*>(surf) + (size_t)y * surf_w + x, &v, 16);
    }
};

template <int R, int CM, bool SURF>
static void run_gauss_stream(const smk::GsArgs& a, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    const int gx = (a.W + smk::kGsTX - 1) / smk::kGsTX, gy = (a.H + a.chunk_rows - 1) / a.chunk_rows;
    const size_t nfl = smk::gs_smem_bytes<R>() / sizeof(float);
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            // "shared memory" starts as NaN: a value that was never staged must not reach a stored output
            std::vector<float> raw(nfl + 4);
            float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw.data()) + 15) & ~(uintptr_t)15);
            for (size_t i = 0; i < nfl; ++i) smem[i] = NAN;
            pthread_barrier_t bar;
            pthread_barrier_init(&bar, nullptr, smk::kGsNT);
            std::vector<std::thread> th;
            th.reserve(smk::kGsNT);
            for (int t = 0; t < smk::kGsNT; ++t)
                th.emplace_back([&, t]() {
                    HostGsCtx cx{t, bx, by, &bar, a.W};
                    if (g_stream_packed) smk::gauss_stream_cta<R, CM, SURF, true>(cx, smem, a, tc, gc);
                    else smk::gauss_stream_cta<R, CM, SURF, false>(cx, smem, a, tc, gc);
                });
            for (auto& x : th) x.join();
            pthread_barrier_destroy(&bar);
        }
}

template <int R>
static void run_gauss_stream_r(int cm, bool surf, const smk::GsArgs& a, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    if (cm == 0) run_gauss_stream<R, smk::GS_NONE, false>(a, tc, gc);
    else if (cm == 1) { if (surf) run_gauss_stream<R, smk::GS_COUNTS, true>(a, tc, gc); else run_gauss_stream<R, smk::GS_COUNTS, false>(a, tc, gc); }
    else { if (surf) run_gauss_stream<R, smk::GS_FLAGS, true>(a, tc, gc); else run_gauss_stream<R, smk::GS_FLAGS, false>(a, tc, gc); }
}

extern "C" int hc_gauss_stream(const float* tin, const void* cin, void* czero, float* tout, float* surf_out, int W, int H,
                               int chunk_rows, int R, const float* weights, int cm, const hc_params* p, int wrap_y)
{
    if (W % 4 != 0 || W < smk::kGsMinW || H < smk::kGsMinRows || R < 1 || R > 8 || chunk_rows < 1) return -1;
    smd::TrailConsts tc{};
    tc.dep = p->pheromone_deposition_amount;
    volatile float d = p->decay_factor * 0.001f;
    tc.decay_sub = d;
    tc.rate = fminf(fmaxf(p->diffusion_rate, 0.0f), 1.0f);
    volatile float om = 1.0f - tc.rate;
    tc.one_minus_rate = om;
    smk::GaussConsts gc{};
    gc.R = R;
    for (int i = 0; i <= 2 * R; ++i) gc.set(i, weights[i]);
    smk::GsArgs a{};
    a.tin = tin; a.cin = cin; a.czero = czero; a.tout = tout;
    a.W = W; a.H = H; a.chunk_rows = chunk_rows; a.wrap_y = wrap_y;
    a.surf = (unsigned long long)reinterpret_cast<uintptr_t>(surf_out); a.surf_row0 = 0;
    const bool surf = surf_out != nullptr;
    switch (R) {
    case 1: run_gauss_stream_r<1>(cm, surf, a, tc, gc); break;
    case 2: run_gauss_stream_r<2>(cm, surf, a, tc, gc); break;
    case 3: run_gauss_stream_r<3>(cm, surf, a, tc, gc); break;
    case 4: run_gauss_stream_r<4>(cm, surf, a, tc, gc); break;
    case 5: run_gauss_stream_r<5>(cm, surf, a, tc, gc); break;
    case 6: run_gauss_stream_r<6>(cm, surf, a, tc, gc); break;
    case 7: run_gauss_stream_r<7>(cm, surf, a, tc, gc); break;
    default: run_gauss_stream_r<8>(cm, surf, a, tc, gc); break;
    }
    return 0;
}

// ---- CTA emulation of the register-streaming Gaussian kernel (gauss_rows.cuh) -------------------------------------
// Same idea; the warp shuffle becomes an exchange through a per-CTA array between two barriers (all threads of the
// emulated CTA execute every shuffle: the kernel's control flow is uniform, only its memory operations are predicated).
struct HostGrCtx : HostGsCtx {
    smk::F4* xch;
    bool warp_may_exit() const { return false; }
    template <int R>
    void neighbours(const smk::F4& v4, float (&v)[2 * R + 4]) const
    {
        xch[t] = v4;
        sync();
        const int lane = t & 31;
        // SHFL up / down by one and by two lanes (a lane without a source keeps its own value)
        const smk::F4 l1 = lane > 0 ? xch[t - 1] : v4, r1 = lane < 31 ? xch[t + 1] : v4;
        const smk::F4 l2 = lane > 1 ? xch[t - 2] : v4, r2 = lane < 30 ? xch[t + 2] : v4;
        sync();
        const float lc[8] = {l2.x, l2.y, l2.z, l2.w, l1.x, l1.y, l1.z, l1.w}, rc[8] = {r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
        for (int i = 0; i < R; ++i) { v[i] = lc[8 - R + i]; v[R + 4 + i] = rc[i]; }
    }
};

template <int R, int CM, bool SURF, int PK>
static void run_gauss_rows_pk(const smk::GsArgs& a, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    const int gx = (a.W + smk::gr_cta_cols<R>() - 1) / smk::gr_cta_cols<R>(), gy = (a.H + a.chunk_rows - 1) / a.chunk_rows;
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            std::vector<smk::F4> xch(smk::kGrNT);
            pthread_barrier_t bar;
            pthread_barrier_init(&bar, nullptr, smk::kGrNT);
            std::vector<std::thread> th;
            th.reserve(smk::kGrNT);
            for (int t = 0; t < smk::kGrNT; ++t)
                th.emplace_back([&, t]() {
                    HostGrCtx cx;
                    cx.t = t; cx.bxv = bx; cx.byv = by; cx.bar = &bar; cx.surf_w = a.W; cx.xch = xch.data();
                    smk::gauss_rows_cta<R, CM, SURF, PK>(cx, a, tc, gc);
                });
            for (auto& x : th) x.join();
            pthread_barrier_destroy(&bar);
        }
}

static int g_rows_packed = 0;
template <int R, int CM, bool SURF>
static void run_gauss_rows(const smk::GsArgs& a, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    if (g_rows_packed >= 2) run_gauss_rows_pk<R, CM, SURF, 2>(a, tc, gc);
    else if (g_rows_packed == 1) run_gauss_rows_pk<R, CM, SURF, 1>(a, tc, gc);
    else run_gauss_rows_pk<R, CM, SURF, 0>(a, tc, gc);
}

template <int R>
static void run_gauss_rows_r(int cm, bool surf, const smk::GsArgs& a, const smd::TrailConsts& tc, const smk::GaussConsts& gc)
{
    if (cm == 0) run_gauss_rows<R, smk::GS_NONE, false>(a, tc, gc);
    else if (cm == 1) { if (surf) run_gauss_rows<R, smk::GS_COUNTS, true>(a, tc, gc); else run_gauss_rows<R, smk::GS_COUNTS, false>(a, tc, gc); }
    else { if (surf) run_gauss_rows<R, smk::GS_FLAGS, true>(a, tc, gc); else run_gauss_rows<R, smk::GS_FLAGS, false>(a, tc, gc); }
}

extern "C" void hc_gauss_rows_set_packed(int level) { g_rows_packed = level; }

extern "C" int hc_gauss_rows(const float* tin, const void* cin, void* czero, float* tout, float* surf_out, int W, int H,
                             int chunk_rows, int R, const float* weights, int cm, const hc_params* p, int wrap_y)
{
    if (W % 4 != 0 || W < smk::kGrMinW || H < smk::kGrMinRows || R < 1 || R > smk::kGrMaxR || chunk_rows < 1) return -1;
    smd::TrailConsts tc{};
    tc.dep = p->pheromone_deposition_amount;
    volatile float d = p->decay_factor * 0.001f;
    tc.decay_sub = d;
    tc.rate = fminf(fmaxf(p->diffusion_rate, 0.0f), 1.0f);
    volatile float om = 1.0f - tc.rate;
    tc.one_minus_rate = om;
    smk::GaussConsts gc{};
    gc.R = R;
    for (int i = 0; i <= 2 * R; ++i) gc.set(i, weights[i]);
    smk::GsArgs a{};
    a.tin = tin; a.cin = cin; a.czero = czero; a.tout = tout;
    a.W = W; a.H = H; a.chunk_rows = chunk_rows; a.wrap_y = wrap_y;
    a.surf = (unsigned long long)reinterpret_cast<uintptr_t>(surf_out); a.surf_row0 = 0;
    const bool surf = surf_out != nullptr;
    switch (R) {
    case 1: run_gauss_rows_r<1>(cm, surf, a, tc, gc); break;
    case 2: run_gauss_rows_r<2>(cm, surf, a, tc, gc); break;
    case 3: run_gauss_rows_r<3>(cm, surf, a, tc, gc); break;
    case 4: run_gauss_rows_r<4>(cm, surf, a, tc, gc); break;
    case 5: run_gauss_rows_r<5>(cm, surf, a, tc, gc); break;
    case 6: run_gauss_rows_r<6>(cm, surf, a, tc, gc); break;
    case 7: run_gauss_rows_r<7>(cm, surf, a, tc, gc); break;
    default: run_gauss_rows_r<8>(cm, surf, a, tc, gc); break;
    }
    return 0;
}


// u8 deposit flags in 8 x 8-cell tiles: the offset function the agent kernel, the trail pass and the display pass share
// (trail_core.cuh), for tests/test_hostcheck.py's layout checks.  out[i] = byte offset of cell (x[i], y[i]) relative to owned row 0.
extern "C" void hc_flag_tile_offsets(const int64_t* x, const int64_t* y, int64_t* out, uint64_t n, int64_t W, int64_t wrap)
{
    for (uint64_t i = 0; i < n; ++i) out[i] = smd::flag_tile_offset<int64_t>(x[i], y[i], W, wrap);
}
