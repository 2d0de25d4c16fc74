// hostcheck.cpp -- TEST-ONLY host instantiation of the engine's __host__ __device__
// arithmetic (device_math.cuh, agent_core.cuh, trail_core.cuh) so that the statement
// sequences the CUDA kernels execute can be compared with the oracle on a machine
// without a GPU.  Never linked into libslime_b200.so and never used as a fallback.
#include <cstdint>
#include <cstddef>
#include "../../slime_mold_b200/csrc/agent_core.cuh"
#include "../../slime_mold_b200/csrc/trail_core.cuh"

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
