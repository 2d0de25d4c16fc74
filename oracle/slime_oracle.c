/*
 * slime_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's
 * simulation shaders, /root/reference/src/compute.wgsl, and of the pass order of
 * /root/reference/src/main.rs:1163-1235.  Never shipped, never on the product
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs load this library.
 *
 * PARITY PIN: the reference has no tests, fixtures or golden vectors
 *  and cannot be built or run in this image (no rustc, no
 * Vulkan/lavapipe), so there is no wgpu run to compare with.  What pins this
 * file to the reference is the reference's own SHADER SOURCE executed here:
 * tests/wgsl_interp.py interprets the text of compute.wgsl / display.wgsl,
 * tests/golden/make_wgsl_golden.py commits its outputs (tests/golden/wgsl_*.npz,
 * with the SHA-256 of the shader file), and tests/test_wgsl_reference.py
 * requires this restatement to reproduce them BIT FOR BIT: so_step_sequential
 * == the shader's invocations run one after another, so_step_phase_split == the
 * shader with every load of a dispatch served from the dispatch-start buffers
 * (dep >= 1), so_display == display.wgsl.  The one thing the shader text does
 * not fix -- the accuracy of sin / cos and of float % -- is the arithmetic spec
 * of sm_oracle_math.h; against numpy's libm the results
 * stay within the stated tolerance.  The Gaussian extension has no reference
 * semantics and stays "parity unpinned".
 * Also: tests/test_oracle_vs_numpy.py (independent numpy transliteration) and
 * tests/test_oracle_kats.py (hand-derived known answers).
 *
 * Update semantics offered:
 *   phase_split : all agents sense the step-start field; deposits accumulate as
 *                 integer per-cell counts k and are merged as
 *                 clamp(t + f32(k)*dep, 0, 1) (== the reference's saturating
 *                 read-modify-write whenever dep >= 1); Jacobi diffusion.
 *                 Deterministic and order-free: this is the engine's semantics.
 *   sequential  : agents 0..N-1 in order, sensing and depositing in place on
 *                 one buffer exactly like compute.wgsl:93-95,140 executed by a
 *                 single thread ("order-fixed reference run").
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off -fopenmp).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "sm_oracle_math.h"

/* == SimSizeUniform, /root/reference/src/main.rs:29-46 (56 bytes, repr(C)) == */
typedef struct {
    uint32_t width, height;
    float decay_factor;
    float agent_jitter;
    float agent_speed_min, agent_speed_max;
    float agent_turn_speed;
    float agent_sensor_angle, agent_sensor_distance;
    float diffusion_rate;
    float pheromone_deposition_amount;
    float blur_radius, blur_sigma;
    uint32_t pad;
} so_params;

int so_params_size(void) { return (int)sizeof(so_params); }

/* Small problems (tests) run faster on a few threads than on all of them. */
void so_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int so_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- thin exports of the arithmetic spec (for math parity tests) ---------- */
void so_sincos_array(const float *x, float *s, float *c, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) so_sincos(x[i], &s[i], &c[i]);
}
void so_fmod_array(const float *a, const float *b, float *r, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) r[i] = so_fmod(a[i], b[i]);
}
void so_div9_array(const float *a, float *r, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) r[i] = a[i] / 9.0f;
}
/* compute.wgsl:117 */
static inline float so_hash01(int32_t idx, float x, float y)
{
    float a = (float)idx * 12.9898f;
    float b = x * 78.233f;
    float c = y * 37.719f;
    float arg = (a + b) + c;
    float s, cc;
    so_sincos(arg, &s, &cc);
    return so_fract(s * 43758.5453f);
}
void so_hash01_array(const int32_t *idx, const float *x, const float *y, float *r, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) r[i] = so_hash01(idx[i], x[i], y[i]);
}

/* ---- initial state, /root/reference/src/main.rs:269-282 ------------------- */
void so_init_agents(float *xyas, uint64_t first_id, uint64_t n, uint32_t W, uint32_t H,
                    float speed_min, float speed_max, uint64_t seed)
{
    const float PI_F = 3.14159274101257324f; /* std::f32::consts::PI */
    float speed_range = speed_max - speed_min;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint64_t id = first_id + (uint64_t)i;
        xyas[4 * i + 0] = so_rand01(seed, id, 0) * (float)W;
        xyas[4 * i + 1] = so_rand01(seed, id, 1) * (float)H;
        xyas[4 * i + 2] = so_rand01(seed, id, 2) * 2.0f * PI_F;
        xyas[4 * i + 3] = speed_min + so_rand01(seed, id, 3) * speed_range;
    }
}

/* reassign_agent_speeds, /root/reference/src/main.rs:101-145 (seeded) */
void so_reassign_speeds(float *xyas, uint64_t first_id, uint64_t n,
                        float speed_min, float speed_max, uint64_t seed)
{
    float speed_range = speed_max - speed_min;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        xyas[4 * i + 3] = speed_min + so_rand01(seed, first_id + (uint64_t)i, 3) * speed_range;
}

/* resize rescale, /root/reference/src/main.rs:985-989 */
void so_rescale_agents(float *xyas, uint64_t n, uint32_t oldW, uint32_t oldH,
                       uint32_t newW, uint32_t newH)
{
    float fx = (float)newW / (float)oldW;
    float fy = (float)newH / (float)oldH;
    for (uint64_t i = 0; i < n; ++i) {
        xyas[4 * i + 0] *= fx;
        xyas[4 * i + 1] *= fy;
    }
}

/* ---- sample_trail_map, compute.wgsl:7-29 ---------------------------------- */
static inline float so_sample(const float *trail, uint32_t W, uint32_t H, float px, float py)
{
    float fx = floorf(px), fy = floorf(py);
    /* x0 < 0 || x1 >= W || y0 < 0 || y1 >= H  (NaN -> outside) */
    if (!(fx >= 0.0f && fx <= (float)W - 2.0f && fy >= 0.0f && fy <= (float)H - 2.0f))
        return 0.0f;
    int64_t x0 = (int64_t)fx, y0 = (int64_t)fy;
    float dx = px - fx, dy = py - fy;
    const float *row0 = trail + (size_t)y0 * W + (size_t)x0;
    const float *row1 = row0 + W;
    float v0 = so_mix(row0[0], row0[1], dx);
    float v1 = so_mix(row1[0], row1[1], dx);
    return so_mix(v0, v1, dy);
}

/* ---- one agent, compute.wgsl:65-144.  Returns the deposit cell (or -1). ---- */
static inline int64_t so_agent_update(float *ag, int32_t agent_index, const float *trail,
                                      const so_params *p)
{
    const uint32_t W = p->width, H = p->height;
    float x = ag[0], y = ag[1], angle = ag[2], speed = ag[3];

    speed = so_clamp(speed, p->agent_speed_min, p->agent_speed_max);      /* :72 */

    float aL = angle - p->agent_sensor_angle;                             /* :75 */
    float aR = angle + p->agent_sensor_angle;                             /* :76 */
    float sL, cL, sR, cR, sC, cC;
    so_sincos(aL, &sL, &cL);
    so_sincos(aR, &sR, &cR);
    so_sincos(angle, &sC, &cC);
    const float sd = p->agent_sensor_distance;
    float vL = so_sample(trail, W, H, x + sd * cL, y + sd * sL);          /* :79-82,93 */
    float vR = so_sample(trail, W, H, x + sd * cR, y + sd * sR);          /* :83-86,94 */
    float vC = so_sample(trail, W, H, x + sd * cC, y + sd * sC);          /* :87-90,95 */

    if (vC > vL && vC > vR) {                                             /* :98 */
    } else if (vL > vR) {                                                 /* :100-104 */
        float target = angle - SO_TAU;
        float diff = target - angle;
        angle += fminf(p->agent_turn_speed, fabsf(diff)) * so_sign(diff);
    } else if (vR > vL) {                                                 /* :105-109 */
        float target = angle + SO_TAU;
        float diff = target - angle;
        angle += fminf(p->agent_turn_speed, fabsf(diff)) * so_sign(diff);
    }

    float rnd = so_hash01(agent_index, x, y);                             /* :117 (pre-move x,y) */
    angle += (rnd * 2.0f - 1.0f) * p->agent_jitter;                       /* :118 */

    angle = so_fmod(angle, SO_TWO_PI);                                    /* :121 */
    if (angle < 0.0f) angle = angle + SO_TWO_PI;                          /* :122 */

    float move = speed * SO_TIME_STEP;                                    /* :125 */
    float sM, cM;
    so_sincos(angle, &sM, &cM);
    x = x + move * cM;                                                    /* :126 */
    y = y + move * sM;                                                    /* :127 */

    x = so_fmod(x, (float)W);                                             /* :130 */
    if (x < 0.0f) x = x + (float)W;                                       /* :131 */
    y = so_fmod(y, (float)H);                                             /* :132 */
    if (y < 0.0f) y = y + (float)H;                                       /* :133 */

    ag[0] = x; ag[1] = y; ag[2] = angle; ag[3] = speed;                   /* :144 */

    /* :136-141 -- i32(x), i32(y) with the bounds check (x == W can occur by rounding) */
    if (x >= 0.0f && x < (float)W && y >= 0.0f && y < (float)H)
        return (int64_t)(int32_t)y * (int64_t)W + (int64_t)(int32_t)x;
    return -1;
}

/* Agent pass, phase_split: trail is read only, deposits counted in counts[].
 * ids == NULL -> agent_index = array position (the reference's linear mapping,
 * valid for N <= 4,194,240). */
void so_agents_phase_split(float *agents, const uint32_t *ids, uint64_t n, const float *trail,
                           uint32_t *counts, const so_params *p)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        int32_t idx = ids ? (int32_t)ids[i] : (int32_t)i;
        int64_t cell = so_agent_update(agents + 4 * i, idx, trail, p);
        if (cell >= 0) __atomic_fetch_add(&counts[cell], 1u, __ATOMIC_RELAXED);
    }
}

/* Agent pass, sequential in-place (single thread by definition). compute.wgsl:140 */
void so_agents_sequential(float *agents, uint64_t n, float *trail, const so_params *p)
{
    for (uint64_t i = 0; i < n; ++i) {
        int64_t cell = so_agent_update(agents + 4 * i, (int32_t)i, trail, p);
        if (cell >= 0)
            trail[cell] = so_clamp(trail[cell] + p->pheromone_deposition_amount, 0.0f, 1.0f);
    }
}

/* ---- trail passes ----------------------------------------------------------- */
/* deposit merge of phase_split: cells with k deposits become clamp(t + f32(k)*dep, 0, 1) */
static inline float so_merge(float t, uint32_t k, float dep)
{
    if (k == 0u) return t;
    return so_clamp(t + (float)k * dep, 0.0f, 1.0f);
}
void so_deposit_merge(float *trail, uint32_t *counts, uint64_t cells, float dep)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)cells; ++i) {
        trail[i] = so_merge(trail[i], counts[i], dep);
        counts[i] = 0u;
    }
}

/* decay_trail, compute.wgsl:148-161 */
static inline float so_decay1(float t, float decay_factor)
{
    float d = decay_factor * 0.001f;                                      /* :159 */
    return fmaxf(t - d, 0.0f);                                            /* :160 */
}
void so_decay(float *trail, uint64_t cells, float decay_factor)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)cells; ++i) trail[i] = so_decay1(trail[i], decay_factor);
}

/* diffuse_trail, compute.wgsl:164-195, Jacobi (all reads from `in`). */
void so_diffuse(const float *in, float *out, uint32_t W, uint32_t H, float diffusion_rate)
{
    float rate = so_clamp(diffusion_rate, 0.0f, 1.0f);                    /* :173 */
    #pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)H; ++y) {
        for (int64_t x = 0; x < (int64_t)W; ++x) {
            float sum = 0.0f;
            for (int dy = -1; dy <= 1; ++dy) {                            /* :181 */
                int64_t ny = (y + dy + H) % H;                            /* :184 */
                for (int dx = -1; dx <= 1; ++dx) {                        /* :182 */
                    int64_t nx = (x + dx + W) % W;                        /* :183 */
                    sum += in[ny * W + nx];                               /* :186 */
                }
            }
            float avg = sum / 9.0f;                                       /* :193 */
            out[y * W + x] = so_mix(in[y * W + x], avg, rate);            /* :194 */
        }
    }
}

/* diffuse_trail executed in place in raster order by one thread (what a fully
 * serialising software rasteriser would do with the reference's racy shader). */
void so_diffuse_inplace_raster(float *t, uint32_t W, uint32_t H, float diffusion_rate)
{
    float rate = so_clamp(diffusion_rate, 0.0f, 1.0f);
    for (int64_t y = 0; y < (int64_t)H; ++y)
        for (int64_t x = 0; x < (int64_t)W; ++x) {
            float sum = 0.0f;
            for (int dy = -1; dy <= 1; ++dy) {
                int64_t ny = (y + dy + H) % H;
                for (int dx = -1; dx <= 1; ++dx) {
                    int64_t nx = (x + dx + W) % W;
                    sum += t[ny * W + nx];
                }
            }
            t[y * W + x] = so_mix(t[y * W + x], sum / 9.0f, rate);
        }
}

/* The engine's fused trail pass: merge -> decay -> Jacobi 3x3 box, out of place.
 * counts may be NULL (diffusion-only, config 5).  counts are cleared. */
void so_trail_pass(const float *in, uint32_t *counts, float *out, float *scratch,
                   uint32_t W, uint32_t H, const so_params *p)
{
    uint64_t cells = (uint64_t)W * H;
    float dep = p->pheromone_deposition_amount;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)cells; ++i) {
        float t = in[i];
        if (counts) { t = so_merge(t, counts[i], dep); counts[i] = 0u; }
        scratch[i] = so_decay1(t, p->decay_factor);
    }
    so_diffuse(scratch, out, W, H, p->diffusion_rate);
}

/* ---- EXTENSION (no reference semantics: compute.wgsl never reads blur_radius /
 * blur_sigma): separable Gaussian of integer radius
 * R >= 1 and sigma, toroidal, replacing the 3x3 mean in the fused pass.
 * weights w[d] = exp(-d^2 / (2 sigma^2)) evaluated in f64, normalised in f64,
 * rounded to f32; horizontal pass then vertical pass, each accumulated in tap
 * order d = -R..R with fmaf(w, v, acc) starting from 0.  "Parity unpinned". */
void so_gauss_weights(float *w, int R, float sigma)
{
    double tmp[2 * 64 + 1], s = 0.0;
    for (int d = -R; d <= R; ++d) {
        tmp[d + R] = exp(-(double)(d * d) / (2.0 * (double)sigma * (double)sigma));
        s += tmp[d + R];
    }
    for (int d = 0; d <= 2 * R; ++d) w[d] = (float)(tmp[d] / s);
}
void so_trail_pass_gauss(const float *in, uint32_t *counts, float *out, float *scratch,
                         float *scratch2, uint32_t W, uint32_t H, const so_params *p,
                         int R, float sigma)
{
    uint64_t cells = (uint64_t)W * H;
    float w[2 * 64 + 1];
    so_gauss_weights(w, R, sigma);
    float dep = p->pheromone_deposition_amount;
    float rate = so_clamp(p->diffusion_rate, 0.0f, 1.0f);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)cells; ++i) {
        float t = in[i];
        if (counts) { t = so_merge(t, counts[i], dep); counts[i] = 0u; }
        scratch[i] = so_decay1(t, p->decay_factor);
    }
    #pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)H; ++y)
        for (int64_t x = 0; x < (int64_t)W; ++x) {
            float acc = 0.0f;
            for (int d = -R; d <= R; ++d) {
                int64_t nx = ((x + d) % (int64_t)W + W) % W;
                acc = fmaf(w[d + R], scratch[y * W + nx], acc);
            }
            scratch2[y * W + x] = acc;
        }
    #pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)H; ++y)
        for (int64_t x = 0; x < (int64_t)W; ++x) {
            float acc = 0.0f;
            for (int d = -R; d <= R; ++d) {
                int64_t ny = ((y + d) % (int64_t)H + H) % H;
                acc = fmaf(w[d + R], scratch2[ny * W + x], acc);
            }
            out[y * W + x] = so_mix(scratch[y * W + x], acc, rate);
        }
}

/* ---- step drivers, pass order of /root/reference/src/main.rs:1163-1235 ------ */
/* phase_split: trail/counts/scratch are W*H; tmp is W*H.  ids may be NULL. */
void so_step_phase_split(float *agents, const uint32_t *ids, uint64_t n, float *trail,
                         uint32_t *counts, float *tmp, float *scratch,
                         const so_params *p, int n_steps)
{
    uint64_t cells = (uint64_t)p->width * p->height;
    for (int s = 0; s < n_steps; ++s) {
        so_agents_phase_split(agents, ids, n, trail, counts, p);          /* main.rs:1164-1181 */
        so_trail_pass(trail, counts, tmp, scratch, p->width, p->height, p); /* :1184-1199,1220-1235 */
        memcpy(trail, tmp, cells * sizeof(float));
    }
}

/* sequential: agents in place, decay in place, then Jacobi diffusion (or raster in-place). */
void so_step_sequential(float *agents, uint64_t n, float *trail, float *tmp,
                        const so_params *p, int n_steps, int inplace_diffuse)
{
    uint64_t cells = (uint64_t)p->width * p->height;
    for (int s = 0; s < n_steps; ++s) {
        so_agents_sequential(agents, n, trail, p);
        so_decay(trail, cells, p->decay_factor);
        if (inplace_diffuse) {
            so_diffuse_inplace_raster(trail, p->width, p->height, p->diffusion_rate);
        } else {
            so_diffuse(trail, tmp, p->width, p->height, p->diffusion_rate);
            memcpy(trail, tmp, cells * sizeof(float));
        }
    }
}

/* The reference's dispatch/index quirk (main.rs:1175-1179 vs compute.wgsl:60):
 * marks, for n agents, how many threads of the reference dispatch map to each
 * index.  Documentation only; the engine uses the linear map. */
void so_reference_dispatch_hits(uint8_t *hits, uint64_t n)
{
    uint64_t wg = (n + 63) / 64;
    uint64_t gx = wg < 65535 ? wg : 65535;
    uint64_t gy = gx ? (wg + gx - 1) / gx : 0;
    for (uint64_t i = 0; i < n; ++i) hits[i] = 0;
    for (uint64_t yy = 0; yy < gy; ++yy)
        for (uint64_t xx = 0; xx < gx * 64; ++xx) {
            uint64_t idx = xx + yy * 65535ull;
            if (idx < n && hits[idx] < 255) hits[idx]++;
        }
}

/* field statistics helper for statistical comparisons */
void so_field_stats(const float *t, uint64_t cells, double *sum, double *sumsq, float *maxv,
                    uint64_t *nonzero)
{
    double s = 0.0, s2 = 0.0; float m = -INFINITY; uint64_t nz = 0;
    for (uint64_t i = 0; i < cells; ++i) {
        s += t[i]; s2 += (double)t[i] * t[i];
        if (t[i] > m) m = t[i];
        if (t[i] != 0.0f) nz++;
    }
    *sum = s; *sumsq = s2; *maxv = m; *nonzero = nz;
}

/* ---------------------------------------------------------------------------
 * Display / colourise pass -- /root/reference/src/display.wgsl:29-86.
 * trail (W x H, f32) -> RGBA8 texture (tw x th), letter-boxed, through a 768-entry LUT laid out
 * as 256 R, 256 G, 256 B (lut_manager.rs:176-178; widened to u32 by main.rs:330-337).
 * Every f32 operation is the literal one of the shader; the rgba8unorm store is the WGSL
 * conversion round(clamp(v, 0, 1) * 255).
 * ------------------------------------------------------------------------- */
static inline uint8_t so_unorm8(float v)
{
    float c = fminf(fmaxf(v, 0.0f), 1.0f);
    return (uint8_t)lrintf(c * 255.0f);
}

void so_display(const float *trail, uint32_t W, uint32_t H, const uint8_t *lut768,
                uint8_t *rgba, uint32_t tw, uint32_t th)
{
    const float sim_w = (float)W, sim_h = (float)H;               /* display.wgsl:48-49 */
    const float tex_w = (float)tw, tex_h = (float)th;             /* :50-51 */
    const float sim_aspect = sim_w / sim_h;                       /* :54 */
    const float tex_aspect = tex_w / tex_h;                       /* :55 */
    float scale, off_x = 0.0f, off_y = 0.0f;                      /* :58-60 */
    if (tex_aspect > sim_aspect) {                                /* :61-64 fit height */
        scale = tex_h / sim_h;
        off_x = (tex_w - sim_w * scale) * 0.5f;
    } else {                                                      /* :65-69 fit width */
        scale = tex_w / sim_w;
        off_y = (tex_h - sim_h * scale) * 0.5f;
    }
#pragma omp parallel for schedule(static)
    for (int64_t py = 0; py < (int64_t)th; ++py) {
        for (uint32_t px = 0; px < tw; ++px) {
            const float fx = ((float)px - off_x) / scale;         /* :72 */
            const float fy = ((float)py - off_y) / scale;         /* :73 */
            uint8_t *o = rgba + 4 * ((size_t)py * tw + px);
            if (fx >= 0.0f && fx < sim_w && fy >= 0.0f && fy < sim_h) {   /* :76 */
                const int32_t x = (int32_t)fx, y = (int32_t)fy;   /* :77-78 */
                const int64_t idx = (int64_t)y * (int64_t)W + x;  /* :79 */
                float t = trail[idx];
                float inten = fminf(fmaxf(t, 0.0f), 1.0f);        /* :80 (NaN -> 0, the spec's clamp) */
                inten = fminf(fmaxf(inten, 0.0f), 1.0f);          /* :31 get_lut_color clamps again */
                const uint32_t li = (uint32_t)(inten * 255.0f);   /* :34 */
                o[0] = so_unorm8((float)lut768[li] / 255.0f);         /* :37 */
                o[1] = so_unorm8((float)lut768[li + 256] / 255.0f);   /* :38 */
                o[2] = so_unorm8((float)lut768[li + 512] / 255.0f);   /* :39 */
                o[3] = so_unorm8(1.0f);
            } else {                                              /* :83-85 black bars */
                o[0] = 0; o[1] = 0; o[2] = 0; o[3] = so_unorm8(1.0f);
            }
        }
    }
}
