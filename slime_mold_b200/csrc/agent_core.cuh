// agent_core.cuh -- per-agent update of /root/reference/src/compute.wgsl:65-144
// (sense -> rotate -> jitter -> move -> wrap -> deposit cell), written once as a
// __host__ __device__ function so the CUDA kernels and the CPU-side hostcheck
// test (tests/hostcheck, test-only) share the exact statement sequence.
#pragma once
#include "device_math.cuh"

namespace smd {

// Uniform, per-launch constants derived on the host from sm_params
// (SimSizeUniform, /root/reference/src/main.rs:29-46).
struct AgentConsts {
    uint32_t W, H;             // global map size
    float Wf, Hf;              // f32(W), f32(H)
    float rcpW, rcpH;          // ~1/W, ~1/H (quotient estimates for fmod_exact)
    float xmax, ymax;          // f32(W) - 2, f32(H) - 2: last valid x0 / y0 of a bilinear tap
    float speed_min, speed_max;
    float turn_speed;
    float sensor_angle, sensor_distance;
    float jitter;
    // strip of the trail this kernel may read (multi-GPU): global rows
    // [row0 - halo, row0 + rows + halo) live at trail + (row - row_base) * W
    int64_t row_base;
    int32_t rows_local;        // rows this rank owns (== H on a single GPU)
    int32_t ghost;             // ghost rows kept above and below the strip (0 on a single GPU)
    int32_t fold_hi, fold_lo;  // seam folding thresholds of local_row()
};

// Local row (relative to the strip's first owned row) of global row `gy`, folded across the
// toroidal seam so that rows just above strip 0 / just below the last strip land in the ghosts.
// Fold to the representative nearest to the strip: with two strips the ghosts can cover the whole
// other strip, so the ghost depth itself cannot be the folding threshold.  fold_hi / fold_lo are
// rows_local + ceil(spare / 2) and -floor(spare / 2) with spare = H - rows_local (host-computed).
SM_HD int32_t local_row(int32_t gy, const AgentConsts& c)
{
    int32_t lr = gy - (int32_t)c.row_base;
    if (lr >= c.fold_hi) lr -= (int32_t)c.H;
    else if (lr < c.fold_lo) lr += (int32_t)c.H;
    return lr;
}

constexpr float kTau = 6.28318530718f;             // compute.wgsl:4
constexpr float kTwoPi = 2.0f * 3.14159265359f;    // compute.wgsl:121
constexpr float kRcpTwoPi = 0.15915494f;
constexpr float kTimeStep = 0.016f;                // compute.wgsl:55

// sample_trail_map, compute.wgsl:7-29.  FETCH(fx, fy, v00, v10, v01, v11) returns the 2x2 footprint
// whose top-left cell is global (fx, fy) (integral floats): four scalar loads from the row-major field (host / LDG
// path) or one texture gather from the block-linear copy (device TEX path) -- raw f32 either way.
template <class FETCH>
SM_HD float sample_trail(const AgentConsts& c, float px, float py, FETCH fetch)
{
    float fx = ::floorf(px), fy = ::floorf(py);
    // x0 < 0 || x1 >= W || y0 < 0 || y1 >= H -> 0 (sensing is NOT toroidal); NaN -> outside
    if (!(fx >= 0.0f && fx <= c.xmax && fy >= 0.0f && fy <= c.ymax)) return 0.0f;
    float dx = sub(px, fx), dy = sub(py, fy);
    float v00, v10, v01, v11;
    fetch(fx, fy, v00, v10, v01, v11);        // fx, fy: integral, inside [0, W-2] x [0, H-2]
    float omdx = sub(1.0f, dx);
    float v0 = mixf_pre(v00, v10, dx, omdx);
    float v1 = mixf_pre(v01, v11, dx, omdx);
    return mixf(v0, v1, dy);
}

// Footprint fetch from the row-major field.  IdxT = int32_t when the strip (with ghosts) has fewer
// than 2^31 cells, else int64_t.  LD(ptr) loads one f32.
template <class IdxT, class LD>
struct FetchLinear {
    const float* trail;      // owned row 0 of this rank's strip
    IdxT W, row_base;
    LD ld;
    SM_HD void operator()(float fx, float fy, float& v00, float& v10, float& v01, float& v11) const
    {
        const IdxT x0 = (IdxT)(int32_t)fx, y0 = (IdxT)(int32_t)fy;
        const float* r0 = trail + ((y0 - row_base) * W + x0);
        const float* r1 = r0 + W;
        v00 = ld(r0); v10 = ld(r0 + 1); v01 = ld(r1); v11 = ld(r1 + 1);
    }
};

// Returns the deposit cell as (cx, cy) with cx < 0 when the deposit is skipped
// (compute.wgsl:138: x == W can occur by rounding).
template <class FETCH>
SM_HD void agent_update(float& x, float& y, float& angle, float& speed, int32_t agent_index,
                        const AgentConsts& c, FETCH fetch, int32_t& cx, int32_t& cy)
{
    speed = clampf(speed, c.speed_min, c.speed_max);                       // :72

    float sL, cL, sR, cR, sC, cC;
    const float aL = sub(angle, c.sensor_angle);                           // :75
    const float aR = add(angle, c.sensor_angle);                           // :76
    if (::fabsf(angle) <= 4096.0f && ::fabsf(c.sensor_angle) <= 4096.0f) {
        // |angle +- sa| <= 8192: the spec's fast path, evaluated without the per-call range test
        sincos_small(aL, sL, cL);
        sincos_small(aR, sR, cR);
        sincos_small(angle, sC, cC);                                       // :77
    } else {
        sincos(aL, sL, cL);
        sincos(aR, sR, cR);
        sincos(angle, sC, cC);
    }
    const float sd = c.sensor_distance;
    float vL = sample_trail(c, add(x, mul(sd, cL)), add(y, mul(sd, sL)), fetch);   // :79-82,93
    float vR = sample_trail(c, add(x, mul(sd, cR)), add(y, mul(sd, sR)), fetch);   // :83-86,94
    float vC = sample_trail(c, add(x, mul(sd, cC)), add(y, mul(sd, sC)), fetch);   // :87-90,95

    if (vC > vL && vC > vR) {                                              // :98
    } else if (vL > vR) {                                                  // :100-104
        float diff = sub(sub(angle, kTau), angle);
        angle = add(angle, mul(::fminf(c.turn_speed, ::fabsf(diff)), signf(diff)));
    } else if (vR > vL) {                                                  // :105-109
        float diff = sub(add(angle, kTau), angle);
        angle = add(angle, mul(::fminf(c.turn_speed, ::fabsf(diff)), signf(diff)));
    }

    // :115-118  angle += (hash*2 - 1) * jitter.  With jitter == +-0 the addend is +-0 (or NaN when x / y
    // are not finite), which changes `angle` only if angle is a zero: in every other case the hash cannot
    // influence the result and is skipped.  Bit-exact, not an approximation.
    const bool hash_is_dead = (c.jitter == 0.0f) && (angle != 0.0f) && (::fabsf(x) <= 1.0e30f) && (::fabsf(y) <= 1.0e30f);
    if (!hash_is_dead) {
        float rnd = hash01(agent_index, x, y);                             // :117 (pre-move x, y)
        angle = add(angle, mul(sub(mul(rnd, 2.0f), 1.0f), c.jitter));      // :118
    }

    angle = fmod_exact(angle, kTwoPi, kRcpTwoPi);                          // :121
    if (angle < 0.0f) angle = add(angle, kTwoPi);                          // :122

    float move = mul(speed, kTimeStep);                                    // :125
    float sM, cM;
    sincos_small(angle, sM, cM);                                           // angle in [0, 2pi] or NaN here
    x = add(x, mul(move, cM));                                             // :126
    y = add(y, mul(move, sM));                                             // :127

    x = fmod_exact(x, c.Wf, c.rcpW);                                       // :130
    if (x < 0.0f) x = add(x, c.Wf);                                        // :131
    y = fmod_exact(y, c.Hf, c.rcpH);                                       // :132
    if (y < 0.0f) y = add(y, c.Hf);                                        // :133

    if (x >= 0.0f && x < c.Wf && y >= 0.0f && y < c.Hf) {                  // :136-138
        cx = (int32_t)x; cy = (int32_t)y;
    } else {
        cx = -1; cy = -1;
    }
}

// seeded start-up fill, /root/reference/src/main.rs:269-282
SM_HD void agent_init(uint64_t seed, uint64_t id, float Wf, float Hf, float speed_min, float speed_max,
                      float& x, float& y, float& angle, float& speed)
{
    const float kPi = 3.14159274101257324f;   // std::f32::consts::PI
    x = mul(rand01(seed, id, 0), Wf);
    y = mul(rand01(seed, id, 1), Hf);
    angle = mul(mul(rand01(seed, id, 2), 2.0f), kPi);
    speed = add(speed_min, mul(rand01(seed, id, 3), sub(speed_max, speed_min)));
}

}  // namespace smd
