// trail_core.cuh -- per-cell arithmetic of the fused trail pass:
//   deposit merge (phase_split form of compute.wgsl:140)
//   decay_trail   (compute.wgsl:148-161)
//   diffuse_trail (compute.wgsl:164-195, Jacobi)
// __host__ __device__ so tests/hostcheck can exercise the same statements on a CPU.
#pragma once
#include "device_math.cuh"

namespace smd {

struct TrailConsts {
    float dep;          // pheromone_deposition_amount
    float decay_sub;    // fl32(decay_factor * 0.001f)              compute.wgsl:159
    float rate;         // clamp(diffusion_rate, 0, 1)              compute.wgsl:173
    float one_minus_rate;
};

// k agents deposited on a cell holding t:  clamp(t + f32(k)*dep, 0, 1); untouched when k == 0.
SM_HD float merge_deposit(float t, uint32_t k, float dep)
{
    float m = clampf(add(t, mul((float)k, dep)), 0.0f, 1.0f);
    return k ? m : t;
}

// compute.wgsl:160   max(t - decay_factor*0.001, 0)
SM_HD float decay_cell(float t, float decay_sub) { return ::fmaxf(sub(t, decay_sub), 0.0f); }

// compute.wgsl:176-194: sum in the order dy = -1..1 (outer), dx = -1..1 (inner),
// starting from 0.0; avg = sum / 9.0; mix(centre, avg, rate).
SM_HD float box9_mix(float a0, float a1, float a2,    // row y-1: x-1, x, x+1
                     float b0, float b1, float b2,    // row y
                     float c0, float c1, float c2,    // row y+1
                     const TrailConsts& tc)
{
    float s = add(0.0f, a0);
    s = add(s, a1); s = add(s, a2);
    s = add(s, b0); s = add(s, b1); s = add(s, b2);
    s = add(s, c0); s = add(s, c1); s = add(s, c2);
    float avg = div9(s);
    return mixf_pre(b1, avg, tc.rate, tc.one_minus_rate);
}

}  // namespace smd
