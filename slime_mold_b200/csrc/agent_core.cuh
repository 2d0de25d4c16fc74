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
    float neg_zero;            // -0.0f, opaque to the compiler (see mul2_nofuse)
    uint32_t zero_bits;        // 0u, opaque to the compiler (see the gather fence in agent_update_impl)
    // strip of the trail this kernel may read (multi-GPU): global rows
    // [row0 - halo, row0 + rows + halo) live at trail + (row - row_base) * W
    int64_t row_base;
    int32_t rows_local;        // rows this rank owns (== H on a single GPU)
    int32_t ghost;             // ghost rows kept above and below the strip (0 on a single GPU)
    int32_t fold_hi, fold_lo;  // seam folding thresholds of local_row()
    int32_t flag_wrap;         // tiled deposit flags (kernels.cuh flag_tile_offset): H on one GPU (row -1 is row H-1), 0 on strips (ghost rows)
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

// sample_trail_map, compute.wgsl:7-29.  FETCH(consts, fx, fy, v00, v10, v01, v11) returns the 2x2 footprint
// whose top-left cell is global (fx, fy) (integral floats): four scalar loads from the row-major field (host / LDG
// path) or one texture gather from the block-linear copy (device TEX path) -- raw f32 either way.  An
// implementation may fetch whether or not the tap passes the reference's bounds test (the texture path
// does: clamped addressing makes any coordinate harmless and an unconditional TLD4 needs no predicate
// bookkeeping) or test and skip the loads (the row-major path must) -- the footprint of an outside tap
// is discarded by zero_unless_inside().
//
// The sample is split in two so that the per-sensor position arithmetic (px - floor, 1 - d) of the left
// and right sensors can run as packed pairs, and so that all three footprints can be in flight at once:
// fetch_footprint() issues the loads, bilinear() takes the footprint and the fractions.
SM_HD bool tap_inside(const AgentConsts& c, float fx, float fy)
{
    // x0 < 0 || x1 >= W || y0 < 0 || y1 >= H -> 0 (sensing is NOT toroidal); NaN -> outside
    return fx >= 0.0f && fx <= c.xmax && fy >= 0.0f && fy <= c.ymax;
}

// inside ? v : 0 -- on the device as ONE predicate chain (4 FSETP + 1 FSEL).  Left to the compiler the
// conjunction becomes four nested selects (8 instructions per sensor); the kernel is issue-bound.
SM_HD float zero_unless_inside(const AgentConsts& c, float fx, float fy, float v)
{
#ifdef __CUDA_ARCH__
    float r;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ge.f32 p, %1, 0f00000000;\n\t"
        "setp.le.and.f32 p, %1, %2, p;\n\t"
        "setp.ge.and.f32 p, %3, 0f00000000, p;\n\t"
        "setp.le.and.f32 p, %3, %4, p;\n\t"
        "selp.f32 %0, %5, 0f00000000, p;\n\t}"
        : "=f"(r) : "f"(fx), "f"(c.xmax), "f"(fy), "f"(c.ymax), "f"(v));
    return r;
#else
    return tap_inside(c, fx, fy) ? v : 0.0f;
#endif
}

// A 2x2 footprint, corner (fx, fy): v[] = {v00, v10, v01, v11}
struct Footprint { float fx, fy, v00, v10, v01, v11; };

template <class FETCH>
SM_HD Footprint fetch_footprint(const AgentConsts& c, float fx, float fy, FETCH fetch)
{
    Footprint f;
    f.fx = fx; f.fy = fy;
    fetch(c, fx, fy, f.v00, f.v10, f.v01, f.v11);   // fx, fy: integral; inside [0, W-2] x [0, H-2] when the tap is inside
    return f;
}

SM_HD float bilinear(const AgentConsts& c, const Footprint& f, float dx, float omdx, float dy, float omdy)
{
    float v0 = mixf_pre(f.v00, f.v10, dx, omdx);                           // :26
    float v1 = mixf_pre(f.v01, f.v11, dx, omdx);                           // :27
    float v = mixf_pre(v0, v1, dy, omdy);                                  // :28
    return zero_unless_inside(c, f.fx, f.fy, v);                           // :14-16
}

template <class FETCH>
SM_HD float sample_trail(const AgentConsts& c, float px, float py, FETCH fetch)
{
    float fx = ::floorf(px), fy = ::floorf(py);
    const Footprint f = fetch_footprint(c, fx, fy, fetch);
    float dx = sub(px, fx), dy = sub(py, fy);
    return bilinear(c, f, dx, sub(1.0f, dx), dy, sub(1.0f, dy));
}

// Footprint fetch from the row-major field.  IdxT = int32_t when the strip (with ghosts) has fewer
// than 2^31 cells, else int64_t.  LD(ptr) loads one f32.
template <class IdxT, class LD>
struct FetchLinear {
    const float* trail;      // owned row 0 of this rank's strip
    IdxT W, row_base;
    LD ld;
    SM_HD void operator()(const AgentConsts& c, float fx, float fy, float& v00, float& v10, float& v01, float& v11) const
    {
        v00 = v10 = v01 = v11 = 0.0f;
        if (tap_inside(c, fx, fy)) {
            const IdxT x0 = (IdxT)(int32_t)fx, y0 = (IdxT)(int32_t)fy;
            const float* r0 = trail + ((y0 - row_base) * W + x0);
            const float* r1 = r0 + W;
            v00 = ld(r0); v10 = ld(r0 + 1); v01 = ld(r1); v11 = ld(r1 + 1);
        }
    }
};

// x % m with the reference's follow-up `if (x < 0) x += m` (compute.wgsl:130-133)
SM_HD float wrap_coord(float v, float m, float rcp_m)
{
    v = fmod_exact(v, m, rcp_m);
    if (v < 0.0f) v = add(v, m);
    return v;
}

// Returns the deposit cell as (cx, cy) with cx < 0 when the deposit is skipped
// (compute.wgsl:138: x == W can occur by rounding).
//
// Statement order and every rounding follow compute.wgsl:65-144; what is specific to this engine is
// only how the independent operations are grouped for issue: the left / right sensor headings,
// positions and fractions run as packed pairs (sincos_small2, add2 / mul2), steering is one
// branch-free expression, and the toroidal wrap is skipped when the moved position is already in
// range (x % W == x for 0 <= x < W, -0 included).
//
// TAME = |angle| <= 4096 and |sensor_angle| <= 4096 (every agent after its first step: headings are wrapped
// into [0, 2pi]).  It licenses two shortcuts that are identities on that range: the spec's small-argument
// sincos without its range test, and sign(diff) == sign(tau) in the turn (diff = (angle +- TAU) - angle is
// within 2^-11 of +-TAU).  The generic instantiation keeps the literal statements.
template <bool TAME, class FETCH>
SM_HD void agent_update_impl(float& x, float& y, float& angle, float& speed, int32_t agent_index,
                             const AgentConsts& c, FETCH fetch, int32_t& cx, int32_t& cy)
{
    speed = clampf(speed, c.speed_min, c.speed_max);                       // :72

    f2 sLR, cLR;           // lo = left sensor, hi = right sensor
    float sC, cC;
    const f2 aLR = mk2(sub(angle, c.sensor_angle), add(angle, c.sensor_angle));   // :75-76
    if (TAME) {
        // |angle +- sa| <= 8192: the spec's fast path, evaluated without the per-call range test
        sincos_small2(aLR, sLR, cLR);
        sincos_small(angle, sC, cC);                                       // :77
    } else {
        sincos(aLR.lo, sLR.lo, cLR.lo);
        sincos(aLR.hi, sLR.hi, cLR.hi);
        sincos(angle, sC, cC);
    }
    const float sd = c.sensor_distance;
    const f2 sd2 = splat2(sd), one2 = splat2(1.0f);
    const f2 pxLR = add2(splat2(x), mul2_nofuse(sd2, cLR, c.neg_zero));                       // :79-86
    const f2 pyLR = add2(splat2(y), mul2_nofuse(sd2, sLR, c.neg_zero));
    const f2 fxLR = mk2(::floorf(pxLR.lo), ::floorf(pxLR.hi));             // :8-9
    const f2 fyLR = mk2(::floorf(pyLR.lo), ::floorf(pyLR.hi));
    const float pxC = add(x, mul(sd, cC)), pyC = add(y, mul(sd, sC));      // :87-90
    const float fxC = ::floorf(pxC), fyC = ::floorf(pyC);
    // all three footprints are requested before any of them is consumed: one exposed memory latency, not three
    Footprint qL = fetch_footprint(c, fxLR.lo, fyLR.lo, fetch);            // :93
    Footprint qR = fetch_footprint(c, fxLR.hi, fyLR.hi, fetch);            // :94
    Footprint qC = fetch_footprint(c, fxC, fyC, fetch);                    // :95
#ifdef __CUDA_ARCH__
    // Scheduling fence, bit-neutral: OR-ing (last gather's word & 0) into one word of the first two footprints makes their
    // consumers depend on the LAST gather, so ptxas cannot consume the first footprint before the other two gathers are
    // issued.  Short of registers it does exactly that in the strip and unrolled instantiations (seen in SASS and in
    // ncu: two or three exposed gather latencies per agent instead of one).  zero_bits is 0 from the parameter block.
    qL.v00 = u2f(f2u(qL.v00) | (f2u(qC.v11) & c.zero_bits));
    qR.v00 = u2f(f2u(qR.v00) | (f2u(qC.v11) & c.zero_bits));
#endif
    const f2 dxLR = sub2(pxLR, fxLR), dyLR = sub2(pyLR, fyLR);             // :18-19
    const f2 mxLR = sub2(one2, dxLR), myLR = sub2(one2, dyLR);             // the (1 - t) of mix()
    const float dxC = sub(pxC, fxC), dyC = sub(pyC, fyC);
    float vL = bilinear(c, qL, dxLR.lo, mxLR.lo, dyLR.lo, myLR.lo);
    float vR = bilinear(c, qR, dxLR.hi, mxLR.hi, dyLR.hi, myLR.hi);
    float vC = bilinear(c, qC, dxC, sub(1.0f, dxC), dyC, sub(1.0f, dyC));

    // :98-112  keep | turn left (towards angle - TAU) | turn right (towards angle + TAU) | keep (vL == vR).
    // angle - TAU == angle + (-TAU) bit for bit, so both turns are one expression in the signed TAU.
    {
        const bool keep = vC > vL && vC > vR;
        const bool left = vL > vR, right = vR > vL;
        const float tau = left ? -kTau : kTau;
        const float diff = sub(add(angle, tau), angle);                    // :101,106
        const float reach = ::fminf(c.turn_speed, ::fabsf(diff));
        // TAME: sign(diff) is -1 for a left turn, +1 for a right turn, and reach * (+-1) == +-reach exactly
        const float turn = TAME ? (left ? -reach : reach) : mul(reach, signf(diff));
        const float turned = add(angle, turn);                             // :104,109
        angle = (!keep && (left || right)) ? turned : angle;
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
    const f2 moved = add2(mk2(x, y), mul2_nofuse(splat2(move), mk2(cM, sM), c.neg_zero));     // :126-127
    x = moved.lo;
    y = moved.hi;

    // :130-133 wrap, :136-138 deposit bounds -- the same four comparisons decide both: a position already
    // inside [0, W) x [0, H) is left untouched by `%` and by the negative fix-up
    bool inside = x >= 0.0f && x < c.Wf && y >= 0.0f && y < c.Hf;
    if (!inside) {
        x = wrap_coord(x, c.Wf, c.rcpW);                                   // :130-131
        y = wrap_coord(y, c.Hf, c.rcpH);                                   // :132-133
        inside = x >= 0.0f && x < c.Wf && y >= 0.0f && y < c.Hf;           // :136-138
    }
    if (inside) {
        cx = (int32_t)x; cy = (int32_t)y;
    } else {
        cx = -1; cy = -1;
    }
}

template <class FETCH>
SM_HD void agent_update(float& x, float& y, float& angle, float& speed, int32_t agent_index,
                        const AgentConsts& c, FETCH fetch, int32_t& cx, int32_t& cy)
{
    if (::fabsf(angle) <= 4096.0f && ::fabsf(c.sensor_angle) <= 4096.0f)
        agent_update_impl<true>(x, y, angle, speed, agent_index, c, fetch, cx, cy);
    else
        agent_update_impl<false>(x, y, angle, speed, agent_index, c, fetch, cx, cy);
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
