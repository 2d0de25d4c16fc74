/*
 * sm_oracle_math.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Bit-defined f32 arithmetic for the CPU restatement of the reference shader
 * /root/reference/src/compute.wgsl.  WGSL leaves sin/cos/mix/fract/% only
 * loosely specified (sin/cos: 2^-11 absolute error on [-pi,pi], nothing outside),
 * and the reference has no golden vectors, so this header *defines* the
 * arithmetic the CUDA engine must reproduce bit for bit (DESIGN.md "Arithmetic
 * spec").  WGSL leaves the accuracy of sin / cos to the backend, so no reference-produced vector can pin these bits; everything
 * around them is pinned to the shader source (tests/test_wgsl_reference.py).
 *
 * Rules: IEEE-754 binary32/binary64, round-to-nearest-even, no contraction
 * (build with -ffp-contract=off), explicit fmaf()/fma() only where written.
 *
 * The CUDA side has its OWN implementation of the same spec
 * (slime_mold_b200/csrc/device_math.cuh); tests compare the two.
 */
#ifndef SM_ORACLE_MATH_H
#define SM_ORACLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

/* compute.wgsl:4   const TAU: f32 = 6.28318530718  -> 0x1.921fb6p+2 */
#define SO_TAU 6.28318530718f
/* compute.wgsl:121 (2.0 * 3.14159265359) evaluated in f32 -> same value as TAU */
#define SO_TWO_PI (2.0f * 3.14159265359f)
/* compute.wgsl:55 */
#define SO_TIME_STEP 0.016f

static inline uint32_t so_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float so_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint64_t so_d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

/* ---- SPEC-SINCOS ---------------------------------------------------------
 * Quadrant reduction x = k*(pi/2) + r, |r| <= pi/4 (+rounding), then degree-7
 * sine / degree-8 cosine polynomials (Cephes single-precision coefficients),
 * max error < 1 ulp on the reduced range (tests/test_oracle_math.py).
 *   |x| <= 8192      : 3-constant Cody-Waite in f32 with fmaf
 *   otherwise        : 2-constant reduction in f64 with fma (hash arguments of
 *                      compute.wgsl:117 reach 1e7..1e10); |x| >= 2^50 is first
 *                      folded by the exact IEEE fmod(x, fl64(2*pi)).
 */
static inline void so_sincos(float x, float *sn, float *cs)
{
    float r;
    uint32_t q;
    if (fabsf(x) <= 8192.0f) {
        const float MAGIC = 12582912.0f;                 /* 1.5 * 2^23 */
        float t = fmaf(x, 0x1.45f306p-1f, MAGIC);        /* fl32(2/pi) */
        q = so_f2u(t) & 3u;
        float k = t - MAGIC;
        r = fmaf(-k, 0x1.921fb6p+0f, x);                 /* fl32(pi/2) */
        r = fmaf(-k, -0x1.777a5cp-25f, r);               /* fl32(pi/2 - P1) */
        r = fmaf(-k, -0x1.ee59dap-50f, r);               /* fl32(pi/2 - P1 - P2) */
    } else {
        const double MAGIC_D = 6755399441055744.0;       /* 1.5 * 2^52 */
        double xd = (double)x;
        if (!(fabs(xd) < 0x1p50)) xd = fmod(xd, 0x1.921fb54442d18p+2); /* NaN/inf -> NaN */
        double td = fma(xd, 0x1.45f306dc9c883p-1, MAGIC_D);
        q = (uint32_t)(so_d2u(td) & 3u);
        double kd = td - MAGIC_D;
        double rd = fma(-kd, 0x1.921fb54442d18p+0, xd);
        rd = fma(-kd, 0x1.1a62633145c07p-54, rd);
        r = (float)rd;
    }
    float s2 = r * r;
    float p = fmaf(-1.9515295891e-4f, s2, 8.3321608736e-3f);
    p = fmaf(p, s2, -1.6666654611e-1f);
    p = p * s2;
    float sinr = fmaf(p, r, r);
    float c = fmaf(2.443315711809948e-5f, s2, -1.388731625493765e-3f);
    c = fmaf(c, s2, 4.166664568298827e-2f);
    c = fmaf(c, s2, -0.5f);
    float cosr = fmaf(c, s2, 1.0f);
    float s_out = (q & 1u) ? cosr : sinr;
    float c_out = (q & 1u) ? sinr : cosr;
    if (q & 2u) s_out = -s_out;
    if ((q + 1u) & 2u) c_out = -c_out;
    *sn = s_out;
    *cs = c_out;
}

/* WGSL float `%`: truncated remainder, sign of the dividend == IEEE fmodf (exact). */
static inline float so_fmod(float a, float b) { return fmodf(a, b); }

/* WGSL clamp(x, lo, hi) = min(max(x, lo), hi); NaN handling = fmaxf/fminf. */
static inline float so_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* WGSL mix(a, b, t) = a*(1-t) + b*t, each operation rounded separately. */
static inline float so_mix(float a, float b, float t)
{
    float one_minus_t = 1.0f - t;
    float lhs = a * one_minus_t;
    float rhs = b * t;
    return lhs + rhs;
}

/* WGSL fract(v) = v - floor(v). */
static inline float so_fract(float v) { return v - floorf(v); }

/* WGSL sign(): 1, -1, 0 (NaN -> 0). */
static inline float so_sign(float v) { return (float)((v > 0.0f) - (v < 0.0f)); }

/* ---- SPEC-RNG (seeded initial state; the reference uses an unseeded
 * rand::random::<f32>(), src/main.rs:273-280 -- we keep its distribution and
 * its 24-bit mantissa convention and make it counter based). */
static inline uint64_t so_mix64(uint64_t z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
static inline float so_rand01(uint64_t seed, uint64_t id, uint32_t stream)
{
    uint64_t z = (id * 4ull + (uint64_t)stream) * 0x9E3779B97F4A7C15ull
               + seed * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
    z = so_mix64(z);
    return (float)(uint32_t)(z >> 40) * 0x1p-24f;        /* [0,1), 24 bits */
}

#endif
