// device_math.cuh -- the engine's implementation of the arithmetic spec
// (DESIGN.md "Arithmetic spec"): bit-defined f32 sin/cos, exact fmod, exact x/9,
// WGSL mix/clamp/fract/sign as used by /root/reference/src/compute.wgsl.
//
// Written independently of oracle/sm_oracle_math.h (the CPU checker); the two are
// compared bit for bit by tests/test_math_parity.py on the GPU and by
// tests/test_hostcheck.py on the CPU (these functions are __host__ __device__ so
// the same source can be exercised without a GPU -- the host instantiation is
// test-only and never linked into the product library's compute path).
//
// Device build flags that this file relies on: --fmad=false (no contraction),
// default -prec-div=true -prec-sqrt=true -ftz=false.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define SM_HD __host__ __device__ __forceinline__
#else
#define SM_HD inline
#endif

namespace smd {

// ---- rounding-explicit primitives (never contracted) -----------------------
SM_HD float mul(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;   // volatile: forbid host-side contraction regardless of flags
    return r;
#endif
}
SM_HD float add(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
SM_HD float sub(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    volatile float r = a - b;
    return r;
#endif
}
SM_HD float fma(float a, float b, float c)
{
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return ::fmaf(a, b, c);
#endif
}
SM_HD uint32_t f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; std::memcpy(&u, &f, 4); return u;
#endif
}
SM_HD float u2f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; std::memcpy(&f, &u, 4); return f;
#endif
}

// ---- packed pairs: two independent f32 lanes per instruction ------------------
// sm_100a has FADD2 / FMUL2 / FFMA2 (PTX add/mul/fma.rn.f32x2): one issue slot, two IEEE binary32
// results, each lane rounded exactly like the scalar instruction (measured on B200,
// tools/microbench/ffma2.cu: 126.6 lane-ops/clk/SM packed vs 98.4 scalar, and the packed form leaves
// the issue slot of the second lane free for other pipes).  The agent kernel is issue-bound, so every
// pair of identical scalar operations on independent values is written as one packed operation.
// Lane semantics are identical to add()/mul()/fma() above -- the host form below IS those functions.
struct alignas(8) f2 { float lo, hi; };
SM_HD f2 mk2(float lo, float hi) { f2 v; v.lo = lo; v.hi = hi; return v; }
SM_HD f2 splat2(float s) { return mk2(s, s); }
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long pk2(f2 v)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.lo), "f"(v.hi));
    return r;
}
__device__ __forceinline__ f2 upk2(unsigned long long r)
{
    f2 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.lo), "=f"(v.hi) : "l"(r));
    return v;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(d);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
    return upk2(d);
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
    return upk2(d);
}
#else
SM_HD f2 add2(f2 a, f2 b) { return mk2(add(a.lo, b.lo), add(a.hi, b.hi)); }
SM_HD f2 mul2(f2 a, f2 b) { return mk2(mul(a.lo, b.lo), mul(a.hi, b.hi)); }
SM_HD f2 fma2(f2 a, f2 b, f2 c) { return mk2(fma(a.lo, b.lo, c.lo), fma(a.hi, b.hi, c.hi)); }
#endif
// Product that must stay a product.  ptxas 12.9 contracts mul.rn.f32x2 followed by add.rn.f32x2 into
// one FFMA2 -- a single rounding -- even under --fmad=false (it honours the flag and the .rn modifiers
// only for the scalar forms).  Wherever a packed product feeds a packed sum the product is therefore
// written as an FFMA2 whose addend is a -0.0 the compiler cannot see (`neg_zero` comes from the kernel's
// parameter block): a*b + (-0.0) rounds once, to RN(a*b), sign of zero included, and an FMA cannot be
// contracted with the following add.  The bit-parity tests against the oracle guard this.
SM_HD f2 mul2_nofuse(f2 a, f2 b, float neg_zero) { return fma2(a, b, splat2(neg_zero)); }
SM_HD f2 neg2(f2 a) { return mk2(-a.lo, -a.hi); }
SM_HD f2 sub2(f2 a, f2 b) { return add2(a, neg2(b)); }      // a - b == a + (-b) bit for bit

// ---- SPEC-SINCOS -------------------------------------------------------------
// x = k*(pi/2) + r by the round-to-integer magic constant; q = k mod 4 read from
// the low mantissa bits.  Fast path: three f32 FMAs against a 3-way split of pi/2
// (exact first step for |k| < 2^13).  Slow path (|x| > 8192: the jitter hash of
// compute.wgsl:117, whose argument reaches 1e7..1e10): the same in f64.
SM_HD void reduce_pio2(float x, float& r, uint32_t& q)
{
    if (::fabsf(x) <= 8192.0f) {
        const float magic = 12582912.0f;                        // 1.5 * 2^23
        float t = fma(x, 0x1.45f306p-1f /* fl32(2/pi) */, magic);
        q = f2u(t) & 3u;
        float k = sub(t, magic);
        float nk = -k;
        r = fma(nk, 0x1.921fb6p+0f /* fl32(pi/2) */, x);
        r = fma(nk, -0x1.777a5cp-25f, r);
        r = fma(nk, -0x1.ee59dap-50f, r);
    } else {
        const double magic = 6755399441055744.0;                // 1.5 * 2^52
        double xd = (double)x;
        if (!(::fabs(xd) < 0x1p50))
            xd = ::fmod(xd, 0x1.921fb54442d18p+2);
#ifdef __CUDA_ARCH__
        double t = __fma_rn(xd, 0x1.45f306dc9c883p-1, magic);
        q = (uint32_t)__double2loint(t) & 3u;
        double k = __dsub_rn(t, magic);
        double rd = __fma_rn(-k, 0x1.921fb54442d18p+0, xd);
        rd = __fma_rn(-k, 0x1.1a62633145c07p-54, rd);
#else
        double t = ::fma(xd, 0x1.45f306dc9c883p-1, magic);
        uint64_t tb; std::memcpy(&tb, &t, 8);
        q = (uint32_t)tb & 3u;
        volatile double kv = t - magic;
        double k = kv;
        double rd = ::fma(-k, 0x1.921fb54442d18p+0, xd);
        rd = ::fma(-k, 0x1.1a62633145c07p-54, rd);
#endif
        r = (float)rd;
    }
}

// sin / cos of k*(pi/2) + r from sin(r), cos(r): q = k (only its two low bits are used).
SM_HD void quadrant_fix(float sinr, float cosr, uint32_t q, float& sn, float& cs)
{
    float so = (q & 1u) ? cosr : sinr;
    float co = (q & 1u) ? sinr : cosr;
    // sign flips as XOR on the sign bit (q&2 -> sin, (q+1)&2 -> cos); shifting by 30 drops the bits above
    sn = u2f(f2u(so) ^ (((q << 30)) & 0x80000000u));
    cs = u2f(f2u(co) ^ (((q << 30) + 0x40000000u) & 0x80000000u));
}

SM_HD void sincos_poly(float r, uint32_t q, float& sn, float& cs)
{
    float s2 = mul(r, r);
    float p = fma(-1.9515295891e-4f, s2, 8.3321608736e-3f);
    p = fma(p, s2, -1.6666654611e-1f);
    p = mul(p, s2);
    float sinr = fma(p, r, r);
    float c = fma(2.443315711809948e-5f, s2, -1.388731625493765e-3f);
    c = fma(c, s2, 4.166664568298827e-2f);
    c = fma(c, s2, -0.5f);
    float cosr = fma(c, s2, 1.0f);
    quadrant_fix(sinr, cosr, q, sn, cs);
}

SM_HD void sincos(float x, float& sn, float& cs)
{
    float r; uint32_t q;
    reduce_pio2(x, r, q);
    sincos_poly(r, q, sn, cs);
}

// Branch-free form for callers that have already established |x| <= 8192 (or x is NaN, for which
// both paths of the spec return NaN): identical results to sincos(), no range test.
SM_HD void sincos_small(float x, float& sn, float& cs)
{
    const float magic = 12582912.0f;
    float t = fma(x, 0x1.45f306p-1f, magic);
    uint32_t q = f2u(t) & 3u;
    float nk = -sub(t, magic);
    float r = fma(nk, 0x1.921fb6p+0f, x);
    r = fma(nk, -0x1.777a5cp-25f, r);
    r = fma(nk, -0x1.ee59dap-50f, r);
    sincos_poly(r, q, sn, cs);
}

// Two angles at once (the left / right sensor headings): the same statement sequence as
// sincos_small() on each lane, so the results are bit-identical to two scalar calls.
SM_HD void sincos_small2(f2 x, f2& sn, f2& cs)
{
    const f2 magic = splat2(12582912.0f);
    f2 t = fma2(x, splat2(0x1.45f306p-1f), magic);
    const uint32_t tl = f2u(t.lo), th = f2u(t.hi);
    f2 nk = neg2(add2(t, splat2(-12582912.0f)));              // -(t - magic), the sign of a zero matters (x = -0)
    f2 r = fma2(nk, splat2(0x1.921fb6p+0f), x);
    r = fma2(nk, splat2(-0x1.777a5cp-25f), r);
    r = fma2(nk, splat2(-0x1.ee59dap-50f), r);
    f2 s2 = mul2(r, r);
    f2 p = fma2(splat2(-1.9515295891e-4f), s2, splat2(8.3321608736e-3f));
    p = fma2(p, s2, splat2(-1.6666654611e-1f));
    p = mul2(p, s2);
    f2 sinr = fma2(p, r, r);
    f2 c = fma2(splat2(2.443315711809948e-5f), s2, splat2(-1.388731625493765e-3f));
    c = fma2(c, s2, splat2(4.166664568298827e-2f));
    c = fma2(c, s2, splat2(-0.5f));
    f2 cosr = fma2(c, s2, splat2(1.0f));
    quadrant_fix(sinr.lo, cosr.lo, tl, sn.lo, cs.lo);
    quadrant_fix(sinr.hi, cosr.hi, th, sn.hi, cs.hi);
}

// ---- exact truncated remainder (WGSL float %, == IEEE fmodf) ----------------
// b > 0, finite, normal; rcp_b ~ 1/b (any rounding: the quotient is corrected).
SM_HD float fmod_exact(float a, float b, float rcp_b)
{
    float aa = ::fabsf(a);
    float r;
    if (aa < b) {
        r = aa;
    } else if (aa <= mul(b, 4194304.0f)) {          // quotient < 2^22: estimate off by at most 1
        float qf = ::truncf(mul(aa, rcp_b));
        r = fma(-qf, b, aa);
        if (r < 0.0f) { qf = sub(qf, 1.0f); r = fma(-qf, b, aa); }
        else if (r >= b) { qf = add(qf, 1.0f); r = fma(-qf, b, aa); }
    } else {
        r = ::fmodf(aa, b);                           // huge / inf / NaN: library routine (exact)
    }
    return ::copysignf(r, a);
}

// ---- exact s / 9.0f (Markstein correction; verified against IEEE division for EVERY
// non-negative binary32 value incl. subnormals by tests/exhaustive_div9.c; negative
// values follow by symmetry except -0.0, which the 9-tap sum can never produce because
// it starts from +0.0) ------------------------------------------------------------
SM_HD float div9(float s)
{
    const float c = 0x1.c71c72p-4f;   // fl32(1/9)
    float q = mul(s, c);
    float r = fma(-q, 9.0f, s);
    float q2 = fma(r, c, q);
    return (::fabsf(s) <= 3.402823466e38f) ? q2 : s;   // inf / 9 = inf, NaN stays NaN
}

// ---- WGSL builtins ----------------------------------------------------------
SM_HD float clampf(float x, float lo, float hi) { return ::fminf(::fmaxf(x, lo), hi); }
SM_HD float mixf(float a, float b, float t)
{
    return add(mul(a, sub(1.0f, t)), mul(b, t));
}
SM_HD float mixf_pre(float a, float b, float t, float one_minus_t)
{
    return add(mul(a, one_minus_t), mul(b, t));
}
SM_HD float fractf(float v) { return sub(v, ::floorf(v)); }
SM_HD float signf(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

// compute.wgsl:117  fract(sin(f32(idx)*12.9898 + x*78.233 + y*37.719) * 43758.5453)
SM_HD float hash01(int32_t idx, float x, float y)
{
    float a = mul((float)idx, 12.9898f);
    float b = mul(x, 78.233f);
    float c = mul(y, 37.719f);
    float arg = add(add(a, b), c);
    float s, cc;
    sincos(arg, s, cc);
    return fractf(mul(s, 43758.5453f));
}

// ---- SPEC-RNG (seeded initial state) ----------------------------------------
SM_HD uint64_t mix64(uint64_t z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
SM_HD float rand01(uint64_t seed, uint64_t id, uint32_t stream)
{
    uint64_t z = (id * 4ull + (uint64_t)stream) * 0x9E3779B97F4A7C15ull
               + seed * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
    z = mix64(z);
    return mul((float)(uint32_t)(z >> 40), 0x1p-24f);
}

}  // namespace smd
