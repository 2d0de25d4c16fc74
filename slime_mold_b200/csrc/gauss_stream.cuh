// gauss_stream.cuh -- EXTENSION (no reference semantics): decay + separable Gaussian, radius 1-8, as ONE
// streaming pass.  A CTA owns a column strip of TX cells and walks DOWN a chunk of rows in batches of B rows:
//
//   stage   the B x (TX + 2RA) cells of the batch -- requested one batch ahead, so their DRAM latency hides behind the
//           arithmetic of the previous batch -- are merged with the deposits, decayed and parked in a ring of 2B rows (D)
//   h-blur  row taps for the B new rows, four outputs per item from a register window filled by LDS.128 -> ring Hb
//   v-blur  the B output rows whose last tap row has just arrived (they lag the newest row by R): four columns x B/4
//           rows per thread, every Hb row read once per thread and spread over the outputs it feeds; mix with the
//           decayed centre; one 16-byte store per four cells (plus the sampler's block-linear copy / the retired
//           deposit marks in a full step)
//
// Against the tile kernel it replaces (k_gauss_fused, kernels.cuh): no vertical halo is recomputed (the tile kernel
// ran the row taps for 32 + 2R rows per 32 output rows), the loads of the next batch overlap the taps of this one,
// and the rings are 34-68 KB whatever the radius.  HBM traffic: 8 B/cell + 2RA/TX column halo + 2R/chunk row halo.
//
// Arithmetic: per output exactly the oracle's statements -- acc = 0.0f; acc = fma(w[d], v[d], acc) for d = -R..R, rows
// after columns, then mix(decayed centre, acc, rate) -- so the bits equal those of the oracle, the tile kernel and the
// two-pass form (tests/test_gpu_parity.py, and on the CPU through tests/hostcheck's CTA emulation of THIS source).
//
// The body is written against a small context (thread / block index, barrier, global loads, surface store) so that
// tests/hostcheck can run the very same statements on host threads; the product only ever instantiates DevCtx.
#pragma once
#include <cstdint>
#include "trail_core.cuh"

namespace smk {

#if defined(__CUDACC__)
#define SM_KD __device__ __forceinline__
#define SM_HDC __host__ __device__ constexpr
#define SM_HDC_INLINE __host__ __device__ inline
// A value the compiler must keep in a register from here on: kernel parameters are otherwise re-read from the constant bank
// (LDC / LDCU) wherever they are used, and those reads share a scoreboard with the global loads around them -- ncu showed
// each LDG of the prefetch waiting for the previous one to RETURN because an LDC sat between them.
#define SM_OPAQUE32(x) asm volatile("" : "+r"(x))
#define SM_OPAQUE64(p) asm volatile("" : "+l"(p))
using F4 = float4;
using U4 = uint4;
#else
#define SM_KD inline
#define SM_HDC constexpr
#define SM_HDC_INLINE inline
#define SM_OPAQUE32(x) ((void)0)
#define SM_OPAQUE64(p) ((void)0)
struct alignas(16) F4 { float x, y, z, w; };
struct alignas(16) U4 { uint32_t x, y, z, w; };
#endif

struct GaussConsts {
    int R;
    float w[17];
    smd::f2 w2[17];         // (w[d], w[d]): the FFMA2 operand of the packed column taps (gauss_rows.cuh)
    SM_HDC_INLINE void set(int d, float v) { w[d] = v; w2[d].lo = v; w2[d].hi = v; }
};

// deposit representation seen by the pass (same values as CM_* in kernels.cuh)
enum { GS_NONE = 0, GS_COUNTS = 1, GS_FLAGS = 2 };

constexpr int kGsTX = 256;      // output columns per CTA
constexpr int kGsNT = 256;      // threads per CTA
template <int R> SM_HDC int gs_ra() { return (R + 3) / 4 * 4; }                  // halo columns staged per side (whole float4s)
template <int R> SM_HDC int gs_batch() { return 16; }   // rows per batch; 2R <= B (batches of 8 rows for radius <= 4: three barriers per 8 rows were 29 % of the stall samples)
template <int R> SM_HDC int gs_dc() { return kGsTX + 2 * gs_ra<R>() + 4; }       // row stride of D (floats): 16-byte multiple, +4 against bank conflicts
template <int R> SM_HDC size_t gs_smem_bytes() { return sizeof(float) * 2 * gs_batch<R>() * (size_t)(gs_dc<R>() + kGsTX); }
// smallest map the kernel takes: a float4 never straddles the seam, one fold per coordinate is enough
constexpr int kGsMinW = kGsTX + 32, kGsMinRows = 64;

struct GsArgs {
    const float* tin;
    const void* cin;        // GS_COUNTS: u32 per cell, GS_FLAGS: u8 per cell (deposits of this step)
    void* czero;            // the other deposit buffer: this pass retires (zeroes) the cells it owns
    float* tout;
    int W, H;               // row length, owned rows (the whole map on one GPU, the strip on several)
    int wrap_y;             // 1: rows wrap toroidally inside the buffer (single GPU); 0: strip with at least R ghost rows
                            //    above and below the owned rows (filled by the neighbours before the pass)
    int chunk_rows;         // output rows per CTA (blockIdx.y)
    unsigned long long surf;   // block-linear copy of the output for the TEX sampler (SURF instantiations)
    int surf_row0;
};

struct GsFoldYes { static constexpr bool value = true; };      // prefetch instantiations: rows folded across the seam, or not
struct GsFoldNo { static constexpr bool value = false; };

template <int R, int CM, bool SURF, bool PK, class Ctx>
SM_KD void gauss_stream_cta(const Ctx& cx, float* __restrict__ gsm, const GsArgs& a, const smd::TrailConsts& tc, const GaussConsts& gc)
{
    constexpr int TX = kGsTX, NT = kGsNT, B = gs_batch<R>(), RA = gs_ra<R>(), DC = gs_dc<R>();
    constexpr int NR = 2 * B;                       // ring rows (power of two)
    constexpr int CG = TX / 4;                      // float4 column groups of the tile (64)
    constexpr int RPK = NT / CG;                    // rows one sweep of the CTA's threads covers (4)
    constexpr int PC = B / RPK;                     // centre pieces per thread and batch (4): row 4k + tid/64, column group tid%64
    constexpr int HP = RA / 2;                      // halo float4s per row (left RA/4 + right RA/4)
    constexpr int PER = PC + 1;                     // + one halo piece for the first B*HP threads: row tid/HP, halo column tid%HP
    static_assert(2 * R <= B && (NR & (NR - 1)) == 0 && NT % CG == 0 && B % RPK == 0 && B * HP <= NT, "ring / piece geometry");
    float* D = gsm;                    // [NR][DC]  decayed cells; column c <-> map column x0 - RA + c; stream row s in slot s & (NR-1)
    float* Hb = gsm + NR * DC;         // [NR][TX]  row-blurred cells
    const int tid = cx.tid();
    const int W = a.W, H = a.H;
    const int x0 = cx.bx() * TX;
    const int yc0 = cx.by() * a.chunk_rows;
    const int nrows = (H - yc0 < a.chunk_rows) ? H - yc0 : a.chunk_rows;
    const int S = nrows + 2 * R;       // stream rows: map rows yc0 - R .. yc0 + nrows + R - 1
    const int nb = (S + B - 1) / B;

    const float* tin = a.tin;
    const uint32_t* cin32 = static_cast<const uint32_t*>(a.cin);
    const uint8_t* cin8 = static_cast<const uint8_t*>(a.cin);
    int Wq = W, Hq = H;
    SM_OPAQUE64(tin); SM_OPAQUE64(cin32); SM_OPAQUE64(cin8); SM_OPAQUE32(Wq); SM_OPAQUE32(Hq);

    // ---- the pieces this thread stages are the same for every batch: shifts and masks only, no division ----
    const int cg = tid & (CG - 1), rsub = tid / CG;
    const bool own_col = x0 + 4 * cg < W;                      // ragged last tile: columns past the map are staged (as the
    int gxc = x0 + 4 * cg;                                      // toroidal neighbours they are) but never stored
    if (gxc >= W) gxc -= W;
    const bool has_halo = tid < B * HP;
    const int hrow = tid / HP, hq = tid % HP;                   // HP is 2 or 4
    const int hcol = hq < HP / 2 ? 4 * hq : RA + TX + 4 * (hq - HP / 2);   // float column of D: left halo | right halo
    int gxh = x0 - RA + hcol;
    if (gxh < 0) gxh += W; else if (gxh >= W) gxh -= W;

    F4 t4[PER];
    U4 k4[CM == GS_COUNTS ? PER : 1];
    uint32_t kf[CM == GS_FLAGS ? PER : 1];
    // a piece past the end of the stream is not loaded (predicate): its registers keep whatever they held -- finite
    // values, zero at first -- and the ring rows they are parked in never reach a stored output
#pragma unroll
    for (int k = 0; k < PER; ++k) { t4[k].x = t4[k].y = t4[k].z = t4[k].w = 0.0f; }
#pragma unroll
    for (int k = 0; k < (CM == GS_COUNTS ? PER : 1); ++k) { k4[k].x = k4[k].y = k4[k].z = k4[k].w = 0u; }
#pragma unroll
    for (int k = 0; k < (CM == GS_FLAGS ? PER : 1); ++k) kf[k] = 0u;
    int pgy[PC];                       // buffer rows of the centre pieces in flight (their deposit marks are retired by the stage)

    // Does this chunk touch the toroidal seam at all?  (Uniform per CTA; all but the first and last chunk of a map do not.)
    const bool seam = a.wrap_y && (yc0 - R < 0 || yc0 + nrows + R > H);

    // All addresses first, then the loads back to back (predicated, no branch, nothing between them).
    auto prefetch = [&](int kb, auto fold_tag) {
        constexpr bool FOLD = decltype(fold_tag)::value;
        int64_t off[PER];
        bool valid[PER];
        const int nleft = S - kb * B;                 // stream rows left, this batch included
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int r = k < PC ? RPK * k + rsub : hrow;
            valid[k] = (k < PC || has_halo) && r < nleft;
            int gy = yc0 - R + kb * B + r;            // strips: rows -R .. -1 and H .. H+R-1 are ghost rows of the buffer
            if (FOLD) { if (gy < 0) gy += Hq; else if (gy >= Hq) gy -= Hq; }
            if (k < PC) pgy[k] = gy;
            off[k] = (int64_t)gy * Wq + (k < PC ? gxc : gxh);
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            cx.ld4(t4[k], tin + off[k], valid[k]);
            if (CM == GS_COUNTS) cx.ldu4(k4[k], cin32 + off[k], valid[k]);
            if (CM == GS_FLAGS) cx.ldu1(kf[k], reinterpret_cast<const uint32_t*>(cin8 + off[k]), valid[k]);
        }
    };

    if (seam) prefetch(0, GsFoldYes{}); else prefetch(0, GsFoldNo{});
    for (int kb = 0; kb < nb; ++kb) {
        const int blk = (kb & 1) * B;
        // ---- stage: registers -> D ring (merge, decay); retire the deposit marks of the cells this CTA owns ----
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (k < PC || has_halo) {
                const int r = k < PC ? RPK * k + rsub : hrow;
                F4 t = t4[k];
                if (CM != GS_NONE) {
                    if (CM == GS_COUNTS) {
                        t.x = smd::merge_deposit(t.x, k4[k].x, tc.dep); t.y = smd::merge_deposit(t.y, k4[k].y, tc.dep);
                        t.z = smd::merge_deposit(t.z, k4[k].z, tc.dep); t.w = smd::merge_deposit(t.w, k4[k].w, tc.dep);
                    } else {                                         // clamp(t + k*dep, 0, 1) == 1 for dep >= 1, t >= 0
                        t.x = (kf[k] & 0xffu) ? 1.0f : t.x; t.y = (kf[k] & 0xff00u) ? 1.0f : t.y;
                        t.z = (kf[k] & 0xff0000u) ? 1.0f : t.z; t.w = (kf[k] & 0xff000000u) ? 1.0f : t.w;
                    }
                    if (k < PC) {                                    // only centre pieces are cells this CTA owns
                        const int o = kb * B + r - R;                // output row (chunk-relative) this stream row is the centre of
                        if (own_col && o >= 0 && o < nrows) {
                            const int64_t off = (int64_t)pgy[k < PC ? k : 0] * W + gxc;
                            if (CM == GS_COUNTS) { U4 z; z.x = z.y = z.z = z.w = 0u; *reinterpret_cast<U4*>(static_cast<uint32_t*>(a.czero) + off) = z; }
                            else *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(a.czero) + off) = 0u;
                        }
                    }
                }
                t.x = smd::decay_cell(t.x, tc.decay_sub); t.y = smd::decay_cell(t.y, tc.decay_sub);
                t.z = smd::decay_cell(t.z, tc.decay_sub); t.w = smd::decay_cell(t.w, tc.decay_sub);
                *reinterpret_cast<F4*>(D + (blk + r) * DC + (k < PC ? RA + 4 * cg : hcol)) = t;
            }
        }
        if (kb + 1 < nb) { if (seam) prefetch(kb + 1, GsFoldYes{}); else prefetch(kb + 1, GsFoldNo{}); }   // in flight while this batch is blurred
        cx.sync();

        // ---- h-blur: Hb[row][xs .. xs+3]; output j taps D columns xs + j + (RA - R) + d, d = 0 .. 2R ----
        {
            constexpr int NV = (RA + R + 4 + 3) / 4 * 4, SH = RA - R;
            // item i of a thread: row RPK*i + rsub of the batch, columns 4*cg .. 4*cg+3 -- fixed offsets from one base address
            const float* dsrc = D + (blk + rsub) * DC + 4 * cg;
            float* hdst = Hb + (blk + rsub) * TX + 4 * cg;
#pragma unroll
            for (int i = 0; i < PC; ++i) {
                float v[NV];
                const F4* src = reinterpret_cast<const F4*>(dsrc + i * RPK * DC);
#pragma unroll
                for (int q = 0; q < NV / 4; ++q) {
                    const F4 f = src[q];
                    v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
                }
                float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
                if (PK) {
                    // taps whose window offset is even read the aligned pairs (v[i], v[i+1]), (v[i+2], v[i+3]) the LDS.128s
                    // delivered: one FFMA2 (uniform-register weight pair) for two outputs; the others stay scalar
                    smd::f2 a01 = smd::mk2(0.0f, 0.0f), a23 = smd::mk2(0.0f, 0.0f);
#pragma unroll
                    for (int d = 0; d <= 2 * R; ++d) {
                        if ((SH + d) % 2 == 0) {
                            a01 = smd::fma2(gc.w2[d], smd::mk2(v[SH + d], v[SH + d + 1]), a01);
                            a23 = smd::fma2(gc.w2[d], smd::mk2(v[SH + d + 2], v[SH + d + 3]), a23);
                        } else {
                            const float w = gc.w[d];
                            a01.lo = smd::fma(w, v[SH + d], a01.lo);
                            a01.hi = smd::fma(w, v[SH + d + 1], a01.hi);
                            a23.lo = smd::fma(w, v[SH + d + 2], a23.lo);
                            a23.hi = smd::fma(w, v[SH + d + 3], a23.hi);
                        }
                    }
                    a0 = a01.lo; a1 = a01.hi; a2 = a23.lo; a3 = a23.hi;
                } else {
#pragma unroll
                    for (int d = 0; d <= 2 * R; ++d) {
                        const float w = gc.w[d];
                        a0 = smd::fma(w, v[SH + d], a0);
                        a1 = smd::fma(w, v[SH + d + 1], a1);
                        a2 = smd::fma(w, v[SH + d + 2], a2);
                        a3 = smd::fma(w, v[SH + d + 3], a3);
                    }
                }
                F4 o;
                o.x = a0; o.y = a1; o.z = a2; o.w = a3;
                *reinterpret_cast<F4*>(hdst + i * RPK * TX) = o;
            }
        }
        cx.sync();

        // ---- v-blur: outputs o_first .. o_first + VR - 1 (chunk-relative rows) of columns 4*c4 .. 4*c4+3 ----
        {
            constexpr int VR = B / 4;                // TX/4 = 64 column groups x 4 row groups = 256 threads
            const int c4 = cg, rg = rsub;
            const int o_first = kb * B - 2 * R + rg * VR;
            const int gx = x0 + 4 * c4;
            if (o_first + VR > 0 && o_first < nrows && gx < W) {
                float acc[VR][4];
#pragma unroll
                for (int j = 0; j < VR; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
                // stream row o + d holds tap d of output o: rows arrive in tap order, so each accumulator sees d = 0 .. 2R
                // ring slot of stream row o_first + i as a byte offset: one add and one mask per row (TX * 4 bytes per row,
                // the column offset sits below the row bits and passes through the mask)
                const uint32_t hoff0 = ((uint32_t)(o_first & (NR - 1)) * TX + 4u * (uint32_t)c4) * 4u;
#pragma unroll
                for (int i = 0; i < VR + 2 * R; ++i) {
                    const F4 h = *reinterpret_cast<const F4*>(reinterpret_cast<const char*>(Hb) +
                                                              ((hoff0 + (uint32_t)i * (TX * 4u)) & (uint32_t)(NR * TX * 4 - 1)));
#pragma unroll
                    for (int j = 0; j < VR; ++j) {
                        const int d = i - j;
                        if (d >= 0 && d <= 2 * R) {
                            if (PK) {                                // column pairs, weight pair from a uniform register
                                const smd::f2 p01 = smd::fma2(gc.w2[d], smd::mk2(h.x, h.y), smd::mk2(acc[j][0], acc[j][1]));
                                const smd::f2 p23 = smd::fma2(gc.w2[d], smd::mk2(h.z, h.w), smd::mk2(acc[j][2], acc[j][3]));
                                acc[j][0] = p01.lo; acc[j][1] = p01.hi; acc[j][2] = p23.lo; acc[j][3] = p23.hi;
                            } else {
                                const float w = gc.w[d];
                                acc[j][0] = smd::fma(w, h.x, acc[j][0]);
                                acc[j][1] = smd::fma(w, h.y, acc[j][1]);
                                acc[j][2] = smd::fma(w, h.z, acc[j][2]);
                                acc[j][3] = smd::fma(w, h.w, acc[j][3]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < VR; ++j) {
                    const int o = o_first + j;
                    if (o >= 0 && o < nrows) {
                        const F4 c = *reinterpret_cast<const F4*>(D + ((o + R) & (NR - 1)) * DC + RA + 4 * c4);
                        F4 out;
                        out.x = smd::mixf_pre(c.x, acc[j][0], tc.rate, tc.one_minus_rate);
                        out.y = smd::mixf_pre(c.y, acc[j][1], tc.rate, tc.one_minus_rate);
                        out.z = smd::mixf_pre(c.z, acc[j][2], tc.rate, tc.one_minus_rate);
                        out.w = smd::mixf_pre(c.w, acc[j][3], tc.rate, tc.one_minus_rate);
                        const int gy = yc0 + o;
                        *reinterpret_cast<F4*>(a.tout + (int64_t)gy * W + gx) = out;
                        if (SURF) cx.surf_write(out, a.surf, gx, gy + a.surf_row0);
                    }
                }
            }
        }
        cx.sync();          // the next batch overwrites ring block (kb+1)&1, which this v-blur was still reading
    }
}

}  // namespace smk
