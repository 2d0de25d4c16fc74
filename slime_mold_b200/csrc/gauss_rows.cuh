// gauss_rows.cuh -- EXTENSION (no reference semantics): decay + separable Gaussian, radius 1-8, as a register-only
// streaming pass, in the mould of k_trail_rows (the 3x3 pass that runs at the HBM roofline): no shared memory, no
// barriers.  Every thread owns four consecutive columns and walks DOWN a chunk of rows:
//
//   load     rows are requested in groups (all 2R+1 rows of a batch for R <= 4, six at a time above), every 16-byte
//            load of a group (+ the deposit marks) issued before the first use
//   stage    merge deposits, decay -> D (four cells)
//   h-blur   the R neighbours on each side come from the adjacent lanes by SHFL; the outermost lane(s) of a warp are
//            HALO lanes -- one per side for R <= 4, two above: they load, decay and feed their neighbours but own no
//            output -- so a warp produces 120 (112) columns from 128 loaded ones and there is no edge case anywhere: the
//            re-loaded columns hit in L1 / L2
//   v-blur   a row of h-blurred cells is tap d of the output d rows above it, so the taps of an output arrive in the
//            oracle's order d = 0 .. 2R: P = 2R+1 live accumulator quads, each finished (mixed with its decayed centre,
//            stored) and reset once per P rows.  The batch is unrolled P times, so every accumulator's role in every
//            row is static -- no register is ever moved or indexed dynamically.  PK >= 1 instantiations run these taps
//            as FFMA2 on column pairs (the weight pair (w, w) is a uniform-register operand: no per-thread cost); PK == 2
//            also packs the half of the row taps whose operands are aligned pairs.
//
// Arithmetic per output: exactly the oracle's statements (acc = 0.0f; acc = fma(w[d], v[d], acc), d = -R..R; rows after
// columns; mix(decayed centre, acc, rate)) -- a packed lane rounds like the scalar instruction -- so the bits equal the
// oracle's, the tile kernel's and the streaming kernel's (tests/test_gpu_parity.py; on the CPU through tests/hostcheck's
// CTA emulation of THIS source).
//
// HBM traffic: 8 B/cell + 2R/chunk_rows of re-read halo rows.  Registers grow with R (accumulators 4P, centres 4(R+1),
// rows in flight): the engine uses this kernel for the radii where it measured faster than the shared-memory streaming
// kernel (gauss_stream.cuh).
#pragma once
#include "gauss_stream.cuh"

namespace smk {

constexpr int kGrNT = 128;                 // threads per CTA
constexpr int kGrMinW = 128, kGrMinRows = 16;     // one fold per coordinate is enough
constexpr int kGrMaxR = 8;                 // two halo lanes hold eight columns
template <int R> SM_HDC int gr_halo_lanes() { return (R + 3) / 4; }
template <int R> SM_HDC int gr_warp_cols() { return (32 - 2 * gr_halo_lanes<R>()) * 4; }      // output columns per warp
template <int R> SM_HDC int gr_cta_cols() { return (kGrNT / 32) * gr_warp_cols<R>(); }
template <int R> SM_HDC int gr_load_rows() { return R <= 4 ? 2 * R + 1 : 6; }                 // rows requested at a time

template <int R, int CM, bool SURF, int PK, class Ctx>
SM_KD void gauss_rows_cta(const Ctx& cx, const GsArgs& a, const smd::TrailConsts& tc, const GaussConsts& gc)
{
    static_assert(R >= 1 && R <= kGrMaxR, "radius");
    constexpr int P = 2 * R + 1, HL = gr_halo_lanes<R>(), LB = gr_load_rows<R>();
    const int tid = cx.tid();
    const int lane = tid & 31, warp = tid >> 5;
    const int W = a.W, H = a.H;
    const int wx = (cx.bx() * (kGrNT / 32) + warp) * gr_warp_cols<R>();      // first output column of this warp
    if (wx >= W && cx.warp_may_exit()) return;
    const int y_begin = cx.by() * a.chunk_rows;
    const int nrows = (H - y_begin < a.chunk_rows) ? H - y_begin : a.chunk_rows;
    const int S = nrows + 2 * R;               // stream rows: map rows y_begin - R .. y_begin + nrows + R - 1
    const int nb = (S + P - 1) / P;

    int gx = wx - 4 * HL + 4 * lane;           // column of this thread's first cell, before folding
    const bool out_lane = lane >= HL && lane <= 31 - HL && gx < W;
    const bool ld_lane = gx < W + 4 * HL && wx < W;     // up to the right halo of the last output lane
    if (gx < 0) gx += W; else if (gx >= W) gx -= W;

    const float* tin = a.tin;
    const uint32_t* cin32 = static_cast<const uint32_t*>(a.cin);
    const uint8_t* cin8 = static_cast<const uint8_t*>(a.cin);
    int Wq = W, Hq = H;
    SM_OPAQUE64(tin); SM_OPAQUE64(cin32); SM_OPAQUE64(cin8); SM_OPAQUE32(Wq); SM_OPAQUE32(Hq);

    F4 acc[P];                                 // accumulator of the output whose chunk-relative row is == slot (mod P)
    F4 dprev[R];                               // decayed rows s0 - R .. s0 - 1 (centres of the outputs finished early in a batch)
    F4 t4[LB];
    U4 k4[CM == GS_COUNTS ? LB : 1];
    uint32_t kf[CM == GS_FLAGS ? LB : 1];
#pragma unroll
    for (int i = 0; i < P; ++i) { acc[i].x = acc[i].y = acc[i].z = acc[i].w = 0.0f; }
#pragma unroll
    for (int i = 0; i < LB; ++i) { t4[i].x = t4[i].y = t4[i].z = t4[i].w = 0.0f; }
#pragma unroll
    for (int i = 0; i < R; ++i) { dprev[i].x = dprev[i].y = dprev[i].z = dprev[i].w = 0.0f; }
#pragma unroll
    for (int i = 0; i < (CM == GS_COUNTS ? LB : 1); ++i) { k4[i].x = k4[i].y = k4[i].z = k4[i].w = 0u; }
#pragma unroll
    for (int i = 0; i < (CM == GS_FLAGS ? LB : 1); ++i) kf[i] = 0u;

    for (int b = 0; b < nb; ++b) {
        const int s0 = b * P;
        F4 drow[P];
        // ---- one row at a time; u is static, so are all accumulator roles ----
#pragma unroll
        for (int u = 0; u < P; ++u) {
            if (u % LB == 0) {
                // all addresses of the group, then its loads back to back (a row past the stream is not loaded: its
                // registers keep finite stale values that never reach a stored output)
                constexpr int NG = LB;
                int64_t off[NG];
                bool valid[NG];
#pragma unroll
                for (int q = 0; q < NG; ++q) {
                    if (u + q < P) {
                        int gy = y_begin - R + s0 + u + q;     // strips: rows -R .. -1 and H .. H+R-1 are ghost rows of the buffer
                        if (a.wrap_y) { if (gy < 0) gy += Hq; else if (gy >= Hq) gy -= Hq; }
                        valid[q] = ld_lane && s0 + u + q < S;
                        off[q] = (int64_t)gy * Wq + gx;
                    }
                }
#pragma unroll
                for (int q = 0; q < NG; ++q) {
                    if (u + q < P) {
                        cx.ld4(t4[q], tin + off[q], valid[q]);
                        if (CM == GS_COUNTS) cx.ldu4(k4[CM == GS_COUNTS ? q : 0], cin32 + off[q], valid[q]);
                        if (CM == GS_FLAGS) cx.ldu1(kf[CM == GS_FLAGS ? q : 0], reinterpret_cast<const uint32_t*>(cin8 + off[q]), valid[q]);
                    }
                }
            }
            const int s = s0 + u;
            constexpr int LBB = LB;
            const int q = u % LBB;
            F4 t = t4[q];
            if (CM == GS_COUNTS) {
                const U4 k = k4[CM == GS_COUNTS ? q : 0];
                t.x = smd::merge_deposit(t.x, k.x, tc.dep); t.y = smd::merge_deposit(t.y, k.y, tc.dep);
                t.z = smd::merge_deposit(t.z, k.z, tc.dep); t.w = smd::merge_deposit(t.w, k.w, tc.dep);
            }
            if (CM == GS_FLAGS) {                              // clamp(t + k*dep, 0, 1) == 1 for dep >= 1, t >= 0
                const uint32_t k = kf[CM == GS_FLAGS ? q : 0];
                t.x = (k & 0xffu) ? 1.0f : t.x; t.y = (k & 0xff00u) ? 1.0f : t.y;
                t.z = (k & 0xff0000u) ? 1.0f : t.z; t.w = (k & 0xff000000u) ? 1.0f : t.w;
            }
            if (CM != GS_NONE) {                               // this row is the centre of output row s - R: retire its marks
                const int oc = s - R;
                if (out_lane && oc >= 0 && oc < nrows) {
                    const int64_t offc = (int64_t)(y_begin + oc) * W + gx;       // inside the chunk: never folded
                    if (CM == GS_COUNTS) { U4 z; z.x = z.y = z.z = z.w = 0u; *reinterpret_cast<U4*>(static_cast<uint32_t*>(a.czero) + offc) = z; }
                    else *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(a.czero) + offc) = 0u;
                }
            }
            t.x = smd::decay_cell(t.x, tc.decay_sub); t.y = smd::decay_cell(t.y, tc.decay_sub);
            t.z = smd::decay_cell(t.z, tc.decay_sub); t.w = smd::decay_cell(t.w, tc.decay_sub);
            drow[u] = t;

            // h-blur: v[] = R cells of the lanes to the left | own four | R cells of the lanes to the right; output j taps v[j + d]
            float v[2 * R + 4];
            cx.template neighbours<R>(t, v);
            v[R] = t.x; v[R + 1] = t.y; v[R + 2] = t.z; v[R + 3] = t.w;
            float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f, h3 = 0.0f;
            if (PK >= 2) {
                // the taps with an even offset read the aligned pairs (v[d], v[d+1]), (v[d+2], v[d+3]): one FFMA2 for two
                // outputs; the odd ones stay scalar.  Every output still sees its taps in the order d = 0 .. 2R.
                smd::f2 h01 = smd::mk2(0.0f, 0.0f), h23 = smd::mk2(0.0f, 0.0f);
#pragma unroll
                for (int d = 0; d < P; ++d) {
                    if (d % 2 == 0) {
                        h01 = smd::fma2(gc.w2[d], smd::mk2(v[d], v[d + 1]), h01);
                        h23 = smd::fma2(gc.w2[d], smd::mk2(v[d + 2], v[d + 3]), h23);
                    } else {
                        const float w = gc.w[d];
                        h01.lo = smd::fma(w, v[d], h01.lo);
                        h01.hi = smd::fma(w, v[d + 1], h01.hi);
                        h23.lo = smd::fma(w, v[d + 2], h23.lo);
                        h23.hi = smd::fma(w, v[d + 3], h23.hi);
                    }
                }
                h0 = h01.lo; h1 = h01.hi; h2 = h23.lo; h3 = h23.hi;
            } else {
#pragma unroll
                for (int d = 0; d < P; ++d) {
                    const float w = gc.w[d];
                    h0 = smd::fma(w, v[d], h0);
                    h1 = smd::fma(w, v[d + 1], h1);
                    h2 = smd::fma(w, v[d + 2], h2);
                    h3 = smd::fma(w, v[d + 3], h3);
                }
            }
            // v-blur: this row is tap d of the output d rows above stream row s (chunk-relative output row s - d)
            if (PK >= 1) {
                const smd::f2 h01 = smd::mk2(h0, h1), h23 = smd::mk2(h2, h3);
#pragma unroll
                for (int d = 0; d < P; ++d) {
                    constexpr int PP = P;
                    const int slot = (u - d + PP) % PP;
                    const smd::f2 a01 = smd::fma2(gc.w2[d], h01, smd::mk2(acc[slot].x, acc[slot].y));
                    const smd::f2 a23 = smd::fma2(gc.w2[d], h23, smd::mk2(acc[slot].z, acc[slot].w));
                    acc[slot].x = a01.lo; acc[slot].y = a01.hi; acc[slot].z = a23.lo; acc[slot].w = a23.hi;
                }
            } else {
#pragma unroll
                for (int d = 0; d < P; ++d) {
                    constexpr int PP = P;
                    const int slot = (u - d + PP) % PP;
                    const float w = gc.w[d];
                    acc[slot].x = smd::fma(w, h0, acc[slot].x);
                    acc[slot].y = smd::fma(w, h1, acc[slot].y);
                    acc[slot].z = smd::fma(w, h2, acc[slot].z);
                    acc[slot].w = smd::fma(w, h3, acc[slot].w);
                }
            }
            // the output that has just received its last tap: chunk-relative row s - 2R, centre = stream row s - R
            {
                const int slot = (u + 1) % P;
                const int o = s - 2 * R;
                const F4 c = u >= R ? drow[u >= R ? u - R : 0] : dprev[u < R ? u : 0];
                if (out_lane && o >= 0 && o < nrows) {
                    F4 out;
                    out.x = smd::mixf_pre(c.x, acc[slot].x, tc.rate, tc.one_minus_rate);
                    out.y = smd::mixf_pre(c.y, acc[slot].y, tc.rate, tc.one_minus_rate);
                    out.z = smd::mixf_pre(c.z, acc[slot].z, tc.rate, tc.one_minus_rate);
                    out.w = smd::mixf_pre(c.w, acc[slot].w, tc.rate, tc.one_minus_rate);
                    const int gy = y_begin + o;
                    *reinterpret_cast<F4*>(a.tout + (int64_t)gy * W + gx) = out;
                    if (SURF) cx.surf_write(out, a.surf, gx, gy + a.surf_row0);
                }
                acc[slot].x = acc[slot].y = acc[slot].z = acc[slot].w = 0.0f;
            }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) dprev[i] = drow[P - R + i];
    }
}

}  // namespace smk
