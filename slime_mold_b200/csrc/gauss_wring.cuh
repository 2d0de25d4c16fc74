// gauss_wring.cuh -- EXTENSION (no reference semantics), EXPERIMENT (selected with SM_GAUSS_KERNEL=wring; not measured
// yet): decay + separable Gaussian, radius 1-8, for the radii where k_gauss_rows runs out of registers.
//
// k_gauss_rows keeps 2R+1 accumulator quads per thread (68 registers at R = 8: two resident CTAs per SM); k_gauss_stream
// keeps its state in shared memory but pays three CTA barriers per batch.  This kernel is the rows kernel with the column
// taps' state moved to shared memory WITHOUT any synchronisation: a lane only ever reads back the shared-memory words it
// wrote itself (the column taps are vertical), so shared memory is used as a statically addressed extension of the lane's
// register file -- a private ring of h-blurred rows, 16 bytes per row and lane, conflict-free LDS.128 / STS.128.
//
//   per group of four rows:  load (4 x 16 B, issued together) -> merge / decay -> row taps (neighbours by SHFL, halo lanes as
//   in gauss_rows.cuh) -> STS into the ring;  then the four outputs whose last tap has just arrived: 2R+4 ring rows are
//   read once each and spread over the outputs they feed (four accumulator quads), the decayed centre is recomputed from
//   a re-load of the output's own cell (an L1 / L2 hit: the row went by R rows ago), mix, store.
//
// The ring holds L = 2R+4 rows rounded up to a multiple of four.  A group's four rows never wrap (stores at immediate
// offsets from the group's base); the 2R+4 rows of the tap window wrap at most once (one compare + select per LDS.128).
// Arithmetic per output: the oracle's statements in the oracle's order, as everywhere else.
#pragma once
#include "gauss_rows.cuh"

namespace smk {

constexpr int kGwNT = 128;
template <int R> SM_HDC int gw_ring_rows() { return (2 * R + 4 + 3) / 4 * 4; }
template <int R> SM_HDC size_t gw_smem_bytes() { return (size_t)(kGwNT / 32) * gw_ring_rows<R>() * 128 * sizeof(float); }

template <int R, int CM, bool SURF, int PK, class Ctx>
SM_KD void gauss_wring_cta(const Ctx& cx, float* __restrict__ smem, const GsArgs& a, const smd::TrailConsts& tc, const GaussConsts& gc)
{
    static_assert(R >= 1 && R <= kGrMaxR, "radius");
    constexpr int P = 2 * R + 1, HL = gr_halo_lanes<R>(), L = gw_ring_rows<R>();
    const int tid = cx.tid();
    const int lane = tid & 31, warp = tid >> 5;
    const int W = a.W, H = a.H;
    const int wx = (cx.bx() * (kGwNT / 32) + warp) * gr_warp_cols<R>();      // first output column of this warp
    if (wx >= W && cx.warp_may_exit()) return;
    const int y_begin = cx.by() * a.chunk_rows;
    const int nrows = (H - y_begin < a.chunk_rows) ? H - y_begin : a.chunk_rows;
    const int S = nrows + 2 * R;               // stream rows: map rows y_begin - R .. y_begin + nrows + R - 1
    const int n_groups = (S + 3) / 4;

    int gx = wx - 4 * HL + 4 * lane;           // column of this thread's first cell, before folding
    const bool out_lane = lane >= HL && lane <= 31 - HL && gx < W;
    const bool ld_lane = gx < W + 4 * HL && wx < W;
    if (gx < 0) gx += W; else if (gx >= W) gx -= W;

    // this lane's private column of the warp's ring: row r at ring + r * 128 floats
    float* ring = smem + (size_t)warp * L * 128 + 4 * lane;

    const float* tin = a.tin;
    const uint32_t* cin32 = static_cast<const uint32_t*>(a.cin);
    const uint8_t* cin8 = static_cast<const uint8_t*>(a.cin);
    int Wq = W, Hq = H;
    SM_OPAQUE64(tin); SM_OPAQUE64(cin32); SM_OPAQUE64(cin8); SM_OPAQUE32(Wq); SM_OPAQUE32(Hq);

    F4 t4[4];
    U4 k4[CM == GS_COUNTS ? 4 : 1];
    uint32_t kf[CM == GS_FLAGS ? 4 : 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) { t4[i].x = t4[i].y = t4[i].z = t4[i].w = 0.0f; }
#pragma unroll
    for (int i = 0; i < (CM == GS_COUNTS ? 4 : 1); ++i) { k4[i].x = k4[i].y = k4[i].z = k4[i].w = 0u; }
#pragma unroll
    for (int i = 0; i < (CM == GS_FLAGS ? 4 : 1); ++i) kf[i] = 0u;
    // the ring starts as zeros: slots read before they are written only feed outputs above the chunk, which are not stored
#pragma unroll
    for (int r = 0; r < L; ++r) { F4 z; z.x = z.y = z.z = z.w = 0.0f; *reinterpret_cast<F4*>(ring + r * 128) = z; }

    auto stage = [&](F4 t, const U4& k, uint32_t f) -> F4 {      // merge the deposits, decay
        if (CM == GS_COUNTS) {
            t.x = smd::merge_deposit(t.x, k.x, tc.dep); t.y = smd::merge_deposit(t.y, k.y, tc.dep);
            t.z = smd::merge_deposit(t.z, k.z, tc.dep); t.w = smd::merge_deposit(t.w, k.w, tc.dep);
        }
        if (CM == GS_FLAGS) {                                      // clamp(t + k*dep, 0, 1) == 1 for dep >= 1, t >= 0
            t.x = (f & 0xffu) ? 1.0f : t.x; t.y = (f & 0xff00u) ? 1.0f : t.y;
            t.z = (f & 0xff0000u) ? 1.0f : t.z; t.w = (f & 0xff000000u) ? 1.0f : t.w;
        }
        t.x = smd::decay_cell(t.x, tc.decay_sub); t.y = smd::decay_cell(t.y, tc.decay_sub);
        t.z = smd::decay_cell(t.z, tc.decay_sub); t.w = smd::decay_cell(t.w, tc.decay_sub);
        return t;
    };

    int sb = 0;                                             // ring slot of the group's first row: 4 * g mod L
    for (int g = 0; g < n_groups; ++g) {
        {
            const int s0 = 4 * g;                           // first stream row of the group
            {
                // ---- the group's four rows: addresses, then the loads back to back ----
                {
                    int64_t off[4];
                    bool valid[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int gy = y_begin - R + s0 + j;     // strips: rows -R .. -1 and H .. H+R-1 are ghost rows of the buffer
                        if (a.wrap_y) { if (gy < 0) gy += Hq; else if (gy >= Hq) gy -= Hq; }
                        valid[j] = ld_lane && s0 + j < S;
                        off[j] = (int64_t)gy * Wq + gx;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        cx.ld4(t4[j], tin + off[j], valid[j]);
                        if (CM == GS_COUNTS) cx.ldu4(k4[CM == GS_COUNTS ? j : 0], cin32 + off[j], valid[j]);
                        if (CM == GS_FLAGS) cx.ldu1(kf[CM == GS_FLAGS ? j : 0], reinterpret_cast<const uint32_t*>(cin8 + off[j]), valid[j]);
                    }
                }
                // ---- row taps of the four rows -> ring ----
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const F4 t = stage(t4[j], k4[CM == GS_COUNTS ? j : 0], kf[CM == GS_FLAGS ? j : 0]);
                    float v[2 * R + 4];
                    cx.template neighbours<R>(t, v);
                    v[R] = t.x; v[R + 1] = t.y; v[R + 2] = t.z; v[R + 3] = t.w;
                    F4 h;
                    if (PK >= 2) {
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
                        h.x = h01.lo; h.y = h01.hi; h.z = h23.lo; h.w = h23.hi;
                    } else {
                        float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f, h3 = 0.0f;
#pragma unroll
                        for (int d = 0; d < P; ++d) {
                            const float w = gc.w[d];
                            h0 = smd::fma(w, v[d], h0);
                            h1 = smd::fma(w, v[d + 1], h1);
                            h2 = smd::fma(w, v[d + 2], h2);
                            h3 = smd::fma(w, v[d + 3], h3);
                        }
                        h.x = h0; h.y = h1; h.z = h2; h.w = h3;
                    }
                    *reinterpret_cast<F4*>(ring + (sb + j) * 128) = h;
                }
                // ---- the four outputs whose last tap row is in this group: chunk-relative rows o0 .. o0 + 3 ----
                const int o0 = s0 - 2 * R;
                if (o0 + 3 >= 0 && o0 < nrows) {            // (uniform)
                    // their centres: the outputs' own cells, re-loaded (the rows went by R rows ago: L1 / L2 hits)
                    F4 c4[4];
                    U4 ck[CM == GS_COUNTS ? 4 : 1];
                    uint32_t cf[CM == GS_FLAGS ? 4 : 1];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int o = o0 + j;
                        const bool ok = out_lane && o >= 0 && o < nrows;
                        const int64_t offc = (int64_t)(y_begin + (ok ? o : 0)) * Wq + gx;
                        c4[j].x = c4[j].y = c4[j].z = c4[j].w = 0.0f;
                        cx.ld4(c4[j], tin + offc, ok);
                        if (CM == GS_COUNTS) { ck[j].x = ck[j].y = ck[j].z = ck[j].w = 0u; cx.ldu4(ck[CM == GS_COUNTS ? j : 0], cin32 + offc, ok); }
                        if (CM == GS_FLAGS) { cf[j] = 0u; cx.ldu1(cf[CM == GS_FLAGS ? j : 0], reinterpret_cast<const uint32_t*>(cin8 + offc), ok); }
                    }
                    // column taps: ring row i of the window is stream row s0 - 2R + i, tap d = i - j of output j
                    float acc[4][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
                    // first row of the window: slot (sb - 2R) mod L = sb + 4 + (L - 2R - 4), folded once
                    int w0 = sb + 4 + (L - 2 * R - 4);
                    if (w0 >= L) w0 -= L;
                    const float* wp = ring + w0 * 128;
                    const float* const wend = ring + L * 128;
#pragma unroll
                    for (int i = 0; i < 2 * R + 4; ++i) {
                        const float* q = wp + i * 128;
                        if (q >= wend) q -= L * 128;
                        const F4 h = *reinterpret_cast<const F4*>(q);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int d = i - j;
                            if (d >= 0 && d <= 2 * R) {
                                if (PK >= 1) {
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
                    for (int j = 0; j < 4; ++j) {
                        const int o = o0 + j;
                        if (out_lane && o >= 0 && o < nrows) {
                            const F4 c = stage(c4[j], ck[CM == GS_COUNTS ? j : 0], cf[CM == GS_FLAGS ? j : 0]);
                            F4 out;
                            out.x = smd::mixf_pre(c.x, acc[j][0], tc.rate, tc.one_minus_rate);
                            out.y = smd::mixf_pre(c.y, acc[j][1], tc.rate, tc.one_minus_rate);
                            out.z = smd::mixf_pre(c.z, acc[j][2], tc.rate, tc.one_minus_rate);
                            out.w = smd::mixf_pre(c.w, acc[j][3], tc.rate, tc.one_minus_rate);
                            const int gy = y_begin + o;
                            const int64_t offo = (int64_t)gy * W + gx;
                            *reinterpret_cast<F4*>(a.tout + offo) = out;
                            if (SURF) cx.surf_write(out, a.surf, gx, gy + a.surf_row0);
                            if (CM == GS_COUNTS) { U4 z; z.x = z.y = z.z = z.w = 0u; *reinterpret_cast<U4*>(static_cast<uint32_t*>(a.czero) + offo) = z; }
                            if (CM == GS_FLAGS) *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(a.czero) + offo) = 0u;
                        }
                    }
                }
            }
        }
        sb += 4;
        if (sb == L) sb = 0;
    }
}

}  // namespace smk
