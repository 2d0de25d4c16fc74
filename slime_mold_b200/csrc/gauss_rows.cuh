// gauss_rows.cuh -- EXTENSION (no reference semantics): decay + separable Gaussian for SMALL radii as a register-only
// streaming pass, in the mould of k_trail_rows (the 3x3 pass that runs at the HBM roofline): no shared memory, no
// barriers.  Every thread owns four consecutive columns and walks DOWN a chunk of rows:
//
//   load     a batch of P = 2R+1 rows (one 16-byte load per row, + the deposit marks), all issued before the first use
//   stage    merge deposits, decay -> D (four cells)
//   h-blur   the R neighbours on each side come from the adjacent lanes by SHFL; lanes 0 and 31 of a warp are HALO
//            lanes (they load, decay and feed their neighbours but own no output), so a warp produces 120 columns from
//            128 loaded ones and there is no edge case anywhere: the ~6 % of re-loaded columns hit in L1 / L2
//   v-blur   a row of h-blurred cells is tap d of the output d rows above it, so the taps of an output arrive in the
//            oracle's order d = 0 .. 2R: P live accumulator quads, each finished (mixed with its decayed centre, stored)
//            and reset once per P rows.  The batch is unrolled P times, so every accumulator's role in every row is
//            static -- no register is ever moved or indexed dynamically.
//
// Arithmetic per output: exactly the oracle's statements (acc = 0.0f; acc = fma(w[d], v[d], acc), d = -R..R; rows after
// columns; mix(decayed centre, acc, rate)), so the bits equal the oracle's, the tile kernel's and the streaming
// kernel's (tests/test_gpu_parity.py; on the CPU through tests/hostcheck's CTA emulation of THIS source).
//
// HBM traffic: 8 B/cell + 2R/chunk_rows of re-read halo rows.  Registers grow with R (accumulators 4P, rows in flight
// 5P): the engine uses this kernel for the radii where it measured faster than the shared-memory streaming kernel.
#pragma once
#include "gauss_stream.cuh"

namespace smk {

constexpr int kGrNT = 128;                 // threads per CTA
constexpr int kGrWarpCols = 120;           // output columns per warp (lanes 1..30)
constexpr int kGrCtaCols = (kGrNT / 32) * kGrWarpCols;
constexpr int kGrMinW = 128, kGrMinRows = 16;     // one fold per coordinate is enough
constexpr int kGrMaxR = 4;                 // the halo lane holds four columns

template <int R, int CM, bool SURF, class Ctx>
SM_KD void gauss_rows_cta(const Ctx& cx, const GsArgs& a, const smd::TrailConsts& tc, const GaussConsts& gc)
{
    static_assert(R >= 1 && R <= kGrMaxR, "radius");
    constexpr int P = 2 * R + 1;
    const int tid = cx.tid();
    const int lane = tid & 31, warp = tid >> 5;
    const int W = a.W, H = a.H;
    const int wx = (cx.bx() * (kGrNT / 32) + warp) * kGrWarpCols;      // first output column of this warp
    if (wx >= W && cx.warp_may_exit()) return;
    const int y_begin = cx.by() * a.chunk_rows;
    const int nrows = (H - y_begin < a.chunk_rows) ? H - y_begin : a.chunk_rows;
    const int S = nrows + 2 * R;               // stream rows: map rows y_begin - R .. y_begin + nrows + R - 1
    const int nb = (S + P - 1) / P;

    int gx = wx - 4 + 4 * lane;                // column of this thread's first cell, before folding
    const bool out_lane = lane >= 1 && lane <= 30 && gx < W;
    const bool ld_lane = gx < W + 4 && wx < W; // up to the right halo of the last output lane
    if (gx < 0) gx += W; else if (gx >= W) gx -= W;

    const float* tin = a.tin;
    const uint32_t* cin32 = static_cast<const uint32_t*>(a.cin);
    const uint8_t* cin8 = static_cast<const uint8_t*>(a.cin);
    int Wq = W, Hq = H;
    SM_OPAQUE64(tin); SM_OPAQUE64(cin32); SM_OPAQUE64(cin8); SM_OPAQUE32(Wq); SM_OPAQUE32(Hq);

    F4 acc[P];                                 // accumulator of the output whose chunk-relative row is == slot (mod P)
    F4 dprev[R];                               // decayed rows s0 - R .. s0 - 1 (centres of the outputs finished early in a batch)
    F4 t4[P];
    U4 k4[CM == GS_COUNTS ? P : 1];
    uint32_t kf[CM == GS_FLAGS ? P : 1];
#pragma unroll
    for (int i = 0; i < P; ++i) { acc[i].x = acc[i].y = acc[i].z = acc[i].w = 0.0f; t4[i] = acc[i]; }
#pragma unroll
    for (int i = 0; i < R; ++i) { dprev[i].x = dprev[i].y = dprev[i].z = dprev[i].w = 0.0f; }
#pragma unroll
    for (int i = 0; i < (CM == GS_COUNTS ? P : 1); ++i) { k4[i].x = k4[i].y = k4[i].z = k4[i].w = 0u; }
#pragma unroll
    for (int i = 0; i < (CM == GS_FLAGS ? P : 1); ++i) kf[i] = 0u;

    for (int b = 0; b < nb; ++b) {
        const int s0 = b * P;
        // ---- all addresses, then the loads back to back (a row past the stream is not loaded: its registers keep
        //      finite stale values that never reach a stored output) ----
        int64_t off[P];
        bool valid[P];
#pragma unroll
        for (int u = 0; u < P; ++u) {
            int gy = y_begin - R + s0 + u;     // strips: rows -R .. -1 and H .. H+R-1 are ghost rows of the buffer
            if (a.wrap_y) { if (gy < 0) gy += Hq; else if (gy >= Hq) gy -= Hq; }
            valid[u] = ld_lane && s0 + u < S;
            off[u] = (int64_t)gy * Wq + gx;
        }
#pragma unroll
        for (int u = 0; u < P; ++u) {
            cx.ld4(t4[u], tin + off[u], valid[u]);
            if (CM == GS_COUNTS) cx.ldu4(k4[u], cin32 + off[u], valid[u]);
            if (CM == GS_FLAGS) cx.ldu1(kf[u], reinterpret_cast<const uint32_t*>(cin8 + off[u]), valid[u]);
        }
        // ---- one row at a time; u is static, so are all accumulator roles ----
        F4 drow[P];
#pragma unroll
        for (int u = 0; u < P; ++u) {
            const int s = s0 + u;
            F4 t = t4[u];
            if (CM == GS_COUNTS) {
                t.x = smd::merge_deposit(t.x, k4[u].x, tc.dep); t.y = smd::merge_deposit(t.y, k4[u].y, tc.dep);
                t.z = smd::merge_deposit(t.z, k4[u].z, tc.dep); t.w = smd::merge_deposit(t.w, k4[u].w, tc.dep);
            }
            if (CM == GS_FLAGS) {                              // clamp(t + k*dep, 0, 1) == 1 for dep >= 1, t >= 0
                t.x = (kf[u] & 0xffu) ? 1.0f : t.x; t.y = (kf[u] & 0xff00u) ? 1.0f : t.y;
                t.z = (kf[u] & 0xff0000u) ? 1.0f : t.z; t.w = (kf[u] & 0xff000000u) ? 1.0f : t.w;
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

            // h-blur: v[] = R cells of the left lane | own four | R cells of the right lane; output j taps v[j + d]
            float v[2 * R + 4];
            cx.template neighbours<R>(t, v);
            v[R] = t.x; v[R + 1] = t.y; v[R + 2] = t.z; v[R + 3] = t.w;
            float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f, h3 = 0.0f;
#pragma unroll
            for (int d = 0; d < P; ++d) {
                const float w = gc.w[d];
                h0 = smd::fma(w, v[d], h0);
                h1 = smd::fma(w, v[d + 1], h1);
                h2 = smd::fma(w, v[d + 2], h2);
                h3 = smd::fma(w, v[d + 3], h3);
            }
            // v-blur: this row is tap d of the output d rows above stream row s (chunk-relative output row s - d)
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
