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

// u8 deposit flags in 8 x 8-cell tiles of 64 bytes (FLAGS == 2 / CM_FLAGS_TILED; W % 8 == 0, owned rows % 8 == 0; strips: peer-store path).
// Why: between two cell sorts the agents of a warp drift apart by a dozen pixels, and in a row-major field every lane's
// flag then lies in its own 32-byte sector -- the byte store costs k_agents 255 us of 1550 at BASELINE configs[2] and 17 of
// 174 at configs[1] (A/B builds with a second, dummy store: tools/r2/gpu_31.sh, profiles/r2_probe_deposit_layout.jsonl).
// In tiles a 3 x 3-tile neighbourhood is 18 sectors instead of 24 rows x lanes.
// Layout: with y' = (y - 1) mod H, tile (y' >> 3, x >> 3) starts at ((y' >> 3) * W/8 + (x >> 3)) * 64 and holds cell (x, y) at
// byte (y' & 7) * 8 + (x & 7).  The row grid is shifted by one because the trail pass requests rows y+1 .. y+4 per batch
// (the window's NEXT rows): with the shift those are one aligned 32-byte sector per tile -- 4 rows x 8 columns -- which a lane
// pair fetches with one 16-byte load each and splits by SHFL (kernels.cuh: k_trail_rows), at the sector efficiency of the row-major field.
// y is relative to the owned row 0 the base pointer addresses.  wrap = H on one GPU (row -1 is row H - 1); 0 on strips, where
// y' = -1 and the rows beyond are ghost rows of the same buffer (the strip's ghost + pad depth is a multiple of 8, so owned row 0
// starts a tile row there too): IdxT is signed, >> and & floor.
template <class IdxT>
SM_HD IdxT flag_tile_offset(IdxT x, IdxT y, IdxT W, IdxT wrap)
{
    IdxT yp = y - 1;
    if (yp < 0) yp += wrap;
    return ((yp >> 3) * W + (yp & 7)) * 8 + (x >> 3) * 64 + (x & 7);
}

}  // namespace smd
