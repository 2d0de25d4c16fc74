// kernels.cuh -- sm_100a CUDA kernels of the slime-mold step loop.
//
//   k_agents        compute.wgsl `main`          (:57-145)  one agent per thread
//   k_trail_rows    `decay_trail`+`diffuse_trail` (:148-195) fused with the deposit
//                   merge, register sliding window, float4 row streaming
//   k_trail_generic same arithmetic, one cell per thread, any W (ragged sizes)
//   k_tile_*        periodic counting sort of agents by map tile (gather locality)
//
// HBM layout (DESIGN.md "Data layout"): agents are SoA-of-vectors -- one float4
// (x, y, angle, speed) array plus a u32 persistent-index array, so one LDG.128 /
// STG.128 per agent, coalesced; the trail is row-major f32 with `ghost` rows above
// and below the owned strip (0 on a single GPU); deposits are u32 per-cell counts in
// two alternating buffers (the trail pass zeroes the one the next step will use).
#pragma once
#include <cuda_runtime.h>
#include "agent_core.cuh"
#include "trail_core.cuh"
#include "gauss_stream.cuh"
#include "gauss_rows.cuh"

namespace smk {

using smd::AgentConsts;
using smd::TrailConsts;

struct LdgF32 {
    __device__ __forceinline__ float operator()(const float* p) const { return __ldg(p); }
};

// 2x2 footprint by ONE texture gather from the block-linear copy of the trail (cudaArray with
// cudaArrayTextureGather; point sampling, unnormalised coordinates -> raw f32 texels, no filtering
// arithmetic).  At (x0 + 1, y0 + 1) -- the common corner of the four texels -- the footprint is
// {x0, x0+1} x {y0, y0+1}; components come back as (x0,y0+1), (x0+1,y0+1), (x0+1,y0), (x0,y0)
// (verified at sm_create by k_gather_probe: the engine refuses the TEX path otherwise).
struct FetchTex {
    cudaTextureObject_t tex;
    float row_off1;          // (array row of global row 0) + 1 = ghost + pad - row_base + 1, exact in f32
    __device__ __forceinline__ void operator()(const AgentConsts&, float fx, float fy, float& v00, float& v10, float& v01, float& v11) const
    {
        // fetched whether or not the tap is inside the map (clamped addressing; the caller discards the
        // footprint of an outside tap).  Coordinates stay in float: for an inside tap fx, fy are integral
        // and < 2^17, so the sums are exact
        float4 g = tex2Dgather<float4>(tex, fx + 1.0f, fy + row_off1, 0);
        v01 = g.x; v11 = g.y; v10 = g.z; v00 = g.w;
    }
};
// Measured and dropped in round 2 (profiles/README.md, "sensor distance 225"): a second copy of the field shifted by
// (+4, +2) texels in the same array, read by the footprints that straddle a 128-byte line (8 x 4 texels) of the first
// copy -- 34 % of them at sensor distance 225, where no two footprints share a line and the kernel is bound by the L1's
// miss-request port.  tools/microbench/gather_rate.cu: 1.44 -> 1.0 SM-cycles per incoherent footprint; in the kernel:
// 12 % fewer L2 requests, 1-9 % less time, paid back by the trail pass writing the second copy.

// Row-major trail rows -> the block-linear copy the TEX sampler reads (ghost rows after an exchange).
// A kernel instead of cudaMemcpy2DToArray: it stays on the compute engine (no copy-engine hand-off
// inside the step loop).  Two row ranges per launch: [r0a, r0a + n) and [r0b, r0b + n) (buffer rows).
static __global__ void __launch_bounds__(256)
k_rows_to_surface(const float* __restrict__ base, cudaSurfaceObject_t surf, uint32_t W, int32_t r0a, int32_t r0b, int32_t n)
{
    const uint64_t total = 2ull * (uint64_t)n * W;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % W);
        const uint64_t r = i / W;
        const int32_t row = r < (uint64_t)n ? r0a + (int32_t)r : r0b + (int32_t)(r - n);
        surf2Dwrite(base[(size_t)row * W + x], surf, (int)(x * 4u), row);
    }
}

// out[0..3] = gather at the corner of texels (1,1),(2,1),(1,2),(2,2) of a probe array holding T[y][x] = 10*y + x
static __global__ void k_gather_probe(cudaTextureObject_t tex, float* out)
{
    FetchTex f{tex, 1.0f};
    float v00, v10, v01, v11;
    f(AgentConsts{}, 1.0f, 1.0f, v00, v10, v01, v11);
    out[0] = v00; out[1] = v10; out[2] = v01; out[3] = v11;
}

// ---------------------------------------------------------------------------
// agents
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
k_init_agents(float4* __restrict__ agents, uint32_t* __restrict__ ids, uint64_t n, uint64_t first_id,
              uint64_t seed, float Wf, float Hf, float speed_min, float speed_max)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t id = first_id + i;
    float4 a;
    smd::agent_init(seed, id, Wf, Hf, speed_min, speed_max, a.x, a.y, a.z, a.w);
    agents[i] = a;
    ids[i] = (uint32_t)id;
}

static __global__ void __launch_bounds__(256)
k_reassign_speeds(float4* __restrict__ agents, const uint32_t* __restrict__ ids, uint64_t n,
                  uint64_t seed, float speed_min, float speed_max)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r = smd::rand01(seed, (uint64_t)ids[i], 3);
    agents[i].w = smd::add(speed_min, smd::mul(r, smd::sub(speed_max, speed_min)));
}

static __global__ void __launch_bounds__(256)
k_rescale_agents(float4* __restrict__ agents, uint64_t n, float fx, float fy)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = agents[i];
    a.x = smd::mul(a.x, fx);
    a.y = smd::mul(a.y, fy);
    agents[i] = a;
}

// field statistics accumulator (sm_trail_statistics): filled by the full-step trail pass as it writes the new
// field (k_trail_rows<CM != none>), or by k_trail_stats when the field was produced some other way
struct StatsAcc { double sum, sum_sq; unsigned long long nonzero; unsigned int max_bits; unsigned int pad; };

constexpr uint32_t kDeadAgent = 0xFFFFFFFFu;   // multi-GPU: slot whose agent migrated away (dropped by the next sort)

// How the agent kernel takes part in the multi-GPU exchange.
//   XM_SINGLE : one GPU, nothing to exchange
//   XM_NCCL   : deposits that land in a neighbour's rows go to the local ghost rows, leavers to local
//               fixed-size messages; exchange.cu ships both with NCCL send/recv
//   XM_P2P    : the kernel writes straight into the neighbour's HBM over NVLink (CUDA IPC peer
//               mappings): deposits into the neighbour's deposit field, leavers into the neighbour's
//               arrival buffer -- no staging, no collective; exchange.cu only runs two flag barriers
enum { XM_SINGLE = 0, XM_NCCL = 1, XM_P2P = 2 };

// Leaver records: [u64 count][u64 pad][float4 a[cap]][u32 id[cap]] -- a local message (XM_NCCL) or the
// neighbour's arrival buffer (XM_P2P).  All counters live on the device: a step needs no host round trip.
struct LeaverBufs {
    float4* send_a[2];                   // 0: towards rank-1 (up), 1: towards rank+1 (down)
    uint32_t* send_id[2];
    unsigned long long* send_count[2];   // record counters (message headers)
    unsigned long long* overflow;        // sticky error flag, checked by the host at the next sync point
    unsigned long long* left_count;      // XM_P2P: how many agents left this rank during the step
    const unsigned long long* slots_in_use;   // device-side count of the slots in use (the grid covers the host's upper bound of it)
    uint32_t cap;
    // boundary-first stepping on strips (exchange.cu: p2p_step_split): this launch covers the slots [0, split) and
    // [split + skip, slots in use) -- the agents near the strip edges at the last sort plus the arrivals since --
    // while the slots in between are stepped by an XM_SINGLE launch beside it.  skip == 0: every slot.
    unsigned long long split, skip;
    // XM_P2P: deposit fields of the two ring neighbours (pointer to THEIR owned row 0) and the row
    // count of the upper neighbour (a deposit on my row -k lands on its row rows_up - k)
    void* peer_dep[2];
    int32_t rows_up;
};

// One agent: compute.wgsl:57-145 on slot i, then the deposit (and, on strips, the hand-over of an agent whose
// new row belongs to a neighbour).  `deposits` points at owned row 0 of this rank's strip (global row
// c.row_base); ghost rows sit at negative / >= rows offsets.
// FLAGS: deposits are u8 "somebody deposited here" marks written with plain stores instead of u32
// counts bumped with RED atomics -- exact whenever dep >= 1 and the field is non-negative, because
// clamp(t + k*dep, 0, 1) == 1 for every k >= 1 (all shipped presets: dep = 1.0).
// Strips, rare path: the new cell belongs to a ring neighbour (or there is no deposit cell at all).
// (Not a real call: passing the parameter blocks by reference to a __noinline__ function makes every thread copy them to
// local memory at kernel entry -- measured 215 -> 292 us.)
template <int XM, class IdxT, int FLAGS>
__device__ __forceinline__ void agent_leaves_strip(float4* __restrict__ agents, uint32_t* __restrict__ ids, uint64_t i,
                                                const float4 a, const uint32_t id, const int32_t cx, const int32_t cy,
                                                void* __restrict__ deposits, const AgentConsts& c, const LeaverBufs& lv)
{
    int32_t lr;
    if (cx >= 0) {
        lr = smd::local_row(cy, c);                       // folded across the toroidal seam
        int32_t lrd = lr;
        void* base = deposits;
        bool ok = true;
        if (XM == XM_NCCL) ok = lrd >= -c.ghost && lrd < c.rows_local + c.ghost;
        if (XM == XM_P2P) {
            // write into the neighbour's field over NVLink
            if (lrd < 0) { base = lv.peer_dep[0]; lrd += lv.rows_up; ok = lrd >= 0; }
            else { base = lv.peer_dep[1]; lrd -= c.rows_local; ok = lrd < c.ghost; }
        }
        if (ok) {
            const IdxT off = (IdxT)lrd * (IdxT)c.W + (IdxT)cx;
            // (tiled flags: lrd is relative to owned row 0 of whichever strip `base` addresses -- every strip has the same geometry)
            if (FLAGS == 2) static_cast<uint8_t*>(base)[smd::flag_tile_offset<IdxT>((IdxT)cx, (IdxT)lrd, (IdxT)c.W, (IdxT)0)] = 1;
            else if (FLAGS) static_cast<uint8_t*>(base)[off] = 1;
            else if (XM == XM_P2P) atomicAdd_system(static_cast<uint32_t*>(base) + off, 1u);
            else atomicAdd(static_cast<uint32_t*>(base) + off, 1u);
        }
    } else {
        // no deposit cell (x == W / y == H rounding corner, non-finite state): owner row clamped like the host
        const int32_t oy = !(a.y >= 0.0f) ? 0 : (a.y >= c.Hf ? (int32_t)c.H - 1 : (int32_t)a.y);
        lr = smd::local_row(oy, c);
    }
    if (lr < 0 || lr >= c.rows_local) {
        // migration: hand the agent to the neighbour that owns its row
        // (no dynamic indexing of the parameter arrays: that would force a per-thread local copy)
        const bool up = lr < 0;
        unsigned long long* cnt = up ? lv.send_count[0] : lv.send_count[1];
        float4* sa = up ? lv.send_a[0] : lv.send_a[1];
        uint32_t* si = up ? lv.send_id[0] : lv.send_id[1];
        unsigned long long slot = (XM == XM_P2P) ? atomicAdd_system(cnt, 1ull) : atomicAdd(cnt, 1ull);
        if (slot < lv.cap) {
            sa[slot] = a;
            si[slot] = id;
            ids[i] = kDeadAgent;
            if (XM == XM_P2P) atomicAdd(lv.left_count, 1ull);
        } else {
            atomicExch(lv.overflow, 1ull);       // staging overflow: reported by the host
        }
    }
}

#ifndef SM_DEPOSIT_MATCH_ANY
#define SM_DEPOSIT_MATCH_ANY 0        // 1: __match_any_sync-aggregated counts.  Measured in round 2 (tools/r2/gpu_17.sh, deposit 0.3): the plain
                                      // RED is FASTER -- config 2: k_agents 181 us (RED) vs 214 us (match_any); config 3: 1597 vs 1628.  After the
                                      // cell sort a warp's 32 agents hit ~25 distinct cells, so a group rarely has more than one or two lanes, and
                                      // MATCH.ANY costs more issue slots than the fire-and-forget REDs it saves; L2 atomic throughput is not a limit
                                      // here (the sort's histogram, where a warp shares one or two keys, is where aggregation pays)
#endif
using smd::flag_tile_offset;     // trail_core.cuh: u8 deposit flags in 8 x 8-cell tiles (FLAGS == 2 / CM_FLAGS_TILED)

template <int XM, class IdxT, int FLAGS, bool AGG = (SM_DEPOSIT_MATCH_ANY != 0)>
__device__ __forceinline__ void finish_agent_slot(float4* __restrict__ agents, uint32_t* __restrict__ ids, uint64_t i,
                                                  const float4 a, const uint32_t id, const int32_t cx, const int32_t cy,
                                                  void* __restrict__ deposits, const AgentConsts& c, const LeaverBufs& lv)
{
    constexpr bool MULTI = XM != XM_SINGLE;
    agents[i] = a;
    // Fast path (every agent on one GPU, all but the strip-boundary agents otherwise): the new cell is
    // on a row this rank owns -> local deposit, no migration.  One subtract + one unsigned compare.
    const int32_t lr = cy - (int32_t)c.row_base;
    const bool interior = cx >= 0 && (!MULTI || (uint32_t)lr < (uint32_t)c.rows_local);
    if (interior) {
        // deposit: order-free (phase_split form of compute.wgsl:140)
        const IdxT off = (IdxT)lr * (IdxT)c.W + (IdxT)cx;
        if (FLAGS == 2) static_cast<uint8_t*>(deposits)[flag_tile_offset<IdxT>((IdxT)cx, (IdxT)lr, (IdxT)c.W, (IdxT)c.flag_wrap)] = 1;
        else if (FLAGS) static_cast<uint8_t*>(deposits)[off] = 1;
        else if (AGG) {
            // warp-aggregated count (fractional deposits): the lanes of the warp that hit the same cell elect a leader that adds
            // the size of the group -- after the cell sort the lanes of a warp share a few tiles, at one or more agents per cell
            const uint32_t active = __activemask();
            const uint32_t group = __match_any_sync(active, (unsigned long long)off);
            if ((threadIdx.x & 31u) == (uint32_t)(__ffs(group) - 1)) atomicAdd(static_cast<uint32_t*>(deposits) + off, (uint32_t)__popc(group));
        }
        else atomicAdd(static_cast<uint32_t*>(deposits) + off, 1u);
    }
    if (MULTI && !interior) agent_leaves_strip<XM, IdxT, FLAGS>(agents, ids, i, a, id, cx, cy, deposits, c, lv);
}

#ifndef SM_AGENTS_MIN_BLOCKS
#define SM_AGENTS_MIN_BLOCKS 5        // 5 x 256 threads: 48 registers -- room for three texture gathers + the next agent's state in flight
                                      // (measured on config 2, us per launch: 8 CTAs/32 regs 178, 6/40 171, 5/48 164, 4/52 179)
#endif
#ifndef SM_AGENTS_PER_THREAD
#define SM_AGENTS_PER_THREAD 4        // agents one thread steps in sequence (a CTA owns 256 * this many consecutive slots); 1: 181 us, 2: 175, 4: 164, 8: 164
#endif
constexpr int kAgentsPerThread = SM_AGENTS_PER_THREAD;
// Small populations take fewer agents per thread (a launch argument) so that the grid still fills the machine:
// 1 M agents are 977 CTAs at 4 per thread -- 1.3 waves of 148 x 5 -- but 3906 at 1 per thread.
static inline int agents_per_thread_for(uint64_t n, int num_sms)
{
    int apt = kAgentsPerThread;
    while (apt > 1 && n / (256ull * apt) < 4ull * 5ull * (uint64_t)num_sms) apt >>= 1;
    return apt;
}

__device__ __forceinline__ void load_agent_slot(const float4* agents, const uint32_t* ids, uint64_t i, float4& a, uint32_t& id)
{
    // asm volatile: the loads stay where they are written (ahead of the previous agent's arithmetic) --
    // left to the compiler a plain load is sunk to its first use, which exposes the full DRAM latency.
    // (evict-first hints on this stream were measured in rounds 1 and 2 and changed nothing: 1667 vs 1673 us at config 3)
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(agents + i));
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(id) : "l"(ids + i));
}

template <int XM, class IdxT, class FETCH, int FLAGS>
__device__ __forceinline__ void step_agent_slot(float4* __restrict__ agents, uint32_t* __restrict__ ids, uint64_t i,
                                                float4 a, uint32_t id, const FETCH& fetch, void* __restrict__ deposits,
                                                const AgentConsts& c, const LeaverBufs& lv)
{
    int32_t cx, cy;
    smd::agent_update(a.x, a.y, a.z, a.w, (int32_t)id, c, fetch, cx, cy);
    finish_agent_slot<XM, IdxT, FLAGS>(agents, ids, i, a, id, cx, cy, deposits, c, lv);
}

// (A software-pipelined form of this kernel -- agent j+1 sensed, its three gathers in flight, while agent j is moved -- was
// built and measured in round 2: long-scoreboard stalls fell from 6.3 to 4.2 per issue, but the kernel issues 73 % of the
// time either way, the pipelined loop executed 10 % more instructions and ran at 4 instead of 5 CTAs per SM: 172 -> 198 us
// at config 2.  The kernel is bound by instruction issue, not by gather latency; profiles/README.md.)
#ifndef SM_AGENTS_STRIP_MIN_BLOCKS
#define SM_AGENTS_STRIP_MIN_BLOCKS 5  // the strip instantiations (XM != XM_SINGLE) carry a little more state through the loop: at 48 registers
                                      // ptxas consumes the first gather before issuing the other two (two exposed latencies per agent); 4 CTAs /
                                      // 53 registers restores the schedule -- A/B in tools/r2/gpu_14.sh
#endif
template <int XM, class IdxT, class FETCH, int FLAGS>
static __global__ void __launch_bounds__(256, (XM == XM_SINGLE ? SM_AGENTS_MIN_BLOCKS : SM_AGENTS_STRIP_MIN_BLOCKS))
k_agents(float4* __restrict__ agents, uint32_t* __restrict__ ids, uint64_t n,
         const FETCH fetch, void* __restrict__ deposits, const AgentConsts c,
         const LeaverBufs lv, StatsAcc* __restrict__ stats_to_zero, const int agents_per_thread)
{
    constexpr bool MULTI = XM != XM_SINGLE;
    // the trail pass that follows this launch accumulates the field statistics of the step: start it from zero
    if (blockIdx.x == 0 && threadIdx.x == 0) *stats_to_zero = StatsAcc{0.0, 0.0, 0ull, 0u, 0u};
    // A CTA steps 256 * agents_per_thread consecutive slots, thread t taking slots t, t + 256, ... (coalesced).
    // The state of the next slot is requested before the current one is stepped, so only the first
    // load of a thread waits for DRAM; consecutive slots are neighbours in the cell-sorted order, so
    // their footprints also reuse this SM's L1.
    uint64_t i = (uint64_t)blockIdx.x * (256u * (uint32_t)agents_per_thread) + threadIdx.x;
    if (MULTI && lv.skip) {
        // boundary launch: a CTA's 256 * agents_per_thread consecutive indices map to slots on one side of the skipped
        // range (the host rounds `split` up to a whole CTA), so the slot arithmetic below is unchanged
        if (i >= lv.split) i += lv.skip;
    }
    if (i >= n) return;                      // n: the host's (upper bound of the) slots in use
    float4 a_next;
    uint32_t id_next;
    load_agent_slot(agents, ids, i, a_next, id_next);       // MULTI: possibly past the slots in use, always inside the allocation
    // MULTI: the host only knows an upper bound of the slots in use between two sorts (it grows by the migration
    // capacity every step); the exact count lives on the device -- CTAs past it leave without stepping anything.  Read
    // AFTER the first state load was issued: the two latencies overlap (they used to add up at the start of every CTA).
    // Every slot past the live ones holds kDeadAgent (kept so by the sort and by k_append_arrivals).
    if (MULTI) {
        n = min(n, (uint64_t)*lv.slots_in_use);
        if (i >= n) return;
    }
#pragma unroll 1
    for (int j = 0; j < agents_per_thread; ++j) {
        float4 a = a_next;
        const uint32_t id = id_next;
        const uint64_t i_next = i + 256u;
        const bool more = (j + 1 < agents_per_thread) && i_next < n;
        if (more) load_agent_slot(agents, ids, i_next, a_next, id_next);
        if (!MULTI || id != kDeadAgent)
            step_agent_slot<XM, IdxT, FETCH, FLAGS>(agents, ids, i, a, id, fetch, deposits, c, lv);
        if (!more) break;
        i = i_next;
    }
}

// (A fully unrolled form of this kernel -- four agents per thread, one base address per array with immediate offsets, 32-bit
// slot numbers, no register moves between iterations, the generic heading path as a real call -- was built and measured in
// round 2: 8.6 % fewer warp instructions (132.9 M vs 145.3 M at config 2), and the same time to within 1 %: the issue rate
// fell from 72 % to 63-66 % (longer gather stalls, instruction-cache misses of the four bodies).  profiles/README.md.)

// ---------------------------------------------------------------------------
// SM_FLAG_SEM_INPLACE: the reference's racy passes, statement for statement, on ONE live buffer
// ---------------------------------------------------------------------------
// Volatile loads: every sensor tap and every deposit read sees whatever the buffer holds at that moment (no L1-stale
// copies beyond what the hardware's own races produce) -- the point of this mode is the race, not its speed.
struct LdLive {
    __device__ __forceinline__ float operator()(const float* p) const { return *reinterpret_cast<const volatile float*>(p); }
};

// compute.wgsl:57-145 as written: sense the live map, move, then `trail[i] = clamp(trail[i] + dep, 0, 1)` NON-atomically.
template <class IdxT>
static __global__ void __launch_bounds__(256)
k_agents_inplace(float4* __restrict__ agents, const uint32_t* __restrict__ ids, uint64_t n, float* trail, const AgentConsts c, const float dep)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = agents[i];
    const uint32_t id = ids[i];
    int32_t cx, cy;
    const smd::FetchLinear<IdxT, LdLive> fetch{trail, (IdxT)c.W, (IdxT)0, LdLive()};
    smd::agent_update(a.x, a.y, a.z, a.w, (int32_t)id, c, fetch, cx, cy);
    agents[i] = a;
    if (cx >= 0) {
        volatile float* cell = trail + ((IdxT)cy * (IdxT)c.W + (IdxT)cx);
        const float cur = *cell;                                           // :140 read ...
        *cell = smd::clampf(smd::add(cur, dep), 0.0f, 1.0f);               // ... modify, write: the reference's race
    }
}

// compute.wgsl:148-161, in place (element-wise: no race)
static __global__ void __launch_bounds__(256)
k_decay_inplace(float* trail, uint64_t cells, const float decay_sub)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (uint64_t)gridDim.x * blockDim.x)
        trail[i] = smd::decay_cell(trail[i], decay_sub);
}

// compute.wgsl:164-195, in place: the nine taps come from the live buffer (some of them already rewritten by other
// invocations of this very dispatch), workgroups of 16 x 16 like the reference's
static __global__ void __launch_bounds__(256)
k_diffuse_inplace(float* trail, const uint32_t W, const uint32_t H, const TrailConsts tc)
{
    const uint32_t x = blockIdx.x * 16u + (threadIdx.x & 15u), y = blockIdx.y * 16u + (threadIdx.x >> 4);
    if (x >= W || y >= H) return;
    const volatile float* t = trail;
    float v[9];
    int j = 0;
    for (int dy = -1; dy <= 1; ++dy) {
        const size_t ry = (size_t)((y + H + dy) % H);
        for (int dx = -1; dx <= 1; ++dx) v[j++] = t[ry * W + (size_t)((x + W + dx) % W)];
    }
    trail[(size_t)y * W + x] = smd::box9_mix(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], tc);
}

// ---------------------------------------------------------------------------
// trail: merge deposits -> decay -> 3x3 toroidal mean -> mix, out of place
// ---------------------------------------------------------------------------
struct TrailGeom {
    uint32_t W;          // row length (cells); fast kernel requires W % 4 == 0
    uint32_t rows;       // owned rows
    uint32_t rows_per_chunk;
    uint32_t y_first, y_last;   // owned rows [y_first, y_last) this launch covers (the whole strip, or its interior when
                                // the exchange is overlapped with the interior pass) ...
    uint32_t chunks1;           // ... in blockIdx.y < chunks1; chunks past that cover a second band [y_first2, y_last2)
    uint32_t y_first2, y_last2; // (the two boundary bands of a strip in one launch)
    int wrap_y;          // 1: rows wrap toroidally inside the buffer (single GPU); 0: ghost rows
    cudaSurfaceObject_t surf;   // block-linear copy of the output for the TEX sampler (0 = none)
    int surf_row0;              // array row of owned row 0
    int surf_pairs;             // 1: surface rows are written in pairs as whole sectors (A/B switch SM_SURF_PAIRS)
};

__device__ __forceinline__ int64_t row_index(int64_t y, const TrailGeom& g)
{
    if (g.wrap_y) {
        if (y < 0) y += g.rows;
        else if (y >= (int64_t)g.rows) y -= g.rows;
    }
    return y;
}

// Deposit representation seen by the trail pass: none (diffusion only), u32 counts, u8 flags.
enum { CM_NONE = 0, CM_COUNTS = 1, CM_FLAGS = 2, CM_FLAGS_TILED = 3 };   // CM_FLAGS_TILED: flag_tile_offset() layout (k_trail_rows, k_display)
static_assert((int)CM_NONE == (int)GS_NONE && (int)CM_COUNTS == (int)GS_COUNTS && (int)CM_FLAGS == (int)GS_FLAGS, "deposit representation tags");

struct RawRow {
    float4 t;
    uint4 k;       // CM_COUNTS: four counts; CM_FLAGS: .x holds the four flag bytes
    float te;      // edge cell (left for lane 0, right for the last lane of a row segment)
    uint32_t ke;
};

template <int CM>
__device__ __forceinline__ float trail_cell(float t, uint32_t k, const TrailConsts& tc)
{
    if (CM == CM_COUNTS) t = smd::merge_deposit(t, k, tc.dep);
    if (CM == CM_FLAGS || CM == CM_FLAGS_TILED) t = k ? 1.0f : t;          // clamp(t + k*dep, 0, 1) with dep >= 1, t >= 0
    return smd::decay_cell(t, tc.decay_sub);
}

// Each thread owns 4 consecutive columns and walks down `rows_per_chunk` rows with a
// 3-row register window of decayed values; the horizontal neighbours come from the
// adjacent lanes by shuffle (warp-edge lanes fetch one extra cell).  Loads for
// UNROLL rows are issued before any of them is consumed.
//
// Requirements (checked by the host): W % 4 == 0 and (W / 4) % 32 != 1, so that a lane is never
// both the left edge (lane 0) and the right edge (last column group) of its warp.
// SURF: 0 = row-major output only, 1 = also the sampler's block-linear copy, 2 = also its shifted second copy
// u32 counts carry four more registers per row in flight than u8 flags: those instantiations run 6 CTAs per SM instead of 8
// rather than spill (fractional deposits only; every shipped preset deposits 1.0 and takes the flag path)
template <int CM, int SURF, int UNROLL, bool STATS>
static __global__ void __launch_bounds__(128, (CM == CM_COUNTS ? 6 : 8))
k_trail_rows(const float* __restrict__ tin, const void* __restrict__ cin_v,
             void* __restrict__ czero_v, float* __restrict__ tout,
             const TrailGeom g, const TrailConsts tc, StatsAcc* __restrict__ stats)
{
    // STATS instantiations also reduce the statistics of the field they write -- sum, sum of squares, non-zero
    // cells, maximum -- so that a host loop that reads the step's result every frame (bench.py's e2e leg,
    // sm_trail_statistics) does not pay a second sweep over the map (18 us at 4096^2).  The reduction costs this
    // kernel ~6 us there (its epilogue is amortised over only rows_per_chunk rows), so the engine selects it
    // only while the host is actually asking for statistics (sm_engine::stats_interest).
    double st_s = 0.0, st_s2 = 0.0;
    unsigned int st_nz = 0u;
    float st_m = 0.0f;
    const uint32_t grp = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t x0 = grp * 4u;
    const bool active = x0 < g.W;
    const uint32_t lane = threadIdx.x & 31u;
    const bool left_edge = active && (lane == 0u);
    const bool right_edge = active && (lane == 31u || x0 + 4u >= g.W);
    const bool edge = left_edge || right_edge;
    // column of the one extra cell an edge lane fetches
    const uint32_t xe = left_edge ? ((x0 == 0u) ? g.W - 1u : x0 - 1u) : ((x0 + 4u >= g.W) ? 0u : x0 + 4u);

    const bool band2 = blockIdx.y >= g.chunks1;
    const int y_begin = band2 ? (int)(g.y_first2 + (blockIdx.y - g.chunks1) * g.rows_per_chunk) : (int)(g.y_first + blockIdx.y * g.rows_per_chunk);
    const int y_end = min(y_begin + (int)g.rows_per_chunk, (int)(band2 ? g.y_last2 : g.y_last));
    const size_t W = g.W;
    // halo rows: only these two can wrap (single GPU) -- rows inside the chunk never do
    const int y_top = (g.wrap_y && y_begin == 0) ? (int)g.rows - 1 : y_begin - 1;
    const int y_bot = (g.wrap_y && y_end == (int)g.rows) ? 0 : y_end;

    const float* tp = tin + x0;
    const float* te = tin + xe;
    const uint32_t* cin = static_cast<const uint32_t*>(cin_v);
    const uint8_t* fin = static_cast<const uint8_t*>(cin_v);
    constexpr bool TILED = CM == CM_FLAGS_TILED;
    // tiled flags (flag_tile_offset): byte offset of a cell = row part (CTA-uniform per row) + column part (per thread, fixed)
    const uint32_t ftile0 = (x0 >> 3) << 6, fcole = ((xe >> 3) << 6) + (xe & 7u);
    const bool odd = (lane & 1u) != 0u;        // lanes 2j and 2j+1 own the two halves of one tile's columns
    auto frow_p = [&](int yp) { return ((ptrdiff_t)(yp >> 3) * (ptrdiff_t)W + (ptrdiff_t)(yp & 7)) * 8; };   // yp = y - 1 (mod H on one GPU)
    // CM_FLAGS_TILED launches consist of chunks of 4 or 8 rows that start on a multiple of their size: the four rows a batch
    // requests are one sector per tile, and the chunk's share of the OTHER flag buffer is zeroed with whole-sector stores
    // up front instead of four bytes per row
    // (a property of the whole launch, guaranteed by the host -- sm_engine::flags_tiled() -- because the bulk zeroing covers
    // rows y_begin+1 .. y_begin+n, which only tiles the field if every chunk does it)
    static_assert(!TILED || UNROLL == 4, "tiled flags: a batch is one sector of four rows");
    const int n_rows = y_end - y_begin;
    if (TILED && active) {
        // yp rows [y_begin, y_begin + n_rows) of this thread's tile: n_rows * 8 bytes, this lane's half of them
        uint8_t* z = static_cast<uint8_t*>(czero_v) + frow_p(y_begin) + ftile0 + (lane & 1u) * (uint32_t)(n_rows * 4);
        *reinterpret_cast<uint4*>(z) = make_uint4(0u, 0u, 0u, 0u);
        if (n_rows == 8) *reinterpret_cast<uint4*>(z + 16) = make_uint4(0u, 0u, 0u, 0u);
        // strips: y' = -1 (owned row 0) is no chunk's share -- on one GPU it is row H - 1's place, here a ghost tile row
        if (!g.wrap_y && y_begin == 0)
            *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(czero_v) + frow_p(-1) + ftile0 + (x0 & 7u)) = 0u;
    }

    // Addressing, measured (round 2, tools/r2/gpu_27.sh, profiles/r2_probe_trail_addressing.jsonl): ~100 of this kernel's 214
    // warp instructions per 128-cell row compute addresses (a 64-bit multiply per load, the halo row re-derived).  Deriving
    // every address from one running per-thread cell index instead cut the loop to 182 instructions per row and made the
    // pass SLOWER -- 4096^2: 49.7 -> 53.9 us at 64 registers (48 B spilled), 55.4 us at 7 CTAs per SM without spills; 8192^2:
    // 170 -> 184 us.  The kernel is bound by memory-level parallelism, not issue: ptxas hoists this form's loads further.
    auto issue = [&](int y, RawRow& r) {
        const ptrdiff_t off = (ptrdiff_t)y * (ptrdiff_t)W;
        r.t = make_float4(0.f, 0.f, 0.f, 0.f);
        r.te = 0.f;
        if (CM != CM_NONE) {
            r.k = make_uint4(0u, 0u, 0u, 0u);
            r.ke = 0u;
        }
        if (active) {
            r.t = __ldg(reinterpret_cast<const float4*>(tp + off));
            if (CM == CM_COUNTS) r.k = __ldg(reinterpret_cast<const uint4*>(cin + off + x0));
            if (CM == CM_FLAGS) r.k.x = __ldg(reinterpret_cast<const uint32_t*>(fin + off + x0));
        }
        if (edge) {
            r.te = __ldg(te + off);
            if (CM == CM_COUNTS) r.ke = __ldg(cin + off + xe);
            if (CM == CM_FLAGS) r.ke = __ldg(fin + off + xe);
        }
    };
    // d[0] = column x0-1, d[1..4] = own columns, d[5] = column x0+4 (all decayed)
    auto finish = [&](const RawRow& r, float (&d)[6]) {
        if (CM == CM_FLAGS || TILED) {
            d[1] = trail_cell<CM>(r.t.x, r.k.x & 0xffu, tc);
            d[2] = trail_cell<CM>(r.t.y, r.k.x & 0xff00u, tc);
            d[3] = trail_cell<CM>(r.t.z, r.k.x & 0xff0000u, tc);
            d[4] = trail_cell<CM>(r.t.w, r.k.x & 0xff000000u, tc);
        } else {
            d[1] = trail_cell<CM>(r.t.x, r.k.x, tc);
            d[2] = trail_cell<CM>(r.t.y, r.k.y, tc);
            d[3] = trail_cell<CM>(r.t.z, r.k.z, tc);
            d[4] = trail_cell<CM>(r.t.w, r.k.w, tc);
        }
        const float de = trail_cell<CM>(r.te, r.ke, tc);
        const float from_left = __shfl_up_sync(0xffffffffu, d[4], 1);
        const float from_right = __shfl_down_sync(0xffffffffu, d[1], 1);
        d[0] = left_edge ? de : from_left;
        d[5] = right_edge ? de : from_right;
    };

    float prev[6], cur[6], next[6];
    {
        RawRow r0, r1;
        issue(y_top, r0);
        issue(y_begin, r1);
        if (TILED) {
            // y' of y_top and y_begin: the last two rows of the sector above the chunk's first one (of the last sector of the
            // field when the chunk starts at row 0) -- its second half, fetched by both lanes of a pair
            const ptrdiff_t pb = frow_p(((g.wrap_y && y_begin == 0) ? (int)g.rows : y_begin) - 4) + 16;
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (active) q = __ldg(reinterpret_cast<const uint4*>(fin + pb + ftile0));
            r0.k.x = odd ? q.y : q.x;
            r1.k.x = odd ? q.w : q.z;
            if (edge) {
                r0.ke = __ldg(fin + pb + fcole);
                r1.ke = __ldg(fin + pb + 8 + fcole);
            }
        }
        finish(r0, prev);
        finish(r1, cur);
    }
    // Surface writes as WHOLE sectors.  A 32-byte sector of the block-linear array is 4 texels x 2 rows, so a warp-wide
    // SUST.128 along one row touches 32 half sectors -- and costs as much as streaming twice the bytes
    // (tools/microbench/surfcopy.cu, 8192^2: row-wise 118.6 us per pass, whole sectors 85.0, a plain linear copy 82.6).
    // Rows are therefore written in pairs: the lower half-warp hands its odd row to the upper one and takes the upper
    // half-warp's even row (one SHFL.BFLY per component), so that each of the two SUSTs of a pair covers both rows of 16
    // column groups.  Needs the pair to start on an even ARRAY row (warp-uniform; otherwise rows are written one by one).
    const bool lo_half = lane < 16u;
    const int xp = lo_half ? (int)x0 + 64 : (int)x0 - 64;           // first column of the partner lane (lane ^ 16)
    const bool p_active = (uint32_t)xp < g.W;
    const bool pair_mode = SURF && g.surf_pairs && (((y_begin + g.surf_row0) & 1) == 0);
    auto surf_pair = [&](const float4& oa, const float4& ob, int ya, bool va, bool vb) {
        const float4 send = lo_half ? ob : oa;
        float4 recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 16);
        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 16);
        recv.z = __shfl_xor_sync(0xffffffffu, send.z, 16);
        recv.w = __shfl_xor_sync(0xffffffffu, send.w, 16);
        const int row = ya + g.surf_row0;
        // first SUST: column groups of the lower half-warp, both rows; second: those of the upper half-warp
        const float4 d1 = lo_half ? oa : recv, d2 = lo_half ? recv : ob;
        const int x1 = lo_half ? (int)x0 : xp, x2 = lo_half ? xp : (int)x0;
        const int r1 = lo_half ? row : row + 1;
        const bool p1 = lo_half ? (active && va) : (p_active && vb), p2 = lo_half ? (p_active && va) : (active && vb);
        if (p1) surf2Dwrite(d1, g.surf, x1 * 4, r1);
        if (p2) surf2Dwrite(d2, g.surf, x2 * 4, r1);
    };
    static_assert(UNROLL % 2 == 0, "rows are written to the surface in pairs");
    for (int y = y_begin; y < y_end; y += UNROLL) {
        RawRow raw[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int yy = y + u + 1;
            issue(yy >= y_end ? y_bot : yy, raw[u]);     // past the chunk: harmless re-load of the halo row
        }
        if (TILED) {
            // rows y+1 .. y+4 are rows y' = y .. y+3 of the tiled flags (also across the wrap: y_bot = 0 <-> y' = H - 1): one
            // 32-byte sector per tile.  The even lane of a pair fetches its first half (two rows x 8 columns), the odd lane
            // the second; each keeps its own four columns and hands the neighbour's over.
            const ptrdiff_t gbase = frow_p(y);
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (active) q = __ldg(reinterpret_cast<const uint4*>(fin + gbase + ftile0 + ((lane & 1u) << 4)));
            const uint32_t ra = __shfl_xor_sync(0xffffffffu, odd ? q.x : q.y, 1);
            const uint32_t rb = __shfl_xor_sync(0xffffffffu, odd ? q.z : q.w, 1);
            raw[0].k.x = odd ? ra : q.x;
            raw[1].k.x = odd ? rb : q.z;
            raw[2 % UNROLL].k.x = odd ? q.y : ra;
            raw[3 % UNROLL].k.x = odd ? q.w : rb;
            if (edge) {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) raw[u].ke = __ldg(fin + gbase + u * 8 + fcole);
            }
        }
        float4 o_even = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            finish(raw[u], next);
            const bool row_ok = y + u < y_end;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok && active) {
                o.x = smd::box9_mix(prev[0], prev[1], prev[2], cur[0], cur[1], cur[2], next[0], next[1], next[2], tc);
                o.y = smd::box9_mix(prev[1], prev[2], prev[3], cur[1], cur[2], cur[3], next[1], next[2], next[3], tc);
                o.z = smd::box9_mix(prev[2], prev[3], prev[4], cur[2], cur[3], cur[4], next[2], next[3], next[4], tc);
                o.w = smd::box9_mix(prev[3], prev[4], prev[5], cur[3], cur[4], cur[5], next[3], next[4], next[5], tc);
                const size_t off = (size_t)(y + u) * W + x0;
                *reinterpret_cast<float4*>(tout + off) = o;
                if (STATS) {
                    // the four cells of a row are combined in f32 (pairwise: <= 3 roundings at magnitude <= 4, i.e. a
                    // relative error below 2e-7 per quad, random in sign) and only the quad enters the f64 accumulators:
                    // the FP64 pipe is narrow, 18 FP64 instructions per row cost this kernel 6 us, 4 cost < 1 us
                    const float q = (o.x + o.y) + (o.z + o.w);
                    const float q2 = fmaf(o.x, o.x, o.y * o.y) + fmaf(o.z, o.z, o.w * o.w);
                    st_s += (double)q;
                    st_s2 += (double)q2;
                    st_nz += min(__float_as_uint(o.x), 1u) + min(__float_as_uint(o.y), 1u) + min(__float_as_uint(o.z), 1u) +
                             min(__float_as_uint(o.w), 1u);          // the field is >= +0 here: non-zero <=> any bit set
                    st_m = fmaxf(fmaxf(st_m, fmaxf(o.x, o.y)), fmaxf(o.z, o.w));
                }
                if (CM == CM_COUNTS) *reinterpret_cast<uint4*>(static_cast<uint32_t*>(czero_v) + off) = make_uint4(0u, 0u, 0u, 0u);
                if (CM == CM_FLAGS) *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(czero_v) + off) = 0u;
                // keep the block-linear copy the agent kernel gathers from in step (4 B/cell extra)
                if (SURF && !pair_mode) surf2Dwrite(o, g.surf, (int)(x0 * 4u), y + u + g.surf_row0);
            }
            if (SURF && pair_mode) {                     // warp-uniform branch: every lane takes part in the exchange
                if ((u & 1) == 0) o_even = o;
                else {
                    const bool even_ok = y + u - 1 < y_end;
                    surf_pair(o_even, o, y + u - 1, even_ok, row_ok);
                }
            }
#pragma unroll
            for (int j = 0; j < 6; ++j) { prev[j] = cur[j]; cur[j] = next[j]; }
        }
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st_s += __shfl_down_sync(0xffffffffu, st_s, o);
            st_s2 += __shfl_down_sync(0xffffffffu, st_s2, o);
            st_nz += __shfl_down_sync(0xffffffffu, st_nz, o);
            st_m = fmaxf(st_m, __shfl_down_sync(0xffffffffu, st_m, o));
        }
        __shared__ double sh_s[4], sh_s2[4];
        __shared__ unsigned int sh_nz[4];
        __shared__ float sh_m[4];
        const uint32_t warp = threadIdx.x >> 5;
        if (lane == 0u) { sh_s[warp] = st_s; sh_s2[warp] = st_s2; sh_nz[warp] = st_nz; sh_m[warp] = st_m; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (uint32_t w = 1; w < (blockDim.x >> 5); ++w) { st_s += sh_s[w]; st_s2 += sh_s2[w]; st_nz += sh_nz[w]; st_m = fmaxf(st_m, sh_m[w]); }
            atomicAdd(&stats->sum, st_s);
            atomicAdd(&stats->sum_sq, st_s2);
            atomicAdd(&stats->nonzero, (unsigned long long)st_nz);
            atomicMax(&stats->max_bits, __float_as_uint(st_m));   // valid ordering: the field is non-negative after decay
        }
    }
}

// Any W, H >= 1 (ragged sizes, W % 4 != 0): one cell per thread, nine direct loads.
template <int CM>
static __global__ void __launch_bounds__(256)
k_trail_generic(const float* __restrict__ tin, const void* __restrict__ cin_v,
                void* __restrict__ czero_v, float* __restrict__ tout,
                const TrailGeom g, const TrailConsts tc, int64_t y_first)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t y = y_first + blockIdx.y;
    if (x >= (int64_t)g.W) return;
    float v[9];
    int j = 0;
    for (int dy = -1; dy <= 1; ++dy) {
        int64_t ry = y + dy;
        if (g.wrap_y) ry = (ry + g.rows) % (int64_t)g.rows;
        for (int dx = -1; dx <= 1; ++dx) {
            int64_t nx = (x + dx + g.W) % (int64_t)g.W;
            int64_t off = ry * (int64_t)g.W + nx;
            uint32_t k = 0u;
            if (CM == CM_COUNTS) k = __ldg(static_cast<const uint32_t*>(cin_v) + off);
            if (CM == CM_FLAGS) k = __ldg(static_cast<const uint8_t*>(cin_v) + off);
            v[j++] = trail_cell<CM>(__ldg(tin + off), k, tc);
        }
    }
    const int64_t off = y * (int64_t)g.W + x;
    const float o = smd::box9_mix(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], tc);
    tout[off] = o;
    if (g.surf) surf2Dwrite(o, g.surf, (int)(x * 4), (int)y + g.surf_row0);
    if (CM == CM_COUNTS) static_cast<uint32_t*>(czero_v)[off] = 0u;
    if (CM == CM_FLAGS) static_cast<uint8_t*>(czero_v)[off] = 0;
}

// ---------------------------------------------------------------------------
// EXTENSION: separable Gaussian blur (no reference semantics, "parity unpinned").
// Pass 1 (horizontal) fuses merge + decay and writes the decayed field D and the
// row-blurred field Hb; pass 2 (vertical) mixes D with the column blur of Hb.
// ---------------------------------------------------------------------------
// GaussConsts { R, w[17] }: gauss_stream.cuh

template <bool HAS_COUNTS>
static __global__ void __launch_bounds__(256)
k_gauss_h(const float* __restrict__ tin, const uint32_t* __restrict__ cin, uint32_t* __restrict__ czero,
          float* __restrict__ dec, float* __restrict__ hb, const TrailGeom g, const TrailConsts tc,
          const GaussConsts gc, int64_t y_first)
{
    extern __shared__ float srow[];                    // blockDim.x + 2R decayed values
    const int64_t y = y_first + blockIdx.y;
    const int64_t off = y * (int64_t)g.W;
    const int64_t xb = (int64_t)blockIdx.x * blockDim.x;
    const int R = gc.R;
    for (int j = threadIdx.x; j < (int)blockDim.x + 2 * R; j += blockDim.x) {
        int64_t x = xb + j - R;
        x = ((x % (int64_t)g.W) + g.W) % (int64_t)g.W;
        float t = __ldg(tin + off + x);
        if (HAS_COUNTS) t = smd::merge_deposit(t, __ldg(cin + off + x), tc.dep);
        srow[j] = smd::decay_cell(t, tc.decay_sub);
    }
    __syncthreads();
    const int64_t x = xb + threadIdx.x;
    if (x >= (int64_t)g.W) return;
    float acc = 0.0f;
    for (int d = 0; d <= 2 * R; ++d) acc = smd::fma(gc.w[d], srow[threadIdx.x + d], acc);
    dec[off + x] = srow[threadIdx.x + R];
    hb[off + x] = acc;
}

static __global__ void __launch_bounds__(256)
k_gauss_v(const float* __restrict__ dec, const float* __restrict__ hb, float* __restrict__ tout,
          uint32_t* __restrict__ czero, int has_counts, const TrailGeom g, const TrailConsts tc,
          const GaussConsts gc, int64_t y_first)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t y = y_first + blockIdx.y;
    if (x >= (int64_t)g.W) return;
    float acc = 0.0f;
    for (int d = -gc.R; d <= gc.R; ++d) {
        int64_t ry = y + d;
        if (g.wrap_y) ry = ((ry % (int64_t)g.rows) + g.rows) % (int64_t)g.rows;
        acc = smd::fma(gc.w[d + gc.R], __ldg(hb + ry * (int64_t)g.W + x), acc);
    }
    const int64_t off = y * (int64_t)g.W + x;
    tout[off] = smd::mixf_pre(dec[off], acc, tc.rate, tc.one_minus_rate);
    if (has_counts) czero[off] = 0u;
}

// Fused form of the two passes above (the default for maps of at least 160 x 64 cells): one kernel per pass,
// tiles with halos staged in shared memory, 8 B/cell of HBM traffic plus the halo re-reads instead of 20+.
//   phase 1  (TY + 2R) x (TX + 2R) cells: merge + decay -> D (shared)
//   phase 2  horizontal taps for all TY + 2R rows, 4 outputs per thread from a register window -> Hb (shared)
//   phase 3  vertical taps, one column and 16 rows per thread from a register window; mix with D; coalesced store
// The per-cell statement sequences (tap order d = -R..R, one FMA per tap starting from 0.0f, the final mix)
// are those of k_gauss_h / k_gauss_v and of the oracle, so the results are bit-identical to the two-pass form.
constexpr int kGaussTX = 128, kGaussTY = 32;

template <int R>
__host__ __device__ constexpr int gauss_ra() { return (R + 3) / 4 * 4; }                        // halo columns staged per side: R rounded up to whole float4s
template <int R>
__host__ __device__ constexpr int gauss_dcols() { return kGaussTX + 2 * gauss_ra<R>() + 4; }   // row stride of D (floats), 16-byte multiple, +4 against bank conflicts
template <int R>
__host__ __device__ constexpr size_t gauss_smem_bytes()
{
    return sizeof(float) * (size_t)(kGaussTY + 2 * R) * (size_t)(gauss_dcols<R>() + kGaussTX);
}

// Requires W % 4 == 0, W >= 160, rows >= 64 (the host falls back to the two-pass form otherwise).
template <int R, bool HAS_COUNTS>
static __global__ void __launch_bounds__(256)
k_gauss_fused(const float* __restrict__ tin, const uint32_t* __restrict__ cin, uint32_t* __restrict__ czero,
              float* __restrict__ tout, const TrailGeom g, const TrailConsts tc, const GaussConsts gc)
{
    constexpr int TX = kGaussTX, TY = kGaussTY, RW = TY + 2 * R, RA = gauss_ra<R>(), DC = gauss_dcols<R>();
    constexpr int DC4 = (TX + 2 * RA) / 4;          // float4 columns staged per row
    extern __shared__ __align__(16) float gsm[];
    float* D = gsm;                    // [RW][DC]   decayed cells, column c <-> map column x0 - RA + c
    float* Hb = gsm + RW * DC;         // [RW][TX]   row-blurred cells, row r <-> map row y0 - R + r
    const int W = (int)g.W, H = (int)g.rows;
    const int x0 = (int)blockIdx.x * TX, y0 = (int)blockIdx.y * TY;

    // ---- phase 1: every thread requests all its 16-byte pieces of the tile before it touches any of them ----
    {
        constexpr int N4 = RW * DC4, PER = (N4 + 255) / 256;
        float4 t4[PER];
        uint4 k4[HAS_COUNTS ? PER : 1];
        int64_t goff[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int e = (int)threadIdx.x + k * 256;
            goff[k] = -1;
            if (e < N4) {
                const int r = e / DC4, c4 = e - r * DC4;
                int gy = y0 - R + r;
                if (gy < 0) gy += H; else if (gy >= H) gy -= H;          // toroidal (rows >= TY + 2R: one fold is enough)
                int gx = x0 - RA + 4 * c4;
                if (gx < 0) gx += W; else if (gx >= W) gx -= W;          // W % 4 == 0: a float4 never straddles the seam
                goff[k] = (int64_t)gy * W + gx;
                t4[k] = __ldg(reinterpret_cast<const float4*>(tin + goff[k]));
                if (HAS_COUNTS) k4[k] = __ldg(reinterpret_cast<const uint4*>(cin + goff[k]));
            }
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int e = (int)threadIdx.x + k * 256;
            if (e < N4) {
                const int r = e / DC4, c4 = e - r * DC4;
                float4 t = t4[k];
                if (HAS_COUNTS) {
                    t.x = smd::merge_deposit(t.x, k4[k].x, tc.dep); t.y = smd::merge_deposit(t.y, k4[k].y, tc.dep);
                    t.z = smd::merge_deposit(t.z, k4[k].z, tc.dep); t.w = smd::merge_deposit(t.w, k4[k].w, tc.dep);
                    // cells this block owns: retire their counts in the buffer the next step deposits into
                    const bool own = r >= R && r < R + TY && y0 + (r - R) < H && 4 * c4 >= RA && 4 * c4 < RA + TX && x0 + (4 * c4 - RA) < W;
                    if (own) *reinterpret_cast<uint4*>(czero + goff[k]) = make_uint4(0u, 0u, 0u, 0u);
                }
                t.x = smd::decay_cell(t.x, tc.decay_sub); t.y = smd::decay_cell(t.y, tc.decay_sub);
                t.z = smd::decay_cell(t.z, tc.decay_sub); t.w = smd::decay_cell(t.w, tc.decay_sub);
                *reinterpret_cast<float4*>(D + r * DC + 4 * c4) = t;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: Hb[r][xs .. xs+3]; output j taps D columns xs + j + (RA - R) + d, d = 0 .. 2R ----
    for (int item = threadIdx.x; item < RW * (TX / 4); item += 256) {
        const int r = item / (TX / 4), xs = (item % (TX / 4)) * 4;
        constexpr int NV = (4 + RA + R + 3) / 4 * 4, SH = RA - R;
        float v[NV];
        const float4* src = reinterpret_cast<const float4*>(D + r * DC + xs);
#pragma unroll
        for (int q = 0; q < NV / 4; ++q) {
            const float4 f = src[q];
            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
        for (int d = 0; d <= 2 * R; ++d) {
            const float w = gc.w[d];
            a0 = smd::fma(w, v[SH + d], a0);
            a1 = smd::fma(w, v[SH + d + 1], a1);
            a2 = smd::fma(w, v[SH + d + 2], a2);
            a3 = smd::fma(w, v[SH + d + 3], a3);
        }
        *reinterpret_cast<float4*>(Hb + r * TX + xs) = make_float4(a0, a1, a2, a3);
    }
    __syncthreads();

    // ---- phase 3: column x, rows [16*half, 16*half + 16) ----
    {
        const int x = threadIdx.x & (TX - 1), half = threadIdx.x / TX;
        constexpr int ROWS = TY / 2;
        float v[ROWS + 2 * R];
#pragma unroll
        for (int k = 0; k < ROWS + 2 * R; ++k) v[k] = Hb[(half * ROWS + k) * TX + x];
        const int gx = x0 + x;
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            float acc = 0.0f;
#pragma unroll
            for (int d = 0; d <= 2 * R; ++d) acc = smd::fma(gc.w[d], v[j + d], acc);
            const int gy = y0 + half * ROWS + j;
            if (gx < W && gy < H)
                tout[(int64_t)gy * W + gx] = smd::mixf_pre(D[(half * ROWS + j + R) * DC + x + RA], acc, tc.rate, tc.one_minus_rate);
        }
    }
}

// Streaming form (gauss_stream.cuh): the device context of gauss_stream_cta and the kernel around it.
struct GsDevCtx {
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
    __device__ __forceinline__ int bx() const { return (int)blockIdx.x; }
    __device__ __forceinline__ int by() const { return (int)blockIdx.y; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    // predicated read-only loads as volatile asm: they are issued where they are written, back to back (no branch, no
    // constant-bank read between them); a piece that is not `valid` keeps its old register contents
    __device__ __forceinline__ void ld4(float4& v, const float* p, bool valid) const
    {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
                     : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) : "l"(p), "r"((int)valid));
    }
    __device__ __forceinline__ void ldu4(uint4& v, const uint32_t* p, bool valid) const
    {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                     : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "l"(p), "r"((int)valid));
    }
    __device__ __forceinline__ void ldu1(uint32_t& v, const uint32_t* p, bool valid) const
    {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p ld.global.nc.u32 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"((int)valid));
    }
    __device__ __forceinline__ void surf_write(float4 v, unsigned long long surf, int x, int y) const
    {
        surf2Dwrite(v, (cudaSurfaceObject_t)surf, x * 4, y);
    }
};

template <int R, int CM, bool SURF, bool PK = false>
static __global__ void __launch_bounds__(kGsNT, (gs_batch<R>() == 8 ? 4 : 3))     // rings: 34 KB (batches of 8 rows) / 67 KB (16)
k_gauss_stream(const GsArgs a, const TrailConsts tc, const GaussConsts gc)
{
    extern __shared__ __align__(16) float gs_smem[];
    gauss_stream_cta<R, CM, SURF, PK>(GsDevCtx{}, gs_smem, a, tc, gc);
}

// Register-streaming form (gauss_rows.cuh): device context and kernel.
struct GrDevCtx : GsDevCtx {
    __device__ __forceinline__ bool warp_may_exit() const { return true; }      // no barriers anywhere in the kernel
    // v[0 .. R-1] = the R cells to the left of this thread's four, v[R+4 .. 2R+3] = the R cells to their right: the
    // nearest four of each side live in the adjacent lane, the rest (R > 4) two lanes away
    template <int R>
    __device__ __forceinline__ void neighbours(const float4& t, float (&v)[2 * R + 4]) const
    {
        const float c[4] = {t.x, t.y, t.z, t.w};
        constexpr int FAR = R > 4 ? R - 4 : 0, NEAR = R - FAR;
#pragma unroll
        for (int i = 0; i < FAR; ++i) v[i] = __shfl_up_sync(0xffffffffu, c[4 - FAR + i], 2);
#pragma unroll
        for (int i = 0; i < NEAR; ++i) v[FAR + i] = __shfl_up_sync(0xffffffffu, c[4 - NEAR + i], 1);
#pragma unroll
        for (int i = 0; i < NEAR; ++i) v[R + 4 + i] = __shfl_down_sync(0xffffffffu, c[i], 1);
#pragma unroll
        for (int i = 0; i < FAR; ++i) v[R + 4 + NEAR + i] = __shfl_down_sync(0xffffffffu, c[i], 2);
    }
};

template <int R> constexpr int gr_min_blocks() { return R == 1 ? 8 : R == 2 ? 5 : R == 3 ? 4 : R == 4 ? 3 : 2; }

template <int R, int CM, bool SURF, int PK>
static __global__ void __launch_bounds__(kGrNT, gr_min_blocks<R>())
k_gauss_rows(const GsArgs a, const TrailConsts tc, const GaussConsts gc)
{
    gauss_rows_cta<R, CM, SURF, PK>(GrDevCtx{}, a, tc, gc);
}

// ---------------------------------------------------------------------------
// periodic cell sort (counting sort by tile key), carries the persistent index
// ---------------------------------------------------------------------------
struct TileGeom {
    uint32_t shift_x, shift_y;   // tile = 2^shift_x x 2^shift_y cells
    uint32_t tiles_x, tiles_y;   // over the strip
    uint32_t W;
    int64_t row_base;            // global row of local row 0
    uint32_t rows;               // owned rows
};

// (Sort keys that also group by heading sector, or number the tiles super-tile by super-tile, were measured in round 2 on
// configs [1] and [2] and changed nothing: gpurun_out/r2 probe, profiles/README.md.)
__device__ __forceinline__ uint32_t tile_key(float x, float y, const TileGeom& t)
{
    int32_t cx = (int32_t)x;                                   // x in [0, W] (W by rounding)
    int64_t cy = (int64_t)(int32_t)y - t.row_base;
    if (cx < 0) cx = 0;
    if (cx >= (int32_t)t.W) cx = t.W - 1;
    if (cy < 0) cy = 0;
    if (cy >= (int64_t)t.rows) cy = t.rows - 1;
    const uint32_t tx = (uint32_t)cx >> t.shift_x, ty = (uint32_t)cy >> t.shift_y;
    return ty * t.tiles_x + tx;
}

// Lanes of a warp that target the same tile are combined into one atomic (the agents are nearly
// sorted already, so a warp usually holds one or two distinct keys).  Returns this lane's slot
// offset within the group and the group size; the leader lane is the lowest lane of the group.
__device__ __forceinline__ void warp_group(uint32_t key, uint32_t& rank_in_group, uint32_t& group_size, bool& leader,
                                           uint32_t& group_mask)
{
    group_mask = __match_any_sync(0xffffffffu, key);
    const uint32_t lane = threadIdx.x & 31u;
    rank_in_group = __popc(group_mask & ((1u << lane) - 1u));
    group_size = __popc(group_mask);
    leader = rank_in_group == 0u;
}

static __global__ void __launch_bounds__(256)
k_tile_hist(const float4* __restrict__ agents, const uint32_t* __restrict__ ids, uint64_t n,
            uint32_t* __restrict__ hist, const TileGeom t)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key = 0xFFFFFFFFu;                                  // invalid / dead lanes group together
    if (i < n && ids[i] != kDeadAgent) {
        float4 a = agents[i];
        key = tile_key(a.x, a.y, t);
    }
    uint32_t r, sz, gm; bool leader;
    warp_group(key, r, sz, leader, gm);
    if (leader && key != 0xFFFFFFFFu) atomicAdd(hist + key, sz);
}

// exclusive scan of `n` u32 in three phases (block sums -> scan of sums -> add).
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;    // per thread
static __global__ void __launch_bounds__(kScanBlock)
k_scan_block(const uint32_t* in, uint32_t* out, uint32_t* __restrict__ block_sums, uint32_t n)
{
    __shared__ uint32_t warp_sums[kScanBlock / 32];
    const uint32_t base = (blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
    uint32_t v[kScanItems];
    uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        v[j] = (base + j < n) ? in[base + j] : 0u;
        tsum += v[j];
    }
    // inclusive scan of tsum across the block
    uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += up;
    }
    if (lane == 31u) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t ws = (lane < kScanBlock / 32) ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t up = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= (uint32_t)o) ws += up;
        }
        if (lane < kScanBlock / 32) warp_sums[lane] = ws;   // inclusive warp totals
    }
    __syncthreads();
    uint32_t excl = inc - tsum + (warp ? warp_sums[warp - 1] : 0u);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if (base + j < n) out[base + j] = excl;
        excl += v[j];
    }
    if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = excl;
}
static __global__ void __launch_bounds__(1024)
k_scan_sums(uint32_t* __restrict__ sums, uint32_t n)   // single block, in-place exclusive scan
{
    __shared__ uint32_t carry;
    __shared__ uint32_t warp_sums[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = (i < n) ? sums[i] : 0u;
        uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += up;
        }
        if (lane == 31u) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t ws = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t up = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= (uint32_t)o) ws += up;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        uint32_t excl = inc - v + (warp ? warp_sums[warp - 1] : 0u) + carry;
        if (i < n) sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
static __global__ void __launch_bounds__(kScanBlock)
k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ block_sums, uint32_t n)
{
    const uint32_t base = (blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
    const uint32_t add = block_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
        if (base + j < n) out[base + j] += add;
}

static __global__ void __launch_bounds__(256)
k_tile_scatter(const float4* __restrict__ agents, const uint32_t* __restrict__ ids, uint64_t n,
               uint32_t* __restrict__ cursor, float4* __restrict__ agents_out, uint32_t* __restrict__ ids_out,
               const TileGeom t)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key = 0xFFFFFFFFu;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t id = kDeadAgent;
    if (i < n) {
        id = ids[i];
        if (id != kDeadAgent) {
            a = agents[i];
            key = tile_key(a.x, a.y, t);
        }
    }
    uint32_t r, sz, gm; bool leader;
    warp_group(key, r, sz, leader, gm);
    uint32_t base = 0;
    if (leader && key != 0xFFFFFFFFu) base = atomicAdd(cursor + key, sz);
    base = __shfl_sync(gm, base, __ffs(gm) - 1);
    if (key != 0xFFFFFFFFu) {
        agents_out[base + r] = a;
        ids_out[base + r] = id;
    }
}

// ---------------------------------------------------------------------------
// field statistics (the per-step "result" a host harness reads back)
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
k_trail_stats(const float* __restrict__ t, uint64_t cells, StatsAcc* __restrict__ acc)
{
    // float4 streaming when the field is 16-byte aligned and a multiple of 4 cells; f32 partial sums per
    // thread would lose bits on 1e7+ cells, so every lane accumulates in f64 (the kernel stays HBM-bound)
    double s = 0.0, s2 = 0.0;
    unsigned long long nz = 0;
    float m = 0.0f;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    auto take = [&](float v) {
        s += (double)v;
        s2 += (double)v * (double)v;
        nz += (v != 0.0f);
        m = fmaxf(m, v);
    };
    if ((cells & 3ull) == 0 && (reinterpret_cast<uintptr_t>(t) & 15u) == 0) {
        const float4* t4 = reinterpret_cast<const float4*>(t);
        const uint64_t n4 = cells >> 2;
        uint64_t i = tid;
        for (; i + 3 * nthreads < n4; i += 4 * nthreads) {          // four independent 16-byte loads in flight
            const float4 a = __ldg(t4 + i), b = __ldg(t4 + i + nthreads), c = __ldg(t4 + i + 2 * nthreads),
                         d = __ldg(t4 + i + 3 * nthreads);
            take(a.x); take(a.y); take(a.z); take(a.w);
            take(b.x); take(b.y); take(b.z); take(b.w);
            take(c.x); take(c.y); take(c.z); take(c.w);
            take(d.x); take(d.y); take(d.z); take(d.w);
        }
        for (; i < n4; i += nthreads) {
            const float4 a = __ldg(t4 + i);
            take(a.x); take(a.y); take(a.z); take(a.w);
        }
    } else {
        for (uint64_t i = tid; i < cells; i += nthreads) take(t[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        nz += __shfl_down_sync(0xffffffffu, nz, o);
        m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    }
    __shared__ double sh_s[8], sh_s2[8];
    __shared__ unsigned long long sh_nz[8];
    __shared__ float sh_m[8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0u) { sh_s[warp] = s; sh_s2[warp] = s2; sh_nz[warp] = nz; sh_m[warp] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { s += sh_s[w]; s2 += sh_s2[w]; nz += sh_nz[w]; m = fmaxf(m, sh_m[w]); }
        atomicAdd(&acc->sum, s);
        atomicAdd(&acc->sum_sq, s2);
        atomicAdd(&acc->nonzero, nz);
        atomicMax(&acc->max_bits, __float_as_uint(m));   // valid ordering for m >= 0
    }
}

// ---------------------------------------------------------------------------
// display / colourise pass -- /root/reference/src/display.wgsl:29-86
// trail -> letter-boxed RGBA8 frame through the 768-byte planar LUT (256 R, 256 G, 256 B).
// One thread per 4 texels of a frame row: 4 trail gathers (a frame row maps to one trail row, so they
// coalesce whenever the frame is not magnified), 3 LUT bytes each from shared memory, one 16-byte store.
// 4 B/cell read + 4 B/texel written; HBM-bound.
// ---------------------------------------------------------------------------
struct DisplayGeom {
    uint32_t W, H;            // simulation size (global)
    uint32_t tw, th;          // frame size
    uint32_t row_base;        // global row of row 0 of `trail` / `dep` (strips; 0 on one GPU)
    uint32_t py_first;        // frame row of blockIdx.y == 0 (strips render the frame rows that show their map rows)
    float sim_w, sim_h;       // f32(W), f32(H)                    display.wgsl:48-49
    float scale, off_x, off_y;   // display.wgsl:58-69, computed once on the host with the same f32 operations
};

// What the frame shows.  The reference draws BETWEEN its decay and diffuse dispatches (main.rs:1184-1217): the field
// D = decay(merge(T_prev, deposits)).  This engine fuses decay + diffuse, so D is never stored -- but after a full step both
// of its inputs are still in HBM (the ping-pong field the pass read, and the step's deposit buffer, which only the NEXT
// trail pass retires), and the display pass recomputes D per texel with the very statements of the trail pass
// (trail_cell<CM>).  cm == CM_NONE: show `trail` as it is (after an upload, a clear, a diffusion-only pass ...).
struct DisplaySrc {
    const float* trail;       // cm == CM_NONE: the current field; else the field the last step started from
    const void* dep;          // the last step's deposits: u32 counts (CM_COUNTS) or u8 flags (CM_FLAGS, CM_FLAGS_TILED)
    int cm;
    int flag_wrap;            // CM_FLAGS_TILED: as AgentConsts::flag_wrap
    TrailConsts tc;           // deposit amount / decay of that step
};

__device__ __forceinline__ uint32_t display_texel(const DisplaySrc& src, const uint8_t* lut, const DisplayGeom& g,
                                                  uint32_t px, float fy, bool row_inside)
{
    const float fx = __fdiv_rn(smd::sub((float)px, g.off_x), g.scale);               // :72
    if (!(row_inside && fx >= 0.0f && fx < g.sim_w)) return 0xFF000000u;             // :83-85 black, alpha 1
    const int32_t x = (int32_t)fx, y = (int32_t)fy;                                  // :77-78
    const size_t idx = (size_t)(y - (int32_t)g.row_base) * g.W + x;
    float t = __ldg(src.trail + idx);                                                // :79
    if (src.cm == CM_COUNTS) t = trail_cell<CM_COUNTS>(t, __ldg(static_cast<const uint32_t*>(src.dep) + idx), src.tc);
    else if (src.cm == CM_FLAGS) t = trail_cell<CM_FLAGS>(t, __ldg(static_cast<const uint8_t*>(src.dep) + idx), src.tc);
    else if (src.cm == CM_FLAGS_TILED)
        t = trail_cell<CM_FLAGS>(t, __ldg(static_cast<const uint8_t*>(src.dep) +
                                          flag_tile_offset<ptrdiff_t>((ptrdiff_t)x, (ptrdiff_t)y - (ptrdiff_t)g.row_base, (ptrdiff_t)g.W, (ptrdiff_t)src.flag_wrap)), src.tc);
    const float inten = smd::clampf(smd::clampf(t, 0.0f, 1.0f), 0.0f, 1.0f);         // :80 and :31
    const uint32_t li = (uint32_t)smd::mul(inten, 255.0f);                           // :34
    // :37-39 f32(lut)/255 stored as rgba8unorm = round(v * 255) = the LUT byte itself for every byte value
    // (tests/test_display.py proves it for all 256 through the oracle's literal conversion)
    return (uint32_t)lut[li] | ((uint32_t)lut[li + 256] << 8) | ((uint32_t)lut[li + 512] << 16) | 0xFF000000u;
}

static __global__ void __launch_bounds__(256)
k_display(const DisplaySrc src, const uint8_t* __restrict__ lut768, uint32_t* __restrict__ rgba, const DisplayGeom g)
{
    __shared__ uint8_t lut[768];
    for (uint32_t i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = lut768[i];
    __syncthreads();
    const uint32_t py = g.py_first + blockIdx.y;
    const float fy = __fdiv_rn(smd::sub((float)py, g.off_y), g.scale);               // :73
    const bool row_inside = fy >= 0.0f && fy < g.sim_h;                              // :76
    const uint32_t px0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (px0 >= g.tw) return;
    uint32_t* row = rgba + (size_t)blockIdx.y * g.tw;
    if (px0 + 4u <= g.tw && (g.tw & 3u) == 0u) {
        uint4 o;
        o.x = display_texel(src, lut, g, px0, fy, row_inside);
        o.y = display_texel(src, lut, g, px0 + 1u, fy, row_inside);
        o.z = display_texel(src, lut, g, px0 + 2u, fy, row_inside);
        o.w = display_texel(src, lut, g, px0 + 3u, fy, row_inside);
        *reinterpret_cast<uint4*>(row + px0) = o;
    } else {
        for (uint32_t px = px0; px < g.tw && px < px0 + 4u; ++px) row[px] = display_texel(src, lut, g, px, fy, row_inside);
    }
}

// ---------------------------------------------------------------------------
// arithmetic-spec probes (sm_test_math)
// ---------------------------------------------------------------------------
static __global__ void k_test_math(int what, const float* a, const float* b, const int32_t* iv,
                            float* o0, float* o1, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (what) {
    case 0: { float s, c; smd::sincos(a[i], s, c); o0[i] = s; o1[i] = c; break; }
    case 1: { float bb = b[i]; o0[i] = smd::fmod_exact(a[i], bb, __frcp_rn(bb)); break; }
    case 2: o0[i] = smd::div9(a[i]); break;
    case 3: o0[i] = smd::hash01(iv[i], a[i], b[i]); break;
    default: break;
    }
}

}  // namespace smk
