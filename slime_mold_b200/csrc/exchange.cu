// exchange.cu -- multi-GPU strip exchange: one process per GPU, ring neighbours only, NCCL
// point-to-point over NVLink on the engine's stream.
//
// The map is cut into `world` horizontal strips; rank r owns rows
// [r*H/G, (r+1)*H/G) and the agents standing on them.  Per step (protocol validated on
// the CPU by tests/strip_model.py against the single-domain oracle):
//
//   k_agents<true>     senses from trail rows own +- g (ghost rows), counts deposits into own +- m,
//                      stages agents whose new row belongs to a neighbour (leavers)
//   exchange_counts()  to each neighbour: the m ghost count rows on its side + my own boundary row
//                      (contiguous (m+1) x W u32)
//   k_trail_rows       merge -> decay -> 3x3 mean on the owned rows (ghost row +-1 now complete)
//   exchange_trail_ghosts() + migrate_agents(): my new top / bottom g trail rows -> the neighbours'
//                      ghost rows; the two fixed-size leaver messages -> the neighbours, whose
//                      k_append_arrivals adds them behind a device-side cursor.  No host round trip per
//                      step: slot counts live on the device and are read back only when the agents are
//                      sorted (every 16 steps), downloaded or counted.
//
// g = ceil(sensor_distance) + 3,  m = ceil(speed_max * 0.016) + 1.
// The toroidal seam (strip 0 <-> strip G-1) carries diffusion and motion; sensing is not
// toroidal (compute.wgsl:14-16), which the kernel handles by its global bounds check.
// With world == 2 both neighbours are the same peer: sends are issued [to up, to down] and
// receives [from down, from up], which is the order NCCL matches them in.
#include "engine.h"

#include <nccl.h>      // types and prototypes only: the library is bound at run time (see nccl_api below)
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

// NCCL is resolved with dlopen/dlsym instead of a link-time dependency: a process that also uses
// torch must end up with ONE libnccl.so.2 (torch bundles 2.28, the system has 2.27 -- loading the
// older one first breaks `import torch`).  Order: $SM_NCCL_LIB, an already-loaded libnccl.so.2
// (RTLD_NOLOAD: torch's, if torch was imported first), then the default search path.
namespace {
struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    void* handle = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

int nccl_load()
{
    if (g_nccl.ok) return SM_OK;
    void* h = nullptr;
    const char* env = getenv("SM_NCCL_LIB");
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return sm_fail(SM_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.handle = h;
#define SM_SYM(name)                                                                    \
    g_nccl.name = reinterpret_cast<decltype(g_nccl.name)>(dlsym(h, "nccl" #name));      \
    if (!g_nccl.name) return sm_fail(SM_ERR_NCCL, "libnccl.so.2 lacks nccl" #name)
    SM_SYM(GetUniqueId); SM_SYM(CommInitRank); SM_SYM(CommDestroy); SM_SYM(Send); SM_SYM(Recv); SM_SYM(AllGather);
    SM_SYM(GroupStart); SM_SYM(GroupEnd); SM_SYM(GetErrorString);
#undef SM_SYM
    g_nccl.ok = true;
    return SM_OK;
}
}  // namespace
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclAllGather g_nccl.AllGather
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclGetErrorString g_nccl.GetErrorString

#define SM_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return sm_fail(err__ == cudaErrorMemoryAllocation ? SM_ERR_OOM : SM_ERR_CUDA,     \
                           "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                           cudaGetErrorString(err__));                                         \
    } while (0)
#define SM_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t res__ = (call);                                                           \
        if (res__ != ncclSuccess)                                                              \
            return sm_fail(SM_ERR_NCCL, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,   \
                           ncclGetErrorString(res__));                                         \
    } while (0)
#define SM_TRY(expr)                    \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != SM_OK) return rc__; \
    } while (0)

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

namespace smk {

// dst[i] += src[i] over a contiguous range of deposit counts
static __global__ void __launch_bounds__(256)
k_counts_add(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

// dst[i] |= src[i] over a contiguous range of deposit flags
static __global__ void __launch_bounds__(256)
k_flags_or(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] |= src[i];
}

// Appends the agents of the two received messages behind the device-side cursor counters[0].
static __global__ void __launch_bounds__(256)
k_append_arrivals(const uint8_t* __restrict__ msg_from_down, const uint8_t* __restrict__ msg_from_up, uint32_t cap,
                  float4* __restrict__ agents, uint32_t* __restrict__ ids, unsigned long long* __restrict__ counters,
                  uint64_t cap_local)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2u * cap) return;
    const uint32_t which = j / cap, k = j % cap;
    const unsigned long long c0 = min(*reinterpret_cast<const unsigned long long*>(msg_from_down), (unsigned long long)cap);
    const unsigned long long c1 = min(*reinterpret_cast<const unsigned long long*>(msg_from_up), (unsigned long long)cap);
    const uint8_t* msg = which ? msg_from_up : msg_from_down;
    if (k >= (which ? c1 : c0)) return;
    const unsigned long long pos = counters[0] + (which ? c0 : 0ull) + k;
    if (pos >= cap_local) { atomicExch(counters + 2, 2ull); return; }
    agents[pos] = reinterpret_cast<const float4*>(msg + 16)[k];
    ids[pos] = reinterpret_cast<const uint32_t*>(msg + 16 + (size_t)cap * sizeof(float4))[k];
}

// One thread: advance the cursors, retire this step's outgoing message headers.
static __global__ void k_bump_counters(unsigned long long* counters, const uint8_t* msg_from_down, const uint8_t* msg_from_up,
                                       unsigned long long* send_up, unsigned long long* send_down, uint32_t cap,
                                       uint64_t cap_local)
{
    const unsigned long long c0 = min(*reinterpret_cast<const unsigned long long*>(msg_from_down), (unsigned long long)cap);
    const unsigned long long c1 = min(*reinterpret_cast<const unsigned long long*>(msg_from_up), (unsigned long long)cap);
    const unsigned long long left = min(*send_up, (unsigned long long)cap) + min(*send_down, (unsigned long long)cap);
    unsigned long long arrived = c0 + c1;
    if (counters[0] + arrived > cap_local) arrived = cap_local - counters[0];
    counters[0] += arrived;
    counters[1] = counters[1] + arrived - left;
    *send_up = 0ull;
    *send_down = 0ull;
}

// ---- XM_P2P: flag barrier between ring neighbours --------------------------------------------
// Window layout (one per rank, IPC-mapped by both neighbours):
//   [0]  u32 seq written by my UP neighbour      [4]  u32 seq written by my DOWN neighbour
//   [64 ...] arrival buffer filled by up, then arrival buffer filled by down
// A barrier = "tell both neighbours I reached sequence number s, wait until both told me the same".
// Everything this rank wrote into a neighbour's memory earlier in the stream is fenced before the flag.
// The wait gives up after ~20 s and raises the sticky error flag instead of hanging the GPU.
// Fences: what the neighbours read after the barrier was written by EARLIER kernels of this stream (the agent
// pass, the ghost-row push), and a kernel boundary already orders those writes before anything this kernel
// stores.  The barrier itself therefore needs one release fence ahead of the flag stores (for the few words
// this very thread wrote) and one acquire fence behind the wait.  g_fence_mode (SM_BARRIER_FENCE): 0 = three
// sequentially-consistent system fences (the first version: each one costs ~10 us while the interior trail
// pass is streaming beside it), 1 = acq_rel system fences (default), 2 = device-scope fences only (A/B).
__device__ int g_fence_mode = 1;

__device__ __forceinline__ void barrier_fence()
{
    if (g_fence_mode == 0) __threadfence_system();
    else if (g_fence_mode == 1) asm volatile("fence.acq_rel.sys;" ::: "memory");
    else __threadfence();
}

__device__ __forceinline__ void ring_barrier(volatile uint32_t* mine, volatile uint32_t* up_slot, volatile uint32_t* down_slot,
                                             uint32_t seq, unsigned long long* err)
{
    barrier_fence();
    *up_slot = seq;        // I am the DOWN neighbour of `up`: its slot [4]
    *down_slot = seq;      // I am the UP neighbour of `down`: its slot [0]
    if (g_fence_mode == 0) __threadfence_system();
    const long long t0 = clock64();
    while ((int32_t)(mine[0] - seq) < 0 || (int32_t)(mine[1] - seq) < 0) {
        __nanosleep(100);
        if (clock64() - t0 > 40000000000ll) { atomicExch(err, 3ull); break; }   // ~20 s: the neighbour is gone
    }
    barrier_fence();
}

// The waiting half of ring_barrier alone (CTAs that did not announce).
__device__ __forceinline__ void ring_wait(volatile uint32_t* mine, uint32_t seq, unsigned long long* err)
{
    const long long t0 = clock64();
    while ((int32_t)(mine[0] - seq) < 0 || (int32_t)(mine[1] - seq) < 0) {
        __nanosleep(100);
        if (clock64() - t0 > 40000000000ll) { atomicExch(err, 3ull); break; }
    }
    barrier_fence();
}

// Barrier 1 (all neighbours finished their agent pass: their deposits and leavers have landed here),
// then pull the one deposit row of each neighbour that the 3x3 blur of my boundary rows needs.
// Small CTAs on purpose (128 threads, a few of them): this kernel runs beside the interior trail pass, whose
// 128-thread CTAs refill every slot an SM frees -- a 1024-thread CTA found no SM with that many free threads
// until the interior pass had drained (measured: 41 us of queueing per step).  CTA 0 announces this rank;
// every CTA waits for both neighbours on its own (no CTA depends on another one of this grid).
template <class T>
static __global__ void __launch_bounds__(128)
k_barrier_pull(uint32_t* window, uint32_t* up_window, uint32_t* down_window, uint32_t seq, unsigned long long* err,
               T* my_row_above, const T* up_last_row, T* my_row_below, const T* down_first_row, uint32_t W)
{
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) ring_barrier(window, up_window + 1, down_window + 0, seq, err);
        else ring_wait(window, seq, err);
    }
    __syncthreads();
    // Remote reads over NVLink cost a round trip each (a few us beside a streaming kernel): the grid is sized so
    // that a thread has ONE 16-byte read per row in flight, not a loop of dependent narrow ones.
    const size_t row_bytes = (size_t)W * sizeof(T);
    const bool wide = (row_bytes % 16 == 0) &&
                      ((reinterpret_cast<uintptr_t>(my_row_above) | reinterpret_cast<uintptr_t>(up_last_row) |
                        reinterpret_cast<uintptr_t>(my_row_below) | reinterpret_cast<uintptr_t>(down_first_row)) % 16 == 0);
    if (wide) {
        const size_t n16 = row_bytes / 16;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
            const uint4 a = reinterpret_cast<const uint4*>(up_last_row)[i];
            const uint4 b = reinterpret_cast<const uint4*>(down_first_row)[i];
            reinterpret_cast<uint4*>(my_row_above)[i] = a;
            reinterpret_cast<uint4*>(my_row_below)[i] = b;
        }
    } else {
        for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
            my_row_above[x] = up_last_row[x];
            my_row_below[x] = down_first_row[x];
        }
    }
}

// The same for u8 flags stored in 8 x 8-cell tiles (kernels.cuh flag_tile_offset): a row is W / 8 pieces of 8 bytes, 64 bytes apart.
static __global__ void __launch_bounds__(128)
k_barrier_pull_tiled(uint32_t* window, uint32_t* up_window, uint32_t* down_window, uint32_t seq, unsigned long long* err,
                     uint8_t* my_row_above, const uint8_t* up_last_row, uint8_t* my_row_below, const uint8_t* down_first_row, uint32_t pieces)
{
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) ring_barrier(window, up_window + 1, down_window + 0, seq, err);
        else ring_wait(window, seq, err);
    }
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pieces; i += gridDim.x * blockDim.x) {
        const size_t o = (size_t)i * 64u;
        const uint2 a = *reinterpret_cast<const uint2*>(up_last_row + o);
        const uint2 b = *reinterpret_cast<const uint2*>(down_first_row + o);
        *reinterpret_cast<uint2*>(my_row_above + o) = a;
        *reinterpret_cast<uint2*>(my_row_below + o) = b;
    }
}

static __global__ void k_barrier_only(uint32_t* window, uint32_t* up_window, uint32_t* down_window, uint32_t seq,
                                      unsigned long long* err)
{
    ring_barrier(window, up_window + 1, down_window + 0, seq, err);
}

// Push my new top / bottom g trail rows into the neighbours' ghost rows (float4 stores over NVLink).
static __global__ void __launch_bounds__(256)
k_push_rows(const float4* __restrict__ my_top, float4* __restrict__ up_bottom_ghost,
            const float4* __restrict__ my_bottom, float4* __restrict__ down_top_ghost, uint64_t n4)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        up_bottom_ghost[i] = my_top[i];
        down_top_ghost[i] = my_bottom[i];
    }
}
static __global__ void __launch_bounds__(256)
k_push_rows_scalar(const float* __restrict__ my_top, float* __restrict__ up_bottom_ghost,
                   const float* __restrict__ my_bottom, float* __restrict__ down_top_ghost, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        up_bottom_ghost[i] = my_top[i];
        down_top_ghost[i] = my_bottom[i];
    }
}

// Barrier 2: bookkeeping of this step's migration (one thread), then the barrier that lets every
// rank start its next agent pass (ghost rows pushed, arrival buffers drained and reset).
static __global__ void k_bump_barrier(unsigned long long* counters, uint8_t* arr_from_down, uint8_t* arr_from_up, uint32_t cap,
                                      uint64_t cap_local, uint32_t* window, uint32_t* up_window, uint32_t* down_window,
                                      uint32_t seq)
{
    unsigned long long* c0p = reinterpret_cast<unsigned long long*>(arr_from_down);
    unsigned long long* c1p = reinterpret_cast<unsigned long long*>(arr_from_up);
    const unsigned long long c0 = min(*c0p, (unsigned long long)cap), c1 = min(*c1p, (unsigned long long)cap);
    unsigned long long arrived = c0 + c1;
    if (counters[0] + arrived > cap_local) arrived = cap_local - counters[0];
    counters[0] += arrived;
    counters[1] = counters[1] + arrived - counters[3];
    counters[3] = 0ull;
    *c0p = 0ull;
    *c1p = 0ull;
    ring_barrier(window, up_window + 1, down_window + 0, seq, counters + 2);
}

// Seeded start-up fill restricted to one strip: every rank walks all agent indices and keeps
// the agents whose row it owns (same counter-based generator as k_init_agents).
static __global__ void __launch_bounds__(256)
k_init_agents_strip(float4* __restrict__ agents, uint32_t* __restrict__ ids, uint64_t n_global, uint64_t seed,
                    float Wf, float Hf, uint32_t H, float speed_min, float speed_max, uint32_t row0, uint32_t rows,
                    unsigned long long* __restrict__ counter, uint64_t cap)
{
    uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_global) return;
    float4 a;
    smd::agent_init(seed, id, Wf, Hf, speed_min, speed_max, a.x, a.y, a.z, a.w);
    uint32_t r = !(a.y >= 0.0f) ? 0u : (a.y >= Hf ? H - 1u : (uint32_t)a.y);
    if (r < row0 || r >= row0 + rows) return;
    unsigned long long slot = atomicAdd(counter, 1ull);
    if (slot < cap) {
        agents[slot] = a;
        ids[slot] = (uint32_t)id;
    }
}

}  // namespace smk

// ---------------------------------------------------------------------------
static int halo_depths(const sm_engine* e, uint32_t* g, uint32_t* m)
{
    float sd = fabsf(e->params.agent_sensor_distance);
    float mv = fabsf(e->params.agent_speed_max) * 0.016f;
    if (!(sd < 60000.0f) || !(mv < 60000.0f)) return sm_fail(SM_ERR_BAD_ARG, "sensor distance / speed too large for strips");
    *g = (uint32_t)ceilf(sd) + 3u;
    *m = (uint32_t)ceilf(mv) + 1u;
    if (e->cfg.flags & SM_FLAG_GAUSSIAN_BLUR) {          // extension: the blur reads R rows of the neighbours (R <= 8)
        const float br = e->params.blur_radius;
        const uint32_t R = (br >= 1.0f && br <= 8.5f) ? (uint32_t)lroundf(br) : 8u;
        if (*g < R) *g = R;
    }
    if (*g > e->ghost || *m > e->ghost || *m + 1 > e->rows)
        return sm_fail(SM_ERR_BAD_ARG, "strip of %u rows with %u ghost rows is too thin: sensing needs %u, motion needs %u",
                       e->rows, e->ghost, *g, *m);
    return SM_OK;
}

extern "C" int sm_comm_unique_id(uint8_t id[SM_COMM_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == SM_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id) return sm_fail(SM_ERR_BAD_ARG, "null id");
    SM_TRY(nccl_load());
    ncclUniqueId u;
    SM_NCCL(ncclGetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return SM_OK;
}

extern "C" int sm_comm_init(sm_engine* e, const uint8_t id[SM_COMM_ID_BYTES])
{
    if (!e || !id) return sm_fail(SM_ERR_BAD_ARG, "null argument");
    if (e->world < 2) return sm_fail(SM_ERR_STATE, "sm_comm_init on a single-GPU engine (world_size == 1)");
    if (e->comm_ready) return sm_fail(SM_ERR_STATE, "communicator already initialised");
    SM_CUDA(cudaSetDevice(e->device));
    SM_TRY(nccl_load());
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm;
    SM_NCCL(ncclCommInitRank(&comm, e->world, u, e->rank));
    e->comm = comm;

    // staging: received count rows (2 x (ghost + 1) rows), migration messages, device counters
    e->counts_xchg_rows = (uint64_t)e->ghost + 1;
    SM_CUDA(cudaMalloc(&e->counts_xchg, 2 * e->counts_xchg_rows * e->W * sizeof(uint32_t)));
    // Leavers per step and direction ~ density * W * |v_y| * dt (a few thousand at W = 32768); the whole
    // fixed-size message is sent every step, so keep it small: 64 Ki agents (1.3 MB, ~2 us on NVLink),
    // sm_tuning.migrate_capacity overrides.  Overflow is detected on the device and reported at the next sync point.
    uint64_t cap = 65536;
    if (e->tuning.migrate_capacity) cap = std::max<uint64_t>(1024, e->tuning.migrate_capacity);
    cap = (cap + 3) & ~3ull;
    e->mig_cap = cap;
    e->mig_bytes = 16 + cap * (sizeof(float4) + sizeof(uint32_t));
    for (int d = 0; d < 2; ++d) {
        SM_CUDA(cudaMalloc(&e->mig[d].send, e->mig_bytes));
        SM_CUDA(cudaMalloc(&e->mig[d].recv, e->mig_bytes));
        SM_CUDA(cudaMemset(e->mig[d].send, 0, 16));
        SM_CUDA(cudaMemset(e->mig[d].recv, 0, 16));
    }
    SM_CUDA(cudaMalloc(&e->dev_counters, 8 * sizeof(unsigned long long)));
    SM_CUDA(cudaMemset(e->dev_counters, 0, 8 * sizeof(unsigned long long)));
    SM_CUDA(cudaMallocHost(&e->host_counters, 8 * sizeof(unsigned long long)));
    SM_TRY(e->setup_p2p());
    e->comm_ready = true;
    e->ghost_stale = true;
    SM_TRY(e->mark_tail_dead());
    SM_TRY(e->push_counters());          // agents may have been uploaded before the communicator existed
    return SM_OK;
}

// PROFILING ONLY (sm_tuning.debug_single_rank_strip): lets ONE process run the strip kernels and the overlapped exchange
// under ncu -- k_agents<XM_P2P>, the boundary bands, the flag barriers, the ghost-row push -- by making the engine its own
// ring neighbour: the "peer" views are its own buffers, so every peer store lands in local HBM and every barrier is
// satisfied by the flags this rank wrote itself.  The numbers a profiler reads are those of the kernels; the simulation
// results are meaningless (the agents that leave the strip come back as arrivals in the wrong rows).
int sm_engine::fake_comm_init()
{
    counts_xchg_rows = (uint64_t)ghost + 1;
    SM_CUDA(cudaMalloc(&counts_xchg, 2 * counts_xchg_rows * W * sizeof(uint32_t)));
    mig_cap = 65536;
    mig_bytes = 16 + mig_cap * (sizeof(float4) + sizeof(uint32_t));
    for (int d = 0; d < 2; ++d) {
        SM_CUDA(cudaMalloc(&mig[d].send, mig_bytes));
        SM_CUDA(cudaMalloc(&mig[d].recv, mig_bytes));
        SM_CUDA(cudaMemset(mig[d].send, 0, 16));
        SM_CUDA(cudaMemset(mig[d].recv, 0, 16));
    }
    SM_CUDA(cudaMalloc(&dev_counters, 8 * sizeof(unsigned long long)));
    SM_CUDA(cudaMemset(dev_counters, 0, 8 * sizeof(unsigned long long)));
    SM_CUDA(cudaMallocHost(&host_counters, 8 * sizeof(unsigned long long)));
    const size_t arr_bytes = (mig_bytes + 63) & ~(size_t)63;
    window_arrival_off[0] = 64;
    window_arrival_off[1] = 64 + arr_bytes;
    SM_CUDA(cudaMalloc(&window, 64 + 2 * arr_bytes));
    SM_CUDA(cudaMemset(window, 0, 64 + 2 * arr_bytes));
    for (int d = 0; d < 2; ++d) {
        for (int i = 0; i < 2; ++i) { peer[d].counts[i] = counts_base[i]; peer[d].flags8[i] = flags_base[i]; peer[d].trail[i] = trail_base[i]; }
        peer[d].window = window;
        peer[d].rows = rows;
    }
    fake_multi = true;
    comm_ready = true;
    p2p = true;
    barrier_seq = 0;
    SM_TRY(p2p_streams_init());
    return SM_OK;
}

void sm_engine::comm_destroy()
{
    if (comm && g_nccl.ok) { ncclCommDestroy((ncclComm_t)comm); comm = nullptr; }
    if (counts_xchg) { cudaFree(counts_xchg); counts_xchg = nullptr; }
    for (int d = 0; d < 2; ++d) {
        if (mig[d].send) cudaFree(mig[d].send);
        if (mig[d].recv) cudaFree(mig[d].recv);
        mig[d] = MigrateBuf{};
    }
    for (void* p : ipc_opened) cudaIpcCloseMemHandle(p);
    ipc_opened.clear();
    if (window) { cudaFree(window); window = nullptr; }
    if (side_dbg && side_dbg_n[0]) {
        resolve_timing();
        fprintf(stderr, "[slime_b200 rank %d] side stream per step: boundary agents %.2f us, barrier1+pull %.2f us, bands %.2f us, push+arrivals+barrier2 %.2f us (%llu steps)\n",
                rank, 1e3 * side_dbg_ms[3] / std::max<uint64_t>(1, side_dbg_n[3]), 1e3 * side_dbg_ms[0] / side_dbg_n[0],
                1e3 * side_dbg_ms[1] / std::max<uint64_t>(1, side_dbg_n[1]), 1e3 * side_dbg_ms[2] / std::max<uint64_t>(1, side_dbg_n[2]),
                (unsigned long long)side_dbg_n[0]);
    }
    if (side_stream) { cudaStreamDestroy(side_stream); side_stream = nullptr; }
    if (ev_fork) { cudaEventDestroy(ev_fork); ev_fork = nullptr; }
    if (ev_join) { cudaEventDestroy(ev_join); ev_join = nullptr; }
    if (ev_boundary) { cudaEventDestroy(ev_boundary); ev_boundary = nullptr; }
    if (split_host) { cudaFreeHost(split_host); split_host = nullptr; }
    split_valid = false;
    p2p = false;
    if (dev_counters) { cudaFree(dev_counters); dev_counters = nullptr; }
    if (host_counters) { cudaFreeHost(host_counters); host_counters = nullptr; }
    comm_ready = false;
}

int sm_engine::init_agents_strip(uint64_t seed)
{
    unsigned long long* counter = nullptr;
    SM_CUDA(cudaMalloc(&counter, sizeof(unsigned long long)));
    SM_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
    if (n_global)
        smk::k_init_agents_strip<<<blocks_for(n_global, 256), 256, 0, stream>>>(
            agents[acur], ids[acur], n_global, seed, (float)W, (float)H, H, params.agent_speed_min,
            params.agent_speed_max, row0, rows, counter, cap_local);
    unsigned long long got = 0;
    cudaError_t err = cudaMemcpyAsync(&got, counter, sizeof got, cudaMemcpyDeviceToHost, stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
    cudaFree(counter);
    if (err != cudaSuccess) return sm_fail(SM_ERR_CUDA, "strip init failed: %s", cudaGetErrorString(err));
    if (got > cap_local)
        return sm_fail(SM_ERR_OOM, "strip %d would own %llu agents, capacity %llu", rank, got, (unsigned long long)cap_local);
    n_local = n_live = got;
    identity_order = false;
    SM_TRY(mark_tail_dead());
    SM_TRY(push_counters());
    return SM_OK;
}

// Deposit counts (+ leaver counts).  Leaves counts[ccur] complete on rows [-1, rows].
int sm_engine::exchange_counts()
{
    if (fake_multi) return SM_OK;
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    SM_TRY(tic(3));
    ncclComm_t c = (ncclComm_t)comm;
    const int up = (rank - 1 + world) % world, down = (rank + 1) % world;
    const size_t msg = (size_t)(m + 1) * W;                       // elements per direction
    if (deposit_mode == 2) {
        // u8 flags: same rows, one byte per cell, combined with OR
        uint8_t* f = flags_ptr(ccur);
        uint8_t* from_down = reinterpret_cast<uint8_t*>(counts_xchg);
        uint8_t* from_up = from_down + counts_xchg_rows * W;
        SM_NCCL(ncclGroupStart());
        SM_NCCL(ncclSend(f - (int64_t)m * W, msg, ncclUint8, up, c, stream));                  // rows [-m, 0]
        SM_NCCL(ncclSend(f + (int64_t)(rows - 1) * W, msg, ncclUint8, down, c, stream));       // rows [rows-1, rows+m)
        SM_NCCL(ncclRecv(from_down, msg, ncclUint8, down, c, stream));
        SM_NCCL(ncclRecv(from_up, msg, ncclUint8, up, c, stream));
        SM_NCCL(ncclGroupEnd());
        smk::k_flags_or<<<blocks_for(msg, 256), 256, 0, stream>>>(f - (int64_t)W, from_up, msg);
        smk::k_flags_or<<<blocks_for(msg, 256), 256, 0, stream>>>(f + (int64_t)(rows - m) * W, from_down, msg);
    } else {
        uint32_t* cnt = counts_ptr(ccur);
        uint32_t* from_down = counts_xchg;                            // [down's top ghost m rows | down's own first row]
        uint32_t* from_up = counts_xchg + counts_xchg_rows * W;       // [up's own last row | up's bottom ghost m rows]
        SM_NCCL(ncclGroupStart());
        SM_NCCL(ncclSend(cnt - (int64_t)m * W, msg, ncclUint32, up, c, stream));                 // rows [-m, 0]
        SM_NCCL(ncclSend(cnt + (int64_t)(rows - 1) * W, msg, ncclUint32, down, c, stream));      // rows [rows-1, rows+m)
        SM_NCCL(ncclRecv(from_down, msg, ncclUint32, down, c, stream));
        SM_NCCL(ncclRecv(from_up, msg, ncclUint32, up, c, stream));
        SM_NCCL(ncclGroupEnd());
        // from_up lands on my rows [-1, m): its own last row completes my ghost row -1, its ghost rows my first m rows
        smk::k_counts_add<<<blocks_for(msg, 256), 256, 0, stream>>>(cnt - (int64_t)W, from_up, msg);
        // from_down lands on my rows [rows-m, rows]: its ghost rows my last m rows, its own first row my ghost row `rows`
        smk::k_counts_add<<<blocks_for(msg, 256), 256, 0, stream>>>(cnt + (int64_t)(rows - m) * W, from_down, msg);
    }
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 2;
    SM_TRY(toc());
    return SM_OK;
}

// New trail rows -> neighbours' ghost rows (also used alone after uploads / clears).
int sm_engine::exchange_trail_ghosts()
{
    if (fake_multi) { ghost_stale = false; return SM_OK; }
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    SM_TRY(tic(3));
    ncclComm_t c = (ncclComm_t)comm;
    const int up = (rank - 1 + world) % world, down = (rank + 1) % world;
    float* t = trail_ptr(cur);
    const size_t msg = (size_t)g * W;
    SM_NCCL(ncclGroupStart());
    SM_NCCL(ncclSend(t, msg, ncclFloat, up, c, stream));                                     // my top g rows
    SM_NCCL(ncclSend(t + (int64_t)(rows - g) * W, msg, ncclFloat, down, c, stream));         // my bottom g rows
    SM_NCCL(ncclRecv(t + (int64_t)rows * W, msg, ncclFloat, down, c, stream));               // bottom ghost <- down's top
    SM_NCCL(ncclRecv(t - (int64_t)g * W, msg, ncclFloat, up, c, stream));                    // top ghost <- up's bottom
    SM_NCCL(ncclGroupEnd());
    SM_TRY(toc());
    ghost_stale = false;
    SM_TRY(refresh_tex_ghosts(g));
    return SM_OK;
}

// After the trail pass: new trail rows -> neighbours' ghost rows and the two leaver messages -> the
// neighbours, in ONE NCCL group; arrivals are appended on the device.  Also retires the ghost count
// rows this step used (the trail kernel only zeroes the owned rows of the *other* count buffer).
int sm_engine::migrate_agents()
{
    if (fake_multi) {
        for (int d = 0; d < 2; ++d) SM_CUDA(cudaMemsetAsync(mig[d].send, 0, 16, stream));
        return SM_OK;
    }
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    // counts buffer the agent kernel of THIS step wrote is 1 - ccur now (launch_trail flipped it)
    if (deposit_mode == 2) {
        uint8_t* used = flags_ptr(1 - ccur);
        SM_CUDA(cudaMemsetAsync(used - (int64_t)m * W, 0, (size_t)m * W, stream));
        SM_CUDA(cudaMemsetAsync(used + (int64_t)rows * W, 0, (size_t)m * W, stream));
    } else {
        uint32_t* used = counts_ptr(1 - ccur);
        SM_CUDA(cudaMemsetAsync(used - (int64_t)m * W, 0, (size_t)m * W * sizeof(uint32_t), stream));
        SM_CUDA(cudaMemsetAsync(used + (int64_t)rows * W, 0, (size_t)m * W * sizeof(uint32_t), stream));
    }

    SM_TRY(tic(3));
    ncclComm_t c = (ncclComm_t)comm;
    const int up = (rank - 1 + world) % world, down = (rank + 1) % world;
    float* t = trail_ptr(cur);
    const size_t tmsg = (size_t)g * W;
    SM_NCCL(ncclGroupStart());
    SM_NCCL(ncclSend(t, tmsg, ncclFloat, up, c, stream));
    SM_NCCL(ncclSend(t + (int64_t)(rows - g) * W, tmsg, ncclFloat, down, c, stream));
    SM_NCCL(ncclSend(mig[0].send, mig_bytes, ncclUint8, up, c, stream));
    SM_NCCL(ncclSend(mig[1].send, mig_bytes, ncclUint8, down, c, stream));
    SM_NCCL(ncclRecv(t + (int64_t)rows * W, tmsg, ncclFloat, down, c, stream));
    SM_NCCL(ncclRecv(t - (int64_t)g * W, tmsg, ncclFloat, up, c, stream));
    SM_NCCL(ncclRecv(mig[1].recv, mig_bytes, ncclUint8, down, c, stream));                   // what `down` sent up
    SM_NCCL(ncclRecv(mig[0].recv, mig_bytes, ncclUint8, up, c, stream));                     // what `up` sent down
    SM_NCCL(ncclGroupEnd());
    ghost_stale = false;
    SM_TRY(refresh_tex_ghosts(g));                      // the TEX sampler's copy needs the new ghost rows too
    smk::k_append_arrivals<<<blocks_for(2 * mig_cap, 256), 256, 0, stream>>>(mig[1].recv, mig[0].recv, (uint32_t)mig_cap,
                                                                            agents[acur], ids[acur], dev_counters, cap_local);
    smk::k_bump_counters<<<1, 1, 0, stream>>>(dev_counters, mig[1].recv, mig[0].recv,
                                              reinterpret_cast<unsigned long long*>(mig[0].send),
                                              reinterpret_cast<unsigned long long*>(mig[1].send), (uint32_t)mig_cap, cap_local);
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 2;
    SM_TRY(toc());
    n_upper = std::min<uint64_t>(cap_local, n_upper + 2 * mig_cap);
    return SM_OK;
}

// sm_resize on strips (collective; /root/reference/src/main.rs:954-1015 is a window event, so this is host-mediated and makes
// no attempt to be fast): every rank rescales its live agents exactly as k_rescale_agents does (one IEEE multiply per
// coordinate), keeps those that fall into its NEW strip, and hands the others -- rounding at the strip edges, or rows that
// moved because H' / world does not divide like H / world did -- to everybody through one all-gather; each rank adopts
// what is now its own.  Peer mappings are closed before any buffer is freed (the all-gathers are the barriers), the fields are
// reallocated zeroed at the new size, and the peer-memory path is set up again over the communicator that already exists.
int sm_engine::resize_strips(uint32_t width, uint32_t height)
{
    if (!comm_ready) return sm_fail(SM_ERR_STATE, "multi-GPU engine: call sm_comm_init first");
    if (fake_multi) return sm_fail(SM_ERR_STATE, "sm_resize: not in the single-rank profiling mode");
    if ((uint32_t)world > height) return sm_fail(SM_ERR_BAD_ARG, "more strips than rows");
    const uint32_t n_row0 = (uint32_t)(((uint64_t)rank * height) / world);
    const uint32_t n_rows = (uint32_t)(((uint64_t)(rank + 1) * height) / world) - n_row0;
    const uint32_t n_ghost = std::min(cfg.ghost_rows ? cfg.ghost_rows : 232u, height / (uint32_t)world);
    {   // the same rule as halo_depths(), on the new geometry, before anything is torn down (every rank decides alike)
        uint32_t g = 0, m = 0;
        SM_TRY(halo_depths(this, &g, &m));
        if (g > n_ghost || m > n_ghost || m + 1 > height / (uint32_t)world)
            return sm_fail(SM_ERR_BAD_ARG, "sm_resize: strips of %u rows are too thin (sensing needs %u ghost rows, motion %u)",
                           height / (uint32_t)world, g, m);
    }
    SM_CUDA(cudaStreamSynchronize(stream));
    if (side_stream) SM_CUDA(cudaStreamSynchronize(side_stream));
    volatile float fx = (float)width / (float)W;         // src/main.rs:985-989
    volatile float fy = (float)height / (float)H;

    std::vector<float> keep, orphan;
    std::vector<uint32_t> keep_ids, orphan_ids;
    if (agents_valid) {
        SM_TRY(refresh_counters());
        std::vector<float> a((size_t)n_local * 4);
        std::vector<uint32_t> id((size_t)n_local);
        SM_CUDA(cudaMemcpy(a.data(), agents[acur], n_local * sizeof(float4), cudaMemcpyDeviceToHost));
        SM_CUDA(cudaMemcpy(id.data(), ids[acur], n_local * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < n_local; ++i) {
            if (id[i] == smk::kDeadAgent) continue;
            volatile float x = a[4 * i] * fx, y = a[4 * i + 1] * fy;
            float v[4] = {x, y, a[4 * i + 2], a[4 * i + 3]};
            const uint32_t r = owner_row(y, height);
            const bool mine = r >= n_row0 && r < n_row0 + n_rows;
            (mine ? keep : orphan).insert((mine ? keep : orphan).end(), v, v + 4);
            (mine ? keep_ids : orphan_ids).push_back(id[i]);
        }
    }
    // one u64 from every rank (also serves as a barrier)
    ncclComm_t c = (ncclComm_t)comm;
    auto gather_u64 = [&](unsigned long long v, std::vector<unsigned long long>& out) -> int {
        unsigned long long* d = nullptr;
        SM_CUDA(cudaMalloc(&d, (size_t)(world + 1) * sizeof(unsigned long long)));
        SM_CUDA(cudaMemcpy(d + world, &v, sizeof v, cudaMemcpyHostToDevice));
        SM_NCCL(ncclAllGather(d + world, d, sizeof(unsigned long long), ncclUint8, c, stream));
        SM_CUDA(cudaStreamSynchronize(stream));
        out.resize((size_t)world);
        SM_CUDA(cudaMemcpy(out.data(), d, (size_t)world * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        cudaFree(d);
        return SM_OK;
    };
    // orphans: counts first (also the barrier "every rank has finished its last step") ...
    const unsigned long long mine_n = orphan_ids.size();
    std::vector<unsigned long long> cnt;
    SM_TRY(gather_u64(mine_n, cnt));
    // ... then the payloads, padded to the largest count
    const size_t slot = std::max<size_t>(1, (size_t)*std::max_element(cnt.begin(), cnt.end())) * 20;     // 16 B agent + 4 B id
    uint8_t* d_all = nullptr;
    SM_CUDA(cudaMalloc(&d_all, slot * (size_t)(world + 1)));
    std::vector<uint8_t> msg(slot, 0);
    if (mine_n) {
        memcpy(msg.data(), orphan.data(), (size_t)mine_n * 16);
        memcpy(msg.data() + (size_t)mine_n * 16, orphan_ids.data(), (size_t)mine_n * 4);
    }
    SM_CUDA(cudaMemcpy(d_all + slot * world, msg.data(), slot, cudaMemcpyHostToDevice));
    SM_NCCL(ncclAllGather(d_all + slot * world, d_all, slot, ncclUint8, c, stream));
    SM_CUDA(cudaStreamSynchronize(stream));
    std::vector<uint8_t> all(slot * (size_t)world);
    SM_CUDA(cudaMemcpy(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        const float* pa = reinterpret_cast<const float*>(all.data() + slot * r);
        const uint32_t* pi = reinterpret_cast<const uint32_t*>(all.data() + slot * r + (size_t)cnt[r] * 16);
        for (uint64_t i = 0; i < cnt[r]; ++i) {
            const uint32_t row = owner_row(pa[4 * i + 1], height);
            if (row >= n_row0 && row < n_row0 + n_rows) {
                keep.insert(keep.end(), pa + 4 * i, pa + 4 * i + 4);
                keep_ids.push_back(pi[i]);
            }
        }
    }
    const uint64_t m = keep_ids.size();
    // a strip that cannot hold its share fails the call on EVERY rank, before anything has been changed
    std::vector<unsigned long long> fits;
    SM_TRY(gather_u64(m <= cap_local ? 1ull : 0ull, fits));
    for (int r = 0; r < world; ++r)
        if (!fits[r]) return sm_fail(SM_ERR_OOM, "sm_resize: strip %d would own more agents than its capacity (%llu on this rank, capacity %llu)",
                                     r, (unsigned long long)m, (unsigned long long)cap_local);
    // peer mappings are closed before any buffer they point into is freed (the gather is the barrier "every mapping is closed")
    for (void* p : ipc_opened) cudaIpcCloseMemHandle(p);
    ipc_opened.clear();
    p2p = false;
    SM_TRY(gather_u64(1ull, fits));

    // new fields (zeroed: src/main.rs:999-1015), new tiles, new exchange buffers, peer memory mapped again
    if (window) { cudaFree(window); window = nullptr; }
    if (counts_xchg) { cudaFree(counts_xchg); counts_xchg = nullptr; }
    free_trail();
    if (tex_fallback) { use_tex = true; tex_fallback = false; }
    W = width; H = height; row0 = n_row0; rows = n_rows; ghost = n_ghost;
    cfg.width = width; cfg.height = height;
    params.width = width; params.height = height;
    SM_TRY(alloc_trail());
    SM_TRY(setup_tiles());
    counts_xchg_rows = (uint64_t)ghost + 1;
    SM_CUDA(cudaMalloc(&counts_xchg, 2 * counts_xchg_rows * W * sizeof(uint32_t)));
    SM_TRY(setup_p2p());
    if (m) {
        SM_CUDA(cudaMemcpy(agents[acur], keep.data(), m * sizeof(float4), cudaMemcpyHostToDevice));
        SM_CUDA(cudaMemcpy(ids[acur], keep_ids.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (agents_valid) {
        n_local = n_live = m;
        SM_TRY(mark_tail_dead());
        SM_TRY(push_counters());
        identity_order = false;
    }
    steps_since_sort = sort_interval;
    split_valid = false;
    ghost_stale = true;
    SM_CUDA(cudaStreamSynchronize(stream));
    return SM_OK;
}

// Between two sorts the agent kernel's grid covers n_upper >= (slots in use) slots; the ones that are
// not in use must read as dead.  Arrivals overwrite them from the front (k_append_arrivals).
int sm_engine::mark_tail_dead()
{
    if (world == 1) return SM_OK;
    const uint64_t steps = (uint64_t)std::max<uint32_t>(sort_interval, 1) + 1;
    const uint64_t end = std::min<uint64_t>(cap_local, n_local + steps * 2 * std::max<uint64_t>(mig_cap, 65536));
    if (end > n_local)
        SM_CUDA(cudaMemsetAsync(ids[acur] + n_local, 0xFF, (end - n_local) * sizeof(uint32_t), stream));
    return SM_OK;
}

int sm_engine::refresh_counters()
{
    if (!dev_counters) { SM_CUDA(cudaStreamSynchronize(stream)); return SM_OK; }
    SM_CUDA(cudaMemcpyAsync(host_counters, dev_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    SM_CUDA(cudaStreamSynchronize(stream));
    if (host_counters[2] == 3)
        return sm_fail(SM_ERR_STATE, "rank %d: a ring neighbour did not reach the step barrier within 20 s", rank);
    if (host_counters[2])
        return sm_fail(SM_ERR_OOM, "rank %d: %s overflow (raise sm_tuning.migrate_capacity or the agent capacity)", rank,
                       host_counters[2] == 1 ? "migration message" : "agent array");
    n_local = host_counters[0];
    n_live = host_counters[1];
    n_upper = n_local;
    return SM_OK;
}

int sm_engine::push_counters()
{
    n_upper = n_local;
    if (!dev_counters) return SM_OK;
    host_counters[0] = n_local;
    host_counters[1] = n_live;
    host_counters[2] = 0;
    SM_CUDA(cudaMemcpyAsync(dev_counters, host_counters, 3 * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    SM_CUDA(cudaStreamSynchronize(stream));
    return SM_OK;
}

// ---------------------------------------------------------------------------
// XM_P2P: direct peer writes over NVLink (CUDA IPC) instead of NCCL messages
// ---------------------------------------------------------------------------
namespace {
struct IpcBundle { cudaIpcMemHandle_t h[7]; uint32_t rows; uint32_t pad[15]; };   // counts0/1, flags0/1, trail0/1, window
}

// Exchanges IPC handles of the deposit fields, trail fields and the window with both ring neighbours
// (over the NCCL communicator that already exists) and maps them.  Any failure leaves the engine on
// the NCCL path (p2p = false) -- both are device paths; sm_tuning.exchange = 1 forces it.
int sm_engine::setup_p2p()
{
    p2p = false;
    const bool want = tuning.exchange == 0;
    // the decision must be collective: every rank exchanges a bundle even if it will not use it
    const size_t arr_bytes = (mig_bytes + 63) & ~(size_t)63;
    window_arrival_off[0] = 64;
    window_arrival_off[1] = 64 + arr_bytes;
    const size_t wbytes = 64 + 2 * arr_bytes;
    SM_CUDA(cudaMalloc(&window, wbytes));
    SM_CUDA(cudaMemset(window, 0, wbytes));

    IpcBundle mine{};
    bool ok = want;
    void* ptrs[7] = {counts_base[0], counts_base[1], flags_base[0], flags_base[1], trail_base[0], trail_base[1], window};
    for (int i = 0; i < 7 && ok; ++i)
        if (cudaIpcGetMemHandle(&mine.h[i], ptrs[i]) != cudaSuccess) { cudaGetLastError(); ok = false; }
    mine.rows = ok ? rows : 0u;        // rows == 0 announces "no P2P on this rank"

    IpcBundle* dev = nullptr;          // [0] mine, [1] from up, [2] from down
    SM_CUDA(cudaMalloc(&dev, 3 * sizeof(IpcBundle)));
    SM_CUDA(cudaMemcpy(dev, &mine, sizeof mine, cudaMemcpyHostToDevice));
    ncclComm_t c = (ncclComm_t)comm;
    const int up = (rank - 1 + world) % world, down = (rank + 1) % world;
    SM_NCCL(ncclGroupStart());
    SM_NCCL(ncclSend(dev, sizeof(IpcBundle), ncclUint8, up, c, stream));
    SM_NCCL(ncclSend(dev, sizeof(IpcBundle), ncclUint8, down, c, stream));
    SM_NCCL(ncclRecv(dev + 2, sizeof(IpcBundle), ncclUint8, down, c, stream));
    SM_NCCL(ncclRecv(dev + 1, sizeof(IpcBundle), ncclUint8, up, c, stream));
    SM_NCCL(ncclGroupEnd());
    SM_CUDA(cudaStreamSynchronize(stream));
    IpcBundle got[2];
    SM_CUDA(cudaMemcpy(got, dev + 1, 2 * sizeof(IpcBundle), cudaMemcpyDeviceToHost));
    cudaFree(dev);
    if (!ok || got[0].rows == 0 || got[1].rows == 0) return SM_OK;      // somebody cannot: stay on NCCL

    auto open_all = [&](const IpcBundle& b, PeerView& v) -> bool {
        void* p[7];
        for (int i = 0; i < 7; ++i) {
            if (cudaIpcOpenMemHandle(&p[i], b.h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            ipc_opened.push_back(p[i]);
        }
        v.counts[0] = (uint32_t*)p[0]; v.counts[1] = (uint32_t*)p[1];
        v.flags8[0] = (uint8_t*)p[2]; v.flags8[1] = (uint8_t*)p[3];
        v.trail[0] = (float*)p[4]; v.trail[1] = (float*)p[5];
        v.window = (uint8_t*)p[6];
        v.rows = b.rows;
        return true;
    };
    bool mapped = open_all(got[0], peer[0]);
    if (mapped) {
        if (up == down) peer[1] = peer[0];            // two strips: both neighbours are the same process
        else mapped = open_all(got[1], peer[1]);
    }
    // agree on the outcome (a rank that failed to map must not leave the others spinning on flags)
    unsigned int* flag = nullptr;
    SM_CUDA(cudaMalloc(&flag, 3 * sizeof(unsigned int)));
    unsigned int mineok = mapped ? 1u : 0u;
    SM_CUDA(cudaMemcpy(flag, &mineok, sizeof mineok, cudaMemcpyHostToDevice));
    SM_NCCL(ncclGroupStart());
    SM_NCCL(ncclSend(flag, 1, ncclUint32, up, c, stream));
    SM_NCCL(ncclSend(flag, 1, ncclUint32, down, c, stream));
    SM_NCCL(ncclRecv(flag + 2, 1, ncclUint32, down, c, stream));
    SM_NCCL(ncclRecv(flag + 1, 1, ncclUint32, up, c, stream));
    SM_NCCL(ncclGroupEnd());
    SM_CUDA(cudaStreamSynchronize(stream));
    unsigned int oks[3];
    SM_CUDA(cudaMemcpy(oks, flag, sizeof oks, cudaMemcpyDeviceToHost));
    cudaFree(flag);
    // neighbours-only agreement is enough: a rank uses P2P with its two neighbours, and each of them
    // sees this rank's flag.  For the whole ring to take one path we need a global AND, reached here by
    // world/2 + 1 relay rounds of the same neighbour exchange.
    unsigned int all = oks[0] & oks[1] & oks[2];
    for (int round = 0; round < world / 2 + 1; ++round) {
        unsigned int* f2 = nullptr;
        SM_CUDA(cudaMalloc(&f2, 3 * sizeof(unsigned int)));
        SM_CUDA(cudaMemcpy(f2, &all, sizeof all, cudaMemcpyHostToDevice));
        SM_NCCL(ncclGroupStart());
        SM_NCCL(ncclSend(f2, 1, ncclUint32, up, c, stream));
        SM_NCCL(ncclSend(f2, 1, ncclUint32, down, c, stream));
        SM_NCCL(ncclRecv(f2 + 2, 1, ncclUint32, down, c, stream));
        SM_NCCL(ncclRecv(f2 + 1, 1, ncclUint32, up, c, stream));
        SM_NCCL(ncclGroupEnd());
        SM_CUDA(cudaStreamSynchronize(stream));
        unsigned int r[3];
        SM_CUDA(cudaMemcpy(r, f2, sizeof r, cudaMemcpyDeviceToHost));
        cudaFree(f2);
        all = r[0] & r[1] & r[2];
    }
    p2p = all != 0;
    barrier_seq = 0;
    if (p2p) SM_TRY(p2p_streams_init());
    return SM_OK;
}

// Side stream, fork / join events and barrier flavour of the overlapped exchange.
int sm_engine::p2p_streams_init()
{
    if (!side_stream) {
        // highest priority: the small exchange kernels must not queue behind the pending CTAs of the
        // interior trail pass (the block scheduler drains grids of equal priority in launch order)
        int prio_lo = 0, prio_hi = 0;
        SM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        SM_CUDA(cudaStreamCreateWithPriority(&side_stream, cudaStreamNonBlocking, prio_hi));
        SM_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        SM_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        SM_CUDA(cudaEventCreateWithFlags(&ev_boundary, cudaEventDisableTiming));
        SM_CUDA(cudaMallocHost(&split_host, 2 * sizeof(uint32_t)));
    }
    overlap_enabled = tuning.serial_exchange == 0;
    side_dbg = tuning.debug_side_timing != 0;
    // g_fence_mode: 0 three sequentially-consistent system fences, 1 acq_rel system fences (default), 2 device scope
    const int mode = tuning.barrier_fence == 1 ? 0 : tuning.barrier_fence == 2 ? 2 : 1;
    SM_CUDA(cudaMemcpyToSymbol(smk::g_fence_mode, &mode, sizeof mode));
    return SM_OK;
}

// A bare barrier: used when a rank-local memset must be ordered against the neighbours' next writes.
int sm_engine::p2p_barrier()
{
    ++barrier_seq;
    smk::k_barrier_only<<<1, 1, 0, stream>>>(reinterpret_cast<uint32_t*>(window), reinterpret_cast<uint32_t*>(peer[0].window),
                                             reinterpret_cast<uint32_t*>(peer[1].window), barrier_seq, dev_counters + 2);
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 1;
    return SM_OK;
}

int sm_engine::p2p_after_agents(cudaStream_t st, bool timed)
{
    if (!st) st = stream;
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    if (timed) SM_TRY(tic(3));
    ++barrier_seq;
    uint32_t* w = reinterpret_cast<uint32_t*>(window);
    uint32_t* wu = reinterpret_cast<uint32_t*>(peer[0].window);
    uint32_t* wd = reinterpret_cast<uint32_t*>(peer[1].window);
    const size_t row0_off = (size_t)(ghost + pad_rows) * W;
    // rows of each neighbour's deposit field the trail pass of my edge rows reads: one for the 3x3 box, R for the Gaussian
    // extension (its full steps on strips run the serial order: pull, pass, push)
    uint32_t pr = 1;
    if (cfg.flags & SM_FLAG_GAUSSIAN_BLUR) pr = (uint32_t)std::max(1l, lroundf(params.blur_radius));
    const uint32_t pw = pr * W;                      // contiguous: `pr` whole rows
    // one 16-byte element per thread (flags: 1 byte per cell, counts: 4)
    const uint64_t pull_elems = ((uint64_t)pw * (deposit_mode == 2 ? 1 : 4) + 15) / 16;
    const unsigned pull_blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(256, (pull_elems + 127) / 128));
    if (deposit_mode == 2 && flags_tiled()) {
        // row y of a strip starts at byte ((y' >> 3) * W + (y' & 7)) * 8 past its owned row 0, y' = y - 1 (signed: ghost rows)
        auto row_at = [&](int64_t y) { const int64_t yp = y - 1; return ((yp >> 3) * (int64_t)W + (yp & 7)) * 8; };
        uint8_t* f = flags_ptr(ccur);
        const uint8_t* up_last = peer[0].flags8[ccur] + row0_off + row_at((int64_t)peer[0].rows - 1);
        const uint8_t* down_first = peer[1].flags8[ccur] + row0_off + row_at(0);
        const uint32_t pieces = W / 8;
        const unsigned nb = (unsigned)std::max<uint32_t>(1, std::min<uint32_t>(256, (pieces + 127) / 128));
        smk::k_barrier_pull_tiled<<<nb, 128, 0, st>>>(w, wu, wd, barrier_seq, dev_counters + 2, f + row_at(-1), up_last,
                                                      f + row_at((int64_t)rows), down_first, pieces);
    } else if (deposit_mode == 2) {
        uint8_t* f = flags_ptr(ccur);
        const uint8_t* up_last = peer[0].flags8[ccur] + row0_off + (size_t)(peer[0].rows - pr) * W;
        const uint8_t* down_first = peer[1].flags8[ccur] + row0_off;
        smk::k_barrier_pull<uint8_t><<<pull_blocks, 128, 0, st>>>(w, wu, wd, barrier_seq, dev_counters + 2, f - (int64_t)pr * W, up_last,
                                                            f + (int64_t)rows * W, down_first, pw);
    } else {
        uint32_t* cn = counts_ptr(ccur);
        const uint32_t* up_last = peer[0].counts[ccur] + row0_off + (size_t)(peer[0].rows - pr) * W;
        const uint32_t* down_first = peer[1].counts[ccur] + row0_off;
        smk::k_barrier_pull<uint32_t><<<pull_blocks, 128, 0, st>>>(w, wu, wd, barrier_seq, dev_counters + 2, cn - (int64_t)pr * W, up_last,
                                                             cn + (int64_t)rows * W, down_first, pw);
    }
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 1;
    if (timed) SM_TRY(toc());
    return SM_OK;
}

// my new rows -> the neighbours' ghost rows of THEIR trail[cur] (all ranks flip `cur` in lock step)
int sm_engine::p2p_push_ghosts(cudaStream_t st, uint32_t g)
{
    const size_t row0_off = (size_t)(ghost + pad_rows) * W;
    const float* t = trail_ptr(cur);
    float* up_ghost = peer[0].trail[cur] + row0_off + (size_t)peer[0].rows * W;      // up's bottom ghost rows [rows_up, rows_up + g)
    float* down_ghost = peer[1].trail[cur] + row0_off - (size_t)g * W;               // down's top ghost rows [-g, 0)
    const uint64_t n = (uint64_t)g * W;
    const unsigned nb = (unsigned)std::min<uint64_t>((n / 4 + 255) / 256 + 1, (uint64_t)num_sms);
    if (W % 4 == 0)
        smk::k_push_rows<<<nb, 256, 0, st>>>(reinterpret_cast<const float4*>(t), reinterpret_cast<float4*>(up_ghost),
                                                reinterpret_cast<const float4*>(t + (size_t)(rows - g) * W),
                                                reinterpret_cast<float4*>(down_ghost), n / 4);
    else
        smk::k_push_rows_scalar<<<nb, 256, 0, st>>>(t, up_ghost, t + (size_t)(rows - g) * W, down_ghost, n);
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 1;
    return SM_OK;
}

int sm_engine::p2p_after_trail(cudaStream_t st, bool timed)
{
    const bool own_stream = !st;
    if (!st) st = stream;
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    if (timed) SM_TRY(tic(3));
    SM_TRY(p2p_push_ghosts(st, g));
    // arrivals (written by the neighbours during their agent pass, complete since barrier 1)
    uint8_t* from_up = window + window_arrival_off[0];
    uint8_t* from_down = window + window_arrival_off[1];
    smk::k_append_arrivals<<<blocks_for(2 * mig_cap, 256), 256, 0, st>>>(from_down, from_up, (uint32_t)mig_cap, agents[acur],
                                                                            ids[acur], dev_counters, cap_local);
    ++barrier_seq;
    smk::k_bump_barrier<<<1, 1, 0, st>>>(dev_counters, from_down, from_up, (uint32_t)mig_cap, cap_local,
                                             reinterpret_cast<uint32_t*>(window), reinterpret_cast<uint32_t*>(peer[0].window),
                                             reinterpret_cast<uint32_t*>(peer[1].window), barrier_seq);
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 2;
    ghost_stale = false;
    if (own_stream) SM_TRY(refresh_tex_ghosts(g));      // the TEX sampler's copy needs the new ghost rows too (overlapped form: after the join)
    if (timed) SM_TRY(toc());
    n_upper = std::min<uint64_t>(cap_local, n_upper + 2 * mig_cap);
    return SM_OK;
}

// ---------------------------------------------------------------------------
// Overlapped exchange (the default on strips that are tall enough)
// ---------------------------------------------------------------------------
// Only the rows within `band` of a strip edge depend on the neighbours during a step: remote deposits
// land within m rows of the edge, the 3x3 blur of the edge row needs one deposit row of the neighbour,
// and the g rows next to each edge are what the neighbours need back as ghost rows.  So after the agent
// pass the main stream runs the trail pass over the interior rows at once, while the side stream runs
//   barrier 1 + deposit-row pull -> trail pass over the two boundary bands -> ghost-row push over NVLink
//   -> arrivals -> barrier 2,
// and the main stream joins it before refreshing the sampler's ghost rows.  Nothing of the exchange but
// the join and that refresh is left on the critical path.
uint32_t sm_engine::overlap_band()
{
    uint32_t g = 0, m = 0;
    if (halo_depths(this, &g, &m) != SM_OK) return 0;
    TrailPass p;
    if (trail_plan(true, p) != SM_OK || !p.fast) return 0;
    const uint32_t rpc = p.g.rows_per_chunk;
    uint32_t band = std::max(g, m + 1);
    band = (band + rpc - 1) / rpc * rpc;
    if ((uint64_t)rows < 2ull * band + rpc) return 0;          // no interior worth splitting off
    return band;
}

bool sm_engine::overlap_ok()
{
    if (cfg.flags & SM_FLAG_GAUSSIAN_BLUR) return false;   // the Gaussian extension runs the serial order (pass, then ghost exchange)
    return overlap_enabled && p2p && world > 1 && side_stream && overlap_band() > 0;
}

// Diffusion-only pass on strips (sm_diffuse_only): interior rows on the main stream; boundary bands, ghost-row
// push and ONE barrier beside them.  The barrier of pass p-1 is what makes the neighbours' ghost buffers
// safe to overwrite in pass p (they finished reading them in their boundary bands of pass p-1).
int sm_engine::p2p_diffuse_overlapped()
{
    const uint32_t band = overlap_band();
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    TrailPass p;
    SM_TRY(trail_plan(false, p));
    SM_CUDA(cudaEventRecord(ev_fork, stream));
    SM_CUDA(cudaStreamWaitEvent(side_stream, ev_fork, 0));
    SM_TRY(tic(1));
    SM_TRY(trail_launch_rows(p, band, rows - band, stream));
    SM_TRY(toc());
    SM_TRY(trail_launch_rows(p, 0, band, side_stream, rows - band, rows));
    trail_done(false);
    SM_TRY(p2p_push_ghosts(side_stream, g));
    ++barrier_seq;
    smk::k_barrier_only<<<1, 1, 0, side_stream>>>(reinterpret_cast<uint32_t*>(window), reinterpret_cast<uint32_t*>(peer[0].window),
                                                  reinterpret_cast<uint32_t*>(peer[1].window), barrier_seq, dev_counters + 2);
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 1;
    SM_CUDA(cudaEventRecord(ev_join, side_stream));
    SM_TRY(tic(3));
    SM_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
    SM_TRY(toc());
    ghost_stale = false;
    return SM_OK;
}

int sm_engine::p2p_trail_overlapped()
{
    const uint32_t band = overlap_band();
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    TrailPass p;
    SM_TRY(trail_plan(true, p));
    // fork: the side stream starts where the agent pass ends
    SM_CUDA(cudaEventRecord(ev_fork, stream));
    SM_CUDA(cudaStreamWaitEvent(side_stream, ev_fork, 0));
    // side (launched first, so that its one-CTA barrier kernel is resident before the interior pass fills the
    // machine): barrier 1 -- the neighbours' deposits and leavers have landed -- and the deposit-row pull
    if (side_dbg) SM_TRY(tic(10, side_stream));
    SM_TRY(p2p_after_agents(side_stream, false));
    if (side_dbg) SM_TRY(toc(side_stream));
    // main: interior rows -- every input is local and final once this rank's agent pass is done
    SM_TRY(tic(1));
    SM_TRY(trail_launch_rows(p, band, rows - band, stream));
    SM_TRY(toc());
    // side: the two boundary bands in one launch
    if (side_dbg) SM_TRY(tic(11, side_stream));
    SM_TRY(trail_launch_rows(p, 0, band, side_stream, rows - band, rows));
    if (side_dbg) { SM_TRY(toc(side_stream)); SM_TRY(tic(12, side_stream)); }
    trail_done(true);                                          // host bookkeeping: cur / ccur flip
    stats_fused_valid = p.stats;                               // interior + both bands accumulated this rank's statistics
    if (stats_interest) --stats_interest;
    // side: push the new boundary rows, append arrivals, barrier 2
    SM_TRY(p2p_after_trail(side_stream, false));
    if (side_dbg) SM_TRY(toc(side_stream));
    SM_CUDA(cudaEventRecord(ev_join, side_stream));
    // join; what is timed as "exchange" is the part that is NOT hidden
    SM_TRY(tic(3));
    SM_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
    SM_TRY(refresh_tex_ghosts(g));
    SM_TRY(toc());
    return SM_OK;
}

// ---------------------------------------------------------------------------
// Boundary-first step (the default on strips with an interior; engine.cu: plan_split)
// ---------------------------------------------------------------------------
// "launch boundary-row work first -> start exchange -> interior agents / blur -> wait -> finish edges".
//   side stream (high priority)                                   main stream
//   k_agents<XM_P2P> over the boundary slots + arrivals            k_agents<XM_SINGLE> over the interior slots
//   barrier 1 + deposit-row pull                                       (the long kernel: everything on the left hides behind it)
//   trail pass over the two boundary bands
//   ghost-row push, arrivals, barrier 2                            wait(boundary agents) -> trail pass over the interior rows
//                                                                  wait(side) -> sampler ghost rows
// Why the two agent launches are independent, and the bands independent of the interior agents: plan_split's margin.
// The boundary bands are written into the sampler array while the interior agents still gather from it -- rows those
// agents cannot reach (margin includes the sensing reach).
int sm_engine::p2p_step_split()
{
    const uint32_t band = overlap_band();
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    SM_TRY(prepare_agents());                                      // deposit mode, sampler copy: once, before the fork
    SM_CUDA(cudaEventRecord(ev_fork, stream));
    SM_CUDA(cudaStreamWaitEvent(side_stream, ev_fork, 0));
    // side: boundary agents first -- their deposits and leavers are what the neighbours wait for
    if (side_dbg) SM_TRY(tic(13, side_stream));
    SM_TRY(launch_agents(2, side_stream));
    if (side_dbg) SM_TRY(toc(side_stream));
    SM_CUDA(cudaEventRecord(ev_boundary, side_stream));
    // main: interior agents
    SM_TRY(launch_agents(1, stream));
    // side: barrier 1 + pull
    if (side_dbg) SM_TRY(tic(10, side_stream));
    SM_TRY(p2p_after_agents(side_stream, false));
    if (side_dbg) SM_TRY(toc(side_stream));
    TrailPass p;
    SM_TRY(trail_plan(true, p));
    // main: interior rows, once the boundary agents' deposits are in as well
    SM_CUDA(cudaStreamWaitEvent(stream, ev_boundary, 0));
    SM_TRY(tic(1));
    SM_TRY(trail_launch_rows(p, band, rows - band, stream));
    SM_TRY(toc());
    // side: the two boundary bands in one launch
    if (side_dbg) SM_TRY(tic(11, side_stream));
    SM_TRY(trail_launch_rows(p, 0, band, side_stream, rows - band, rows));
    if (side_dbg) { SM_TRY(toc(side_stream)); SM_TRY(tic(12, side_stream)); }
    trail_done(true);
    stats_fused_valid = p.stats;
    if (stats_interest) --stats_interest;
    SM_TRY(p2p_after_trail(side_stream, false));
    if (side_dbg) SM_TRY(toc(side_stream));
    SM_CUDA(cudaEventRecord(ev_join, side_stream));
    SM_TRY(tic(3));
    SM_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
    SM_TRY(refresh_tex_ghosts(g));
    SM_TRY(toc());
    return SM_OK;
}
