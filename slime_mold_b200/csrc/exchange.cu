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
//                      (contiguous (m+1) x W u32), and my leaver count; one host sync to learn the
//                      arrival counts
//   k_trail_rows       merge -> decay -> 3x3 mean on the owned rows (ghost row +-1 now complete)
//   exchange_trail_ghosts() + migrate_agents(): my new top / bottom g trail rows -> the neighbours'
//                      ghost rows, leavers -> the tail of the neighbour's agent arrays
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
    SM_SYM(GetUniqueId); SM_SYM(CommInitRank); SM_SYM(CommDestroy); SM_SYM(Send); SM_SYM(Recv);
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

    // staging: received count rows (2 x (ghost + 1) rows), leaver buffers, counters
    e->counts_xchg_rows = (uint64_t)e->ghost + 1;
    SM_CUDA(cudaMalloc(&e->counts_xchg, 2 * e->counts_xchg_rows * e->W * sizeof(uint32_t)));
    // leavers per step ~ (agents per row) * m rows; generous: 1/8 of the strip's capacity, at least 64k
    uint64_t cap = std::max<uint64_t>(e->cap_local / 8, 65536);
    cap = std::min<uint64_t>(cap, 0x7fffffffull);
    for (int d = 0; d < 2; ++d) {
        e->mig[d].cap = cap;
        SM_CUDA(cudaMalloc(&e->mig[d].send_a, cap * sizeof(float4)));
        SM_CUDA(cudaMalloc(&e->mig[d].send_id, cap * sizeof(uint32_t)));
    }
    SM_CUDA(cudaMalloc(&e->mig_counters, 8 * sizeof(unsigned long long)));
    SM_CUDA(cudaMemset(e->mig_counters, 0, 8 * sizeof(unsigned long long)));
    SM_CUDA(cudaMallocHost(&e->mig_counters_host, 8 * sizeof(unsigned long long)));
    e->comm_ready = true;
    e->ghost_stale = true;
    return SM_OK;
}

void sm_engine::comm_destroy()
{
    if (comm && g_nccl.ok) { ncclCommDestroy((ncclComm_t)comm); comm = nullptr; }
    if (counts_xchg) { cudaFree(counts_xchg); counts_xchg = nullptr; }
    for (int d = 0; d < 2; ++d) {
        if (mig[d].send_a) cudaFree(mig[d].send_a);
        if (mig[d].send_id) cudaFree(mig[d].send_id);
        mig[d] = MigrateBuf{};
    }
    if (mig_counters) { cudaFree(mig_counters); mig_counters = nullptr; }
    if (mig_counters_host) { cudaFreeHost(mig_counters_host); mig_counters_host = nullptr; }
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
    return SM_OK;
}

// Deposit counts (+ leaver counts).  Leaves counts[ccur] complete on rows [-1, rows].
int sm_engine::exchange_counts()
{
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    SM_TRY(tic(3));
    ncclComm_t c = (ncclComm_t)comm;
    const int up = (rank - 1 + world) % world, down = (rank + 1) % world;
    uint32_t* cnt = counts_ptr(ccur);
    const size_t msg = (size_t)(m + 1) * W;                       // u32 elements per direction
    uint32_t* from_down = counts_xchg;                            // [down's top ghost m rows | down's own first row]
    uint32_t* from_up = counts_xchg + counts_xchg_rows * W;       // [up's own last row | up's bottom ghost m rows]
    SM_NCCL(ncclGroupStart());
    SM_NCCL(ncclSend(cnt - (int64_t)m * W, msg, ncclUint32, up, c, stream));                 // rows [-m, 0]
    SM_NCCL(ncclSend(cnt + (int64_t)(rows - 1) * W, msg, ncclUint32, down, c, stream));      // rows [rows-1, rows+m)
    SM_NCCL(ncclSend(mig_counters + 0, 1, ncclUint64, up, c, stream));
    SM_NCCL(ncclSend(mig_counters + 1, 1, ncclUint64, down, c, stream));
    SM_NCCL(ncclRecv(from_down, msg, ncclUint32, down, c, stream));
    SM_NCCL(ncclRecv(from_up, msg, ncclUint32, up, c, stream));
    SM_NCCL(ncclRecv(mig_counters + 2, 1, ncclUint64, down, c, stream));                     // arrivals from down
    SM_NCCL(ncclRecv(mig_counters + 3, 1, ncclUint64, up, c, stream));                       // arrivals from up
    SM_NCCL(ncclGroupEnd());
    // from_up lands on my rows [-1, m): its own last row completes my ghost row -1, its ghost rows my first m rows
    smk::k_counts_add<<<blocks_for(msg, 256), 256, 0, stream>>>(cnt - (int64_t)W, from_up, msg);
    // from_down lands on my rows [rows-m, rows]: its ghost rows my last m rows, its own first row my ghost row `rows`
    smk::k_counts_add<<<blocks_for(msg, 256), 256, 0, stream>>>(cnt + (int64_t)(rows - m) * W, from_down, msg);
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 2;
    SM_CUDA(cudaMemcpyAsync(mig_counters_host, mig_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    SM_TRY(toc());
    return SM_OK;
}

// New trail rows -> neighbours' ghost rows.
int sm_engine::exchange_trail_ghosts()
{
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
    return SM_OK;
}

// Leavers -> neighbours; arrivals appended to the live arrays.  Also retires the ghost count rows
// this step used (the trail kernel only zeroes the owned rows of the *other* count buffer).
int sm_engine::migrate_agents()
{
    uint32_t g = 0, m = 0;
    SM_TRY(halo_depths(this, &g, &m));
    // counts buffer the agent kernel of THIS step wrote is 1 - ccur now (launch_trail flipped it)
    uint32_t* used = counts_ptr(1 - ccur);
    SM_CUDA(cudaMemsetAsync(used - (int64_t)m * W, 0, (size_t)m * W * sizeof(uint32_t), stream));
    SM_CUDA(cudaMemsetAsync(used + (int64_t)rows * W, 0, (size_t)m * W * sizeof(uint32_t), stream));

    SM_CUDA(cudaStreamSynchronize(stream));            // the one host round trip per step: leaver / arrival counts
    const unsigned long long leave_up = mig_counters_host[0], leave_down = mig_counters_host[1];
    const unsigned long long arr_down = mig_counters_host[2], arr_up = mig_counters_host[3];
    if (mig_counters_host[4])
        return sm_fail(SM_ERR_OOM, "migration staging overflow on rank %d (%llu up, %llu down, capacity %llu)", rank,
                       leave_up, leave_down, (unsigned long long)mig[0].cap);
    if (n_local + arr_down + arr_up > cap_local)
        return sm_fail(SM_ERR_OOM, "rank %d agent capacity exceeded: %llu + %llu arrivals > %llu", rank,
                       (unsigned long long)n_local, arr_down + arr_up, (unsigned long long)cap_local);
    SM_TRY(tic(3));
    ncclComm_t c = (ncclComm_t)comm;
    const int up = (rank - 1 + world) % world, down = (rank + 1) % world;
    float4* a = agents[acur];
    uint32_t* id = ids[acur];
    SM_NCCL(ncclGroupStart());
    if (leave_up) {
        SM_NCCL(ncclSend(mig[0].send_a, leave_up * 4, ncclFloat, up, c, stream));
        SM_NCCL(ncclSend(mig[0].send_id, leave_up, ncclUint32, up, c, stream));
    }
    if (leave_down) {
        SM_NCCL(ncclSend(mig[1].send_a, leave_down * 4, ncclFloat, down, c, stream));
        SM_NCCL(ncclSend(mig[1].send_id, leave_down, ncclUint32, down, c, stream));
    }
    if (arr_down) {
        SM_NCCL(ncclRecv(a + n_local, arr_down * 4, ncclFloat, down, c, stream));
        SM_NCCL(ncclRecv(id + n_local, arr_down, ncclUint32, down, c, stream));
    }
    if (arr_up) {
        SM_NCCL(ncclRecv(a + n_local + arr_down, arr_up * 4, ncclFloat, up, c, stream));
        SM_NCCL(ncclRecv(id + n_local + arr_down, arr_up, ncclUint32, up, c, stream));
    }
    SM_NCCL(ncclGroupEnd());
    SM_CUDA(cudaMemsetAsync(mig_counters, 0, 8 * sizeof(unsigned long long), stream));
    SM_TRY(toc());
    n_local += arr_down + arr_up;
    n_live = n_live - leave_up - leave_down + arr_down + arr_up;
    return SM_OK;
}
