// engine.cu -- C ABI (include/slime_b200.h) over the sm_100a kernels in kernels.cuh.
//
// Replaces, for the simulation half of the reference, what
// /root/reference/src/pipeline_manager.rs and src/bind_group_manager.rs set up and
// what src/main.rs:1163-1235 dispatches every frame.  No CPU fallback: every entry
// point that needs the device fails with SM_ERR_NO_DEVICE / SM_ERR_CUDA otherwise.
#include "../../include/slime_b200.h"
#include "kernels.cuh"
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

thread_local std::string g_sm_err;

int sm_fail(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_sm_err = buf;
    return code;
}

#define SM_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return sm_fail(err__ == cudaErrorMemoryAllocation ? SM_ERR_OOM : SM_ERR_CUDA,     \
                           "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                           cudaGetErrorString(err__));                                         \
    } while (0)

#define SM_TRY(expr)                    \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != SM_OK) return rc__; \
    } while (0)

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// ---------------------------------------------------------------------------
// derived constants
// ---------------------------------------------------------------------------
smd::AgentConsts sm_engine::agent_consts() const
{
    smd::AgentConsts c{};
    c.W = W; c.H = H;
    c.Wf = (float)W; c.Hf = (float)H;
    c.rcpW = 1.0f / c.Wf; c.rcpH = 1.0f / c.Hf;
    c.neg_zero = -0.0f;
    c.zero_bits = 0u;
    c.xmax = (float)W - 2.0f; c.ymax = (float)H - 2.0f;
    c.speed_min = params.agent_speed_min; c.speed_max = params.agent_speed_max;
    c.turn_speed = params.agent_turn_speed;
    c.sensor_angle = params.agent_sensor_angle;
    c.sensor_distance = params.agent_sensor_distance;
    c.jitter = params.agent_jitter;
    c.row_base = (int64_t)row0;
    c.rows_local = (int32_t)rows;
    c.ghost = (int32_t)ghost;
    const int32_t spare = (int32_t)H - (int32_t)rows;
    c.fold_hi = (int32_t)rows + (spare + 1) / 2;
    c.fold_lo = -(spare / 2);
    c.flag_wrap = world == 1 ? (int32_t)H : 0;
    return c;
}

smd::TrailConsts sm_engine::trail_consts() const
{
    smd::TrailConsts t{};
    t.dep = params.pheromone_deposition_amount;
    volatile float d = params.decay_factor * 0.001f;            // compute.wgsl:159 (f32 product)
    t.decay_sub = d;
    t.rate = fminf(fmaxf(params.diffusion_rate, 0.0f), 1.0f);   // compute.wgsl:173
    volatile float om = 1.0f - t.rate;
    t.one_minus_rate = om;
    return t;
}

// ---------------------------------------------------------------------------
// timing helpers
// ---------------------------------------------------------------------------
int sm_engine::tic(int kind, cudaStream_t st)
{
    if (!st) st = stream;
    if (!timing_enabled) return SM_OK;
    if (ev_used == ev_pool.size()) {
        if (ev_pool.size() >= 4096) SM_TRY(resolve_timing());
        else {
            EvPair p{};
            SM_CUDA(cudaEventCreate(&p.a));
            SM_CUDA(cudaEventCreate(&p.b));
            ev_pool.push_back(p);
        }
    }
    ev_pool[ev_used].kind = kind;
    SM_CUDA(cudaEventRecord(ev_pool[ev_used].a, st));
    return SM_OK;
}
int sm_engine::toc(cudaStream_t st)
{
    if (!timing_enabled) return SM_OK;
    if (!st) st = stream;
    SM_CUDA(cudaEventRecord(ev_pool[ev_used].b, st));
    ev_used++;
    return SM_OK;
}
int sm_engine::resolve_timing()
{
    if (ev_used == 0) return SM_OK;
    SM_CUDA(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < ev_used; ++i) {
        float ms = 0.f;
        SM_CUDA(cudaEventElapsedTime(&ms, ev_pool[i].a, ev_pool[i].b));
        switch (ev_pool[i].kind) {
        case 0: timing.agents_ms += ms; timing.agent_launches++; break;
        case 1: timing.trail_ms += ms; timing.trail_launches++; break;
        case 2: timing.sort_ms += ms; timing.sort_launches++; break;
        case 3: timing.exchange_ms += ms; timing.exchange_launches++; break;
        default:                                     // >= 10: side-stream pieces of the overlapped exchange (diagnostics only)
            if (ev_pool[i].kind - 10 < 8) { side_dbg_ms[ev_pool[i].kind - 10] += ms; side_dbg_n[ev_pool[i].kind - 10]++; }
            break;
        }
    }
    ev_used = 0;
    return SM_OK;
}

// ---------------------------------------------------------------------------
// allocation
// ---------------------------------------------------------------------------
int sm_engine::alloc_trail()
{
    const size_t cells = field_cells();
    for (int i = 0; i < 2; ++i) {
        SM_CUDA(cudaMalloc(&trail_base[i], cells * sizeof(float)));
        SM_CUDA(cudaMemsetAsync(trail_base[i], 0, cells * sizeof(float), stream));
        SM_CUDA(cudaMalloc(&counts_base[i], cells * sizeof(uint32_t)));
        SM_CUDA(cudaMemsetAsync(counts_base[i], 0, cells * sizeof(uint32_t), stream));
        SM_CUDA(cudaMalloc(&flags_base[i], cells));
        SM_CUDA(cudaMemsetAsync(flags_base[i], 0, cells, stream));
    }
    cur = 0; ccur = 0;
    stats_fused_valid = false;
    frame_pre_valid = false;
    trail_nonneg = true;
    deposit_mode = 0;
    if (use_tex && setup_tex() != SM_OK) {
        // no block-linear copy for this map (above the texture-gather size limit, or out of memory for the extra
        // 4 B/cell): the agent kernel samples the row-major field with LDGs instead -- same bits, still the CUDA path
        cudaGetLastError();
        free_tex();
        use_tex = false;
        tex_fallback = true;
    }
    return SM_OK;
}
void sm_engine::free_trail()
{
    free_tex();
    for (int i = 0; i < 2; ++i) {
        if (trail_base[i]) cudaFree(trail_base[i]);
        if (counts_base[i]) cudaFree(counts_base[i]);
        if (flags_base[i]) cudaFree(flags_base[i]);
        trail_base[i] = nullptr; counts_base[i] = nullptr; flags_base[i] = nullptr;
    }
    if (gauss_dec) cudaFree(gauss_dec);
    if (gauss_hb) cudaFree(gauss_hb);
    gauss_dec = gauss_hb = nullptr;
}
int sm_engine::alloc_agents(uint64_t capacity)
{
    free_agents();
    cap_local = std::max<uint64_t>(capacity, 1);
    for (int i = 0; i < 2; ++i) {
        SM_CUDA(cudaMalloc(&agents[i], cap_local * sizeof(float4)));
        SM_CUDA(cudaMalloc(&ids[i], cap_local * sizeof(uint32_t)));
    }
    acur = 0;
    return SM_OK;
}
void sm_engine::free_agents()
{
    for (int i = 0; i < 2; ++i) {
        if (agents[i]) cudaFree(agents[i]);
        if (ids[i]) cudaFree(ids[i]);
        agents[i] = nullptr; ids[i] = nullptr;
    }
}
// ---- texture-gather sampler: block-linear copy of the trail -------------------------------------
static int make_tex(cudaArray_t arr, cudaTextureObject_t* tex)
{
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    SM_CUDA(cudaCreateTextureObject(tex, &rd, &td, nullptr));
    return SM_OK;
}

// Does tex2Dgather return the footprint in the component order FetchTex assumes?
static int gather_probe_ok(bool* ok)
{
    *ok = false;
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    float* out = nullptr;
    float host[16], got[4] = {0, 0, 0, 0};
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) host[y * 4 + x] = 10.0f * y + x;
    cudaError_t err = cudaMallocArray(&arr, &fd, 4, 4, cudaArrayTextureGather);
    if (err == cudaSuccess) err = cudaMemcpy2DToArray(arr, 0, 0, host, 16, 16, 4, cudaMemcpyHostToDevice);
    int rc = SM_OK;
    if (err == cudaSuccess) rc = make_tex(arr, &tex);
    if (err == cudaSuccess && rc == SM_OK) err = cudaMalloc(&out, 16);
    if (err == cudaSuccess && rc == SM_OK) {
        smk::k_gather_probe<<<1, 1>>>(tex, out);
        err = cudaMemcpy(got, out, 16, cudaMemcpyDeviceToHost);
    }
    if (tex) cudaDestroyTextureObject(tex);
    if (arr) cudaFreeArray(arr);
    if (out) cudaFree(out);
    if (err != cudaSuccess) { cudaGetLastError(); return SM_OK; }      // probe unavailable -> LDG sampler
    *ok = (got[0] == 11.0f && got[1] == 12.0f && got[2] == 21.0f && got[3] == 22.0f);
    return rc;
}

int sm_engine::setup_tex()
{
    free_tex();
    const size_t htot = rows + 2 * (size_t)(ghost + pad_rows);
    int gw = 0, gh = 0;       // cudaArrayTextureGather arrays are limited (32768 x 32768 on sm_100)
    SM_CUDA(cudaDeviceGetAttribute(&gw, cudaDevAttrMaxTexture2DGatherWidth, device));
    SM_CUDA(cudaDeviceGetAttribute(&gh, cudaDevAttrMaxTexture2DGatherHeight, device));
    if ((size_t)W > (size_t)gw || htot > (size_t)gh)
        return sm_fail(SM_ERR_BAD_ARG, "map %u x %zu exceeds the texture-gather limit %d x %d", W, htot, gw, gh);
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
    SM_CUDA(cudaMallocArray(&trail_arr, &fd, W, htot, cudaArraySurfaceLoadStore | cudaArrayTextureGather));
    SM_TRY(make_tex(trail_arr, &trail_tex));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = trail_arr;
    SM_CUDA(cudaCreateSurfaceObject(&trail_surf, &rd));
    arr_stale = true;
    return SM_OK;
}
void sm_engine::free_tex()
{
    if (trail_tex) { cudaDestroyTextureObject(trail_tex); trail_tex = 0; }
    if (trail_surf) { cudaDestroySurfaceObject(trail_surf); trail_surf = 0; }
    if (trail_arr) { cudaFreeArray(trail_arr); trail_arr = nullptr; }
}
int sm_engine::refresh_tex(int64_t local_row_begin, int64_t n_rows)
{
    if (!use_tex || n_rows <= 0) return SM_OK;
    const int64_t arr_row = local_row_begin + (int64_t)(ghost + pad_rows);
    SM_CUDA(cudaMemcpy2DToArrayAsync(trail_arr, 0, (size_t)arr_row, trail_ptr(cur) + local_row_begin * (int64_t)W,
                                     (size_t)W * 4, (size_t)W * 4, (size_t)n_rows, cudaMemcpyDeviceToDevice, stream));
    return SM_OK;
}

int sm_engine::refresh_tex_ghosts(uint32_t g)
{
    if (!use_tex || arr_stale || g == 0) return SM_OK;
    const int32_t off = (int32_t)(ghost + pad_rows);
    const uint64_t total = 2ull * g * W;
    const unsigned nb = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)num_sms * 4);
    smk::k_rows_to_surface<<<nb, 256, 0, stream>>>(trail_base[cur], trail_surf, W, off - (int32_t)g, off + (int32_t)rows, (int32_t)g);
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 1;
    return SM_OK;
}

int sm_engine::setup_tiles()
{
    if (tile_hist) { cudaFree(tile_hist); tile_hist = nullptr; }
    if (tile_sums) { cudaFree(tile_sums); tile_sums = nullptr; }
    tiles.shift_x = tuning.tile_shift_x ? std::min(tuning.tile_shift_x, 10u) : 3u;
    tiles.shift_y = tuning.tile_shift_y ? std::min(tuning.tile_shift_y, 10u) : 3u;
    tiles.W = W;
    tiles.rows = rows;
    tiles.row_base = (int64_t)row0;
    tiles.tiles_x = (W + (1u << tiles.shift_x) - 1) >> tiles.shift_x;
    tiles.tiles_y = (rows + (1u << tiles.shift_y) - 1) >> tiles.shift_y;
    n_tiles = (uint64_t)tiles.tiles_x * tiles.tiles_y;
    if (n_tiles >= (1ull << 31)) return sm_fail(SM_ERR_BAD_ARG, "too many sort tiles");
    n_scan_blocks = (uint32_t)((n_tiles + smk::kScanBlock * smk::kScanItems - 1) / (smk::kScanBlock * smk::kScanItems));
    SM_CUDA(cudaMalloc(&tile_hist, n_tiles * sizeof(uint32_t)));
    SM_CUDA(cudaMalloc(&tile_sums, (size_t)n_scan_blocks * sizeof(uint32_t)));
    return SM_OK;
}

// ---------------------------------------------------------------------------
// deposit representation
// ---------------------------------------------------------------------------
bool sm_engine::flag_mode() const
{
    if (no_flags) return false;
    if ((cfg.flags & SM_FLAG_GAUSSIAN_BLUR) && !gauss_fast_ok()) return false;   // the tile / two-pass Gaussian kernels only merge counts
    return trail_nonneg && params.pheromone_deposition_amount >= 1.0f;
}

// Both representations are all-zero while idle.  The buffer the last agents pass wrote is only
// zeroed by the NEXT trail pass of the same mode, so a mode change has to retire it explicitly.
int sm_engine::switch_deposit_mode(int mode)
{
    if (deposit_mode != 0 && deposit_mode != mode) {
        const size_t cells = field_cells();
        for (int i = 0; i < 2; ++i) {
            if (deposit_mode == 1) SM_CUDA(cudaMemsetAsync(counts_base[i], 0, cells * sizeof(uint32_t), stream));
            else SM_CUDA(cudaMemsetAsync(flags_base[i], 0, cells, stream));
        }
        // the neighbours write into these fields during their agent pass: nobody may start it before
        // every rank has finished wiping (all ranks switch at the same step)
        if (world > 1 && p2p) SM_TRY(p2p_barrier());
    }
    deposit_mode = mode;
    return SM_OK;
}

// ---------------------------------------------------------------------------
// kernels: launch helpers
// ---------------------------------------------------------------------------
int sm_engine::sort_agents()
{
    if (world > 1) SM_TRY(refresh_counters());
    if (n_local == 0) return SM_OK;
    SM_TRY(tic(2));
    SM_CUDA(cudaMemsetAsync(tile_hist, 0, n_tiles * sizeof(uint32_t), stream));
    smk::k_tile_hist<<<blocks_for(n_local, 256), 256, 0, stream>>>(agents[acur], ids[acur], n_local, tile_hist, tiles);
    smk::k_scan_block<<<n_scan_blocks, smk::kScanBlock, 0, stream>>>(tile_hist, tile_hist, tile_sums, (uint32_t)n_tiles);
    smk::k_scan_sums<<<1, 1024, 0, stream>>>(tile_sums, n_scan_blocks);
    smk::k_scan_add<<<n_scan_blocks, smk::kScanBlock, 0, stream>>>(tile_hist, tile_sums, (uint32_t)n_tiles);
    smk::k_tile_scatter<<<blocks_for(n_local, 256), 256, 0, stream>>>(agents[acur], ids[acur], n_local, tile_hist,
                                                                      agents[1 - acur], ids[1 - acur], tiles);
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 5;
    acur = 1 - acur;
    identity_order = false;
    if (world > 1) {                      // the scatter dropped the dead (migrated-away) slots
        n_local = n_live;
        SM_TRY(mark_tail_dead());
        SM_TRY(plan_split());             // boundary-first stepping: which slots are "interior" until the next sort
        SM_TRY(push_counters());          // (synchronises: the two slot offsets plan_split requested have arrived)
        finish_split();
    }
    SM_TRY(toc());
    return SM_OK;
}

// Boundary-first stepping on strips (exchange.cu: p2p_step_split).  The sort leaves the agents ordered by tile row, so the
// agents of the first / last `t` tile rows of the strip are the slots [0, a) and [b, n): a and b are two entries of the
// scanned tile histogram (after the scatter, tile_hist[k] is the END offset of tile k).  An agent that starts the sort
// interval at least `margin` rows away from both strip edges can, until the next sort, neither leave the strip, nor
// deposit outside it or into the boundary bands, nor sense a row of the boundary bands:
//   margin = band (rows the boundary trail pass rewrites) + g (sensing reach) + drift (rows it can move in sort_interval steps).
// Those agents are stepped by the single-GPU instantiation on the main stream while the boundary agents, the barriers,
// the boundary bands and the ghost-row push run beside them on the side stream.
int sm_engine::plan_split()
{
    split_valid = false;
    split_pending = false;
    if (!p2p || !overlap_ok() || tuning.serial_exchange || tuning.no_boundary_first) return SM_OK;
    const uint32_t band = overlap_band();
    if (band == 0) return SM_OK;
    const float vmax = fmaxf(fabsf(params.agent_speed_min), fabsf(params.agent_speed_max));
    const float sd = fabsf(params.agent_sensor_distance);
    if (!(vmax < 1.0e5f) || !(sd < 6.0e4f)) return SM_OK;
    const uint32_t drift = (uint32_t)ceilf(vmax * 0.016f * (float)(sort_interval + 1)) + 2u;
    const uint32_t margin = band + ((uint32_t)ceilf(sd) + 3u) + drift;
    const uint32_t th = 1u << tiles.shift_y;
    const uint32_t t = (margin + th - 1) / th;                    // tile rows per side
    if (2ull * t + 2 > tiles.tiles_y) return SM_OK;               // no interior worth a second launch
    // END offsets of the last tile of tile row t-1 and of tile row tiles_y-t-1
    const size_t ia = (size_t)t * tiles.tiles_x - 1, ib = (size_t)(tiles.tiles_y - t) * tiles.tiles_x - 1;
    SM_CUDA(cudaMemcpyAsync(split_host, tile_hist + ia, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    SM_CUDA(cudaMemcpyAsync(split_host + 1, tile_hist + ib, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    split_pending = true;
    return SM_OK;
}
void sm_engine::finish_split()
{
    if (!split_pending) return;
    split_pending = false;
    const uint64_t a = ((uint64_t)split_host[0] + 1023) / 1024 * 1024;      // whole CTAs of the boundary launch (kernels.cuh)
    const uint64_t b = split_host[1];
    if (b > a && b <= n_local && b - a >= 64 * 1024) {                       // a second launch has to be worth its latency
        split_a = a; split_b = b;
        split_valid = true;
    }
}

// What has to happen once per step before any agent kernel of the step is launched (main stream).
int sm_engine::prepare_agents()
{
    if (world > 1) n_local = n_upper;            // grid bound; the kernel reads the exact count on the device
    // The deposit mode of the step is a collective decision (a rank whose strip is empty still merges what its
    // neighbours deposited into it, and takes part in the barrier of a mode change)
    SM_TRY(switch_deposit_mode(flag_mode() ? 2 : 1));
    if (use_tex && arr_stale) {                       // the array lost track of trail[cur]: one full copy
        SM_TRY(refresh_tex(-(int64_t)(ghost + pad_rows), (int64_t)rows + 2 * (int64_t)(ghost + pad_rows)));
        arr_stale = false;
    }
    return SM_OK;
}

// part 0: every slot (one launch).  Strips, boundary-first stepping (exchange.cu: p2p_step_split):
// part 1: the interior slots [split_a, split_b) with the single-GPU instantiation -- agents that cannot leave the strip
//         or deposit outside it before the next sort;  part 2: everything else with the strip instantiation.
int sm_engine::launch_agents(int part, cudaStream_t st)
{
    if (!st) st = stream;
    if (part == 0) SM_TRY(prepare_agents());
    const bool flags = deposit_mode == 2;
    uint64_t first = 0, count = n_local;
    if (part == 1) { first = split_a; count = split_b - split_a; }
    if (part == 2) count = split_a + (n_local - split_b);
    if (count == 0) {
        // the trail pass accumulates its statistics from zero whether or not an agent kernel ran
        if (part != 2) SM_CUDA(cudaMemsetAsync(stats_dev, 0, sizeof(smk::StatsAcc), st));
        return SM_OK;
    }
    if (part != 2) SM_TRY(tic(0, st));
    smk::LeaverBufs lv{};
    const bool idx32 = (uint64_t)field_cells() < (1ull << 31);
    const int apt = smk::agents_per_thread_for(count, num_sms);
    const unsigned nb = blocks_for(count, 256u * (unsigned)apt);        // a CTA steps 256 * apt consecutive slots
    float4* a = agents[acur] + first;
    uint32_t* id = ids[acur] + first;
    void* dep = flags ? (void*)flags_ptr(ccur) : (void*)counts_ptr(ccur);
    const smd::AgentConsts ac = agent_consts();
    const bool multi = world > 1 && part != 1;
    if (multi && p2p) {
        // leavers and out-of-strip deposits go straight into the neighbours' HBM (CUDA IPC mappings)
        for (int d = 0; d < 2; ++d) {
            // I am the DOWN neighbour of `up` (d == 0) and the UP neighbour of `down` (d == 1)
            uint8_t* arr = peer[d].window + (d == 0 ? window_arrival_off[1] : window_arrival_off[0]);
            lv.send_count[d] = reinterpret_cast<unsigned long long*>(arr);
            lv.send_a[d] = reinterpret_cast<float4*>(arr + 16);
            lv.send_id[d] = reinterpret_cast<uint32_t*>(arr + 16 + mig_cap * sizeof(float4));
            const size_t row0_off = (size_t)(ghost + pad_rows) * W;
            lv.peer_dep[d] = flags ? (void*)(peer[d].flags8[ccur] + row0_off) : (void*)(peer[d].counts[ccur] + row0_off);
        }
        lv.rows_up = (int32_t)peer[0].rows;
        lv.overflow = dev_counters + 2;
        lv.left_count = dev_counters + 3;
        lv.slots_in_use = dev_counters;
        lv.cap = (uint32_t)mig_cap;
    } else if (multi) {
        for (int d = 0; d < 2; ++d) {
            lv.send_count[d] = reinterpret_cast<unsigned long long*>(mig[d].send);
            lv.send_a[d] = reinterpret_cast<float4*>(mig[d].send + 16);
            lv.send_id[d] = reinterpret_cast<uint32_t*>(mig[d].send + 16 + mig_cap * sizeof(float4));
        }
        lv.overflow = dev_counters + 2;
        lv.slots_in_use = dev_counters;
        lv.cap = (uint32_t)mig_cap;
    }
    uint64_t n_arg = count;
    if (part == 2) {
        // launch indices [0, split_a) are slots [0, split_a); indices beyond map to slots >= split_b.  split_a is a multiple
        // of 1024 (sort_agents), so no CTA's 256 * apt consecutive indices straddle the skipped range
        lv.split = split_a;
        lv.skip = split_b - split_a;
        n_arg = n_local;                                                   // the kernel compares SLOT numbers with it
    }
    auto launch = [&](auto fetch, auto idx_tag) {
        using F = decltype(fetch);
        using I = decltype(idx_tag);
        if (multi && p2p) {
            if (flags && flags_tiled()) smk::k_agents<smk::XM_P2P, I, F, 2><<<nb, 256, 0, st>>>(a, id, n_arg, fetch, dep, ac, lv, (smk::StatsAcc*)stats_dev, apt);
            else if (flags) smk::k_agents<smk::XM_P2P, I, F, true><<<nb, 256, 0, st>>>(a, id, n_arg, fetch, dep, ac, lv, (smk::StatsAcc*)stats_dev, apt);
            else smk::k_agents<smk::XM_P2P, I, F, false><<<nb, 256, 0, st>>>(a, id, n_arg, fetch, dep, ac, lv, (smk::StatsAcc*)stats_dev, apt);
        } else if (multi) {
            if (flags) smk::k_agents<smk::XM_NCCL, I, F, true><<<nb, 256, 0, st>>>(a, id, n_arg, fetch, dep, ac, lv, (smk::StatsAcc*)stats_dev, apt);
            else smk::k_agents<smk::XM_NCCL, I, F, false><<<nb, 256, 0, st>>>(a, id, n_arg, fetch, dep, ac, lv, (smk::StatsAcc*)stats_dev, apt);
        } else {
            if (flags && flags_tiled()) smk::k_agents<smk::XM_SINGLE, I, F, 2><<<nb, 256, 0, st>>>(a, id, n_arg, fetch, dep, ac, lv, (smk::StatsAcc*)stats_dev, apt);
            else if (flags) smk::k_agents<smk::XM_SINGLE, I, F, true><<<nb, 256, 0, st>>>(a, id, n_arg, fetch, dep, ac, lv, (smk::StatsAcc*)stats_dev, apt);
            else smk::k_agents<smk::XM_SINGLE, I, F, false><<<nb, 256, 0, st>>>(a, id, n_arg, fetch, dep, ac, lv, (smk::StatsAcc*)stats_dev, apt);
        }
    };
    if (use_tex) {
        const smk::FetchTex f{trail_tex, (float)((int32_t)(ghost + pad_rows) - (int32_t)row0 + 1)};
        if (idx32) launch(f, int32_t{});
        else launch(f, int64_t{});
    } else if (idx32) {
        launch(smd::FetchLinear<int32_t, smk::LdgF32>{trail_ptr(cur), (int32_t)W, (int32_t)row0, smk::LdgF32()}, int32_t{});
    } else {
        launch(smd::FetchLinear<int64_t, smk::LdgF32>{trail_ptr(cur), (int64_t)W, (int64_t)row0, smk::LdgF32()}, int64_t{});
    }
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 1;
    if (part != 2) SM_TRY(toc(st));
    return SM_OK;
}

// What one trail pass reads and writes (fixed for the pass; the launch may be split into row bands).
int sm_engine::trail_plan(bool has_counts, TrailPass& p)
{
    p.tc = trail_consts();
    p.g = smk::TrailGeom{};
    p.g.W = W; p.g.rows = rows; p.g.wrap_y = (world == 1) ? 1 : 0;
    p.g.y_first = 0; p.g.y_last = rows; p.g.chunks1 = 0xFFFFFFFFu; p.g.y_first2 = p.g.y_last2 = 0;
    // the full step keeps the sampler's block-linear copy in step; other passes just mark it stale
    const bool write_surf = use_tex && has_counts && (!(cfg.flags & SM_FLAG_GAUSSIAN_BLUR) || gauss_fast_ok());
    p.g.surf = write_surf ? trail_surf : 0;
    p.g.surf_row0 = (int)(ghost + pad_rows);
    p.g.surf_pairs = surf_pairs ? 1 : 0;
    if (!write_surf) arr_stale = true;
    p.tin = trail_ptr(cur);
    p.tout = trail_ptr(1 - cur);
    // deposit representation the agents pass of this step used (none for diffusion-only)
    p.cm = !has_counts ? smk::CM_NONE : (deposit_mode == 2 ? (flags_tiled() ? smk::CM_FLAGS_TILED : smk::CM_FLAGS) : smk::CM_COUNTS);
    const bool cm_flags = p.cm == smk::CM_FLAGS || p.cm == smk::CM_FLAGS_TILED;
    p.cin = cm_flags ? (const void*)flags_ptr(ccur) : (const void*)counts_ptr(ccur);
    p.czero = cm_flags ? (void*)flags_ptr(1 - ccur) : (void*)counts_ptr(1 - ccur);
    p.fast = trail_rows_kernel_ok();
    p.stats = has_counts && p.fast && stats_interest > 0;   // fused statistics only while the host keeps asking for them
    p.g.rows_per_chunk = p.fast ? trail_rows_per_chunk(has_counts) : 1u;
    return SM_OK;
}

uint32_t sm_engine::trail_rows_per_chunk(bool has_counts) const
{
    // 8-16 rows per chunk measured best from 4096^2 to 32768^2 (profiles/README.md): the 2/rpc halo
    // re-reads hit L2; small maps take shorter chunks to keep >= 4 CTAs per SM
    uint64_t rpc = 8;
    // diffusion-only passes on maps far beyond the L2: 16 rows per chunk measured 4-6 % faster from 16384^2 up (halo rows
    // are 2/16 instead of 2/8 of the reads, and they no longer hit L2); 32768^2: 0.789 -> 0.835 of nominal 8 TB/s
    if (!has_counts && (uint64_t)W * rows >= (1ull << 28)) rpc = 16;
    const unsigned bx = blocks_for(W / 4, 128);
    while (rpc > 4 && (uint64_t)bx * ((rows + rpc - 1) / rpc) < (uint64_t)num_sms * 4) rpc /= 2;
    if (rpc_override > 0) rpc = rpc_override;
    return (uint32_t)rpc;
}

// Rows [y_first, y_last) of the pass on stream `st` (fast kernel only).
int sm_engine::trail_launch_rows(const TrailPass& p, uint32_t y_first, uint32_t y_last, cudaStream_t st,
                                 uint32_t y_first2, uint32_t y_last2)
{
    if (y_first >= y_last) return SM_OK;
    smk::TrailGeom g = p.g;
    const unsigned bs = 128;
    const uint32_t rpc = g.rows_per_chunk;
    g.y_first = y_first; g.y_last = y_last;
    g.chunks1 = (y_last - y_first + rpc - 1) / rpc;
    g.y_first2 = y_first2; g.y_last2 = y_last2;
    const uint32_t chunks2 = y_last2 > y_first2 ? (y_last2 - y_first2 + rpc - 1) / rpc : 0u;
    dim3 grid(blocks_for(W / 4, bs), (unsigned)(g.chunks1 + chunks2));
    if (p.cm == smk::CM_FLAGS_TILED && !((rpc == 4 || rpc == 8) && y_first % rpc == 0 && y_last % rpc == 0 &&
                                         (chunks2 == 0 || (y_first2 % rpc == 0 && y_last2 % rpc == 0))))
        return sm_fail(SM_ERR_STATE, "internal: tiled deposit flags need whole chunks of 4 or 8 rows");
    smk::StatsAcc* acc = (smk::StatsAcc*)stats_dev;
    auto go = [&](auto cm_tag, auto surf_tag, auto stats_tag) {
        constexpr int CMv = decltype(cm_tag)::value;
        constexpr bool SURFv = decltype(surf_tag)::value, STATSv = decltype(stats_tag)::value;
        smk::k_trail_rows<CMv, SURFv, 4, STATSv><<<grid, bs, 0, st>>>(p.tin, CMv == smk::CM_NONE ? nullptr : p.cin,
                                                                       CMv == smk::CM_NONE ? nullptr : p.czero, p.tout, g, p.tc, acc);
    };
    using std::integral_constant;
    using T = std::true_type; using F = std::false_type;
    auto go_cm = [&](auto cm_tag) {
        if (!g.surf) { if (p.stats) go(cm_tag, F{}, T{}); else go(cm_tag, F{}, F{}); }
        else         { if (p.stats) go(cm_tag, T{}, T{}); else go(cm_tag, T{}, F{}); }
    };
    if (p.cm == smk::CM_NONE) go(integral_constant<int, smk::CM_NONE>{}, F{}, F{});
    else if (p.cm == smk::CM_COUNTS) go_cm(integral_constant<int, smk::CM_COUNTS>{});
    else if (p.cm == smk::CM_FLAGS_TILED) go_cm(integral_constant<int, smk::CM_FLAGS_TILED>{});
    else go_cm(integral_constant<int, smk::CM_FLAGS>{});
    SM_CUDA(cudaGetLastError());
    timing.kernel_launches += 1;
    return SM_OK;
}

void sm_engine::trail_done(bool has_counts)
{
    stats_fused_valid = false;    // the callers that ran a full-step k_trail_rows pass set it again
    cur = 1 - cur;
    if (has_counts) ccur = 1 - ccur;
    trail_nonneg = true;          // decay clamps at 0 (NaN included), the mix of non-negatives is non-negative
}

int sm_engine::launch_trail(bool has_counts)
{
    TrailPass p;
    SM_TRY(trail_plan(has_counts, p));
    SM_TRY(tic(1));
    if (cfg.flags & SM_FLAG_GAUSSIAN_BLUR) {
        SM_TRY(launch_gauss(has_counts, p));
    } else if (p.fast) {
        SM_TRY(trail_launch_rows(p, 0, rows, stream));
    } else {
        for (uint32_t y0 = 0; y0 < rows; y0 += 32768) {
            uint32_t ny = std::min<uint32_t>(32768, rows - y0);
            dim3 grid(blocks_for(W, 256), ny);
            if (p.cm == smk::CM_COUNTS) smk::k_trail_generic<smk::CM_COUNTS><<<grid, 256, 0, stream>>>(p.tin, p.cin, p.czero, p.tout, p.g, p.tc, (int64_t)y0);
            else if (p.cm == smk::CM_FLAGS) smk::k_trail_generic<smk::CM_FLAGS><<<grid, 256, 0, stream>>>(p.tin, p.cin, p.czero, p.tout, p.g, p.tc, (int64_t)y0);
            else smk::k_trail_generic<smk::CM_NONE><<<grid, 256, 0, stream>>>(p.tin, nullptr, nullptr, p.tout, p.g, p.tc, (int64_t)y0);
            timing.kernel_launches += 1;
        }
    }
    SM_CUDA(cudaGetLastError());
    SM_TRY(toc());
    trail_done(has_counts);
    stats_fused_valid = p.stats;
    if (has_counts && stats_interest) --stats_interest;
    return SM_OK;
}

// SM_FLAG_SEM_INPLACE: the reference's three dispatches on one live buffer (header: opt-in, statistics-only parity).
int sm_engine::step_inplace()
{
    float* t = trail_ptr(cur);
    const smd::AgentConsts ac = agent_consts();
    const smd::TrailConsts tc = trail_consts();
    if (n_local) {
        SM_TRY(tic(0));
        if ((uint64_t)field_cells() < (1ull << 31))
            smk::k_agents_inplace<int32_t><<<blocks_for(n_local, 256), 256, 0, stream>>>(agents[acur], ids[acur], n_local, t, ac, tc.dep);
        else
            smk::k_agents_inplace<int64_t><<<blocks_for(n_local, 256), 256, 0, stream>>>(agents[acur], ids[acur], n_local, t, ac, tc.dep);
        SM_TRY(toc());
    }
    SM_TRY(tic(1));
    const uint64_t cells = (uint64_t)rows * W;
    smk::k_decay_inplace<<<(unsigned)std::min<uint64_t>((cells + 255) / 256, (uint64_t)num_sms * 32), 256, 0, stream>>>(t, cells, tc.decay_sub);
    smk::k_diffuse_inplace<<<dim3(blocks_for(W, 16), blocks_for(rows, 16)), 256, 0, stream>>>(t, W, rows, tc);
    SM_CUDA(cudaGetLastError());
    SM_TRY(toc());
    timing.kernel_launches += n_local ? 3 : 2;
    arr_stale = true;                   // this mode never samples the block-linear copy
    stats_fused_valid = false;
    trail_nonneg = true;
    frame_pre_valid = false;            // the frame shows the field as it stands
    steps_since_sort++;
    timing.steps++;
    return SM_OK;
}

// One frame of src/main.rs:1163-1235: (sort) -> agents -> decay + diffuse (+ the strip exchange).
int sm_engine::step_once()
{
    if (cfg.flags & SM_FLAG_SEM_INPLACE) {
        if (sort_interval && steps_since_sort >= sort_interval) {
            SM_TRY(sort_agents());        // locality only: the jitter hash uses the persistent index
            steps_since_sort = 0;
        }
        return step_inplace();
    }
    if (world > 1 && ghost_stale) SM_TRY(exchange_trail_ghosts());
    if (sort_interval && steps_since_sort >= sort_interval) {
        SM_TRY(sort_agents());
        steps_since_sort = 0;
    }
    if (world > 1 && p2p && split_valid && overlap_ok()) {
        SM_TRY(p2p_step_split());                        // boundary agents + exchange beside the interior agents
        steps_since_sort++;
        timing.steps++;
        frame_pre_valid = true;
        frame_tc = trail_consts();
        return SM_OK;
    }
    SM_TRY(launch_agents());                             // src/main.rs:1164-1181
    if (world > 1 && p2p && overlap_ok()) {
        SM_TRY(p2p_trail_overlapped());                  // trail pass with the exchange hidden behind its interior rows
    } else {
        if (world > 1) SM_TRY(p2p ? p2p_after_agents() : exchange_counts());
        SM_TRY(launch_trail(true));                      // src/main.rs:1184-1199 + 1220-1235
        if (world > 1) SM_TRY(p2p ? p2p_after_trail() : migrate_agents());
    }
    steps_since_sort++;
    timing.steps++;
    frame_pre_valid = true;
    frame_tc = trail_consts();
    return SM_OK;
}

// ---------------------------------------------------------------------------
// CUDA-graph replay of whole sort periods (single GPU)
// ---------------------------------------------------------------------------
// The reference submits three command buffers per frame (main.rs:1181, 1199, 1235); this engine launches two kernels per
// step, and on small maps (BASELINE config 1: 1 M agents on 1920 x 1080, both kernels a few microseconds) the step is bound
// by launch latency, not by the kernels.  sm_step(n) therefore replays a captured graph of one full period of the host
// state machine -- 2 x sort_interval steps including their two cell sorts (after which the ping-pong indices of the trail,
// the deposit buffers and the agent arrays are back where they started), or 2 steps when sorting is off -- whenever the
// engine is in steady state: same parameters, same buffers, nothing pending (no stale sampler copy, no deposit-mode
// switch, no statistics request, no per-kernel timing).  Anything else runs the ordinary launch path; results are the
// same launches in the same order either way.
uint32_t sm_engine::graph_period() const
{
    if (!graph_enabled || world != 1 || timing_enabled || stats_interest || (cfg.flags & (SM_FLAG_GAUSSIAN_BLUR | SM_FLAG_SEM_INPLACE))) return 0;
    return sort_interval ? 2u * sort_interval : 2u;
}

sm_engine::GraphKey sm_engine::graph_key_now() const
{
    GraphKey k;
    memset(&k, 0, sizeof k);                  // compared with memcmp: padding must be defined
    k.params = params;
    k.n_local = n_local;
    k.acur = acur; k.cur = cur; k.ccur = ccur; k.deposit_mode = deposit_mode;
    k.sort_interval = sort_interval;
    k.agents0 = agents[0]; k.trail0 = trail_base[0]; k.arr = trail_arr;
    k.use_tex = use_tex;
    k.rpc_override = rpc_override; k.force_generic = force_generic; k.no_flags = no_flags;
    return k;
}

// Steady state, and aligned with the period (the next step starts with a sort)?
bool sm_engine::graph_ready()
{
    if (n_local == 0 || !agents_valid) return false;
    if (sort_interval && steps_since_sort < sort_interval) return false;
    if (use_tex && arr_stale) return false;
    if (deposit_mode != (flag_mode() ? 2 : 1)) return false;       // the next agent pass would switch (and wipe) buffers
    return true;
}

int sm_engine::graph_steps()
{
    const uint32_t period = graph_period();
    const GraphKey now = graph_key_now();
    if (step_graph && memcmp(&now, &step_graph_key, sizeof now) != 0) {
        cudaGraphExecDestroy(step_graph);
        step_graph = nullptr;
    }
    if (!step_graph) {
        // capture = run the ordinary host path once with the stream in capture mode (nothing executes yet)
        const uint64_t launches0 = timing.kernel_launches;
        SM_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        int rc = SM_OK;
        for (uint32_t i = 0; i < period && rc == SM_OK; ++i) rc = step_once();
        cudaGraph_t g = nullptr;
        const cudaError_t err = cudaStreamEndCapture(stream, &g);
        if (rc != SM_OK) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }
        if (err != cudaSuccess) { if (g) cudaGraphDestroy(g); return sm_fail(SM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(err)); }
        const cudaError_t ierr = cudaGraphInstantiate(&step_graph, g, 0);
        cudaGraphDestroy(g);
        if (ierr != cudaSuccess) { step_graph = nullptr; return sm_fail(SM_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ierr)); }
        step_graph_launches = timing.kernel_launches - launches0;
        memcpy(&step_graph_key, &now, sizeof now);
        const GraphKey after = graph_key_now();                     // the period must bring the host state back
        if (memcmp(&after, &now, sizeof now) != 0 || (sort_interval && steps_since_sort != sort_interval)) {
            cudaGraphExecDestroy(step_graph);
            step_graph = nullptr;
            graph_enabled = false;
            return sm_fail(SM_ERR_STATE, "internal: a captured period did not restore the engine's launch state");
        }
    } else {
        // replay: the host state after a period equals the state before it; only the counters move
        timing.steps += period;
        timing.kernel_launches += step_graph_launches;
        frame_pre_valid = true;
        frame_tc = trail_consts();
    }
    SM_CUDA(cudaGraphLaunch(step_graph, stream));
    return SM_OK;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

const char* sm_last_error(void) { return g_sm_err.c_str(); }

void sm_version(int* major, int* minor)
{
    if (major) *major = SM_VERSION_MAJOR;
    if (minor) *minor = SM_VERSION_MINOR;
}

int sm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

static int check_device(int device, int* num_sms)
{
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n == 0) {
        cudaGetLastError();
        return sm_fail(SM_ERR_NO_DEVICE, "no CUDA device available (%s); this engine has no CPU fallback",
                       err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    }
    if (device < 0 || device >= n) return sm_fail(SM_ERR_BAD_ARG, "device %d out of range (0..%d)", device, n - 1);
    int major = 0, minor = 0;
    SM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    SM_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10)
        return sm_fail(SM_ERR_NO_DEVICE, "device %d is sm_%d%d; this library contains sm_100a code only", device, major, minor);
    SM_CUDA(cudaSetDevice(device));
    if (num_sms) SM_CUDA(cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, device));
    return SM_OK;
}

int sm_create(sm_engine** out, const sm_config* cfg)
{
    if (!out || !cfg) return sm_fail(SM_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (cfg->width == 0 || cfg->height == 0) return sm_fail(SM_ERR_BAD_ARG, "map size must be non-zero");
    if (cfg->width > 65536 || cfg->height > 65536) return sm_fail(SM_ERR_BAD_ARG, "map size above 65536 is not supported");
    if (cfg->agent_count >= (1ull << 31)) return sm_fail(SM_ERR_BAD_ARG, "agent_count must be < 2^31 (i32 index of compute.wgsl:60)");
    if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size)
        return sm_fail(SM_ERR_BAD_ARG, "bad rank/world_size %d/%d", cfg->rank, cfg->world_size);
    if ((uint32_t)cfg->world_size > cfg->height) return sm_fail(SM_ERR_BAD_ARG, "more strips than rows");
    if ((cfg->flags & SM_FLAG_SEM_INPLACE) && (cfg->world_size != 1 || (cfg->flags & SM_FLAG_GAUSSIAN_BLUR)))
        return sm_fail(SM_ERR_BAD_ARG, "SM_FLAG_SEM_INPLACE is single-GPU, 3x3 box only");
    int num_sms = 0;
    SM_TRY(check_device(cfg->device, &num_sms));

    sm_engine* e = new sm_engine();
    e->cfg = *cfg;
    e->device = cfg->device;
    e->num_sms = num_sms;
    e->W = cfg->width; e->H = cfg->height;
    e->rank = cfg->rank; e->world = cfg->world_size;
    // strip r owns rows [r*H/G, (r+1)*H/G)
    e->row0 = (uint32_t)(((uint64_t)e->rank * e->H) / e->world);
    uint32_t row1 = (uint32_t)(((uint64_t)(e->rank + 1) * e->H) / e->world);
    e->rows = row1 - e->row0;
    e->ghost = 0;
    if (e->world > 1) {
        uint32_t want = cfg->ghost_rows ? cfg->ghost_rows : 232u;   // >= ceil(225)+2 (Snake/Mesh presets) + slack
        uint32_t min_rows = e->H / e->world;                    // thinnest strip
        e->ghost = std::min(want, min_rows);
        e->pad_rows = 16;
    }
    e->n_global = cfg->agent_count;
    // config 2: 16 -> 230.1, 24 -> 225.4, 32 -> 225.7, 48 -> 225.1 us/step (the agent kernel slows down as the order decays)
    e->sort_interval = cfg->sort_interval ? cfg->sort_interval : 24u;
    if (cfg->flags & SM_FLAG_NO_SORT) e->sort_interval = 0;
    if (e->world > 1 && e->sort_interval == 0) e->sort_interval = 16;   // strips need the sort to compact migrated-away slots
    // measurement switches: sm_config.tuning, all zero by default (no environment variable is read here)
    const sm_tuning& tn = cfg->tuning;
    e->tuning = tn;
    {
        bool probe = false;
        if (tn.sampler == 0) { int prc = gather_probe_ok(&probe); if (prc != SM_OK) { delete e; return prc; } }
        e->use_tex = tn.sampler == 0 && probe;
    }
    e->force_generic = tn.generic_trail_kernel != 0;
    e->no_flags = tn.deposit_counts_only != 0;
    e->rpc_override = (int)tn.trail_rows_per_chunk;
    e->surf_pairs = tn.surface_row_writes == 0;
    e->graph_enabled = tn.no_step_graph == 0;
    // Gaussian extension: the register-streaming kernel (gauss_rows.cuh) up to radius gauss_rows_max_r (5), the shared-memory
    // streaming kernel (gauss_stream.cuh) above, the tile kernel (k_gauss_fused) for maps the first two do not take (W % 4 != 0,
    // tiny).  Measured on 8192^2-16384^2 (profiles/), fraction of the HBM peak: rows 0.89-0.95 (R 1-2), 0.78-0.85 (R 3-4),
    // 0.65-0.69 (R 5); stream 0.51-0.69; tile 0.31-0.61.  rows and stream let a Gaussian full step keep the u8 deposit flags
    // and the sampler copy (6.8e10 vs 4.4e10 agent-steps/s).
    e->gauss_two_pass = tn.gauss_kernel == 4;
    e->gauss_stream = tn.gauss_kernel != 3 && tn.gauss_kernel != 4;
    e->gauss_rows = tn.gauss_kernel == 0 || tn.gauss_kernel == 1;
    e->gauss_rows_max_r = tn.gauss_kernel == 1 ? 5 : (tn.gauss_rows_max_radius ? (int)std::min(tn.gauss_rows_max_radius, 5u) : 5);
    e->gauss_rows_packing = (int)tn.gauss_rows_packing;

    // defaults = Settings::default(), /root/reference/src/settings.rs:8-27
    sm_params p{};
    p.width = e->W; p.height = e->H;
    p.decay_factor = 10.0f; p.agent_jitter = 0.0f;
    p.agent_speed_min = 30.0f; p.agent_speed_max = 50.0f;
    p.agent_turn_speed = 0.43f; p.agent_sensor_angle = 0.3f; p.agent_sensor_distance = 20.0f;
    p.diffusion_rate = 1.0f; p.pheromone_deposition_amount = 1.0f;
    p.blur_radius = 2.0f; p.blur_sigma = 1.0f;
    e->params = p;

    auto fail = [&](int rc) { sm_destroy(e); return rc; };
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(sm_fail(SM_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError())));
    int rc;
    if ((rc = e->alloc_trail()) != SM_OK) return fail(rc);
    uint64_t cap = e->n_global;
    // strips: twice the even share plus slack; never more than every agent plus slack -- the slack is what arrivals need
    // between two sorts even on a rank that already holds everybody (slots of departed agents are only reclaimed by the sort)
    if (e->world > 1) cap = std::min<uint64_t>(e->n_global, 2 * ((e->n_global + e->world - 1) / e->world)) + (1u << 20);
    if ((rc = e->alloc_agents(cap)) != SM_OK) return fail(rc);
    if ((rc = e->setup_tiles()) != SM_OK) return fail(rc);
    if (cudaMalloc(&e->stats_dev, sizeof(smk::StatsAcc)) != cudaSuccess ||
        cudaMallocHost(&e->stats_host, sizeof(smk::StatsAcc)) != cudaSuccess)       // pinned: the per-frame read-back is a 32-byte DMA
        return fail(sm_fail(SM_ERR_OOM, "cudaMalloc(stats) failed"));
    e->n_local = 0;
    e->agents_valid = false;
    if (e->world > 1 && tn.debug_single_rank_strip) {
        if ((rc = e->fake_comm_init()) != SM_OK) return fail(rc);
    }
    if (cudaStreamSynchronize(e->stream) != cudaSuccess)
        return fail(sm_fail(SM_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(cudaGetLastError())));
    *out = e;
    return SM_OK;
}

int sm_destroy(sm_engine* e)
{
    if (!e) return SM_OK;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->comm_destroy();
    e->free_trail();
    e->free_agents();
    if (e->tile_hist) cudaFree(e->tile_hist);
    if (e->tile_sums) cudaFree(e->tile_sums);
    if (e->stats_dev) cudaFree(e->stats_dev);
    if (e->stats_host) cudaFreeHost(e->stats_host);
    if (e->lut_dev) cudaFree(e->lut_dev);
    if (e->frame_dev) cudaFree(e->frame_dev);
    if (e->step_graph) cudaGraphExecDestroy(e->step_graph);
    for (auto& p : e->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return SM_OK;
}

#define SM_ENTER(e)                                                    \
    if (!(e)) return sm_fail(SM_ERR_BAD_ARG, "null engine handle");    \
    SM_CUDA(cudaSetDevice((e)->device))

int sm_set_params(sm_engine* e, const sm_params* p)
{
    SM_ENTER(e);
    if (!p) return sm_fail(SM_ERR_BAD_ARG, "null params");
    if (p->width != e->W || p->height != e->H)
        return sm_fail(SM_ERR_BAD_ARG, "params size %ux%u does not match the engine's map %ux%u (use sm_resize)",
                       p->width, p->height, e->W, e->H);
    if (e->world > 1) {
        // ghost depth needed for sensing: taps reach floor(x + sd*cos) + 1
        float sd = fabsf(p->agent_sensor_distance);
        if (!(sd + 3.0f <= (float)e->ghost))
            return sm_fail(SM_ERR_BAD_ARG, "sensor distance %g needs %d ghost rows; strips provide %u", sd, (int)ceilf(sd) + 3, e->ghost);
    }
    const sm_params before = e->params;
    e->params = *p;
    if (e->world > 1 && memcmp(&before, p, sizeof before) != 0) {
        e->split_valid = false;                          // the interior / boundary split was planned for the old speeds and reach
        e->steps_since_sort = e->sort_interval;          // (every rank gets the same call: the sort stays collective)
    }
    if (const int rc = e->check_gauss(false); rc != SM_OK) {   // extension: a bad radius / sigma is refused here, not in the middle of a step
        e->params = before;
        return rc;
    }
    return SM_OK;
}

int sm_get_params(sm_engine* e, sm_params* p)
{
    SM_ENTER(e);
    if (!p) return sm_fail(SM_ERR_BAD_ARG, "null params");
    *p = e->params;
    return SM_OK;
}

uint64_t sm_agent_count(sm_engine* e) { return e ? e->n_global : 0; }
uint64_t sm_local_agent_count(sm_engine* e)
{
    if (!e) return 0;
    if (e->world > 1) { if (e->refresh_counters() != SM_OK) return 0; return e->n_live; }
    return e->n_local;
}


int sm_upload_agents(sm_engine* e, const float* xyas, uint64_t first, uint64_t n)
{
    SM_ENTER(e);
    if (!xyas && n) return sm_fail(SM_ERR_BAD_ARG, "null agent array");
    if (first + n > e->n_global) return sm_fail(SM_ERR_BAD_ARG, "agent range [%llu, %llu) exceeds agent_count %llu",
                                                (unsigned long long)first, (unsigned long long)(first + n), (unsigned long long)e->n_global);
    SM_CUDA(cudaStreamSynchronize(e->stream));
    if (e->world == 1) {
        if (!e->agents_valid && !(first == 0 && n == e->n_global)) {
            // first partial upload into uninitialised storage: zero-fill the rest deterministically
            SM_CUDA(cudaMemsetAsync(e->agents[e->acur], 0, e->n_global * sizeof(float4), e->stream));
            e->identity_order = false;
        }
        if (!e->identity_order) {
            if (e->agents_valid) SM_TRY(e->restore_identity_order());
            else SM_TRY(e->fill_identity_ids());
        }
        SM_CUDA(cudaMemcpyAsync(e->agents[e->acur] + first, xyas, n * sizeof(float4), cudaMemcpyHostToDevice, e->stream));
        SM_CUDA(cudaStreamSynchronize(e->stream));
        e->n_local = e->n_global;
        e->agents_valid = true;
        e->steps_since_sort = e->sort_interval;   // force a sort before the next step
        return SM_OK;
    }
    if (!(first == 0 && n == e->n_global))
        return sm_fail(SM_ERR_BAD_ARG, "multi-GPU upload must cover all agents (first=0, n=agent_count)");
    std::vector<float> keep;
    std::vector<uint32_t> keep_ids;
    keep.reserve((size_t)(n / e->world) * 4 + 1024);
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t r = owner_row(xyas[4 * i + 1], e->H);
        if (r >= e->row0 && r < e->row0 + e->rows) {
            keep.insert(keep.end(), xyas + 4 * i, xyas + 4 * i + 4);
            keep_ids.push_back((uint32_t)(first + i));
        }
    }
    uint64_t m = keep_ids.size();
    if (m > e->cap_local) return sm_fail(SM_ERR_OOM, "strip %d would own %llu agents, capacity %llu", e->rank,
                                         (unsigned long long)m, (unsigned long long)e->cap_local);
    SM_CUDA(cudaMemcpy(e->agents[e->acur], keep.data(), m * sizeof(float4), cudaMemcpyHostToDevice));
    SM_CUDA(cudaMemcpy(e->ids[e->acur], keep_ids.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice));
    e->n_local = m;
    e->n_live = m;
    if (e->comm_ready) { SM_TRY(e->mark_tail_dead()); SM_TRY(e->push_counters()); }
    e->agents_valid = true;
    e->identity_order = false;
    e->steps_since_sort = e->sort_interval;
    e->split_valid = false;
    return SM_OK;
}

int sm_download_agents(sm_engine* e, float* xyas, uint64_t first, uint64_t n, uint64_t* n_owned)
{
    SM_ENTER(e);
    if (!xyas && n) return sm_fail(SM_ERR_BAD_ARG, "null agent array");
    if (first + n > e->n_global) return sm_fail(SM_ERR_BAD_ARG, "agent range exceeds agent_count");
    if (!e->agents_valid) return sm_fail(SM_ERR_STATE, "agents were never initialised or uploaded");
    if (e->world == 1) {
        if (!e->identity_order) SM_TRY(e->restore_identity_order());
        SM_CUDA(cudaMemcpyAsync(xyas, e->agents[e->acur] + first, n * sizeof(float4), cudaMemcpyDeviceToHost, e->stream));
        SM_CUDA(cudaStreamSynchronize(e->stream));
        if (n_owned) *n_owned = n;
        return SM_OK;
    }
    SM_TRY(e->refresh_counters());
    std::vector<float> a((size_t)e->n_local * 4);
    std::vector<uint32_t> id((size_t)e->n_local);
    SM_CUDA(cudaMemcpy(a.data(), e->agents[e->acur], e->n_local * sizeof(float4), cudaMemcpyDeviceToHost));
    SM_CUDA(cudaMemcpy(id.data(), e->ids[e->acur], e->n_local * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    uint64_t cnt = 0;
    for (uint64_t i = 0; i < e->n_local; ++i) {
        uint64_t g = id[i];
        if (g >= first && g < first + n) {
            memcpy(xyas + 4 * (g - first), a.data() + 4 * i, 16);
            cnt++;
        }
    }
    if (n_owned) *n_owned = cnt;
    return SM_OK;
}

int sm_init_agents(sm_engine* e, uint64_t seed)
{
    SM_ENTER(e);
    const float Wf = (float)e->W, Hf = (float)e->H;
    if (e->world == 1) {
        if (e->n_global)
            smk::k_init_agents<<<blocks_for(e->n_global, 256), 256, 0, e->stream>>>(
                e->agents[e->acur], e->ids[e->acur], e->n_global, 0, seed, Wf, Hf,
                e->params.agent_speed_min, e->params.agent_speed_max);
        SM_CUDA(cudaGetLastError());
        e->n_local = e->n_global;
        e->identity_order = true;
    } else {
        SM_TRY(e->init_agents_strip(seed));
    }
    e->agents_valid = true;
    e->steps_since_sort = e->sort_interval;
    e->split_valid = false;
    return SM_OK;
}

int sm_set_agent_count(sm_engine* e, uint64_t n, uint64_t seed)
{
    SM_ENTER(e);
    if (n == 0) return sm_fail(SM_ERR_BAD_ARG, "agent_count must be >= 1 (src/main.rs:738-739)");
    if (n >= (1ull << 31)) return sm_fail(SM_ERR_BAD_ARG, "agent_count must be < 2^31");
    SM_CUDA(cudaStreamSynchronize(e->stream));
    e->n_global = n;
    e->cfg.agent_count = n;
    uint64_t cap = n;
    if (e->world > 1) cap = std::min<uint64_t>(n, 2 * ((n + e->world - 1) / e->world)) + (1u << 20);
    if (cap > e->cap_local) SM_TRY(e->alloc_agents(cap));
    e->agents_valid = false;
    return sm_init_agents(e, seed);
}

int sm_reassign_speeds(sm_engine* e, uint64_t seed)
{
    SM_ENTER(e);
    if (!e->agents_valid) return sm_fail(SM_ERR_STATE, "agents were never initialised or uploaded");
    if (e->world > 1) SM_TRY(e->refresh_counters());
    if (e->n_local)
        smk::k_reassign_speeds<<<blocks_for(e->n_local, 256), 256, 0, e->stream>>>(
            e->agents[e->acur], e->ids[e->acur], e->n_local, seed, e->params.agent_speed_min, e->params.agent_speed_max);
    SM_CUDA(cudaGetLastError());
    return SM_OK;
}

int sm_clear_trail(sm_engine* e)
{
    SM_ENTER(e);
    const size_t cells = e->field_cells();
    SM_CUDA(cudaMemsetAsync(e->trail_base[e->cur], 0, cells * sizeof(float), e->stream));
    e->ghost_stale = true;
    e->arr_stale = true;
    e->trail_nonneg = true;
    e->stats_fused_valid = false;
    e->frame_pre_valid = false;
    return SM_OK;
}

static int clip_rows(sm_engine* e, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, size_t pitch,
                     uint32_t* ya, uint32_t* yb)
{
    if ((uint64_t)x0 + w > e->W || (uint64_t)y0 + h > e->H) return sm_fail(SM_ERR_BAD_ARG, "rectangle outside the map");
    if (pitch < w) return sm_fail(SM_ERR_BAD_ARG, "pitch smaller than the rectangle width");
    *ya = std::max(y0, e->row0);
    *yb = std::min(y0 + h, e->row0 + e->rows);
    return SM_OK;
}

int sm_upload_trail(sm_engine* e, const float* src, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, size_t pitch)
{
    SM_ENTER(e);
    if (!src) return sm_fail(SM_ERR_BAD_ARG, "null source");
    uint32_t ya = 0, yb = 0;
    SM_TRY(clip_rows(e, x0, y0, w, h, pitch, &ya, &yb));
    // strips: the deposit representation of the next step must be the same on every rank, whichever strips the
    // rectangle touches -- every rank falls back to counts until the next trail pass (this used to sit
    // inside the intersection test, so ranks the rectangle missed kept flags and lost cross-strip deposits)
    if (e->world > 1) e->trail_nonneg = false;
    if (ya < yb && w) {
        if (e->trail_nonneg) {
            bool ok = true;
            for (uint32_t y = ya; y < yb && ok; ++y) {
                const float* rowp = src + (size_t)(y - y0) * pitch;
                for (uint32_t x = 0; x < w; ++x)
                    if (!(rowp[x] >= 0.0f)) { ok = false; break; }    // negative or NaN: counts mode until the next pass
            }
            e->trail_nonneg = ok;
        }
        float* dst = e->trail_ptr(e->cur) + (size_t)(ya - e->row0) * e->W + x0;
        SM_CUDA(cudaMemcpy2DAsync(dst, (size_t)e->W * 4, src + (size_t)(ya - y0) * pitch, pitch * 4, (size_t)w * 4,
                                  yb - ya, cudaMemcpyHostToDevice, e->stream));
        SM_CUDA(cudaStreamSynchronize(e->stream));
    }
    e->ghost_stale = true;
    e->arr_stale = true;
    e->stats_fused_valid = false;
    e->frame_pre_valid = false;
    return SM_OK;
}

int sm_download_trail(sm_engine* e, float* dst, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, size_t pitch)
{
    SM_ENTER(e);
    if (!dst) return sm_fail(SM_ERR_BAD_ARG, "null destination");
    uint32_t ya = 0, yb = 0;
    SM_TRY(clip_rows(e, x0, y0, w, h, pitch, &ya, &yb));
    if (ya < yb && w) {
        const float* src = e->trail_ptr(e->cur) + (size_t)(ya - e->row0) * e->W + x0;
        SM_CUDA(cudaMemcpy2DAsync(dst + (size_t)(ya - y0) * pitch, pitch * 4, src, (size_t)e->W * 4, (size_t)w * 4,
                                  yb - ya, cudaMemcpyDeviceToHost, e->stream));
    }
    SM_CUDA(cudaStreamSynchronize(e->stream));
    return SM_OK;
}

int sm_trail_statistics(sm_engine* e, sm_trail_stats* out)
{
    SM_ENTER(e);
    if (!out) return sm_fail(SM_ERR_BAD_ARG, "null output");
    e->stats_interest = 64;        // the next full-step passes reduce the statistics on the way (see k_trail_rows)
    if (!e->stats_fused_valid) {
        // the field was not produced by a full-step trail pass (upload, clear, diffusion-only, ...): one sweep over it
        SM_CUDA(cudaMemsetAsync(e->stats_dev, 0, sizeof(smk::StatsAcc), e->stream));
        const uint64_t cells = (uint64_t)e->rows * e->W;
        unsigned nb = (unsigned)std::min<uint64_t>((cells / 16 + 255) / 256 + 1, (uint64_t)e->num_sms * 16);
        smk::k_trail_stats<<<nb, 256, 0, e->stream>>>(e->trail_ptr(e->cur), cells, (smk::StatsAcc*)e->stats_dev);
        SM_CUDA(cudaGetLastError());
        e->timing.kernel_launches += 1;
    }
    SM_CUDA(cudaMemcpyAsync(e->stats_host, e->stats_dev, sizeof(smk::StatsAcc), cudaMemcpyDeviceToHost, e->stream));
    SM_CUDA(cudaStreamSynchronize(e->stream));
    const smk::StatsAcc h = *static_cast<const smk::StatsAcc*>(e->stats_host);
    out->sum = h.sum; out->sum_sq = h.sum_sq; out->nonzero = h.nonzero;
    memcpy(&out->max, &h.max_bits, 4);
    out->_pad = 0;
    return SM_OK;
}

// ---- snapshot / restore ---------------------------------------------------------------
// File: SnapshotHeader, sm_params (56 B), n_agents x u32 persistent index, n_agents x float4, rows x W f32.
struct SnapshotHeader {
    char magic[8];                 // "SLIMB200"
    uint32_t version;              // 1
    uint32_t width, height;
    uint32_t world, rank, row0, rows;
    uint32_t reserved;
    uint64_t n_global, n_agents;   // agent_count of the engine; agents owned by this rank (stored here)
    uint64_t steps;                // steps taken since creation (informational)
};

static std::string snapshot_path(const sm_engine* e, const char* path)
{
    std::string p(path);
    if (e->world > 1) p += ".rank" + std::to_string(e->rank);
    return p;
}

int sm_save_snapshot(sm_engine* e, const char* path)
{
    SM_ENTER(e);
    if (!path || !*path) return sm_fail(SM_ERR_BAD_ARG, "null snapshot path");
    if (!e->agents_valid) return sm_fail(SM_ERR_STATE, "agents were never initialised or uploaded");
    SM_TRY(e->refresh_counters());                     // syncs; multi-GPU: slots in use / live agents
    const uint64_t n_slots = e->world == 1 ? e->n_global : e->n_local;
    std::vector<float> a((size_t)n_slots * 4);
    std::vector<uint32_t> id((size_t)n_slots);
    SM_CUDA(cudaMemcpy(a.data(), e->agents[e->acur], n_slots * sizeof(float4), cudaMemcpyDeviceToHost));
    SM_CUDA(cudaMemcpy(id.data(), e->ids[e->acur], n_slots * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    uint64_t m = 0;                                    // drop the slots of agents that migrated away
    for (uint64_t i = 0; i < n_slots; ++i)
        if (id[i] != smk::kDeadAgent) {
            if (m != i) { id[m] = id[i]; memcpy(&a[4 * m], &a[4 * i], 16); }
            ++m;
        }
    std::vector<float> t((size_t)e->rows * e->W);
    SM_CUDA(cudaMemcpy(t.data(), e->trail_ptr(e->cur), t.size() * sizeof(float), cudaMemcpyDeviceToHost));
    SnapshotHeader h{};
    memcpy(h.magic, "SLIMB200", 8);
    h.version = 1; h.width = e->W; h.height = e->H;
    h.world = (uint32_t)e->world; h.rank = (uint32_t)e->rank; h.row0 = e->row0; h.rows = e->rows;
    h.n_global = e->n_global; h.n_agents = m; h.steps = e->timing.steps;
    const std::string fn = snapshot_path(e, path);
    FILE* f = fopen(fn.c_str(), "wb");
    if (!f) return sm_fail(SM_ERR_BAD_ARG, "cannot open %s for writing", fn.c_str());
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(&e->params, sizeof(sm_params), 1, f) == 1;
    ok = ok && (m == 0 || (fwrite(id.data(), sizeof(uint32_t), m, f) == m && fwrite(a.data(), 16, m, f) == m));
    ok = ok && fwrite(t.data(), sizeof(float), t.size(), f) == t.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) return sm_fail(SM_ERR_STATE, "short write to %s", fn.c_str());
    return SM_OK;
}

int sm_load_snapshot(sm_engine* e, const char* path)
{
    SM_ENTER(e);
    if (!path || !*path) return sm_fail(SM_ERR_BAD_ARG, "null snapshot path");
    if (e->world > 1 && !e->comm_ready) return sm_fail(SM_ERR_STATE, "multi-GPU engine: call sm_comm_init first");
    const std::string fn = snapshot_path(e, path);
    FILE* f = fopen(fn.c_str(), "rb");
    if (!f) return sm_fail(SM_ERR_BAD_ARG, "cannot open %s", fn.c_str());
    SnapshotHeader h{};
    sm_params p{};
    bool ok = fread(&h, sizeof h, 1, f) == 1 && fread(&p, sizeof p, 1, f) == 1;
    if (!ok || memcmp(h.magic, "SLIMB200", 8) != 0 || h.version != 1) { fclose(f); return sm_fail(SM_ERR_BAD_ARG, "%s is not a version-1 snapshot", fn.c_str()); }
    if (h.width != e->W || h.height != e->H || h.n_global != e->n_global || h.world != (uint32_t)e->world ||
        h.rank != (uint32_t)e->rank || h.row0 != e->row0 || h.rows != e->rows) {
        fclose(f);
        return sm_fail(SM_ERR_BAD_ARG, "snapshot %s is for a %ux%u map, %llu agents, rank %u of %u (rows %u+%u); this engine differs",
                       fn.c_str(), h.width, h.height, (unsigned long long)h.n_global, h.rank, h.world, h.row0, h.rows);
    }
    const uint64_t cap = e->world == 1 ? e->n_global : e->cap_local;
    if (h.n_agents > cap || (e->world == 1 && h.n_agents != e->n_global)) { fclose(f); return sm_fail(SM_ERR_BAD_ARG, "snapshot agent count does not fit this engine"); }
    std::vector<uint32_t> id((size_t)h.n_agents);
    std::vector<float> a((size_t)h.n_agents * 4), t((size_t)e->rows * e->W);
    ok = h.n_agents == 0 || (fread(id.data(), sizeof(uint32_t), h.n_agents, f) == h.n_agents && fread(a.data(), 16, h.n_agents, f) == h.n_agents);
    ok = ok && fread(t.data(), sizeof(float), t.size(), f) == t.size();
    fclose(f);
    if (!ok) return sm_fail(SM_ERR_BAD_ARG, "snapshot %s is truncated", fn.c_str());
    SM_CUDA(cudaStreamSynchronize(e->stream));
    SM_TRY(sm_set_params(e, &p));
    SM_CUDA(cudaMemcpy(e->agents[e->acur], a.data(), h.n_agents * sizeof(float4), cudaMemcpyHostToDevice));
    SM_CUDA(cudaMemcpy(e->ids[e->acur], id.data(), h.n_agents * sizeof(uint32_t), cudaMemcpyHostToDevice));
    e->n_local = h.n_agents;
    e->n_live = h.n_agents;
    if (e->world > 1 && e->comm_ready) { SM_TRY(e->mark_tail_dead()); SM_TRY(e->push_counters()); }
    e->agents_valid = true;
    e->identity_order = false;
    e->steps_since_sort = e->sort_interval;            // sort before the next step
    e->split_valid = false;
    SM_CUDA(cudaMemcpy(e->trail_ptr(e->cur), t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    bool nonneg = e->world == 1;                        // strips: the decision must be the same on every rank
    for (size_t i = 0; nonneg && i < t.size(); ++i) nonneg = t[i] >= 0.0f;
    e->trail_nonneg = nonneg;
    e->ghost_stale = true;
    e->arr_stale = true;
    e->stats_fused_valid = false;
    e->frame_pre_valid = false;
    return SM_OK;
}

// ---- display pass: /root/reference/src/display.wgsl, LUT layout of lut_manager.rs:162-186 -------------
int sm_set_lut(sm_engine* e, const uint8_t* lut768)
{
    SM_ENTER(e);
    if (!lut768) return sm_fail(SM_ERR_BAD_ARG, "null LUT");
    if (!e->lut_dev) SM_CUDA(cudaMalloc(&e->lut_dev, 768));
    SM_CUDA(cudaMemcpyAsync(e->lut_dev, lut768, 768, cudaMemcpyHostToDevice, e->stream));
    SM_CUDA(cudaStreamSynchronize(e->stream));          // the caller's buffer is only borrowed for the call
    e->lut_set = true;
    return SM_OK;
}

int sm_render_rgba8(sm_engine* e, uint32_t tex_width, uint32_t tex_height, uint8_t* rgba_host)
{
    SM_ENTER(e);
    if (!e->lut_set) return sm_fail(SM_ERR_STATE, "no LUT: call sm_set_lut first");
    if (!rgba_host) return sm_fail(SM_ERR_BAD_ARG, "null frame buffer");
    if (tex_width == 0 || tex_height == 0 || tex_width > 65536 || tex_height > 65535) return sm_fail(SM_ERR_BAD_ARG, "bad frame size");
    size_t texels = (size_t)tex_width * tex_height;
    if (texels > e->frame_cap) {
        if (e->frame_dev) { cudaFree(e->frame_dev); e->frame_dev = nullptr; e->frame_cap = 0; }
        SM_CUDA(cudaMalloc(&e->frame_dev, texels * 4));
        e->frame_cap = texels;
    }
    smk::DisplayGeom g{};
    g.W = e->W; g.H = e->H; g.tw = tex_width; g.th = tex_height;
    {   // display.wgsl:48-69, once per frame instead of once per texel; volatile: no contraction, f32 roundings
        volatile float sim_w = (float)e->W, sim_h = (float)e->H, tex_w = (float)tex_width, tex_h = (float)tex_height;
        volatile float sim_aspect = sim_w / sim_h, tex_aspect = tex_w / tex_h;
        volatile float scale, off_x = 0.0f, off_y = 0.0f, prod, diff;
        if (tex_aspect > sim_aspect) {
            scale = tex_h / sim_h; prod = sim_w * scale; diff = tex_w - prod; off_x = diff * 0.5f;
        } else {
            scale = tex_w / sim_w; prod = sim_h * scale; diff = tex_h - prod; off_y = diff * 0.5f;
        }
        g.sim_w = sim_w; g.sim_h = sim_h; g.scale = scale; g.off_x = off_x; g.off_y = off_y;
    }
    // Strips: this rank draws the frame rows that show its map rows -- fy = (py - off_y) / scale is monotone in py, so they are
    // one contiguous range; the letter-box bars above the map belong to the first strip, those below it to the last.  The
    // other rows of the caller's buffer are left untouched (like sm_download_agents / sm_download_trail on strips).
    uint32_t py0 = 0, py1 = tex_height;
    if (e->world > 1) {
        auto owner_is_me = [&](uint32_t py) {
            volatile float d = (float)py - g.off_y;
            volatile float fy = d / g.scale;                            // display.wgsl:73, the kernel's f32 operations
            const float v = fy;
            uint32_t row;
            if (!(v >= 0.0f)) row = 0;                                  // bar above the map (or NaN)
            else if (!(v < g.sim_h)) row = e->H - 1;                    // bar below
            else row = (uint32_t)v;
            return row >= e->row0 && row < e->row0 + e->rows;
        };
        while (py0 < tex_height && !owner_is_me(py0)) ++py0;
        py1 = py0;
        while (py1 < tex_height && owner_is_me(py1)) ++py1;
    }
    g.row_base = e->row0;
    g.py_first = py0;
    if (py1 == py0) return SM_OK;                                       // a frame smaller than the strip count: nothing of it is mine
    texels = (size_t)tex_width * (py1 - py0);
    dim3 grid(blocks_for((tex_width + 3) / 4, 256), py1 - py0);
    smk::DisplaySrc src{};
    if (e->frame_pre_valid) {
        // the field the reference draws: decay(merge(T_prev, deposits)) of the last step (trail_done flipped cur / ccur)
        src.trail = e->trail_ptr(1 - e->cur);
        src.cm = e->deposit_mode == 2 ? (e->flags_tiled() ? smk::CM_FLAGS_TILED : smk::CM_FLAGS) : smk::CM_COUNTS;
        src.dep = e->deposit_mode == 2 ? (const void*)e->flags_ptr(1 - e->ccur) : (const void*)e->counts_ptr(1 - e->ccur);
        src.flag_wrap = e->world == 1 ? (int)e->H : 0;
        src.tc = e->frame_tc;
    } else {
        src.trail = e->trail_ptr(e->cur);
        src.cm = smk::CM_NONE;
    }
    smk::k_display<<<grid, 256, 0, e->stream>>>(src, e->lut_dev, e->frame_dev, g);
    SM_CUDA(cudaGetLastError());
    e->timing.kernel_launches += 1;
    SM_CUDA(cudaMemcpyAsync(rgba_host + (size_t)py0 * tex_width * 4, e->frame_dev, texels * 4, cudaMemcpyDeviceToHost, e->stream));
    SM_CUDA(cudaStreamSynchronize(e->stream));
    return SM_OK;
}

int sm_resize(sm_engine* e, uint32_t width, uint32_t height)
{
    SM_ENTER(e);
    if (width == 0 || height == 0 || width > 65536 || height > 65536) return sm_fail(SM_ERR_BAD_ARG, "bad map size");
    if (e->world != 1) return e->resize_strips(width, height);      // collective: every rank calls it
    SM_CUDA(cudaStreamSynchronize(e->stream));
    // src/main.rs:985-989: x *= new_w as f32 / old_w as f32
    volatile float fx = (float)width / (float)e->W;
    volatile float fy = (float)height / (float)e->H;
    if (e->agents_valid && e->n_local)
        smk::k_rescale_agents<<<blocks_for(e->n_local, 256), 256, 0, e->stream>>>(e->agents[e->acur], e->n_local, fx, fy);
    SM_CUDA(cudaGetLastError());
    SM_CUDA(cudaStreamSynchronize(e->stream));
    e->free_trail();                                   // src/main.rs:999-1015: new zeroed trail
    if (e->tex_fallback) { e->use_tex = true; e->tex_fallback = false; }   // the new size may fit the gather limits again
    e->W = width; e->H = height;
    e->rows = height; e->row0 = 0;
    e->cfg.width = width; e->cfg.height = height;
    e->params.width = width; e->params.height = height;
    SM_TRY(e->alloc_trail());
    SM_TRY(e->setup_tiles());
    e->steps_since_sort = e->sort_interval;
    SM_CUDA(cudaStreamSynchronize(e->stream));
    return SM_OK;
}

int sm_step(sm_engine* e, uint32_t n_steps)
{
    SM_ENTER(e);
    if (!e->agents_valid) return sm_fail(SM_ERR_STATE, "agents were never initialised or uploaded");
    if (e->world > 1 && !e->comm_ready) return sm_fail(SM_ERR_STATE, "multi-GPU engine: call sm_comm_init first");
    SM_TRY(e->check_gauss(true));                        // nothing is launched on bad blur parameters
    uint32_t s = 0;
    while (s < n_steps) {
        const uint32_t period = e->graph_period();
        if (period && n_steps - s >= period && e->graph_ready()) {
            SM_TRY(e->graph_steps());                    // one whole sort period (or two steps without sorting) per graph launch
            s += period;
        } else {
            SM_TRY(e->step_once());
            ++s;
        }
    }
    return SM_OK;
}

int sm_diffuse_only(sm_engine* e, uint32_t n_passes)
{
    SM_ENTER(e);
    if (e->world > 1 && !e->comm_ready) return sm_fail(SM_ERR_STATE, "multi-GPU engine: call sm_comm_init first");
    SM_TRY(e->check_gauss(false));
    if (n_passes) e->frame_pre_valid = false;
    for (uint32_t s = 0; s < n_passes; ++s) {
        if (e->world > 1 && e->ghost_stale) SM_TRY(e->exchange_trail_ghosts());
        if (e->world > 1 && e->p2p && e->overlap_ok()) {
            SM_TRY(e->p2p_diffuse_overlapped());
        } else {
            SM_TRY(e->launch_trail(false));
            if (e->world > 1) SM_TRY(e->exchange_trail_ghosts());
        }
    }
    return SM_OK;
}

int sm_sync(sm_engine* e)
{
    SM_ENTER(e);
    SM_CUDA(cudaStreamSynchronize(e->stream));
    return SM_OK;
}

int sm_get_timing(sm_engine* e, sm_timing* out)
{
    SM_ENTER(e);
    if (!out) return sm_fail(SM_ERR_BAD_ARG, "null output");
    SM_TRY(e->resolve_timing());
    *out = e->timing;
    return SM_OK;
}
int sm_reset_timing(sm_engine* e)
{
    SM_ENTER(e);
    SM_TRY(e->resolve_timing());
    e->timing = sm_timing{};
    return SM_OK;
}
int sm_set_timing_enabled(sm_engine* e, int enabled)
{
    SM_ENTER(e);
    SM_TRY(e->resolve_timing());
    e->timing_enabled = enabled != 0;
    return SM_OK;
}

void* sm_stream(sm_engine* e) { return e ? (void*)e->stream : nullptr; }

int sm_test_math(int device, int what, const float* a, const float* b, const int32_t* iv, float* o0, float* o1, uint64_t n)
{
    SM_TRY(check_device(device, nullptr));
    if (what < 0 || what > 3 || !a || !o0) return sm_fail(SM_ERR_BAD_ARG, "bad sm_test_math arguments");
    float *da = nullptr, *db = nullptr, *d0 = nullptr, *d1 = nullptr;
    int32_t* di = nullptr;
    int rc = SM_OK;
    auto cleanup = [&]() { cudaFree(da); cudaFree(db); cudaFree(d0); cudaFree(d1); cudaFree(di); };
#define TM(call) do { cudaError_t er = (call); if (er != cudaSuccess) { rc = sm_fail(SM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(er)); cleanup(); return rc; } } while (0)
    TM(cudaMalloc(&da, n * 4)); TM(cudaMalloc(&d0, n * 4)); TM(cudaMalloc(&d1, n * 4));
    TM(cudaMemcpy(da, a, n * 4, cudaMemcpyHostToDevice));
    if (b) { TM(cudaMalloc(&db, n * 4)); TM(cudaMemcpy(db, b, n * 4, cudaMemcpyHostToDevice)); }
    if (iv) { TM(cudaMalloc(&di, n * 4)); TM(cudaMemcpy(di, iv, n * 4, cudaMemcpyHostToDevice)); }
    if ((what == 1 || what == 3) && !b) { cleanup(); return sm_fail(SM_ERR_BAD_ARG, "second operand required"); }
    if (what == 3 && !iv) { cleanup(); return sm_fail(SM_ERR_BAD_ARG, "index operand required"); }
    smk::k_test_math<<<blocks_for(n, 256), 256>>>(what, da, db, di, d0, d1, n);
    TM(cudaGetLastError());
    TM(cudaDeviceSynchronize());
    TM(cudaMemcpy(o0, d0, n * 4, cudaMemcpyDeviceToHost));
    if (o1) TM(cudaMemcpy(o1, d1, n * 4, cudaMemcpyDeviceToHost));
#undef TM
    cleanup();
    return SM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// order restoration (single GPU): agents back to index order
// ---------------------------------------------------------------------------
namespace smk {
__global__ void __launch_bounds__(256)
k_unsort(const float4* __restrict__ agents, const uint32_t* __restrict__ ids, uint64_t n,
         float4* __restrict__ agents_out, uint32_t* __restrict__ ids_out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t id = ids[i];
    agents_out[id] = agents[i];
    ids_out[id] = id;
}
__global__ void __launch_bounds__(256)
k_iota(uint32_t* __restrict__ ids, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = (uint32_t)i;
}
}  // namespace smk

int sm_engine::restore_identity_order()
{
    if (n_local) {
        smk::k_unsort<<<blocks_for(n_local, 256), 256, 0, stream>>>(agents[acur], ids[acur], n_local, agents[1 - acur], ids[1 - acur]);
        SM_CUDA(cudaGetLastError());
        acur = 1 - acur;
    }
    identity_order = true;
    return SM_OK;
}
int sm_engine::fill_identity_ids()
{
    if (n_global) smk::k_iota<<<blocks_for(n_global, 256), 256, 0, stream>>>(ids[acur], n_global);
    SM_CUDA(cudaGetLastError());
    identity_order = true;
    return SM_OK;
}
