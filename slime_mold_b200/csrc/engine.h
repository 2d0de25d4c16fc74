// engine.h -- internal state behind the opaque sm_engine handle of include/slime_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/slime_b200.h"
#include "kernels.cuh"

int sm_fail(int code, const char* fmt, ...);

struct EvPair { cudaEvent_t a, b; int kind; };

// the row that owns an agent at height y (out-of-range and NaN clamp as in the agent kernel's own row rule)
static inline uint32_t owner_row(float y, uint32_t H)
{
    if (!(y >= 0.0f)) return 0;                 // negative / NaN
    if (y >= (float)H) return H - 1;
    return (uint32_t)y;
}

// Per-direction migration message (multi-GPU): [u64 count][u64 pad][float4 a[cap]][u32 id[cap]]
struct MigrateBuf {
    uint8_t* send = nullptr;   // filled by k_agents<true>, sent to the ring neighbour after the trail pass
    uint8_t* recv = nullptr;   // the neighbour's message
};

struct sm_engine {
    sm_config cfg{};
    sm_params params{};
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;

    // geometry: this rank owns global rows [row0, row0 + rows) of a W x H map and keeps
    // `ghost` extra rows above and below (0 on a single GPU, where rows wrap toroidally)
    uint32_t W = 0, H = 0;
    int rank = 0, world = 1;
    uint32_t row0 = 0, rows = 0, ghost = 0;

    // trail: two ping-pong f32 fields, two alternating u32 deposit-count fields
    float* trail_base[2] = {nullptr, nullptr};
    uint32_t* counts_base[2] = {nullptr, nullptr};
    int cur = 0, ccur = 0;
    // u8 deposit flags (same geometry, two alternating buffers): used instead of the counts whenever
    // dep >= 1 and the field is known to be non-negative (then any deposit saturates the cell to 1.0)
    uint8_t* flags_base[2] = {nullptr, nullptr};
    bool trail_nonneg = true;         // every cell of trail[cur] is >= 0 and not NaN
    int deposit_mode = 0;             // mode of the most recent agents pass: 1 counts, 2 flags
    uint8_t* flags_ptr(int i) const { return flags_base[i] + (size_t)(ghost + pad_rows) * W; }
    bool flag_mode() const;
    bool trail_rows_kernel_ok() const { return !(cfg.flags & SM_FLAG_GAUSSIAN_BLUR) && W % 4 == 0 && W >= 8 && (W / 4) % 32 != 1 && !force_generic; }
    // u8 flags stored in 8 x 8-cell tiles (kernels.cuh: flag_tile_offset): one GPU, whole tiles, the streaming trail kernel.
    // Fixed for the lifetime of a map size, so the two flag buffers never mix layouts.
    uint32_t trail_rows_per_chunk(bool has_counts) const;
    bool flags_tiled() const
    {
        // (measured, tools/r2/gpu_32.sh: configs[2] 1807 -> 1679 us/step, configs[1] 222.9 -> 221.3, Snake at configs[1] 388 -> 361;
        // the L2-resident configs[0] loses 1 % -- 20.4 -> 20.6 us -- so maps below 2^23 cells keep the row-major field)
        if (tuning.deposit_flag_layout == 1 || !trail_rows_kernel_ok() || W % 8 != 0 || rows % 8 != 0 ||
            (cfg.flags & SM_FLAG_SEM_INPLACE))
            return false;
        if ((uint64_t)W * rows < (1ull << 23) && tuning.deposit_flag_layout != 2) return false;
        // strips: the peer-store exchange only (NCCL ships whole rows), equal strips whose owned row 0 starts a tile row
        if (world > 1 && !(p2p && H % (8u * (uint32_t)world) == 0 && (ghost + pad_rows) % 8 == 0)) return false;
        const uint32_t rpc = trail_rows_per_chunk(true);      // the full-step pass must run whole chunks of 4 or 8 rows
        return rpc == 4 || rpc == 8;
    }
    int switch_deposit_mode(int mode);
    bool ghost_stale = true;          // ghost rows of trail[cur] need a (re-)exchange
    // block-linear copy of trail[cur] for the texture-gather sampler of the agent kernel
    cudaArray_t trail_arr = nullptr;
    cudaTextureObject_t trail_tex = 0;
    cudaSurfaceObject_t trail_surf = 0;
    bool use_tex = false;             // tuning.sampler == 0 and the gather probe passed
    bool tex_fallback = false;        // use_tex was switched off because this map does not fit a gather array (retried by sm_resize)
    bool arr_stale = true;            // the array does not mirror trail[cur] (upload / clear / diffuse-only ...)
    float* gauss_dec = nullptr;       // extension scratch
    float* gauss_hb = nullptr;
    float* trail_ptr(int i) const { return trail_base[i] + (size_t)(ghost + pad_rows) * W; }        // owned row 0
    uint32_t* counts_ptr(int i) const { return counts_base[i] + (size_t)(ghost + pad_rows) * W; }
    size_t field_cells() const { return (size_t)(rows + 2 * (size_t)(ghost + pad_rows)) * W; }

    // agents: float4 state + u32 persistent index, double buffered for the cell sort
    float4* agents[2] = {nullptr, nullptr};
    uint32_t* ids[2] = {nullptr, nullptr};
    int acur = 0;
    uint64_t n_global = 0, n_local = 0, cap_local = 0;
    uint64_t n_live = 0;              // multi-GPU: n_local minus dead (migrated-away) slots
    bool agents_valid = false;
    bool identity_order = false;      // single GPU: agents[i] is agent i

    // periodic cell sort
    smk::TileGeom tiles{};
    uint64_t n_tiles = 0;
    uint32_t n_scan_blocks = 0;
    uint32_t* tile_hist = nullptr;
    uint32_t* tile_sums = nullptr;
    uint32_t sort_interval = 24, steps_since_sort = 0;

    // measurement switches (sm_config.tuning; 0 = default everywhere) and what sm_create derived from them
    sm_tuning tuning{};
    bool force_generic = false;
    bool no_flags = false;            // always count deposits
    int rpc_override = 0;
    bool gauss_stream = true;         // the streaming Gaussian kernels wherever they apply (gauss_kernel 3 / 4 select the tile / two-pass forms)
    bool gauss_stream_ok() const;
    // the register-streaming kernel for small radii (gauss_rows.cuh): gauss_kernel 1 forces it for every radius it is built
    // for (1-5), 0 = radii up to gauss_rows_max_r, 2 / 3 / 4 = never
    bool gauss_rows = true;
    int gauss_rows_max_r = 5;         // measured (profiles/): 0.64-0.95 of the HBM peak at radius 1-5 against 0.55-0.69 for the streaming kernel
    int gauss_rows_packing = 0;       // 0 auto (FFMA2 taps up to radius 4), 1 scalar taps, 2 packed
    bool gauss_rows_ok() const;
    bool gauss_fast_ok() const { return gauss_rows_ok() || gauss_stream_ok(); }   // kernels that merge u8 flags and write the sampler copy
    bool gauss_two_pass = false;      // the unfused Gaussian passes (A/B; also used for maps below 160 x 64)

    // statistics: accumulator on the device; `stats_fused_valid` = the last thing that changed trail[cur] was a full-step
    // pass of k_trail_rows, which filled it on the way (else sm_trail_statistics runs k_trail_stats)
    void* stats_dev = nullptr;
    void* stats_host = nullptr;       // pinned mirror
    bool stats_fused_valid = false;
    uint32_t stats_interest = 0;      // > 0: the host read statistics within the last 64 steps -> passes run the STATS instantiation

    // the frame the reference would draw is the field between decay and diffuse of the last step (main.rs:1202-1217):
    // recomputed by k_display from the inputs of that step's trail pass while they are still intact
    bool frame_pre_valid = false;     // the last thing that changed the trail was a full step (its inputs are still in HBM)
    smd::TrailConsts frame_tc{};      // deposit amount / decay of that step
    // display pass (display.wgsl): LUT and frame buffer, allocated on first use
    uint8_t* lut_dev = nullptr;
    bool lut_set = false;
    uint32_t* frame_dev = nullptr;
    size_t frame_cap = 0;           // texels

    // timing
    bool timing_enabled = false;
    std::vector<EvPair> ev_pool;
    size_t ev_used = 0;
    sm_timing timing{};

    // multi-GPU (exchange.cu)
    bool comm_ready = false;
    void* comm = nullptr;             // ncclComm_t
    uint32_t* counts_xchg = nullptr;  // recv staging for the deposit-count exchange
    uint64_t counts_xchg_rows = 0;
    MigrateBuf mig[2];                // 0: towards rank-1 (up), 1: towards rank+1 (down)
    uint64_t mig_cap = 0;             // agents per message
    size_t mig_bytes = 0;             // bytes per message
    // device counters: [0] slots in use (n_local), [1] live agents, [2] overflow flag
    unsigned long long* dev_counters = nullptr;
    unsigned long long* host_counters = nullptr;      // pinned mirror
    uint64_t n_upper = 0;             // host-side upper bound of slots in use between sorts
    uint32_t pad_rows = 0;            // physical padding rows beyond the ghosts (multi-GPU memory-safety slack)

    // XM_P2P: CUDA IPC views of the two ring neighbours' buffers (peer[0] = up, peer[1] = down)
    struct PeerView {
        uint32_t* counts[2] = {nullptr, nullptr};   // their counts_base[0/1]
        uint8_t* flags8[2] = {nullptr, nullptr};    // their flags_base[0/1]
        float* trail[2] = {nullptr, nullptr};       // their trail_base[0/1]
        uint8_t* window = nullptr;                  // their barrier flags + arrival buffers
        uint32_t rows = 0;                          // rows they own
    };
    bool p2p = false;
    bool fake_multi = false;          // tuning.debug_single_rank_strip (profiling only): a strip engine whose two ring neighbours are itself
    int fake_comm_init();
    PeerView peer[2];
    std::vector<void*> ipc_opened;    // every pointer obtained from cudaIpcOpenMemHandle
    uint8_t* window = nullptr;        // [64 B barrier flags][arrivals from up][arrivals from down]
    size_t window_arrival_off[2] = {0, 0};          // offset of the arrival buffer filled by up / by down
    uint32_t barrier_seq = 0;
    int setup_p2p();
    int p2p_streams_init();
    int p2p_barrier();
    int p2p_after_agents(cudaStream_t st = nullptr, bool timed = true);   // barrier 1 + pull of the neighbours' boundary deposit rows
    int p2p_after_trail(cudaStream_t st = nullptr, bool timed = true);    // push trail ghost rows, append arrivals, barrier 2
    // Overlapped form of {p2p_after_agents, trail pass, p2p_after_trail}: the interior rows of the trail
    // pass need nothing from the neighbours and run on the main stream at once; barrier 1, the two boundary
    // bands, the ghost-row push, the arrivals and barrier 2 run beside them on `side_stream`.
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap_enabled = true;      // tuning.serial_exchange switches back to the serial order (A/B)
    bool overlap_ok();
    // boundary-first stepping (engine.cu: plan_split, exchange.cu: p2p_step_split)
    bool split_valid = false, split_pending = false;
    uint64_t split_a = 0, split_b = 0;      // interior slots [split_a, split_b) until the next sort
    uint32_t* split_host = nullptr;         // pinned: two slot offsets read back at sort time
    cudaEvent_t ev_boundary = nullptr;
    int plan_split();
    void finish_split();
    int p2p_step_split();
    uint32_t overlap_band();          // rows of each boundary band (multiple of the chunk height; 0 = strip too thin)
    int p2p_trail_overlapped();
    int p2p_diffuse_overlapped();
    int p2p_push_ghosts(cudaStream_t st, uint32_t g);

    // CUDA-graph replay of whole sort periods on one GPU (engine.cu: graph_steps)
    struct GraphKey {
        sm_params params; uint64_t n_local; int acur, cur, ccur, deposit_mode; uint32_t sort_interval;
        const void *agents0, *trail0, *arr; bool use_tex, force_generic, no_flags; int rpc_override;
    };
    bool surf_pairs = true;           // !tuning.surface_row_writes
    bool graph_enabled = true;        // !tuning.no_step_graph
    cudaGraphExec_t step_graph = nullptr;
    GraphKey step_graph_key{};
    uint64_t step_graph_launches = 0;
    uint32_t graph_period() const;    // steps per graph launch (0 = graphs do not apply)
    GraphKey graph_key_now() const;
    bool graph_ready();
    int graph_steps();
    int step_once();
    int step_inplace();               // SM_FLAG_SEM_INPLACE

    smd::AgentConsts agent_consts() const;
    smd::TrailConsts trail_consts() const;

    int tic(int kind, cudaStream_t st = nullptr);
    int toc(cudaStream_t st = nullptr);
    double side_dbg_ms[8] = {0};     // tuning.debug_side_timing: per-piece times of the side stream (printed at comm teardown)
    uint64_t side_dbg_n[8] = {0};
    bool side_dbg = false;
    int resolve_timing();

    int alloc_trail();
    void free_trail();
    int alloc_agents(uint64_t capacity);
    void free_agents();
    int setup_tiles();
    int setup_tex();
    void free_tex();
    int refresh_tex(int64_t local_row_begin, int64_t n_rows);   // linear trail[cur] rows -> array
    int refresh_tex_ghosts(uint32_t g);                         // both ghost bands, by a kernel

    int sort_agents();
    int prepare_agents();
    int launch_agents(int part = 0, cudaStream_t st = nullptr);   // part: 0 all slots, 1 interior (single-GPU kernel), 2 boundary
    struct TrailPass {
        smk::TrailGeom g;
        smd::TrailConsts tc;
        const float* tin; float* tout;
        int cm; const void* cin; void* czero;
        bool fast;                    // k_trail_rows applies (else k_trail_generic / the Gaussian extension)
        bool stats;                   // this pass also reduces the field statistics
    };
    int trail_plan(bool has_counts, TrailPass& p);
    int trail_launch_rows(const TrailPass& p, uint32_t y_first, uint32_t y_last, cudaStream_t st,
                          uint32_t y_first2 = 0, uint32_t y_last2 = 0);   // optional second band in the same launch
    void trail_done(bool has_counts);
    int launch_trail(bool has_counts);
    int launch_gauss(bool has_counts, const TrailPass& p);
    int check_gauss(bool has_counts) const;   // parameter / geometry checks of the Gaussian extension, before anything is launched
    int restore_identity_order();
    int fill_identity_ids();

    // exchange.cu
    int init_agents_strip(uint64_t seed);
    int exchange_counts();
    int exchange_trail_ghosts();
    int migrate_agents();
    int resize_strips(uint32_t width, uint32_t height);   // sm_resize on strips (collective, host-mediated)
    int refresh_counters();           // multi-GPU: sync and read the device counters into n_local / n_live
    int push_counters();              // multi-GPU: write n_local / n_live to the device counters
    int mark_tail_dead();             // multi-GPU: ids[n_local .. bound) = kDeadAgent (slots the grid may cover before the next sort)
    void comm_destroy();
};
