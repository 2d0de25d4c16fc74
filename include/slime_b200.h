/*
 * slime_b200.h -- C ABI of the B200-native slime-mold step engine.
 *
 * Drop-in boundary for the simulation half of Velfi/slime-mold: these entry
 * points replace what the reference does through wgpu in
 *   src/pipeline_manager.rs:20-75   (compute / decay / diffuse pipelines)
 *   src/bind_group_manager.rs:13-65 (compute bind group: agents, trail, uniform)
 * and the call sites of src/main.rs listed beside each function.  Plain C types
 * only; the engine owns all device memory, host pointers are borrowed for the
 * duration of a call.  Every function returns 0 on success or a negative
 * sm_status; sm_last_error() returns the message of the calling thread's last
 * failure.  There is NO CPU fallback: without an sm_100 device every call that
 * needs one fails with SM_ERR_NO_DEVICE.
 *
 * A handle is not thread safe (one caller thread at a time, like the
 * reference's single winit thread).  sm_step()/sm_diffuse_only() are
 * asynchronous with respect to the host until sm_sync() or a download.
 */
#ifndef SLIME_B200_H
#define SLIME_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SM_VERSION_MAJOR 0
#define SM_VERSION_MINOR 2

typedef enum sm_status {
    SM_OK = 0,
    SM_ERR_BAD_ARG = -1,
    SM_ERR_CUDA = -2,
    SM_ERR_NCCL = -3,
    SM_ERR_OOM = -4,
    SM_ERR_NO_DEVICE = -5,
    SM_ERR_STATE = -6
} sm_status;

/* The reference's uniform block, byte for byte:
 * `#[repr(C)] struct SimSizeUniform`, src/main.rs:29-46 (56 bytes; the WGSL view
 * in src/compute.wgsl:37-50 stops at offset 44).  width/height are the GLOBAL
 * map size.  blur_radius / blur_sigma are ignored by the reference's shader and
 * by the engine unless SM_FLAG_GAUSSIAN_BLUR is set (extension). */
typedef struct sm_params {
    uint32_t width;
    uint32_t height;
    float decay_factor;
    float agent_jitter;
    float agent_speed_min;
    float agent_speed_max;
    float agent_turn_speed;
    float agent_sensor_angle;
    float agent_sensor_distance;
    float diffusion_rate;
    float pheromone_deposition_amount;
    float blur_radius;
    float blur_sigma;
    uint32_t _pad;
} sm_params;

enum {
    /* Replace the 3x3 mean of compute.wgsl:181-194 by a separable Gaussian of
     * radius round(blur_radius) (1..8) and sigma blur_sigma.  Extension with no
     * reference semantics ("parity unpinned"). */
    SM_FLAG_GAUSSIAN_BLUR = 1u << 0,
    /* Never re-order agents in memory (disables the periodic cell sort). */
    SM_FLAG_NO_SORT = 1u << 1,
    /* The reference's own RACY semantics instead of the deterministic phase_split: agents sense and deposit on ONE live
     * buffer with the shader's non-atomic read-modify-write (compute.wgsl:93-95, 136-141), decay and diffuse run in place
     * (:148-195, a cell may read neighbours the same dispatch already rewrote).  Results depend on the GPU's scheduling and
     * differ from run to run, exactly like the reference's: this mode cannot be checked bit for bit against anything and is
     * validated by field statistics only (tests/test_gpu_statistics.py).  Single GPU, 3x3 box only; slower than the default
     * path (three passes, LDG sampling).  Opt-in. */
    SM_FLAG_SEM_INPLACE = 1u << 2
};

/* Measurement switches.  Every field: 0 = the engine's default, which is what profiles/ measured fastest.  None of them
 * changes a result bit -- they select between kernels / schedules that are tested bit for bit against each other -- and
 * the library reads NO environment variables for them: a harness that wants A/B runs fills this struct (the Python
 * binding maps SM_* environment variables onto it for the test and bench scripts, slime_mold_b200/_lib.py). */
typedef struct sm_tuning {
    uint32_t sampler;               /* agent kernel's trail sampler: 0 texture gather from a block-linear copy (LDG when the
                                       map exceeds the gather limits), 1 always LDG from the row-major field */
    uint32_t tile_shift_x;          /* cell-sort tiles are 2^x by 2^y cells; 0 = 3 (8 x 8) */
    uint32_t tile_shift_y;
    uint32_t trail_rows_per_chunk;  /* rows one CTA of the fused decay+diffuse kernel walks; 0 = 8 (16 for large diffusion-only maps) */
    uint32_t deposit_counts_only;   /* 1: u32 deposit counts even where u8 flags are exact */
    uint32_t generic_trail_kernel;  /* 1: the one-cell-per-thread trail kernel for every map */
    uint32_t surface_row_writes;    /* 1: the trail pass writes the sampler copy row by row instead of as whole sectors */
    uint32_t no_step_graph;         /* 1: sm_step never replays CUDA graphs of whole sort periods */
    uint32_t gauss_kernel;          /* EXTENSION: 0 auto, 1 register-streaming (rows), 2 shared-memory streaming, 3 tile, 4 two-pass */
    uint32_t gauss_rows_max_radius; /* largest radius the rows kernel takes in auto mode; 0 = 5 */
    uint32_t gauss_rows_packing;    /* FFMA2 taps of the rows kernel: 0 auto, 1 scalar, 2 packed */
    uint32_t gauss_chunk_rows;      /* rows per CTA of the streaming kernels; 0 = chosen per map */
    uint32_t exchange;              /* strips: 0 direct peer stores over NVLink (CUDA IPC), 1 NCCL send/recv */
    uint32_t serial_exchange;       /* strips: 1 = no overlap of the exchange with the interior trail rows */
    uint32_t migrate_capacity;      /* strips: agents per migration message; 0 = 65536 */
    uint32_t barrier_fence;         /* strips: flag-barrier fences: 0 acq_rel.sys, 1 three sc system fences, 2 device scope */
    uint32_t debug_single_rank_strip; /* PROFILING ONLY: a world_size > 1 engine that exchanges with itself (results meaningless) */
    uint32_t debug_side_timing;     /* strips: per-piece CUDA-event times of the side stream, printed at teardown */
    uint32_t no_boundary_first;     /* strips: 1 = one agent launch per step, exchange overlapped with the interior trail rows only */
    uint32_t deposit_flag_layout;   /* u8 deposit flags: 0 auto (>= 2^23 cells per GPU, whole 8 x 8-cell tiles, on strips the peer-store
                                       exchange: tiled; else row-major), 1 always row-major, 2 tiled wherever the geometry allows
                                       (never changes a result bit) */
    uint32_t reserved[4];
} sm_tuning;

typedef struct sm_config {
    uint32_t width;          /* global map width  (cells) */
    uint32_t height;         /* global map height (cells) */
    uint64_t agent_count;    /* GLOBAL number of agents (src/settings.rs:9) */
    int32_t device;          /* CUDA device ordinal of this rank */
    int32_t rank;            /* strip index 0..world_size-1 (0 for a single GPU) */
    int32_t world_size;      /* number of horizontal strips == GPUs (1 = whole map) */
    uint32_t flags;          /* SM_FLAG_* */
    uint32_t sort_interval;  /* steps between agent cell sorts; 0 = engine default (24) */
    uint32_t ghost_rows;     /* strips: ghost rows kept above and below; 0 = 232 (sensor distance 225 + slack) */
    sm_tuning tuning;
} sm_config;

typedef struct sm_timing {       /* CUDA-event totals since sm_reset_timing() */
    double agents_ms;            /* agent kernel (compute.wgsl `main`) */
    double trail_ms;             /* fused decay+blur (`decay_trail` + `diffuse_trail`) */
    double sort_ms;              /* periodic cell sort */
    double exchange_ms;          /* halo exchange + migration (multi-GPU) */
    uint64_t agent_launches, trail_launches, sort_launches, exchange_launches;
    uint64_t steps;
    uint64_t kernel_launches;    /* every kernel the engine launched (counted even with timing off) */
} sm_timing;

typedef struct sm_trail_stats {  /* computed on the device over the owned strip */
    double sum;
    double sum_sq;
    float max;
    uint32_t _pad;
    uint64_t nonzero;
} sm_trail_stats;

typedef struct sm_engine sm_engine;

const char *sm_last_error(void);
void sm_version(int *major, int *minor);
/* Number of CUDA devices with compute capability 10.x (0 => nothing can run). */
int sm_device_count(void);

/* Construction = PipelineManager::new + BindGroupManager::new + the buffer
 * creation of src/main.rs:263-293 (agents uninitialised, trail zeroed). */
int sm_create(sm_engine **out, const sm_config *cfg);
int sm_destroy(sm_engine *e);

/* Multi-GPU (one process per GPU).  Rank 0 obtains an id with
 * sm_comm_unique_id(), the host distributes the 128 bytes out of band (bench.py
 * uses torch.distributed), every rank then calls sm_comm_init().  The engine
 * exchanges halo rows and migrating agents with its two ring neighbours over
 * peer stores over NVLink or NCCL (sm_tuning.exchange).  libnccl.so.2 is resolved at run
 * time: an already loaded copy (e.g. torch's) is reused; the environment variable
 * SM_NCCL_LIB names another one -- the only environment variable the library reads. */
#define SM_COMM_ID_BYTES 128
int sm_comm_unique_id(uint8_t id[SM_COMM_ID_BYTES]);
int sm_comm_init(sm_engine *e, const uint8_t id[SM_COMM_ID_BYTES]);

/* queue.write_buffer(&sim_size_buffer, ..) -- src/main.rs:98, 827-831, 1056.
 * params->width/height must equal the engine's map size. */
int sm_set_params(sm_engine *e, const sm_params *params);
int sm_get_params(sm_engine *e, sm_params *params);

/* Agents are (x, y, angle, speed) f32x4, src/main.rs:263-282, indexed by their
 * persistent GLOBAL index (the `agent_index` of compute.wgsl:60, which also
 * seeds the jitter hash at :117).  Single GPU: any [first, first+n) range.
 * Multi GPU: every rank passes the same global array slice; each rank keeps the
 * agents whose row falls in its strip. */
int sm_upload_agents(sm_engine *e, const float *xyas, uint64_t first, uint64_t n);
/* Writes the agents this rank currently owns into xyas[4*global_index..];
 * entries of agents owned by other ranks are left untouched.  *n_owned (may be
 * NULL) receives the number written. */
int sm_download_agents(sm_engine *e, float *xyas, uint64_t first, uint64_t n, uint64_t *n_owned);
/* Seeded version of the start-up fill of src/main.rs:269-282, generated on the
 * device: x~U[0,W) y~U[0,H) angle~U[0,2pi) speed~U[min,max), counter-based RNG
 * keyed by (seed, agent index). */
int sm_init_agents(sm_engine *e, uint64_t seed);
/* N key, src/main.rs:682-791: new agent count, all agents re-randomised. */
int sm_set_agent_count(sm_engine *e, uint64_t n, uint64_t seed);
/* reassign_agent_speeds, src/main.rs:101-145 (on the device, seeded). */
int sm_reassign_speeds(sm_engine *e, uint64_t seed);
uint64_t sm_agent_count(sm_engine *e);        /* global */
uint64_t sm_local_agent_count(sm_engine *e);  /* owned by this rank */

/* C key, src/main.rs:909-913 */
int sm_clear_trail(sm_engine *e);
/* Rectangle (x0,y0,w,h) in GLOBAL coordinates, row-major f32 with `pitch`
 * floats per host row; only the rows this rank owns are touched. */
int sm_upload_trail(sm_engine *e, const float *src, uint32_t x0, uint32_t y0,
                    uint32_t w, uint32_t h, size_t pitch);
int sm_download_trail(sm_engine *e, float *dst, uint32_t x0, uint32_t y0,
                      uint32_t w, uint32_t h, size_t pitch);
int sm_trail_statistics(sm_engine *e, sm_trail_stats *out);

/* Window resize, src/main.rs:954-1015: agent positions scaled by new/old size,
 * trail replaced by a zeroed map of the new size.  On strips the call is collective (every
 * rank, same arguments): the strip boundaries move with the new height, agents that now
 * belong to another strip are handed over (host-mediated -- a window event, not a hot path),
 * and the peer-memory exchange is set up again over the existing communicator. */
int sm_resize(sm_engine *e, uint32_t width, uint32_t height);

/* ---- display pass -------------------------------------------------
 * Replaces the `display_pipeline` dispatch of src/main.rs:1202-1217 and its shader
 * src/display.wgsl:29-86: the trail map is letter-boxed into a tex_width x tex_height
 * RGBA8 frame through a colour look-up table.
 * sm_set_lut: 768 bytes laid out as 256 red, 256 green, 256 blue -- a reference .lut file
 * verbatim (src/lut_manager.rs:162-186), i.e. LutData.red ++ green ++ blue as main.rs:330-334
 * concatenates them (the reference widens each byte to u32 for its storage buffer).
 * sm_render_rgba8: `rgba` is a caller-owned host buffer of tex_width * tex_height * 4 bytes,
 * row-major, R G B A per texel (the rgba8unorm texture of main.rs:296-314).  Like the
 * reference, which draws between its decay and diffuse dispatches, the frame after an
 * sm_step() shows the field of the last step after deposits and decay, before the blur
 * (recomputed per texel from that step's inputs, which are still in device memory); after
 * anything else that changed the trail (upload, clear, sm_diffuse_only, resize, snapshot
 * load) it shows the trail as it stands.  On strips every rank writes the frame rows that
 * show its own map rows (the bars above / below the map belong to the first / last strip)
 * and leaves the rest of `rgba` untouched: the ranks' frames tile the whole frame. */
int sm_set_lut(sm_engine *e, const uint8_t *lut768);
int sm_render_rgba8(sm_engine *e, uint32_t tex_width, uint32_t tex_height, uint8_t *rgba);

/* ---- snapshot / restore (no counterpart in the reference) ------------
 * Everything a bit-exact continuation needs: parameter block, the owned trail rows and the
 * owned agents with their persistent indices.  On strips every rank writes / reads its own
 * file: `path` gets ".rank<r>" appended when world_size > 1.  sm_load_snapshot requires an
 * engine created with the same map size, agent count and strip layout. */
int sm_save_snapshot(sm_engine *e, const char *path);
int sm_load_snapshot(sm_engine *e, const char *path);

/* One frame of src/main.rs:1163-1235 per step: agents -> decay -> diffuse. */
int sm_step(sm_engine *e, uint32_t n_steps);
/* decay + diffuse only (BASELINE config 5). */
int sm_diffuse_only(sm_engine *e, uint32_t n_passes);
int sm_sync(sm_engine *e);

int sm_get_timing(sm_engine *e, sm_timing *out);
int sm_reset_timing(sm_engine *e);
/* 0 = no per-kernel events (fastest), 1 = CUDA events around every kernel. */
int sm_set_timing_enabled(sm_engine *e, int enabled);

/* Raw stream handle (cudaStream_t) the engine launches on, so a host harness can
 * record its own CUDA events on it. */
void *sm_stream(sm_engine *e);

/* Test hooks for the arithmetic spec (device evaluation of sin/cos, fmod, the
 * jitter hash and x/9): out arrays are host pointers of n floats. */
int sm_test_math(int device, int what, const float *a, const float *b, const int32_t *i,
                 float *out0, float *out1, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* SLIME_B200_H */
