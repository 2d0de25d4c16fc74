//! ffi.rs -- include/slime_b200.h, declaration for declaration.  NOT COMPILED in this repository's image (no rustc);
//! tests/test_rust_shim.py checks names, argument counts and struct fields against the header.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_int, c_void};

pub const SM_OK: c_int = 0;
pub const SM_ERR_BAD_ARG: c_int = -1;
pub const SM_ERR_CUDA: c_int = -2;
pub const SM_ERR_NCCL: c_int = -3;
pub const SM_ERR_OOM: c_int = -4;
pub const SM_ERR_NO_DEVICE: c_int = -5;
pub const SM_ERR_STATE: c_int = -6;

pub const SM_FLAG_GAUSSIAN_BLUR: u32 = 1 << 0;
pub const SM_FLAG_NO_SORT: u32 = 1 << 1;
pub const SM_FLAG_SEM_INPLACE: u32 = 1 << 2;
pub const SM_COMM_ID_BYTES: usize = 128;

/// Opaque engine handle.
#[repr(C)]
pub struct sm_engine {
    _private: [u8; 0],
}

/// `SimSizeUniform` of the reference (src/main.rs:29-46), byte for byte: the reference's own struct can be passed
/// wherever a `*const sm_params` is expected (`&uniform as *const SimSizeUniform as *const sm_params`).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sm_params {
    pub width: u32,
    pub height: u32,
    pub decay_factor: f32,
    pub agent_jitter: f32,
    pub agent_speed_min: f32,
    pub agent_speed_max: f32,
    pub agent_turn_speed: f32,
    pub agent_sensor_angle: f32,
    pub agent_sensor_distance: f32,
    pub diffusion_rate: f32,
    pub pheromone_deposition_amount: f32,
    pub blur_radius: f32,
    pub blur_sigma: f32,
    pub _pad: u32,
}

/// Measurement switches; `Default::default()` (all zero) = the engine's defaults.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sm_tuning {
    pub sampler: u32,
    pub tile_shift_x: u32,
    pub tile_shift_y: u32,
    pub trail_rows_per_chunk: u32,
    pub deposit_counts_only: u32,
    pub generic_trail_kernel: u32,
    pub surface_row_writes: u32,
    pub no_step_graph: u32,
    pub gauss_kernel: u32,
    pub gauss_rows_max_radius: u32,
    pub gauss_rows_packing: u32,
    pub gauss_chunk_rows: u32,
    pub exchange: u32,
    pub serial_exchange: u32,
    pub migrate_capacity: u32,
    pub barrier_fence: u32,
    pub debug_single_rank_strip: u32,
    pub debug_side_timing: u32,
    pub no_boundary_first: u32,
    pub deposit_flag_layout: u32,
    pub reserved: [u32; 4],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sm_config {
    pub width: u32,
    pub height: u32,
    pub agent_count: u64,
    pub device: i32,
    pub rank: i32,
    pub world_size: i32,
    pub flags: u32,
    pub sort_interval: u32,
    pub ghost_rows: u32,
    pub tuning: sm_tuning,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sm_timing {
    pub agents_ms: f64,
    pub trail_ms: f64,
    pub sort_ms: f64,
    pub exchange_ms: f64,
    pub agent_launches: u64,
    pub trail_launches: u64,
    pub sort_launches: u64,
    pub exchange_launches: u64,
    pub steps: u64,
    pub kernel_launches: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sm_trail_stats {
    pub sum: f64,
    pub sum_sq: f64,
    pub max: f32,
    pub _pad: u32,
    pub nonzero: u64,
}

extern "C" {
    pub fn sm_last_error() -> *const c_char;
    pub fn sm_version(major: *mut c_int, minor: *mut c_int);
    pub fn sm_device_count() -> c_int;

    pub fn sm_create(out: *mut *mut sm_engine, cfg: *const sm_config) -> c_int;
    pub fn sm_destroy(e: *mut sm_engine) -> c_int;

    pub fn sm_comm_unique_id(id: *mut u8) -> c_int;
    pub fn sm_comm_init(e: *mut sm_engine, id: *const u8) -> c_int;

    pub fn sm_set_params(e: *mut sm_engine, params: *const sm_params) -> c_int;
    pub fn sm_get_params(e: *mut sm_engine, params: *mut sm_params) -> c_int;

    pub fn sm_upload_agents(e: *mut sm_engine, xyas: *const f32, first: u64, n: u64) -> c_int;
    pub fn sm_download_agents(e: *mut sm_engine, xyas: *mut f32, first: u64, n: u64, n_owned: *mut u64) -> c_int;
    pub fn sm_init_agents(e: *mut sm_engine, seed: u64) -> c_int;
    pub fn sm_set_agent_count(e: *mut sm_engine, n: u64, seed: u64) -> c_int;
    pub fn sm_reassign_speeds(e: *mut sm_engine, seed: u64) -> c_int;
    pub fn sm_agent_count(e: *mut sm_engine) -> u64;
    pub fn sm_local_agent_count(e: *mut sm_engine) -> u64;

    pub fn sm_clear_trail(e: *mut sm_engine) -> c_int;
    pub fn sm_upload_trail(e: *mut sm_engine, src: *const f32, x0: u32, y0: u32, w: u32, h: u32, pitch: usize) -> c_int;
    pub fn sm_download_trail(e: *mut sm_engine, dst: *mut f32, x0: u32, y0: u32, w: u32, h: u32, pitch: usize) -> c_int;
    pub fn sm_trail_statistics(e: *mut sm_engine, out: *mut sm_trail_stats) -> c_int;

    pub fn sm_resize(e: *mut sm_engine, width: u32, height: u32) -> c_int;

    pub fn sm_set_lut(e: *mut sm_engine, lut768: *const u8) -> c_int;
    pub fn sm_render_rgba8(e: *mut sm_engine, tex_width: u32, tex_height: u32, rgba: *mut u8) -> c_int;

    pub fn sm_save_snapshot(e: *mut sm_engine, path: *const c_char) -> c_int;
    pub fn sm_load_snapshot(e: *mut sm_engine, path: *const c_char) -> c_int;

    pub fn sm_step(e: *mut sm_engine, n_steps: u32) -> c_int;
    pub fn sm_diffuse_only(e: *mut sm_engine, n_passes: u32) -> c_int;
    pub fn sm_sync(e: *mut sm_engine) -> c_int;

    pub fn sm_get_timing(e: *mut sm_engine, out: *mut sm_timing) -> c_int;
    pub fn sm_reset_timing(e: *mut sm_engine) -> c_int;
    pub fn sm_set_timing_enabled(e: *mut sm_engine, enabled: c_int) -> c_int;

    pub fn sm_stream(e: *mut sm_engine) -> *mut c_void;

    pub fn sm_test_math(device: c_int, what: c_int, a: *const f32, b: *const f32, i: *const i32,
                        out0: *mut f32, out1: *mut f32, n: u64) -> c_int;
}
