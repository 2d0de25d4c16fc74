//! cuda_backend.rs -- stands where `PipelineManager` (src/pipeline_manager.rs:20-75: compute / decay / diffuse /
//! display pipelines) and `BindGroupManager` (src/bind_group_manager.rs:13-65: agents, trail, uniform, LUT bindings)
//! stand in the reference.  The engine owns the agent buffer, the trail map and the uniform; the host keeps
//! `Settings`, the presets and `SimSizeUniform` exactly as they are.
//!
//! NOT COMPILED in this repository's image (no rustc).  Failure mode = the reference's: it `unwrap()`s / `expect()`s
//! its wgpu calls (src/main.rs:169,182,207,213,226), so a failing engine call panics with the engine's message.
//! There is no wgpu-compute and no CPU fallback behind this type.
use std::ffi::{CStr, CString};
use std::os::raw::c_int;

// Module placement (the reference is a lib crate `slime_mold` -- src/lib.rs:1-3: lut_manager, presets, settings -- plus a
// bin crate, src/main.rs, which imports the lib as `slime_mold::..` (main.rs:4-6) and defines `SimSizeUniform` privately
// (main.rs:29-46)).  This file and ffi.rs are modules OF THE BIN CRATE: `mod ffi; mod cuda_backend;` go next to
// `mod pipeline_manager;` in main.rs (:21-24).  Then `crate::` is the bin crate -- ffi and the private SimSizeUniform of
// main.rs are visible from this child module -- and the settings / LUT types come from the lib crate.
use crate::ffi::*;
use crate::SimSizeUniform;
use slime_mold::lut_manager::LutData;
use slime_mold::settings::Settings;

fn check(rc: c_int, what: &str) {
    if rc != SM_OK {
        let msg = unsafe { CStr::from_ptr(sm_last_error()) }.to_string_lossy().into_owned();
        panic!("slime_b200: {what} failed ({rc}): {msg}");
    }
}

pub struct CudaBackend {
    e: *mut sm_engine,
    width: u32,
    height: u32,
}

impl CudaBackend {
    /// `PipelineManager::new` + `BindGroupManager::new` + the buffer creation of src/main.rs:263-293, 326-368.
    pub fn new(width: u32, height: u32, settings: &Settings) -> Self {
        let cfg = sm_config {
            width,
            height,
            agent_count: settings.agent_count as u64,
            device: 0,
            rank: 0,
            world_size: 1,
            flags: 0,
            sort_interval: 0,
            ghost_rows: 0,
            tuning: sm_tuning::default(),
        };
        let mut e = std::ptr::null_mut();
        check(unsafe { sm_create(&mut e, &cfg) }, "sm_create");
        CudaBackend { e, width, height }
    }

    /// `update_settings`, src/main.rs:83-99: the same 56 bytes `queue.write_buffer(&sim_size_buffer, ..)` sends.
    pub fn write_uniform(&self, u: &SimSizeUniform) {
        const _: () = assert!(std::mem::size_of::<SimSizeUniform>() == std::mem::size_of::<sm_params>());
        check(unsafe { sm_set_params(self.e, u as *const SimSizeUniform as *const sm_params) }, "sm_set_params");
    }

    /// Seeded, on-device version of the start-up fill (src/main.rs:269-282).
    pub fn init_agents(&self, seed: u64) {
        check(unsafe { sm_init_agents(self.e, seed) }, "sm_init_agents");
    }

    /// `queue.write_buffer(&agent_buffer, 0, cast_slice(&agents))` with a host-filled `[x, y, angle, speed]` array.
    pub fn write_agents(&self, agents: &[f32]) {
        check(unsafe { sm_upload_agents(self.e, agents.as_ptr(), 0, (agents.len() / 4) as u64) }, "sm_upload_agents");
    }

    /// The staging-buffer read-back of src/main.rs:121-131.
    pub fn read_agents(&self) -> Vec<f32> {
        let n = unsafe { sm_agent_count(self.e) };
        let mut out = vec![0f32; 4 * n as usize];
        check(unsafe { sm_download_agents(self.e, out.as_mut_ptr(), 0, n, std::ptr::null_mut()) }, "sm_download_agents");
        out
    }

    /// `reassign_agent_speeds`, src/main.rs:101-145, without the GPU -> CPU -> GPU round trip.
    pub fn reassign_agent_speeds(&self, seed: u64) {
        check(unsafe { sm_reassign_speeds(self.e, seed) }, "sm_reassign_speeds");
    }

    /// N key, src/main.rs:682-791.
    pub fn set_agent_count(&self, n: usize, seed: u64) {
        check(unsafe { sm_set_agent_count(self.e, n as u64, seed) }, "sm_set_agent_count");
    }

    /// C key, src/main.rs:909-913.
    pub fn clear_trail(&self) {
        check(unsafe { sm_clear_trail(self.e) }, "sm_clear_trail");
    }

    /// Window resize, src/main.rs:954-1015.
    pub fn resize(&mut self, width: u32, height: u32) {
        check(unsafe { sm_resize(self.e, width, height) }, "sm_resize");
        self.width = width;
        self.height = height;
    }

    /// The three compute passes of one frame, src/main.rs:1163-1235 (agents -> decay -> diffuse).  Asynchronous.
    pub fn step(&self) {
        check(unsafe { sm_step(self.e, 1) }, "sm_step");
    }

    pub fn sync(&self) {
        check(unsafe { sm_sync(self.e) }, "sm_sync");
    }

    /// src/main.rs:330-342: the LUT storage buffer (`LutData.red ++ green ++ blue`).
    pub fn set_lut(&self, lut: &LutData) {
        let mut b = Vec::with_capacity(768);
        b.extend_from_slice(&lut.red);
        b.extend_from_slice(&lut.green);
        b.extend_from_slice(&lut.blue);
        check(unsafe { sm_set_lut(self.e, b.as_ptr()) }, "sm_set_lut");
    }

    /// The display dispatch of src/main.rs:1202-1217: the RGBA8 bytes of the display texture (upload them with
    /// `queue.write_texture`).
    pub fn render(&self, tex_width: u32, tex_height: u32) -> Vec<u8> {
        let mut frame = vec![0u8; tex_width as usize * tex_height as usize * 4];
        check(unsafe { sm_render_rgba8(self.e, tex_width, tex_height, frame.as_mut_ptr()) }, "sm_render_rgba8");
        frame
    }

    pub fn read_trail(&self) -> Vec<f32> {
        let mut t = vec![0f32; self.width as usize * self.height as usize];
        check(
            unsafe { sm_download_trail(self.e, t.as_mut_ptr(), 0, 0, self.width, self.height, self.width as usize) },
            "sm_download_trail",
        );
        t
    }

    pub fn trail_statistics(&self) -> sm_trail_stats {
        let mut s = sm_trail_stats::default();
        check(unsafe { sm_trail_statistics(self.e, &mut s) }, "sm_trail_statistics");
        s
    }

    pub fn save_snapshot(&self, path: &str) {
        let p = CString::new(path).expect("path contains a NUL byte");
        check(unsafe { sm_save_snapshot(self.e, p.as_ptr()) }, "sm_save_snapshot");
    }

    pub fn load_snapshot(&self, path: &str) {
        let p = CString::new(path).expect("path contains a NUL byte");
        check(unsafe { sm_load_snapshot(self.e, p.as_ptr()) }, "sm_load_snapshot");
    }
}

impl Drop for CudaBackend {
    fn drop(&mut self) {
        unsafe { sm_destroy(self.e) };
    }
}
