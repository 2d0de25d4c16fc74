#!/usr/bin/env python
"""Kernel-level sweeps on one GPU (not the headline bench): diffusion GB/s by map size and chunking,
agent-kernel time by sort tile shape / sort interval.  Writes JSON lines to gpurun_out/kernel_sweep.jsonl.

    python tools/bench_kernels.py [diffusion] [agents] [presets]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402,F401
import slime_mold_b200 as sm  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "kernel_sweep.jsonl")
PEAK = 6553.9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def emit(d):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(d) + "\n")
    print(json.dumps(d), flush=True)


def event_time(be, fn):
    stream = torch.cuda.ExternalStream(be.stream_handle)
    be.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1)


def diffusion_sweep():
    for S in (4096, 8192, 16384, 32768):
        for rpc in (0, 8, 16, 32, 64):
            os.environ["SM_TRAIL_ROWS_PER_CHUNK"] = str(rpc)
            be = sm.CudaBackend.new(S, S, agent_count=1)
            be.write_trail(np.random.default_rng(0).random((256, S), dtype=np.float32), y0=0)
            passes = max(10, min(200, int(2e10 / (S * S * 8))))
            be.diffuse_only(5)
            ms = event_time(be, lambda: be.diffuse_only(passes))
            gbs = 8.0 * S * S * passes / (ms * 1e-3) / 1e9
            emit({"sweep": "diffusion", "size": S, "rows_per_chunk": rpc, "passes": passes, "ms_per_pass": ms / passes,
                  "gbs": gbs, "frac_of_measured_peak": gbs / PEAK, "frac_of_8TBs": gbs / 8000.0})
            be.close()
    os.environ.pop("SM_TRAIL_ROWS_PER_CHUNK", None)


def agents_sweep():
    N, W, H = 16_777_216, 4096, 4096
    for (sx, sy) in ((3, 3), (4, 4), (5, 5)):
        for interval in (8, 16, 32):
            os.environ["SM_TILE_SHIFT_X"], os.environ["SM_TILE_SHIFT_Y"] = str(sx), str(sy)
            be = sm.CudaBackend.new(W, H, agent_count=N, sort_interval=interval)
            be.init_agents(1)
            be.step(200)
            steps = 96
            ms = event_time(be, lambda: be.step(steps))
            be.set_timing_enabled(True); be.reset_timing()
            be.step(steps)
            t = be.timing()
            emit({"sweep": "agents", "tile": [1 << sx, 1 << sy], "sort_interval": interval, "ms_per_step": ms / steps,
                  "agent_steps_per_s": N * steps / (ms * 1e-3), "agents_ms": t.agents_ms / t.agent_launches,
                  "trail_ms": t.trail_ms / t.trail_launches, "sort_ms_each": t.sort_ms / max(t.sort_launches, 1),
                  "sort_ms_per_step": t.sort_ms / steps})
            be.close()
    os.environ.pop("SM_TILE_SHIFT_X", None); os.environ.pop("SM_TILE_SHIFT_Y", None)
    # no sort at all (random order forever)
    be = sm.CudaBackend.new(W, H, agent_count=N, flags=sm.SM_FLAG_NO_SORT)
    be.init_agents(1); be.step(200)
    ms = event_time(be, lambda: be.step(48))
    emit({"sweep": "agents", "tile": None, "sort_interval": 0, "ms_per_step": ms / 48, "agent_steps_per_s": N * 48 / (ms * 1e-3)})
    be.close()


def presets_sweep():
    N, W, H = 16_777_216, 4096, 4096
    pm = sm.init_preset_manager()
    for name in pm.get_preset_names():
        be = sm.CudaBackend.new(W, H, pm.get_preset(name).settings, agent_count=N)
        be.init_agents(1)
        ms0 = event_time(be, lambda: be.step(50))             # uniform-random initial state
        be.step(450)
        ms1 = event_time(be, lambda: be.step(100))            # steady state (>= 500 steps in)
        st = be.trail_statistics()
        emit({"sweep": "presets", "preset": name, "initial_agent_steps_per_s": N * 50 / (ms0 * 1e-3),
              "steady_agent_steps_per_s": N * 100 / (ms1 * 1e-3), "steady_ms_per_step": ms1 / 100,
              "trail_mean": st.sum / (W * H), "occupied_frac": st.nonzero / (W * H)})
        be.close()


def gauss_sweep():
    """EXTENSION (no reference semantics): separable Gaussian, radius 2 / 4 / 8, fused shared-memory kernel vs the
    two-pass form (BASELINE config 5 names radii 1-8; radius 1 = the reference's box is the parity mode above)."""
    for S in (8192, 16384):
        for R in (2, 4, 8):
            for kern in ("fused_packed", "fused_scalar", "two_pass"):
                two_pass = kern == "two_pass"
                os.environ["SM_GAUSS_TWO_PASS"] = "1" if two_pass else "0"
                os.environ["SM_GAUSS_PACKED"] = "1" if kern == "fused_packed" else "0"
                s = sm.Settings.default().clone(blur_radius=float(R), blur_sigma=R / 2.0)
                be = sm.CudaBackend.new(S, S, s, agent_count=1, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
                be.write_trail(np.random.default_rng(0).random((256, S), dtype=np.float32), y0=0)
                passes = max(6, min(60, int(6e9 / (S * S * 8))))
                be.diffuse_only(3)
                ms = event_time(be, lambda: be.diffuse_only(passes))
                gbs = 8.0 * S * S * passes / (ms * 1e-3) / 1e9
                emit({"sweep": "gauss", "size": S, "radius": R, "kernel": kern, "passes": passes,
                      "ms_per_pass": ms / passes, "gbs": gbs, "frac_of_measured_peak": gbs / PEAK, "frac_of_8TBs": gbs / 8000.0})
                be.close()
    os.environ.pop("SM_GAUSS_TWO_PASS", None)
    os.environ.pop("SM_GAUSS_PACKED", None)


def gauss_stream_sweep():
    """EXTENSION: the streaming Gaussian kernel (gauss_stream.cuh, SM_GAUSS_KERNEL=stream) against the tile kernel
    (k_gauss_fused): diffusion-only passes for radius 1 / 2 / 4 / 8, chunk heights, and the full step in Gaussian
    mode at config-2 size (stream: u8 flags + sampler copy kept in step; tile: u32 counts + per-step array copy)."""
    for S in (8192, 16384):
        for R in (1, 2, 4, 8):
            for kern, chunk in (("tile", 0), ("stream", 0), ("stream", 128), ("stream", 512)):
                if chunk and S != 8192:
                    continue
                os.environ["SM_GAUSS_KERNEL"] = kern
                os.environ["SM_GAUSS_CHUNK"] = str(chunk)
                s = sm.Settings.default().clone(blur_radius=float(R), blur_sigma=R / 2.0)
                be = sm.CudaBackend.new(S, S, s, agent_count=1, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
                be.write_trail(np.random.default_rng(0).random((256, S), dtype=np.float32), y0=0)
                passes = max(6, min(60, int(6e9 / (S * S * 8))))
                be.diffuse_only(3)
                ms = event_time(be, lambda: be.diffuse_only(passes))
                gbs = 8.0 * S * S * passes / (ms * 1e-3) / 1e9
                emit({"sweep": "gauss_stream", "size": S, "radius": R, "kernel": kern, "chunk": chunk, "passes": passes,
                      "ms_per_pass": ms / passes, "gbs": gbs, "frac_of_measured_peak": gbs / PEAK, "frac_of_8TBs": gbs / 8000.0})
                be.close()
    N, W, H = 16_777_216, 4096, 4096
    for R in (2, 8):
        for kern in ("tile", "stream"):
            os.environ["SM_GAUSS_KERNEL"] = kern
            os.environ["SM_GAUSS_CHUNK"] = "0"
            s = sm.Settings.default().clone(blur_radius=float(R), blur_sigma=R / 2.0)
            be = sm.CudaBackend.new(W, H, s, agent_count=N, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
            be.init_agents(1)
            be.step(100)
            steps = 96
            ms = event_time(be, lambda: be.step(steps))
            be.set_timing_enabled(True); be.reset_timing()
            be.step(48)
            t = be.timing()
            emit({"sweep": "gauss_full_step", "radius": R, "kernel": kern, "ms_per_step": ms / steps,
                  "agent_steps_per_s": N * steps / (ms * 1e-3), "agents_ms": t.agents_ms / t.agent_launches,
                  "trail_ms": t.trail_ms / t.trail_launches, "sort_ms_per_step": t.sort_ms / 48})
            be.close()
    os.environ.pop("SM_GAUSS_KERNEL", None)
    os.environ.pop("SM_GAUSS_CHUNK", None)


def gauss_rows_sweep():
    """EXTENSION: the register-streaming Gaussian kernel (gauss_rows.cuh; scalar and FFMA2-packed column taps) against the
    shared-memory streaming kernel -- BASELINE config 5's radii 1-8 on 8192^2 and 16384^2, chunk heights, and the full
    step in Gaussian mode at config-2 size."""
    for S in (8192, 16384):
        for R in (1, 2, 3, 4, 5, 6, 7, 8):
            variants = [("rows", 0), ("rows_packed", 0)]
            if S == 8192:
                variants.append(("stream", 0))
                if R in (2, 5, 8):
                    variants += [("rows_packed", 64), ("rows_packed", 128), ("rows_packed", 512)]
            for kern, chunk in variants:
                os.environ["SM_GAUSS_KERNEL"] = "rows" if kern.startswith("rows") else kern
                os.environ["SM_GAUSS_ROWS_PACKED"] = "1" if kern == "rows_packed" else "0"
                os.environ["SM_GAUSS_CHUNK"] = str(chunk)
                s = sm.Settings.default().clone(blur_radius=float(R), blur_sigma=R / 2.0)
                be = sm.CudaBackend.new(S, S, s, agent_count=1, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
                be.write_trail(np.random.default_rng(0).random((256, S), dtype=np.float32), y0=0)
                passes = max(6, min(60, int(6e9 / (S * S * 8))))
                be.diffuse_only(3)
                ms = event_time(be, lambda: be.diffuse_only(passes))
                gbs = 8.0 * S * S * passes / (ms * 1e-3) / 1e9
                emit({"sweep": "gauss_rows", "size": S, "radius": R, "kernel": kern, "chunk": chunk, "passes": passes,
                      "ms_per_pass": ms / passes, "gbs": gbs, "frac_of_measured_peak": gbs / PEAK, "frac_of_8TBs": gbs / 8000.0})
                be.close()
    os.environ.pop("SM_GAUSS_ROWS_PACKED", None)
    N, W, H = 16_777_216, 4096, 4096
    for R, kern in ((2, "rows"), (2, "rows_packed"), (4, "rows_packed"), (8, "rows_packed"), (8, "stream")):
        os.environ["SM_GAUSS_KERNEL"] = "rows" if kern.startswith("rows") else kern
        os.environ["SM_GAUSS_ROWS_PACKED"] = "1" if kern == "rows_packed" else "0"
        os.environ["SM_GAUSS_CHUNK"] = "0"
        s = sm.Settings.default().clone(blur_radius=float(R), blur_sigma=R / 2.0)
        be = sm.CudaBackend.new(W, H, s, agent_count=N, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
        be.init_agents(1)
        be.step(100)
        steps = 96
        ms = event_time(be, lambda: be.step(steps))
        be.set_timing_enabled(True); be.reset_timing()
        be.step(48)
        t = be.timing()
        emit({"sweep": "gauss_full_step", "radius": R, "kernel": kern, "ms_per_step": ms / steps,
              "agent_steps_per_s": N * steps / (ms * 1e-3), "agents_ms": t.agents_ms / t.agent_launches,
              "trail_ms": t.trail_ms / t.trail_launches, "sort_ms_per_step": t.sort_ms / 48})
        be.close()
    os.environ.pop("SM_GAUSS_KERNEL", None)
    os.environ.pop("SM_GAUSS_CHUNK", None)
    os.environ.pop("SM_GAUSS_ROWS_PACKED", None)


def gauss_packed_sweep():
    """EXTENSION: A/B of the FFMA2 forms -- rows kernel radius 1-5 with SM_GAUSS_ROWS_PACKED = 0 / 1 / 2, streaming kernel radius
    5-8 with SM_GAUSS_STREAM_PACKED = 0 / 1 -- diffusion-only passes on 8192^2 (and 16384^2 for the candidates)."""
    def run(S, R, kern, env, label):
        os.environ["SM_GAUSS_KERNEL"] = kern
        os.environ["SM_GAUSS_CHUNK"] = "0"
        for k, v in env.items():
            os.environ[k] = v
        s = sm.Settings.default().clone(blur_radius=float(R), blur_sigma=R / 2.0)
        be = sm.CudaBackend.new(S, S, s, agent_count=1, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
        be.write_trail(np.random.default_rng(0).random((256, S), dtype=np.float32), y0=0)
        passes = max(6, min(60, int(6e9 / (S * S * 8))))
        be.diffuse_only(3)
        ms = event_time(be, lambda: be.diffuse_only(passes))
        gbs = 8.0 * S * S * passes / (ms * 1e-3) / 1e9
        emit({"sweep": "gauss_rows", "size": S, "radius": R, "kernel": label, "chunk": 0, "passes": passes,
              "ms_per_pass": ms / passes, "gbs": gbs, "frac_of_measured_peak": gbs / PEAK, "frac_of_8TBs": gbs / 8000.0})
        be.close()
        for k in env:
            os.environ.pop(k, None)
    for S in (8192, 16384):
        for R in (1, 2, 3, 4, 5):
            for pk in (0, 1, 2):
                run(S, R, "rows", {"SM_GAUSS_ROWS_PACKED": str(pk)}, f"rows_pk{pk}")
        for R in (5, 6, 7, 8):
            for pk in (0, 1):
                run(S, R, "stream", {"SM_GAUSS_STREAM_PACKED": str(pk)}, f"stream_pk{pk}")
    os.environ.pop("SM_GAUSS_KERNEL", None)
    os.environ.pop("SM_GAUSS_CHUNK", None)


def gauss_rows_ncu_target():
    """A few passes of the register-streaming kernel at 8192^2 (radius 4, then radius 8; packed column taps): the ncu target."""
    os.environ["SM_GAUSS_KERNEL"] = "rows"
    os.environ["SM_GAUSS_ROWS_PACKED"] = "1"
    for R in (4, 8):
        s = sm.Settings.default().clone(blur_radius=float(R), blur_sigma=R / 2.0)
        be = sm.CudaBackend.new(8192, 8192, s, agent_count=1, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
        be.write_trail(np.random.default_rng(0).random((256, 8192), dtype=np.float32), y0=0)
        be.diffuse_only(3)
        be.sync()
        be.close()
    os.environ.pop("SM_GAUSS_KERNEL", None)


def gauss_ncu_target():
    """A few streaming-kernel passes at 8192^2 (radius 8, then radius 2): the ncu target."""
    os.environ["SM_GAUSS_KERNEL"] = "stream"
    for R in (8, 2):
        s = sm.Settings.default().clone(blur_radius=float(R), blur_sigma=R / 2.0)
        be = sm.CudaBackend.new(8192, 8192, s, agent_count=1, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
        be.write_trail(np.random.default_rng(0).random((256, 8192), dtype=np.float32), y0=0)
        be.diffuse_only(3)
        be.sync()
        be.close()
    os.environ.pop("SM_GAUSS_KERNEL", None)


def diffusion_16k():
    """A handful of diffusion-only passes on a 16384^2 field (1 GiB in, 1 GiB out): the ncu target."""
    S = 16384
    be = sm.CudaBackend.new(S, S, agent_count=1)
    be.write_trail(np.random.default_rng(0).random((256, S), dtype=np.float32), y0=0)
    be.diffuse_only(12)
    be.sync()
    be.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["diffusion", "agents", "presets"]
    if "diffusion16k" in which:
        diffusion_16k()
    t0 = time.time()
    if "diffusion" in which:
        diffusion_sweep()
    if "agents" in which:
        agents_sweep()
    if "presets" in which:
        presets_sweep()
    if "gauss" in which:
        gauss_sweep()
    if "gauss_stream" in which:
        gauss_stream_sweep()
    if "gauss_rows" in which:
        gauss_rows_sweep()
    if "gauss_packed" in which:
        gauss_packed_sweep()
    if "gauss_ncu" in which:
        gauss_ncu_target()
    if "gauss_rows_ncu" in which:
        gauss_rows_ncu_target()
    print("sweeps done in", round(time.time() - t0, 1), "s")
