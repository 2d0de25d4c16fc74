#!/usr/bin/env python
"""bench.py -- headline benchmark of the slime-mold step loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): agent-steps/s of the full step (agents -> decay -> diffuse,
/root/reference/src/main.rs:1163-1235) plus the diffusion kernel's HBM GB/s.

* Headline workload (`value`, `e2e`, `roofline`) = BASELINE.json configs[2], the largest single-GPU configuration:
  100,000,000 agents on an 8192 x 8192 trail map, sensor distance 225 / sensor angle 1.34 (SURVEY.md 8d "config 3"),
  other parameters Default, the reference's 3x3 box blur.  N > 1: the same work per GPU ("weak"): 100 M x N agents
  on 8192 x (8192 N) cells cut into N horizontal strips with halo exchange + migration.
* `initial_state` / `steady_state`: the first steps from the uniform-random start on an empty map, and K steps after
  the spin-up (`value` is the steady-state figure).
* `config2_default`: BASELINE configs[1] (16.7 M agents, 4096^2, Default preset; weak-scaled at N > 1) -- round 1's headline.
* `gaussian_r8` (N = 1): configs[2] with the radius-8 Gaussian EXTENSION instead of the box blur (no reference semantics).
* `parity_n` (N > 1): BEFORE anything is timed, a strip run (512 x 192 N map, 300 k agents, 35 steps) is compared bit for
  bit with the single-domain CPU oracle (checker use of oracle/); a mismatch aborts the run.
* `config4` (N > 1): BASELINE configs[3]: 1 B agents on 32768^2, Default preset, N strips.
* `diffusion`: sm_diffuse_only on a 16384 x 16384 map per GPU (1 GiB per field: larger than the 126 MB L2).
* `e2e`: the frame loop driven the way the reference drives it -- 56-byte uniform from host memory (main.rs:98), one
  frame, a field statistic read back to the host -- one host round trip per step; `e2e_host_state`: the other extreme,
  the whole state crossing PCIe around every step.
* `roofline`: the dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak.
* `cpu_baseline` / `--impl reference`: the CPU restatement of compute.wgsl (oracle/, OpenMP, every host core, the FULL
  workload -- no agent sub-sampling) -- the reference itself (Rust + wgpu) cannot be built or run in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
T_START = time.perf_counter()

CONFIG3 = dict(agents=100_000_000, width=8192, height=8192, sd=225.0, sa=1.34)     # BASELINE configs[2]
CONFIG2 = dict(agents=16_777_216, width=4096, height=4096)                          # BASELINE configs[1]
CONFIG4 = dict(agents=1_000_000_000, width=32768, height=32768)                     # BASELINE configs[3]
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="Default")
    ap.add_argument("--agents", type=int, default=CONFIG3["agents"], help="agents per GPU")
    ap.add_argument("--width", type=int, default=CONFIG3["width"])
    ap.add_argument("--height", type=int, default=CONFIG3["height"], help="map rows per GPU")
    ap.add_argument("--sd", type=float, default=None, help="sensor distance (default: 225 on the headline workload, else the preset's)")
    ap.add_argument("--sa", type=float, default=None, help="sensor angle (default: 1.34 on the headline workload, else the preset's)")
    ap.add_argument("--spinup", type=int, default=96, help="untimed steps between the initial-state and the steady-state measurement")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--gaussian", type=int, default=0, metavar="R", help="EXTENSION: Gaussian blur of radius R on the headline workload (N = 1)")
    ap.add_argument("--only-headline", action="store_true", help="skip the side blocks (config2_default, gaussian_r8, config4, diffusion, host state, cpu)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-state", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--host-state-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=240.0, help="budget of the --impl reference run")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e leg (0 = min(steps, 100))")
    return ap.parse_args()


def headline_settings(args):
    """Preset parameters; on the default (configs[2]) workload the sensor distance / angle of SURVEY.md 8d config 3."""
    import slime_mold_b200 as sm
    s = sm.init_preset_manager().get_preset(args.preset).settings
    is_c3 = (args.agents, args.width, args.height) == (CONFIG3["agents"], CONFIG3["width"], CONFIG3["height"]) and args.preset == "Default"
    sd = args.sd if args.sd is not None else (CONFIG3["sd"] if is_c3 else s.agent_sensor_distance)
    sa = args.sa if args.sa is not None else (CONFIG3["sa"] if is_c3 else s.agent_sensor_angle)
    s = s.clone(agent_sensor_distance=sd, agent_sensor_angle=sa)
    if args.gaussian:
        s = s.clone(blur_radius=float(args.gaussian), blur_sigma=args.gaussian / 2.0)
    return s


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML inside this process (a thread polling every ~2 ms) -- the timed
    region of a short run (20 steps x 2 ms) is over before a separate `nvidia-smi -lms` process prints its first row."""

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []          # (t, sm_mhz, reasons bitmask, power_w)
        self.marks = {}
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = gpu_index
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except Exception:
                    idx = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as exc:      # noqa: BLE001
            self.err = repr(exc)[:120]

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                self.rows.append((time.perf_counter(), float(mhz), int(rs), pw))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()

    def mark(self, name):
        self.marks[name] = time.perf_counter()

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")], "samples": 0}
        self._stop.set()
        self.t.join(timeout=2)
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        t0, t1 = self.marks.get("timed_begin", 0.0), self.marks.get("timed_end", float("inf"))
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        use = inside if inside else self.rows          # the per-kernel pass that follows runs the same kernels
        sm = sorted(r[1] for r in use)
        reasons = set()
        for r in use:
            for name, bit in names:
                if r[2] & bit:
                    reasons.add(name)
        pw = [r[3] for r in use if r[3] is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(reasons),
                "samples": len(inside), "samples_total": len(self.rows), "power_w_max": max(pw) if pw else None,
                "how": "NVML polled every ~2 ms in-process; median over the samples taken inside the timed region"}


# ----------------------------------------------------------------------------------------------
# reference arm: CPU restatement of compute.wgsl on the host cores (oracle/, OpenMP)
# ----------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_params(so, width, height, s):
    return so.make_params(width, height, decay_factor=s.pheromone_decay_factor, agent_jitter=s.agent_jitter,
                          agent_speed_min=s.agent_speed_min, agent_speed_max=s.agent_speed_max,
                          agent_turn_speed=s.agent_turn_speed, agent_sensor_angle=s.agent_sensor_angle,
                          agent_sensor_distance=s.agent_sensor_distance, diffusion_rate=s.pheromone_diffusion_rate,
                          pheromone_deposition_amount=s.pheromone_deposition_amount)


def cpu_reference_run(width, height, agents, s, seed, steps, warmup, budget_s, gaussian=0):
    """Times the oracle (phase_split semantics) on the FULL workload with every host core (torchrun pins OMP_NUM_THREADS=1,
    so the thread count is set explicitly).  The sample is bounded by running fewer STEPS, never fewer agents.
    Returns (agent_steps_per_s, ms_per_step, sample description, cores, steps actually timed)."""
    from oracle import slime_oracle as so            # bench.py's cpu_baseline / reference leg only
    os.environ["SM_ORACLE_NATIVE"] = "1"             # -O3 -march=native build of the same source, made on this box
    so.build()
    cores = host_cores()
    so.set_threads(cores)
    p = oracle_params(so, width, height, s)
    ag = so.init_agents(agents, width, height, s.agent_speed_min, s.agent_speed_max, seed)
    Sim = so.Sim
    if gaussian:
        class Sim(so.Sim):
            """Gaussian extension: the oracle's phase-split agents pass followed by its Gaussian trail pass."""

            def step(self, n=1):
                for _ in range(n):
                    so.agents_phase_split(self.agents, self.trail, self.counts, self.p)
                    self.trail = so.trail_pass(self.trail, self.p, counts=self.counts, gauss_radius=gaussian, gauss_sigma=gaussian / 2.0)
    sim = Sim(p, ag)
    t0 = time.perf_counter()
    sim.step(1)                                       # calibration step (also the first warm-up step)
    t1 = time.perf_counter() - t0
    n_warm = max(warmup - 1, 0)
    n_timed = steps
    if t1 * (n_warm + n_timed) > budget_s:
        n_warm = min(n_warm, 1)
        n_timed = max(2, min(steps, int(budget_s / t1) - n_warm))
    for _ in range(n_warm):
        sim.step(1)
    t0 = time.perf_counter()
    sim.step(n_timed)
    dt = time.perf_counter() - t0
    sample = (f"all {agents} agents on the full {width}x{height} map, {n_timed} timed steps"
              f"{' (of the %d requested: bounded by --ref-seconds / --cpu-seconds)' % steps if n_timed != steps else ''}, seed {seed}, "
              f"sensor distance {s.agent_sensor_distance:g}{', Gaussian blur radius %d' % gaussian if gaussian else ''}, phase_split semantics, "
              f"OpenMP {cores} threads, {so.build_flavour()}")
    return agents * n_timed / dt, 1e3 * dt / n_timed, sample, cores, n_timed


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                        # rank 0 alone runs the CPU arm
    N = max(args.gpus, 1)
    width, height, agents = args.width, args.height * N, args.agents * N
    s = headline_settings(args)
    val, ms, sample, cores, n_timed = cpu_reference_run(width, height, agents, s, args.seed, args.steps, args.warmup,
                                                        budget_s=args.ref_seconds, gaussian=args.gaussian)
    line = {
        "impl": "reference", "metric": "agent_steps_per_sec", "value": val, "unit": "agent-steps/s", "n_gpus": N,
        "steps": args.steps, "steps_timed": n_timed, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, N, s),
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of compute.wgsl (oracle/): the Rust+wgpu reference cannot be built here (no rustc/Vulkan)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, N, s):
    total_agents, total_h = args.agents * N, args.height * N
    per = (args.agents, args.width, args.height)
    if per == (CONFIG3["agents"], CONFIG3["width"], CONFIG3["height"]):
        label = "BASELINE configs[2]" + (" per GPU (weak scaling)" if N > 1 else "")
    elif per == (CONFIG2["agents"], CONFIG2["width"], CONFIG2["height"]):
        label = "BASELINE configs[1]" + (" per GPU (weak scaling)" if N > 1 else "")
    elif (total_agents, args.width, total_h) == (1000000, 1920, 1080):
        label = "BASELINE configs[0]"
    elif (total_agents, args.width, total_h) == (CONFIG4["agents"], CONFIG4["width"], CONFIG4["height"]):
        label = "BASELINE configs[3]"
    else:
        label = "custom size"
    what = (f"{args.agents} agents on a {args.width}x{args.height} trail map{' per GPU' if N > 1 else ''}, preset {args.preset} with "
            f"sensor distance {s.agent_sensor_distance:g} / sensor angle {s.agent_sensor_angle:g}")
    return {
        "workload": f"{label}: {what}" + (f"; x{N} strips" if N > 1 else ""),
        "agents": total_agents, "width": args.width, "height": total_h, "preset": args.preset,
        "sensor_distance": s.agent_sensor_distance, "sensor_angle": s.agent_sensor_angle,
        "parallelism": f"strips{N}" if N > 1 else "single",
        "exchange": (os.environ.get("SM_EXCHANGE") or "p2p") if N > 1 else None,
        "spinup_steps": args.spinup, "seed": args.seed,
        "blur": (f"EXTENSION: separable Gaussian, radius {args.gaussian}, sigma {args.gaussian / 2.0} (no reference semantics)"
                 if args.gaussian else "3x3 box (compute.wgsl:164-195)"),
        "l2": (f"inputs larger than L2: {args.agents * 20 / 1e6:.0f} MB of agent state and {args.width * args.height * 4 / 1e6:.0f} MB of trail "
               f"per GPU are streamed every step (no flush between steps)"
               if args.agents * 20 + args.width * args.height * 4 > 2 * 126e6 else
               "working set near the 126 MB L2: partly L2-resident, not a pure DRAM number; no flush"),
    }


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
class Harness:
    """Rank / device plumbing shared by every block of the run."""

    def __init__(self, N):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.N = N
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if N > 1 and world != N:
            raise SystemExit(f"--gpus {N} needs torchrun with {N} ranks (WORLD_SIZE={world})")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if N > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def engine(self, width, height, settings, agents, flags=0, **kw):
        import slime_mold_b200 as sm
        be = sm.CudaBackend.new(width, height, settings, agent_count=agents, device=self.local_rank, rank=self.rank if self.N > 1 else 0,
                                world_size=self.N, flags=flags, **kw)
        if self.N > 1:
            ids = [be.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(ids, src=0)
            be.comm_init(ids[0])
        return be

    def barrier(self, be=None):
        if be is not None:
            be.sync()
        self.torch.cuda.synchronize()
        if self.N > 1:
            self.dist.barrier()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        if self.N > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, x):
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        if self.N > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())

    def timed(self, be, fn):
        """CUDA events on the engine's stream, barrier + sync on both sides, max over ranks (ms)."""
        stream = self.torch.cuda.ExternalStream(be.stream_handle, device=self.local_rank)
        self.barrier(be)
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        self.barrier(be)
        return self.max_over_ranks(e0.elapsed_time(e1))


def kernel_split(be, steps, local_agents, cells_local, peak):
    """Per-launch CUDA-event times of the agent kernel / trail pass on the same workload, and what they mean against the
    roofline.  Algorithmic bytes: agents 32 B/agent-step + trail sensing read 4 + deposit write 4 B/cell;
    fused decay+diffuse 8 B/cell-pass."""
    be.set_timing_enabled(True)
    be.reset_timing()
    be.step(steps)
    t = be.timing()
    be.set_timing_enabled(False)
    agent_ms = t.agents_ms / max(t.agent_launches, 1)
    trail_ms = t.trail_ms / max(t.trail_launches, 1)
    agent_bytes = 32.0 * local_agents + 8.0 * cells_local
    trail_bytes = 8.0 * cells_local
    kern = {
        "agents": {"ms": agent_ms, "alg_bytes": agent_bytes, "gbs": agent_bytes / (agent_ms * 1e-3) / 1e9 if agent_ms else None},
        "trail": {"ms": trail_ms, "alg_bytes": trail_bytes, "gbs": trail_bytes / (trail_ms * 1e-3) / 1e9 if trail_ms else None},
        "sort_ms_per_step": t.sort_ms / max(t.steps, 1),
        "exchange_ms_per_step": t.exchange_ms / max(t.steps, 1),
        "step_alg_bytes": agent_bytes + trail_bytes,
    }
    for k in ("agents", "trail"):
        kern[k]["frac_of_hbm_peak"] = kern[k]["gbs"] / peak if kern[k]["gbs"] else None
    return kern


def note(h, msg):
    """Progress on stderr (never on stdout: the JSON line stays alone there)."""
    print(f"[bench rank {h.rank} +{time.perf_counter() - T_START:.1f}s] {msg}", file=sys.stderr, flush=True)


def measure(h, name, width, rows_per_gpu, agents_per_gpu, settings, args, steps, spinup, flags=0, want_e2e=False, clocks=None,
            init_steps=24):
    """One workload: initial-state steps, spin-up, warm-up, K timed steps, the per-kernel split (+ the e2e frame loop)."""
    from slime_mold_b200.settings import SimSizeUniform
    N = h.N
    note(h, f"measure: {name}")
    width, height, agents = width, rows_per_gpu * N, agents_per_gpu * N
    be = h.engine(width, height, settings, agents, flags=flags)
    be.init_agents(args.seed)
    be.set_timing_enabled(False)
    if clocks is not None:          # rank 0 only (no collective in here); started early: the first NVML calls take tens of ms
        clocks.start()
    peak, _ = hbm_peak()
    out = {"workload": name, "agents": agents, "width": width, "height": height}
    # initial state: uniform-random agents on an empty map (the first sort is part of it, as in any run)
    init_steps = max(1, min(init_steps, steps))
    ms_init = h.timed(be, lambda: be.step(init_steps))
    out["initial_state"] = {"value": agents * init_steps / (ms_init * 1e-3), "ms_per_step": ms_init / init_steps, "steps": init_steps,
                            "note": "first steps after sm_init_agents on a zeroed trail map (includes the first cell sort)"}
    be.step(max(spinup - init_steps, 0))              # untimed: the network forms; gather locality / deposit contention settle
    be.step(args.warmup)
    be.sync()
    be.reset_timing()
    if clocks is not None:
        clocks.mark("timed_begin")
    ms_total = h.timed(be, lambda: be.step(steps))
    if clocks is not None:
        clocks.mark("timed_end")
    out["gpu_launches"] = int(be.timing().kernel_launches)
    out["value"] = agents * steps / (ms_total * 1e-3)
    out["ms_per_step"] = ms_total / steps
    out["steady_state"] = {"value": out["value"], "ms_per_step": out["ms_per_step"], "steps": steps,
                           "note": f"after {max(spinup, init_steps)} spin-up + {args.warmup} warm-up steps"}
    local_agents = be.local_agent_count
    cells_local = width * rows_per_gpu
    out["kernels"] = kernel_split(be, min(steps, 48), local_agents, cells_local, peak)
    out["step_frac_of_hbm_peak"] = out["kernels"]["step_alg_bytes"] / (out["ms_per_step"] * 1e-3) / 1e9 / peak
    if want_e2e:
        # host-driven frame loop (uniform from host each step, statistic back each step)
        e2e_steps = args.e2e_steps or min(steps, 100)
        uni = SimSizeUniform.new(width, height, settings.pheromone_decay_factor, settings)
        st = be.trail_statistics()                    # arms the fused statistics of the trail pass
        h.barrier(be)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            be.write_uniform(uni)            # 56 B host -> device (kernel parameter block)
            be.step(1)
            st = be.trail_statistics()       # 32 B device -> host, synchronises
        h.barrier(be)
        e2e_s = h.max_over_ranks(time.perf_counter() - t0)
        out["e2e"] = {"value": agents * e2e_steps / e2e_s, "unit": "agent-steps/s", "h2d_bytes_per_step": 56, "d2h_bytes_per_step": 32,
                      "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                      "note": "per step: sm_set_params (56-byte uniform from host), sm_step(1), sm_trail_statistics read back (host sync every step)"}
        out["last_trail_mean"] = st.sum / cells_local
    be.close()
    return out


def parity_before_timing(h, args):
    """N > 1: a strip run against the single-domain CPU oracle, bit for bit, before anything is timed."""
    import numpy as np
    import slime_mold_b200 as sm
    from oracle import slime_oracle as so          # checker
    so.build()
    so.set_threads(max(1, min(8, host_cores() // h.N)))
    N = h.N
    note(h, "parity before timing")
    W, H, n_agents, steps, seed = 512, 192 * N, 300_000, 35, 11
    s = sm.init_preset_manager().get_preset("Default").settings
    sim = so.Sim(oracle_params(so, W, H, s), so.init_agents(n_agents, W, H, s.agent_speed_min, s.agent_speed_max, seed))
    sim.step(steps)
    from slime_mold_b200._lib import tuning_from_env
    agents_equal = trail_equal = True
    total_owned = n_agents
    # twice: with the u8 deposit flags row-major (what a map of this size gets) and in 8x8 tiles (what the timed workloads get)
    for layout in (1, 2):
        tn = tuning_from_env()
        tn.deposit_flag_layout = layout
        be = h.engine(W, H, s, n_agents, tuning=tn)
        be.init_agents(seed)
        be.step(steps)
        a = be.read_agents()
        owned, local = be.last_owned, be.local_agent_count
        t = be.read_trail()
        be.close()
        mine = ~np.isnan(a[:, 0])
        rows = ~np.isnan(t[:, 0])
        agents_equal &= bool(np.array_equal(a[mine].view(np.uint32), sim.agents[mine].view(np.uint32))) and owned == local == int(mine.sum())
        trail_equal &= bool(np.array_equal(t[rows].view(np.uint32), sim.trail[rows].view(np.uint32))) and int(rows.sum()) == H // N
        tt = h.torch.tensor([float(owned)], device="cuda", dtype=h.torch.float64)
        h.dist.all_reduce(tt)
        if int(tt.item()) != n_agents:
            total_owned = int(tt.item())          # reported (and fatal below)
    res = {"ranks": N, "agents_equal": h.min_over_ranks(1.0 if agents_equal else 0.0) == 1.0,
           "trail_equal": h.min_over_ranks(1.0 if trail_equal else 0.0) == 1.0, "agents_owned_total": total_owned, "agents": n_agents,
           "case": f"{W}x{H} map in {N} strips, {n_agents} agents, {steps} steps, Default preset, device-side init seed {seed}, P2P exchange, "
                   f"run twice (u8 deposit flags row-major and in 8x8 tiles); "
                   f"every rank compares the agents and trail rows it owns with the single-domain oracle bit for bit"}
    res["agents_equal"] = res["agents_equal"] and total_owned == n_agents
    if not (res["agents_equal"] and res["trail_equal"]):
        if h.rank == 0:
            print(json.dumps({"impl": "ours", "error": "strip parity FAILED before timing", "parity_n": res}), flush=True)
        raise SystemExit(3)
    return res


def diffusion_block(h, args, size=16384):
    """sm_diffuse_only on a map larger than L2: size x size cells per GPU (1 GiB per field at 16384^2)."""
    import numpy as np
    import slime_mold_b200 as sm
    peak, _ = hbm_peak()
    N = h.N
    note(h, "diffusion block")
    rows = min(size, 65536 // N)                       # the engine's maps are at most 65536 rows: 8 GPUs take 8192 rows each (512 MiB per field)
    be = h.engine(size, rows * N, sm.Settings.default(), 1)
    rng = np.random.default_rng(0)
    band = rng.random((256, size), dtype=np.float32)
    for y0 in range(0, rows * N, 256 * 16):           # sparse non-zero bands: the pass costs the same whatever the values
        be.write_trail(band, y0=y0)
    be.diffuse_only(3)
    passes = 20
    ms = h.timed(be, lambda: be.diffuse_only(passes))
    gbs = 8.0 * size * rows * N * passes / (ms * 1e-3) / 1e9
    be.close()
    return {"gbs": gbs, "frac_of_peak": gbs / (peak * N), "frac_of_nominal_8TBs": gbs / (8000.0 * N), "passes": passes, "ms_per_pass": ms / passes,
            "alg_bytes_per_cell": 8, "map": [size, rows * N],
            "note": f"sm_diffuse_only (decay + 3x3 box, compute.wgsl:148-195) on {size}x{rows} cells per GPU = {size * rows * 4 / 2**20:.0f} MiB per field: "
                    f"larger than the 126 MB L2, a DRAM number"}


def run_ours(args):
    import slime_mold_b200 as sm
    N = max(args.gpus, 1)
    h = Harness(N)
    rank = h.rank
    if args.gaussian and N > 1:
        raise SystemExit("--gaussian is single-GPU only (the extension's full step is not built for strips)")
    settings = headline_settings(args)
    eng_flags = sm.SM_FLAG_GAUSSIAN_BLUR if args.gaussian else 0

    parity = parity_before_timing(h, args) if N > 1 else None

    clocks = ClockSampler(h.local_rank) if rank == 0 else None
    head = measure(h, "headline", args.width, args.height, args.agents, settings, args, args.steps, args.spinup, flags=eng_flags,
                   want_e2e=True, clocks=clocks)
    clk = clocks.stop() if clocks is not None else None
    peak, peak_src = hbm_peak()
    kern = head["kernels"]
    dom = "agents" if kern["agents"]["ms"] >= kern["trail"]["ms"] else "trail"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            key = ("agents_sd225" if settings.agent_sensor_distance > 100 else "agents") if dom == "agents" else "trail"
            traffic = tj.get(key, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_agents" if dom == "agents" else "k_trail_rows", "achieved": kern[dom]["gbs"],
                "peak": peak, "unit": "GB/s", "frac": kern[dom]["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": kern[dom]["alg_bytes"], "ms_per_launch": kern[dom]["ms"],
                "whole_step_frac": head["step_frac_of_hbm_peak"]}

    side = {}
    if not args.only_headline:
        # round 1's headline, for round-over-round comparison: BASELINE configs[1], Default preset (weak-scaled on strips)
        s2 = sm.init_preset_manager().get_preset("Default").settings
        c2 = measure(h, "BASELINE configs[1]: 16,777,216 agents on 4096x4096 per GPU, Default preset", CONFIG2["width"], CONFIG2["height"],
                     CONFIG2["agents"], s2, args, max(args.steps, 96), 200, want_e2e=(N == 1))
        side["config2_default"] = c2
        if N == 1 and not args.gaussian:
            s8 = settings.clone(blur_radius=8.0, blur_sigma=4.0)
            g8 = measure(h, "BASELINE configs[2] with the radius-8 Gaussian EXTENSION (sigma 4; no reference semantics)", args.width, args.height,
                         args.agents, s8, args, min(args.steps, 48), 48, flags=sm.SM_FLAG_GAUSSIAN_BLUR)
            side["gaussian_r8"] = g8
        if N > 1 and not args.no_config4:
            s4 = sm.init_preset_manager().get_preset("Default").settings
            c4 = measure(h, f"BASELINE configs[3]: 1,000,000,000 agents on 32768x32768, Default preset, {N} strips", CONFIG4["width"],
                         CONFIG4["height"] // N, CONFIG4["agents"] // N, s4, args, min(args.steps, 50), 60, init_steps=10)
            c4["target"] = "north star: >= 1e11 agent-steps/s on 8 x B200"
            side["config4"] = c4
        side["diffusion"] = diffusion_block(h, args)

    # ---- the other end of the e2e scale: the HOST owns the whole state (agents + trail uploaded before and read back after
    #      every step, pinned host memory) -- what a caller pays who treats the engine as a stateless operator.  The reference
    #      never does this (its buffers live on the GPU, only the 56-byte uniform crosses per frame), so `e2e` is the
    #      reference-faithful figure; this one is reported beside it.  Child process: whatever happens there, the line is printed.
    host_state = None
    if N == 1 and rank == 0 and not args.no_host_state and not args.only_headline:
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--host-state-only", "--agents", str(args.agents), "--width", str(args.width),
                   "--height", str(args.height), "--preset", args.preset, "--seed", str(args.seed), "--gaussian", str(args.gaussian)]
            if args.sd is not None: cmd += ["--sd", str(args.sd)]
            if args.sa is not None: cmd += ["--sa", str(args.sa)]
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
            host_state = json.loads(res.stdout.strip().splitlines()[-1]) if res.returncode == 0 else {"error": (res.stderr or "")[-200:]}
        except Exception as exc:      # noqa: BLE001
            host_state = {"error": repr(exc)[:200]}

    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline and not args.only_headline:
        v, ms_cpu, sample, cores, n_timed = cpu_reference_run(args.width, args.height, args.agents, settings, args.seed, steps=4, warmup=1,
                                                              budget_s=args.cpu_seconds, gaussian=args.gaussian)
        cpu = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample, "ms_per_step": ms_cpu}

    if rank == 0:
        line = {
            "impl": "ours", "metric": "agent_steps_per_sec", "value": head["value"], "unit": "agent-steps/s", "n_gpus": N,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, N, settings),
            "clocks": clk,
            "initial_state": head["initial_state"], "steady_state": head["steady_state"],
            "e2e": head["e2e"],
            "e2e_host_state": host_state,
            "gpu_launches": head["gpu_launches"],
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernels": kern,
            "parity_n": parity,
            "last_trail_mean": head.get("last_trail_mean"),
        }
        line.update(side)
        print(json.dumps(line), flush=True)
    if N > 1:
        h.dist.destroy_process_group()


def run_host_state(args):
    """Child process of the default run: three steps with the whole state crossing PCIe around each of them."""
    import torch
    import slime_mold_b200 as sm
    from slime_mold_b200.settings import SimSizeUniform
    settings = headline_settings(args)
    agents, width, height = args.agents, args.width, args.height
    be = sm.CudaBackend.new(width, height, settings, agent_count=agents, device=0,
                            flags=sm.SM_FLAG_GAUSSIAN_BLUR if args.gaussian else 0)
    be.init_agents(args.seed)
    be.step(20)
    uni = SimSizeUniform.new(width, height, settings.pheromone_decay_factor, settings)
    n_hs = 3
    pin_a = torch.empty((agents, 4), dtype=torch.float32, pin_memory=True)
    pin_t = torch.empty((height, width), dtype=torch.float32, pin_memory=True)
    a_np, t_np = pin_a.numpy(), pin_t.numpy()
    be.read_agents(out=a_np)
    be.read_trail(out=t_np)
    be.sync()
    t0 = time.perf_counter()
    for _ in range(n_hs):
        be.write_agents(a_np)
        be.write_trail(t_np)
        be.write_uniform(uni)
        be.step(1)
        be.read_agents(out=a_np)
        be.read_trail(out=t_np)
    be.sync()
    hs_s = time.perf_counter() - t0
    be.close()
    print(json.dumps({"value": agents * n_hs / hs_s, "unit": "agent-steps/s", "steps": n_hs, "ms_per_step": hs_s / n_hs * 1e3,
                      "h2d_bytes_per_step": int(a_np.nbytes + t_np.nbytes + 56), "d2h_bytes_per_step": int(a_np.nbytes + t_np.nbytes),
                      "note": "sm_upload_agents + sm_upload_trail + sm_set_params, sm_step(1), sm_download_agents + sm_download_trail "
                              "every step, pinned host buffers: PCIe-bound"}), flush=True)


def main():
    args = parse_args()
    if args.host_state_only:
        run_host_state(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
