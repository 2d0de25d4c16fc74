#!/usr/bin/env python
"""bench.py -- headline benchmark of the slime-mold step loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): agent-steps/s of the full step (agents -> decay -> diffuse,
/root/reference/src/main.rs:1163-1235) plus the diffusion kernel's HBM GB/s.

* N = 1 workload = BASELINE.json configs[1]: 16,777,216 agents on a 4096x4096 trail map,
  Default preset parameters (other presets: --preset / --all-presets).
* N > 1: the same per-GPU work per rank ("weak" scaling): the map grows to 4096*N rows,
  the agents to 16.7M*N, split into N horizontal strips with halo exchange + migration.
* `value`  : agent-steps/s with all state resident in HBM, CUDA events on the engine's
             stream, barrier + sync on both sides, max over ranks.
* `e2e`    : the same loop driven the way the reference drives it every frame -- write the
             56-byte uniform from host memory (main.rs:98), run the frame, read a field
             statistic back to the host -- one host round trip per step.
* `roofline`: the dominant kernel's algorithmic bytes / its CUDA-event time vs the measured
             HBM peak (MEASURED_PEAKS.json).
* `cpu_baseline` / `--impl reference`: the CPU restatement of compute.wgsl (oracle/, OpenMP)
             on this box's host cores -- the reference itself (Rust + wgpu) cannot be built
             or run in this image (no rustc, no Vulkan/lavapipe); see DESIGN.md.
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

CONFIG2 = dict(agents=16_777_216, width=4096, height=4096)
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="Default")
    ap.add_argument("--all-presets", action="store_true", help="also time every preset (extra keys)")
    ap.add_argument("--agents", type=int, default=CONFIG2["agents"], help="agents per GPU")
    ap.add_argument("--width", type=int, default=CONFIG2["width"])
    ap.add_argument("--height", type=int, default=CONFIG2["height"], help="map rows per GPU")
    ap.add_argument("--spinup", type=int, default=200, help="untimed steps before warm-up (network formation)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-state", action="store_true", help="skip the e2e_host_state leg (full state over PCIe every step)")
    ap.add_argument("--host-state-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--independent", action="store_true", help="diagnostic: N ranks, each an independent single-GPU engine")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e leg (0 = min(steps, 100))")
    ap.add_argument("--gaussian", type=int, default=0, metavar="R",
                    help="EXTENSION (BASELINE configs[2] 'large blur radius'): Gaussian blur of radius R (sigma R/2) instead of the "
                         "reference's 3x3 box; single GPU only")
    return ap.parse_args()


def bench_settings(args):
    """Preset parameters, plus the Gaussian extension's radius / sigma when --gaussian R is given."""
    import slime_mold_b200 as sm
    s = sm.init_preset_manager().get_preset(args.preset).settings
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
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # the median over the samples taken while the kernels were running
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm: CPU restatement of compute.wgsl on the host cores (oracle/, OpenMP)
# ----------------------------------------------------------------------------------------------
def cpu_reference_run(width, height, agents, preset, seed, steps, warmup, budget_s, gaussian=0):
    """Times the oracle (phase_split, all host threads) on a bounded sample of the workload.
    Returns (agent_steps_per_s, ms_per_step, sample description, cores)."""
    from oracle import slime_oracle as so            # bench.py's cpu_baseline / reference leg only
    import slime_mold_b200 as sm
    so.build()
    s = sm.init_preset_manager().get_preset(preset).settings
    if gaussian:
        s = s.clone(blur_radius=float(gaussian), blur_sigma=gaussian / 2.0)
    p = so.make_params(width, height, decay_factor=s.pheromone_decay_factor, agent_jitter=s.agent_jitter,
                       agent_speed_min=s.agent_speed_min, agent_speed_max=s.agent_speed_max,
                       agent_turn_speed=s.agent_turn_speed, agent_sensor_angle=s.agent_sensor_angle,
                       agent_sensor_distance=s.agent_sensor_distance, diffusion_rate=s.pheromone_diffusion_rate,
                       pheromone_deposition_amount=s.pheromone_deposition_amount)
    cores = so.max_threads()
    n = agents
    ag = so.init_agents(n, width, height, s.agent_speed_min, s.agent_speed_max, seed)
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
    total = steps + warmup
    frac = 1.0
    if t1 * total > budget_s and total > 0:
        # bounded sample: keep the full map (the trail pass is part of every step) and a prefix of the agents
        frac = max(min(1.0, budget_s / (t1 * total)), 1.0 / 64)
        n = max(int(agents * frac), 1)
        sim = Sim(p, ag[:n].copy(), trail=sim.trail)
    for _ in range(max(warmup - 1, 0)):
        sim.step(1)
    t0 = time.perf_counter()
    sim.step(steps)
    dt = time.perf_counter() - t0
    sample = (f"{n} of {agents} agents ({100.0 * n / agents:.1f}%) on the full {width}x{height} map, {steps} timed steps, "
              f"seed {seed}, preset {preset}{', Gaussian blur radius %d' % gaussian if gaussian else ''}, phase_split semantics, "
              f"OpenMP {cores} threads")
    return n * steps / dt, 1e3 * dt / steps, sample, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                        # rank 0 alone runs the CPU arm
    N = max(args.gpus, 1)
    width, height, agents = args.width, args.height * N, args.agents * N
    val, ms, sample, cores = cpu_reference_run(width, height, agents, args.preset, args.seed, args.steps, args.warmup,
                                               budget_s=150.0, gaussian=args.gaussian)
    line = {
        "impl": "reference", "metric": "agent_steps_per_sec", "value": val, "unit": "agent-steps/s", "n_gpus": N,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, N),
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of compute.wgsl (oracle/): the Rust+wgpu reference cannot be built here (no rustc/Vulkan)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, N):
    total_agents, total_h = args.agents * N, args.height * N
    if (args.agents, args.width, args.height) == (16777216, 4096, 4096):
        label = "BASELINE configs[1]" + (" per GPU (weak scaling)" if N > 1 else "")
    elif (total_agents, args.width, total_h) == (1000000, 1920, 1080):
        label = "BASELINE configs[0]"
    elif (total_agents, args.width, total_h) == (100000000, 8192, 8192):
        label = "BASELINE configs[2]"
    elif (total_agents, args.width, total_h) == (1000000000, 32768, 32768):
        label = "BASELINE configs[3]"
    else:
        label = "custom size"
    return {
        "workload": f"{label}: {args.agents} agents on a {args.width}x{args.height} trail map per GPU, "
                    f"preset {args.preset}; x{N} strips" if N > 1 else
                    f"{label}: {args.agents} agents on a {args.width}x{args.height} trail map, preset {args.preset}",
        "agents": args.agents * N, "width": args.width, "height": args.height * N, "preset": args.preset,
        "parallelism": f"strips{N}" if N > 1 else "single",
        "exchange": (os.environ.get("SM_EXCHANGE") or "p2p") if N > 1 else None,
        "spinup_steps": args.spinup, "seed": args.seed,
        "blur": (f"EXTENSION: separable Gaussian, radius {args.gaussian}, sigma {args.gaussian / 2.0} (no reference semantics)"
                 if args.gaussian else "3x3 box (compute.wgsl:164-195)"),
        "l2": (f"inputs larger than L2: agent state {args.agents * 20 / 1e6:.0f} MB/GPU is streamed every step (no flush between steps)"
               if args.agents * 20 > 126e6 else
               f"working set ({args.agents * 20 / 1e6:.0f} MB of agent state per GPU) fits the 126 MB L2: L2-resident, not a DRAM number; no flush"),
    }


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import slime_mold_b200 as sm
    from slime_mold_b200.settings import SimSizeUniform

    N = max(args.gpus, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if N > 1 and world != N:
        raise SystemExit(f"--gpus {N} needs torchrun with {N} ranks (WORLD_SIZE={world})")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if N > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    width, height, agents = args.width, args.height * N, args.agents * N
    settings = bench_settings(args)
    eng_flags = sm.SM_FLAG_GAUSSIAN_BLUR if args.gaussian else 0
    if args.gaussian and N > 1:
        raise SystemExit("--gaussian is single-GPU only (the extension is not built for strips)")
    if args.independent:
        width, height, agents = args.width, args.height, args.agents
        be = sm.CudaBackend.new(width, height, settings, agent_count=agents, device=local_rank, flags=eng_flags)
    else:
        be = sm.CudaBackend.new(width, height, settings, agent_count=agents, device=local_rank, rank=rank, world_size=N, flags=eng_flags)
    if N > 1 and not args.independent:
        ids = [be.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        be.comm_init(ids[0])
    be.init_agents(args.seed)
    stream = torch.cuda.ExternalStream(be.stream_handle, device=local_rank)

    def barrier():
        be.sync()
        torch.cuda.synchronize()
        if N > 1:
            dist.barrier()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if N > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # spin-up (untimed): let the network form so gather locality / deposit contention are the steady-state ones
    be.step(args.spinup)
    be.step(args.warmup)
    be.sync()

    # ---- value: device-resident throughput ------------------------------------------------
    be.set_timing_enabled(False)
    be.reset_timing()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_total = timed(lambda: be.step(args.steps))
    launches = be.timing().kernel_launches
    clk = clocks.stop() if rank == 0 else None
    value = agents * args.steps / (ms_total * 1e-3)

    # ---- per-kernel durations (CUDA events around every launch), same workload --------------
    be.set_timing_enabled(True)
    be.reset_timing()
    k_steps = min(args.steps, 64)
    be.step(k_steps)
    t = be.timing()
    be.set_timing_enabled(False)
    local_agents = be.local_agent_count
    rows_local = height // N
    cells_local = width * rows_local
    agent_ms = t.agents_ms / max(t.agent_launches, 1)
    trail_ms = t.trail_ms / max(t.trail_launches, 1)
    sort_ms_per_step = t.sort_ms / max(t.steps, 1)
    if os.environ.get("SM_SIDE_TIMING"):      # diagnostics: per-rank kernel split (only rank 0's goes into the JSON line)
        print(f"[rank {rank}] agents {agent_ms * 1e3:.1f} us  trail {trail_ms * 1e3:.1f} us  sort/step {sort_ms_per_step * 1e3:.1f} us  "
              f"exchange/step {t.exchange_ms / max(t.steps, 1) * 1e3:.1f} us  local agents {local_agents}", file=sys.stderr, flush=True)
    peak, peak_src = hbm_peak()
    # algorithmic bytes: agents 32 B/agent-step + trail sensing read 4 + deposit write 4 B/cell;
    # fused decay+diffuse 8 B/cell-pass
    agent_bytes = 32.0 * local_agents + 8.0 * cells_local
    trail_bytes = 8.0 * cells_local
    kern = {
        "agents": {"ms": agent_ms, "alg_bytes": agent_bytes, "gbs": agent_bytes / (agent_ms * 1e-3) / 1e9 if agent_ms else None},
        "trail": {"ms": trail_ms, "alg_bytes": trail_bytes, "gbs": trail_bytes / (trail_ms * 1e-3) / 1e9 if trail_ms else None},
        "sort_ms_per_step": sort_ms_per_step,
        "exchange_ms_per_step": t.exchange_ms / max(t.steps, 1),
    }
    dom = "agents" if agent_ms >= trail_ms else "trail"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_agents" if dom == "agents" else "k_trail_rows", "achieved": kern[dom]["gbs"],
                "peak": peak, "unit": "GB/s", "frac": (kern[dom]["gbs"] / peak) if kern[dom]["gbs"] else None,
                "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_launch": kern[dom]["alg_bytes"],
                "ms_per_launch": kern[dom]["ms"]}

    # ---- diffusion-only GB/s (the other half of BASELINE.json's metric) ---------------------
    diff_passes = 50
    be.diffuse_only(5)
    ms_diff = timed(lambda: be.diffuse_only(diff_passes))
    diff_gbs = 8.0 * cells_local * N * diff_passes / (ms_diff * 1e-3) / 1e9

    # ---- e2e: host-driven frame loop (uniform from host each step, statistic back each step) ---
    e2e_steps = args.e2e_steps or min(args.steps, 100)
    uni = SimSizeUniform.new(width, height, settings.pheromone_decay_factor, settings)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        be.write_uniform(uni)            # 56 B host -> device (kernel parameter block)
        be.step(1)
        st = be.trail_statistics()       # 32 B device -> host, synchronises
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if N > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = agents * e2e_steps / float(e2e_t.item())

    extras = {}
    if args.all_presets and N == 1:
        per = {}
        for name in sm.init_preset_manager().get_preset_names():
            be.update_settings(sm.init_preset_manager().get_preset(name).settings)
            be.init_agents(args.seed); be.clear_trail()
            be.step(args.spinup)
            ms_p = timed(lambda: be.step(args.steps))
            per[name] = agents * args.steps / (ms_p * 1e-3)
        extras["agent_steps_per_sec_by_preset"] = per

    # ---- the other end of the e2e scale: the HOST owns the whole state (agents + trail uploaded before and read back after
    #      every step, pinned host memory) -- what a caller pays who treats the engine as a stateless operator.  The reference
    #      never does this (its buffers live on the GPU, only the 56-byte uniform crosses per frame), so `e2e` above is the
    #      reference-faithful figure; this one is reported beside it.  It runs in a child process (`--host-state-only`): whatever
    #      happens there, the headline line is printed.
    host_state = None
    if N == 1 and rank == 0 and not args.no_host_state:
        try:
            import subprocess
            cmd = [sys.executable, os.path.abspath(__file__), "--host-state-only", "--agents", str(args.agents), "--width", str(args.width),
                   "--height", str(args.height), "--preset", args.preset, "--seed", str(args.seed), "--gaussian", str(args.gaussian)]
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
            host_state = json.loads(res.stdout.strip().splitlines()[-1]) if res.returncode == 0 else {"error": (res.stderr or "")[-200:]}
        except Exception as exc:
            host_state = {"error": repr(exc)[:200]}

    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        v, ms_cpu, sample, cores = cpu_reference_run(width, height, agents, args.preset, args.seed, steps=3, warmup=1,
                                                     budget_s=args.cpu_seconds, gaussian=args.gaussian)
        cpu = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample, "ms_per_step": ms_cpu}

    if rank == 0:
        line = {
            "impl": "ours", "metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": N,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, N),
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": "agent-steps/s", "h2d_bytes_per_step": 56, "d2h_bytes_per_step": 32,
                    "steps": e2e_steps, "note": "per step: sm_set_params (56-byte uniform from host), sm_step(1), "
                                                "sm_trail_statistics read back (host sync every step)"},
            "e2e_host_state": host_state,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernels": kern,
            "diffusion": {"gbs": diff_gbs, "frac_of_peak": diff_gbs / (peak * N), "passes": diff_passes,
                          "ms_per_pass": ms_diff / diff_passes, "alg_bytes_per_cell": 8,
                          "note": "sm_diffuse_only on the same map (4096^2 = 64 MiB is L2-resident on one GPU; "
                                  "see profiles/ for the >L2 sweep)"},
            "last_trail_mean": st.sum / (cells_local),
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    be.close()
    if N > 1:
        dist.destroy_process_group()


def run_host_state(args):
    """Child process of the default run: three steps with the whole state crossing PCIe around each of them."""
    import torch
    import slime_mold_b200 as sm
    from slime_mold_b200.settings import SimSizeUniform
    settings = bench_settings(args)
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
