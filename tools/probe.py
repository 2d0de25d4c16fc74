#!/usr/bin/env python
"""One timed run of the step loop on one GPU with explicit parameters -- the A/B harness of round 2.

    python tools/probe.py --agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 [--preset Default]
                          [--steps 40] [--spinup 60] [--tag name] [--dep 1.0] [--gaussian R]

Prints (and appends to gpurun_out/probe.jsonl) one JSON line: us per step (CUDA events around the stepped region),
per-launch times of the agent kernel / trail pass / sort (CUDA events around every launch, second run), and the
environment switches that were set.  Kernel variants are selected through the engine's environment switches by the caller.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import slime_mold_b200 as sm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=16_777_216)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--height", type=int, default=4096)
    ap.add_argument("--preset", default="Default")
    ap.add_argument("--sd", type=float, default=None)
    ap.add_argument("--sa", type=float, default=None)
    ap.add_argument("--dep", type=float, default=None)
    ap.add_argument("--jitter", type=float, default=None)
    ap.add_argument("--gaussian", type=int, default=0)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--spinup", type=int, default=100)
    ap.add_argument("--sort-interval", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-kernel-split", action="store_true")
    ap.add_argument("--fake-strips", type=int, default=0, metavar="N",
                    help="PROFILING ONLY: rank 0 of an N-strip engine that exchanges with itself (sm_tuning.debug_single_rank_strip): "
                         "the strip kernels and the overlapped exchange in one process, e.g. under ncu; results are meaningless")
    a = ap.parse_args()

    s = sm.init_preset_manager().get_preset(a.preset).settings
    ch = {}
    if a.sd is not None: ch["agent_sensor_distance"] = a.sd
    if a.sa is not None: ch["agent_sensor_angle"] = a.sa
    if a.dep is not None: ch["pheromone_deposition_amount"] = a.dep
    if a.jitter is not None: ch["agent_jitter"] = a.jitter
    if a.gaussian: ch.update(blur_radius=float(a.gaussian), blur_sigma=a.gaussian / 2.0)
    s = s.clone(**ch)
    if a.fake_strips > 1:
        os.environ["SM_FAKE_MULTI"] = "1"
        be = sm.CudaBackend.new(a.width, a.height * a.fake_strips, s, agent_count=a.agents * a.fake_strips, sort_interval=a.sort_interval,
                                rank=0, world_size=a.fake_strips)
    else:
        be = sm.CudaBackend.new(a.width, a.height, s, agent_count=a.agents, sort_interval=a.sort_interval,
                                flags=sm.SM_FLAG_GAUSSIAN_BLUR if a.gaussian else 0)
    be.init_agents(a.seed)
    stream = torch.cuda.ExternalStream(be.stream_handle)

    def timed(fn):
        be.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize()
        return e0.elapsed_time(e1)

    ms_init = timed(lambda: be.step(a.steps))          # uniform-random start
    be.step(max(a.spinup - a.steps, 0))
    ms = timed(lambda: be.step(a.steps))
    out = {"tag": a.tag, "agents": a.agents, "map": [a.width, a.height], "preset": a.preset, "sd": s.agent_sensor_distance,
           "sa": s.agent_sensor_angle, "dep": s.pheromone_deposition_amount, "jitter": s.agent_jitter, "gaussian": a.gaussian,
           "us_per_step": 1e3 * ms / a.steps, "us_per_step_initial": 1e3 * ms_init / a.steps,
           "agent_steps_per_s": a.agents * a.steps / (ms * 1e-3),
           "env": {k: v for k, v in os.environ.items() if k.startswith("SM_")}}
    if not a.no_kernel_split:
        be.set_timing_enabled(True); be.reset_timing()
        be.step(a.steps)
        t = be.timing()
        out.update(agents_us=1e3 * t.agents_ms / max(t.agent_launches, 1), trail_us=1e3 * t.trail_ms / max(t.trail_launches, 1),
                   sort_us_each=1e3 * t.sort_ms / max(t.sort_launches, 1), sort_us_per_step=1e3 * t.sort_ms / max(t.steps, 1))
    st = be.trail_statistics()
    out["trail_mean"] = st.sum / (a.width * a.height)
    be.close()
    line = json.dumps(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.jsonl"), "a") as f:
        f.write(line + "\n")
    print(line, flush=True)


if __name__ == "__main__":
    main()
