#!/usr/bin/env python
"""Field statistics of the ORDER-FIXED reference run at BASELINE config 1:
1,000,000 agents on 1920 x 1080, Default preset, 1000 steps, 5 seeds -- `so_step_sequential` (agents in index order on one
live buffer: the reference's shader executed by a single thread, pinned to the shader source by tests/test_wgsl_reference.py),
with the Jacobi diffusion and with the reference's in-place raster diffusion -- and of the oracle's phase_split run (the
engine's semantics).  Written to tests/golden/statistics_config1.json; tests/test_gpu_statistics.py runs the CUDA engine
on the same seeds and compares.

    python tests/golden/make_statistics_golden.py            # ~20 min of CPU (the sequential runs are single-threaded)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import slime_oracle as so  # noqa: E402

W, H, N = 1920, 1080, 1_000_000
SEEDS = (1, 2, 3, 4, 5)
MARKS = (10, 150, 300, 500, 1000)


def field_stats(t):
    t64 = t.astype(np.float64)
    return {"mean": float(t64.mean()), "var": float(t64.var()), "occupancy": float((t > 0.05).mean()),
            "nonzero": float((t != 0).mean()), "max": float(t.max())}


def quarter():
    """Quarter-scale companion (960 x 540, 250,000 agents, 300 steps, 3 seeds): the corridor the engine's opt-in racy mode
    (SM_FLAG_SEM_INPLACE) has to stay in -- tests/test_gpu_statistics.py::test_inplace_mode_stays_inside_the_family."""
    Wq, Hq, Nq, seeds, marks = 960, 540, 250_000, (1, 2, 3), (10, 150, 300)
    p = so.make_params(Wq, Hq)
    out = {"config": {"width": Wq, "height": Hq, "agents": Nq, "preset": "Default", "seeds": list(seeds), "marks": list(marks)}, "modes": {}}
    for mode, stepper in (("sequential_inplace", lambda sim, n: sim.step_sequential(n, inplace_diffuse=True)),
                          ("phase_split", lambda sim, n: sim.step(n))):
        rows = []
        for seed in seeds:
            sim = so.Sim(p, so.init_agents(Nq, Wq, Hq, 30.0, 50.0, seed))
            done, per_mark = 0, []
            for m in marks:
                stepper(sim, m - done)
                done = m
                per_mark.append(field_stats(sim.trail))
            rows.append(per_mark)
            print(f"quarter {mode} seed {seed}: mean@{marks[-1]} {per_mark[-1]['mean']:.4f}", flush=True)
        out["modes"][mode] = rows
    with open(os.path.join(ROOT, "tests", "golden", "statistics_quarter.json"), "w") as f:
        json.dump(out, f, indent=1)


def main():
    so.build()
    if "--quarter" in sys.argv:
        quarter()
        return
    p = so.make_params(W, H)            # Settings::default()
    out = {"config": {"width": W, "height": H, "agents": N, "preset": "Default", "seeds": list(SEEDS), "marks": list(MARKS)},
           "modes": {}}
    modes = (("sequential_inplace", lambda sim, n: sim.step_sequential(n, inplace_diffuse=True)),
             ("sequential_jacobi", lambda sim, n: sim.step_sequential(n, inplace_diffuse=False)),
             ("phase_split", lambda sim, n: sim.step(n)))
    for mode, stepper in modes:
        rows = []
        for seed in SEEDS:
            t0 = time.time()
            sim = so.Sim(p, so.init_agents(N, W, H, 30.0, 50.0, seed))
            done, per_mark = 0, []
            for m in MARKS:
                stepper(sim, m - done)
                done = m
                per_mark.append(field_stats(sim.trail))
            rows.append(per_mark)
            print(f"{mode} seed {seed}: {time.time() - t0:.0f} s  mean@1000 {per_mark[-1]['mean']:.4f}", flush=True)
        out["modes"][mode] = rows
        with open(os.path.join(ROOT, "tests", "golden", "statistics_config1.json"), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
