"""Multi-GPU strip protocol on CPU: the decomposition exchange.cu implements (halo rows, deposit-count
exchange, agent migration) must reproduce the single-domain oracle bit for bit.
In-process for 2/3/4 strips, and as two real processes over torch.distributed (gloo)."""
import os
import socket
import sys

import numpy as np
import pytest

import strip_model as smod
from conftest import bits_equal, mismatch_report
from presets_util import preset_uniform, random_trail, to_oracle_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_run(so, p, agents, trail, steps):
    sim = so.Sim(p, agents, trail=trail)
    sim.step(steps)
    return sim.agents, sim.trail


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("name", ["Default", "Waves", "Curls"])
def test_strips_equal_single_domain(oracle, world, name):
    W, H, N, steps = 96, 192, 12000, 12
    u = preset_uniform(name, W, H)
    p = to_oracle_params(oracle, u)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, seed=world)
    trail = random_trail(W, H, seed=3)
    ref_a, ref_t = reference_run(oracle, p, ag, trail, steps)
    a, t = smod.run_in_process(oracle, p, ag, trail, world, steps)
    assert bits_equal(a, ref_a), mismatch_report(a, ref_a, "agents")
    assert bits_equal(t, ref_t), mismatch_report(t, ref_t, "trail")


def test_fast_agents_cross_several_rows(oracle):
    # speed 300 -> 4.8 rows per step: deposit / migration reach m = 6 rows
    W, H, N, steps = 64, 128, 6000, 10
    u = preset_uniform("Default", W, H)
    u.agent_speed_min, u.agent_speed_max, u.pheromone_deposition_amount = 250.0, 300.0, 0.3
    p = to_oracle_params(oracle, u)
    ag = oracle.init_agents(N, W, H, 250.0, 300.0, seed=9)
    ref_a, ref_t = reference_run(oracle, p, ag, None, steps)
    a, t = smod.run_in_process(oracle, p, ag, None, 2, steps)
    assert bits_equal(a, ref_a) and bits_equal(t, ref_t)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import slime_oracle as so
    import strip_model
    from presets_util import preset_uniform, random_trail, to_oracle_params
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    W, H, N, steps = 96, 160, 10000, 10
    u = preset_uniform("Waves", W, H)
    p = to_oracle_params(so, u)
    ag = so.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, seed=5)
    trail = random_trail(W, H, seed=2)
    me = strip_model.run_distributed(so, p, ag, trail, steps, dist)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=me.ids, agents=me.agents, rows=me.trail[me.row0:me.row1],
             row0=me.row0, row1=me.row1)
    dist.barrier()
    dist.destroy_process_group()


def test_two_processes_over_gloo(oracle, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    W, H, N, steps = 96, 160, 10000, 10
    u = preset_uniform("Waves", W, H)
    p = to_oracle_params(oracle, u)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, seed=5)
    ref_a, ref_t = reference_run(oracle, p, ag, random_trail(W, H, seed=2), steps)
    a = np.full((N, 4), np.nan, np.float32)
    t = np.empty((H, W), np.float32)
    total = 0
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        a[d["ids"]] = d["agents"]
        t[int(d["row0"]):int(d["row1"])] = d["rows"]
        total += len(d["ids"])
    assert total == N
    assert bits_equal(a, ref_a), mismatch_report(a, ref_a, "agents")
    assert bits_equal(t, ref_t), mismatch_report(t, ref_t, "trail")
