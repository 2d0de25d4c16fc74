"""Multi-GPU strips on real devices (needs >= 2 B200s: run with `gpurun --gpus 2`): one process per
GPU, NCCL halo exchange + deposit-count exchange + agent migration, compared bit for bit with the
single-domain oracle.  -m gpu; skipped when fewer than 2 devices are visible."""
import os
import sys
import time

import numpy as np
import pytest

from conftest import bits_equal, mismatch_report
from presets_util import preset_uniform, random_trail, to_oracle_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    # name: (preset, W, H, N, steps, device_init)
    "waves_upload": ("Waves", 256, 512, 100_000, 40, False),
    "default_devinit": ("Default", 512, 768, 300_000, 35, True),
    "curls_upload": ("Curls", 128, 384, 60_000, 50, False),
    # ghost depth == strip height (the ghosts cover the whole neighbour): seam folding must not depend on it
    "thin_strips": ("Default", 256, 128, 30_000, 40, False),
    "firecracker_devinit": ("Firecracker Trees", 1024, 1024, 1_000_000, 33, True),
    "mode_switch": ("Default", 256, 512, 150_000, 40, False),
    # a trail rectangle that touches only SOME strips, uploaded mid-run, followed by steps -- the deposit
    # representation of the next step must be the same collective decision on every rank
    "partial_upload": ("Default", 256, 512, 150_000, 30, False),
    # a strip with no agents at all (everything starts in the upper half of the map) still merges its neighbours' deposits
    "empty_strip": ("Default", 256, 512, 80_000, 30, False),
    # steps, diffusion-only passes (sm_diffuse_only on strips: overlapped ghost push + one barrier per pass), steps
    "diffuse_mix": ("Sponge", 512, 1024, 200_000, 30, False),
    # per-rank snapshot files mid-run, restored into freshly created engines (new communicator), then continued
    "snapshot_mid": ("Waves", 256, 512, 100_000, 40, False),
    # EXTENSION: Gaussian diffusion-only passes on strips (BASELINE config 5 at 2/4/8 GPUs): the pass reads R ghost rows of
    # each neighbour, refreshed after every pass; radius 3 runs the register-streaming kernel, radius 6 the shared-memory one
    "gauss_rows_diffuse": ("Default", 512, 512, 20_000, 6, False),
    "gauss_stream_diffuse": ("Default", 512, 512, 20_000, 6, False),
}
GAUSS = {"gauss_rows_diffuse": (3, 1.5), "gauss_stream_diffuse": (6, 3.0)}
# EXTENSION, round 2: FULL steps in Gaussian mode on strips (the pass pulls R deposit rows of each neighbour); u8 flags with the
# rows kernel, u32 counts (deposit 0.4) with the streaming kernel
GAUSS_FULL = {"gauss_rows_full": (3, 1.5, 1.0), "gauss_stream_full": (6, 3.0, 0.4)}
CASES.update({
    "gauss_rows_full": ("Default", 512, 512, 120_000, 12, False),
    "gauss_stream_full": ("Default", 512, 512, 120_000, 12, False),
    # the display pass on strips: every rank draws the frame rows that show its map rows; together they are the reference's frame
    "render": ("Waves", 256, 512, 100_000, 9, False),
    # sm_resize on strips (main.rs:954-1015): agents rescaled, zeroed fields of the new size, new strip boundaries (405 rows do
    # not divide evenly, so agents change owner), peer memory mapped again -- then more steps
    "resize": ("Default", 256, 512, 100_000, 24, False),
})
RESIZE_TO = (320, 405)
# u8 deposit flags in 8x8 tiles on strips (kernels.cuh flag_tile_offset; what strips of 2^23 cells and more get by default, forced
# here): peer stores into the neighbours' tiled fields, the tiled pull of the halo deposit rows, the boundary bands, the display pass,
# a deposit-mode change, an empty strip, thin strips (ghost depth == strip height)
for _c in ("default_devinit", "waves_upload", "thin_strips", "mode_switch", "empty_strip", "render", "firecracker_devinit"):
    CASES["tiled_" + _c] = CASES[_c]


def _worker(rank, world, case, out_dir, exchange):
    os.environ["SM_EXCHANGE"] = exchange
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch  # noqa: F401  (first, so the process ends up with torch's libnccl)
    import slime_mold_b200 as sm
    from oracle import slime_oracle as so
    preset, W, H, N, steps, device_init = CASES[case]
    if case.startswith("tiled_"):
        os.environ["SM_FLAG_LAYOUT"] = "tiled"
        case = case[len("tiled_"):]
    s = sm.init_preset_manager().get_preset(preset).settings
    def connect(be, tag):
        idf = os.path.join(out_dir, f"nccl_id_{tag}.bin")
        if rank == 0:
            uid = be.comm_unique_id()
            with open(idf + ".tmp", "wb") as f:
                f.write(uid)
            os.rename(idf + ".tmp", idf)
        else:
            t0 = time.time()
            while not os.path.exists(idf):
                time.sleep(0.05)
                assert time.time() - t0 < 120
            uid = open(idf, "rb").read()
        be.comm_init(uid)

    if case in GAUSS:
        s = s.clone(blur_radius=float(GAUSS[case][0]), blur_sigma=GAUSS[case][1], pheromone_diffusion_rate=0.8)
    if case in GAUSS_FULL:
        s = s.clone(blur_radius=float(GAUSS_FULL[case][0]), blur_sigma=GAUSS_FULL[case][1], pheromone_diffusion_rate=0.8,
                    pheromone_deposition_amount=GAUSS_FULL[case][2])
    be = sm.CudaBackend.new(W, H, s, agent_count=N, device=rank, rank=rank, world_size=world,
                            flags=sm.SM_FLAG_GAUSSIAN_BLUR if (case in GAUSS or case in GAUSS_FULL) else 0)
    connect(be, "a")
    if device_init:
        be.init_agents(seed=11)
    else:
        be.write_agents(so.init_agents(N, W, H, s.agent_speed_min, s.agent_speed_max, 11))
        be.write_trail(random_trail(W, H, seed=4))
    if case == "partial_upload":
        be.step(steps // 2)
        rect = random_trail(64, 40, seed=77) - np.float32(0.25)          # some negative cells: counts mode until the next pass
        be.write_trail(rect, x0=100, y0=20)                              # rows 20..59: the first strip only
        be.step(steps - steps // 2)
    elif case == "empty_strip":
        a0 = so.init_agents(N, W, H, s.agent_speed_min, s.agent_speed_max, 11)
        a0[:, 1] *= np.float32(0.45)                                     # nobody starts in the lower strips
        be.write_agents(a0)
        be.step(steps)
    elif case == "mode_switch":
        # deposit representation changes mid-run (flags <-> counts) on every rank at the same step
        for dep in (1.0, 0.3, 2.0, 0.05, 1.0):
            be.update_settings(s.clone(pheromone_deposition_amount=dep))
            be.step(steps // 5)
    elif case == "snapshot_mid":
        be.step(steps // 2)
        snap = os.path.join(out_dir, "snap.smb")
        be.save_snapshot(snap)                       # writes snap.smb.rank<r>
        be.close()
        be = sm.CudaBackend.new(W, H, sm.Settings.default(), agent_count=N, device=rank, rank=rank, world_size=world)
        connect(be, "b")
        be.load_snapshot(snap)
        be.step(steps - steps // 2)
    elif case in GAUSS:
        be.diffuse_only(steps)
    elif case == "resize":
        be.step(steps // 2)
        be.resize(*RESIZE_TO)
        be.step(steps - steps // 2)
    elif case == "diffuse_mix":
        be.step(steps // 3)
        be.diffuse_only(7)
        be.step(steps // 3)
        be.diffuse_only(1)
        be.step(steps // 3)
    else:
        be.step(steps)
    a = be.read_agents()
    t = be.read_trail()
    frame = np.zeros((1, 1, 4), np.uint8)
    if case == "render":
        lut = np.random.default_rng(3).integers(0, 256, 768).astype(np.uint8)
        be.set_lut(lut)
        frame = np.zeros((200, 333, 4), np.uint8)            # alpha 0 = "not written by this rank"
        be.render(333, 200, out=frame)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), agents=a, trail=t, owned=be.last_owned, local=be.local_agent_count, frame=frame)
    be.close()


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("world", [2, 4, 8])
def test_strips_on_gpus_equal_oracle(oracle, engine_lib, tmp_path, case, world, exchange):
    import slime_mold_b200 as sm
    if sm.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    preset, W, H, N, steps, device_init = CASES[case]
    if H // world < 32:
        pytest.skip("strips too thin for this case")
    if exchange == "nccl" and (case in GAUSS_FULL or case == "render" or case.startswith("tiled_")):
        pytest.skip("peer-store exchange only")
    if exchange == "nccl" and case not in ("waves_upload", "mode_switch", "thin_strips", "diffuse_mix", "gauss_rows_diffuse", "partial_upload", "resize"):
        pytest.skip("NCCL path: three representative cases")
    mp.spawn(_worker, args=(world, case, str(tmp_path), exchange), nprocs=world, join=True)
    if case.startswith("tiled_"):
        case = case[len("tiled_"):]
    u = preset_uniform(preset, W, H)
    ag = oracle.init_agents(N, W, H, u.agent_speed_min, u.agent_speed_max, 11)
    sim = oracle.Sim(to_oracle_params(oracle, u), ag, trail=None if device_init else random_trail(W, H, seed=4))
    if case == "partial_upload":
        sim.step(steps // 2)
        rect = random_trail(64, 40, seed=77) - np.float32(0.25)
        sim.trail[20:60, 100:164] = rect
        sim.step(steps - steps // 2)
    elif case == "empty_strip":
        a0 = ag.copy()
        a0[:, 1] *= np.float32(0.45)
        sim = oracle.Sim(to_oracle_params(oracle, u), a0, trail=random_trail(W, H, seed=4))
        sim.step(steps)
    elif case == "mode_switch":
        import slime_mold_b200 as sm2
        s0 = sm2.init_preset_manager().get_preset(preset).settings
        for dep in (1.0, 0.3, 2.0, 0.05, 1.0):
            ss = s0.clone(pheromone_deposition_amount=dep)
            sim.p = to_oracle_params(oracle, sm2.SimSizeUniform.new(W, H, ss.pheromone_decay_factor, ss))
            sim.step(steps // 5)
    elif case in GAUSS_FULL:
        import slime_mold_b200 as sm2
        R, sigma, dep = GAUSS_FULL[case]
        ss = sm2.init_preset_manager().get_preset(preset).settings.clone(blur_radius=float(R), blur_sigma=sigma, pheromone_diffusion_rate=0.8,
                                                                         pheromone_deposition_amount=dep)
        sim.p = to_oracle_params(oracle, sm2.SimSizeUniform.new(W, H, ss.pheromone_decay_factor, ss))
        for _ in range(steps):
            oracle.agents_phase_split(sim.agents, sim.trail, sim.counts, sim.p)
            sim.trail = oracle.trail_pass(sim.trail, sim.p, counts=sim.counts, gauss_radius=R, gauss_sigma=sigma)
    elif case == "render":
        sim.step(steps - 1)
        oracle.agents_phase_split(sim.agents, sim.trail, sim.counts, sim.p)
        pre = sim.trail.copy()
        oracle.deposit_merge(pre, sim.counts, u.pheromone_deposition_amount)      # (clears the counts)
        oracle.decay(pre, u.decay_factor)
        sim.trail = oracle.diffuse(pre, u.diffusion_rate)
        lut = np.random.default_rng(3).integers(0, 256, 768).astype(np.uint8)
        want = oracle.display(pre, lut, 333, 200)                               # the frame between decay and diffuse (main.rs:1202-1217)
        got = np.zeros_like(want)
        for r in range(world):
            f = np.load(tmp_path / f"rank{r}.npz")["frame"]
            mine = f[..., 3] != 0
            assert not (mine & (got[..., 3] != 0)).any(), "a frame row was drawn by two strips"
            got[mine] = f[mine]
        assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} frame bytes differ"
    elif case in GAUSS:
        import slime_mold_b200 as sm2
        R, sigma = GAUSS[case]
        ss = sm2.init_preset_manager().get_preset(preset).settings.clone(blur_radius=float(R), blur_sigma=sigma, pheromone_diffusion_rate=0.8)
        sim.p = to_oracle_params(oracle, sm2.SimSizeUniform.new(W, H, ss.pheromone_decay_factor, ss))
        for _ in range(steps):
            sim.trail = oracle.trail_pass(sim.trail, sim.p, counts=None, gauss_radius=R, gauss_sigma=sigma)
    elif case == "resize":
        sim.step(steps // 2)
        W2, H2 = RESIZE_TO
        a1 = sim.agents.copy()
        a1[:, 0] *= np.float32(W2) / np.float32(W)                  # main.rs:985-989
        a1[:, 1] *= np.float32(H2) / np.float32(H)
        W, H = W2, H2
        sim = oracle.Sim(to_oracle_params(oracle, preset_uniform(preset, W, H)), a1)
        sim.step(steps - steps // 2)
    elif case == "diffuse_mix":
        for n_steps, n_passes in ((steps // 3, 7), (steps // 3, 1), (steps // 3, 0)):
            sim.step(n_steps)
            for _ in range(n_passes):
                sim.trail = oracle.trail_pass(sim.trail, sim.p)
    else:
        sim.step(steps)
    a = np.full((N, 4), np.nan, np.float32)
    t = np.full((H, W), np.nan, np.float32)
    owned = 0
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        m = ~np.isnan(d["agents"][:, 0])
        assert not (m & ~np.isnan(a[:, 0])).any(), "an agent is owned by two strips"
        a[m] = d["agents"][m]
        tm = ~np.isnan(d["trail"])
        t[tm] = d["trail"][tm]
        owned += int(d["owned"])
        assert int(d["owned"]) == int(d["local"])
    assert owned == N
    assert bits_equal(a, sim.agents), mismatch_report(a, sim.agents, "agents")
    assert bits_equal(t, sim.trail), mismatch_report(t, sim.trail, "trail")
