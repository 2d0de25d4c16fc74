"""Statistical parity at BASELINE config 1 (north star: "deposition-order nondeterminism is handled by
comparing against an order-fixed atomic reference run and by field statistics").

The reference's shaders race (agents sense and deposit on one buffer, `diffuse_trail` blurs in place), so its output is a
family of runs; the engine computes the deterministic phase_split member and is bit-exact against the oracle's
phase_split.  This test runs the CUDA ENGINE at full config-1 size -- 1,000,000 agents on 1920 x 1080, Default preset,
1000 steps, 5 seeds -- and compares the trail field's mean, variance and occupancy at t = 10 / 150 / 300 / 500 / 1000 with

  * the oracle's phase_split run (tests/golden/statistics_config1.json): the same numbers to rounding (bit parity at this
    size is a statistic of 2 M cells, not a sample), and
  * the ORDER-FIXED run (`so_step_sequential`: agents in index order on one live buffer, pinned to the shader source by
    tests/test_wgsl_reference.py), with the Jacobi and with the reference's in-place raster diffusion: within the seed
    noise, thresholds in sigma below.

The goldens take ~55 CPU-minutes (the sequential runs are single-threaded by definition):
tests/golden/make_statistics_golden.py.  -m gpu."""
import json
import os

import numpy as np
import pytest

import slime_mold_b200 as sm

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "statistics_config1.json")
KEYS = ("mean", "var", "occupancy")


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/statistics_config1.json not generated")
    g = json.load(open(GOLD))
    if not all(m in g["modes"] and len(g["modes"][m]) == len(g["config"]["seeds"]) for m in ("sequential_inplace", "sequential_jacobi", "phase_split")):
        pytest.skip("statistics golden incomplete")
    return g


@pytest.fixture(scope="module")
def engine_stats(gold, engine_lib):
    c = gold["config"]
    W, H, N = c["width"], c["height"], c["agents"]
    s = sm.init_preset_manager().get_preset(c["preset"]).settings
    rows = []
    for seed in c["seeds"]:
        with sm.CudaBackend.new(W, H, s, agent_count=N, device=0) as be:
            be.init_agents(seed)                       # the oracle's generator, bit for bit (tests/test_gpu_parity.py)
            done, per_mark = 0, []
            for m in c["marks"]:
                be.step(m - done)
                done = m
                t = be.read_trail()
                t64 = t.astype(np.float64)
                st = be.trail_statistics()             # the device-side reduction must agree with the host's
                assert abs(st.sum / t.size - t64.mean()) < 1e-7 and int(st.nonzero) == int(np.count_nonzero(t))
                per_mark.append({"mean": float(t64.mean()), "var": float(t64.var()), "occupancy": float((t > 0.05).mean())})
            rows.append(per_mark)
    return rows


def _arr(rows, key):
    return np.array([[m[key] for m in per_mark] for per_mark in rows])        # [seed, mark]


def test_engine_reproduces_the_phase_split_statistics(gold, engine_stats):
    """Bit parity, seen through the statistics: same seeds, same semantics -> the same field."""
    for key in KEYS:
        a, b = _arr(engine_stats, key), _arr(gold["modes"]["phase_split"], key)
        assert np.allclose(a, b, rtol=1e-12, atol=1e-15), (key, np.abs(a - b).max())


@pytest.mark.parametrize("mode", ["sequential_jacobi", "sequential_inplace"])
def test_engine_vs_the_order_fixed_run(gold, engine_stats, mode):
    """Engine (phase_split) vs the order-fixed run, 5 seeds each; gap = |difference of the seed means|, sigma = pooled standard
    deviation over seeds.  MEASURED (profiles/README.md, "statistics at config 1"): with 2 M cells per field the seed noise is
    small (sigma/mean 0.4-5 %), so the schedule shows: identical while the map fills (t = 10: gap < 0.1 % of the mean), a
    first-order effect while the network coarsens (t = 150-300: 3-12 % of the mean / occupancy, 6-12 sigma), and 4-12 % /
    1-4 sigma from t = 500 on.  The sequential run is one extreme of the reference's family (every agent sees every earlier
    agent's deposit of the same frame), phase_split the other (none does); a GPU runs the reference's dispatch in between.
    The assertions are those measurements with a margin: they fail if the engine's statistics leave that corridor."""
    marks = gold["config"]["marks"]
    for key in KEYS:
        a, b = _arr(engine_stats, key), _arr(gold["modes"][mode], key)
        for j, t in enumerate(marks):
            gap = abs(a[:, j].mean() - b[:, j].mean())
            sigma = np.sqrt(0.5 * (a[:, j].var(ddof=1) + b[:, j].var(ddof=1)))
            scale = max(abs(a[:, j].mean()), abs(b[:, j].mean()))
            what = f"{mode} {key} t={t}: gap {gap:.5f} ({100 * gap / scale:.1f} %), sigma {sigma:.5f}"
            if key == "var":
                assert gap < 0.20 * scale or gap < 5e-4, what        # t = 10: variances of 4e-4 vs 6e-4 (in-place blur is smoother)
            elif t == marks[0]:
                assert gap < 2e-3 * scale, what
            else:
                assert gap < 0.15 * scale, what
                if t >= 500:
                    assert gap < 5.0 * sigma, what


def test_table(gold, engine_stats, capsys):
    """Prints the comparison table (pytest -s) -- the one committed under profiles/."""
    marks = gold["config"]["marks"]
    with capsys.disabled():
        print("\nstatistic | t | engine (CUDA, phase_split) | sequential, Jacobi | sequential, in-place | gap/sigma (Jacobi, in-place)")
        for key in KEYS:
            a = _arr(engine_stats, key)
            for j, t in enumerate(marks):
                cells = [f"{a[:, j].mean():.4f} +- {a[:, j].std(ddof=1):.4f}"]
                gs = []
                for mode in ("sequential_jacobi", "sequential_inplace"):
                    b = _arr(gold["modes"][mode], key)
                    sigma = np.sqrt(0.5 * (a[:, j].var(ddof=1) + b[:, j].var(ddof=1)))
                    cells.append(f"{b[:, j].mean():.4f} +- {b[:, j].std(ddof=1):.4f}")
                    gs.append(f"{abs(a[:, j].mean() - b[:, j].mean()) / max(sigma, 1e-12):.2f}")
                print(f"{key} | {t} | " + " | ".join(cells) + " | " + ", ".join(gs))


# ---- SM_FLAG_SEM_INPLACE: the reference's racy semantics as an opt-in mode, validated by statistics only ---------------------
QUARTER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "statistics_quarter.json")


def test_inplace_mode_stays_inside_the_family(engine_lib):
    """The racy mode (agents sense and deposit on one live buffer, decay and diffuse in place -- compute.wgsl as written) has no
    bit-exact reference: two runs of it differ.  What can be asserted: its field statistics stay inside the corridor spanned by the
    two deterministic members of the same family -- phase_split (nobody sees a deposit of the same frame) and the order-fixed
    sequential run (everybody sees all earlier ones) -- widened by three seed sigmas and 2 % of the value; and two runs from the
    same seed are close but, in general, not equal.  960 x 540, 250,000 agents, 300 steps, 3 seeds (goldens:
    tests/golden/make_statistics_golden.py --quarter)."""
    if not os.path.exists(QUARTER):
        pytest.skip("tests/golden/statistics_quarter.json not generated")
    g = json.load(open(QUARTER))
    c = g["config"]
    W, H, N = c["width"], c["height"], c["agents"]
    s = sm.init_preset_manager().get_preset(c["preset"]).settings
    rows = []
    for seed in c["seeds"]:
        with sm.CudaBackend.new(W, H, s, agent_count=N, device=0, flags=sm.SM_FLAG_SEM_INPLACE) as be:
            be.init_agents(seed)
            done, per_mark = 0, []
            for m in c["marks"]:
                be.step(m - done)
                done = m
                t = be.read_trail()
                assert np.isfinite(t).all() and t.min() >= 0.0 and t.max() <= 1.0
                per_mark.append({"mean": float(t.astype(np.float64).mean()), "occupancy": float((t > 0.05).mean())})
            a = be.read_agents()
            assert np.isfinite(a).all() and a[:, 0].min() >= 0 and a[:, 0].max() <= W and a[:, 1].min() >= 0 and a[:, 1].max() <= H
            rows.append(per_mark)
    for key in ("mean", "occupancy"):
        e = _arr(rows, key)
        ps, sq = _arr(g["modes"]["phase_split"], key), _arr(g["modes"]["sequential_inplace"], key)
        for j, t in enumerate(c["marks"]):
            lo, hi = min(ps[:, j].mean(), sq[:, j].mean()), max(ps[:, j].mean(), sq[:, j].mean())
            sigma = max(ps[:, j].std(ddof=1), sq[:, j].std(ddof=1), e[:, j].std(ddof=1))
            slack = 3.0 * sigma + 0.02 * hi
            assert lo - slack <= e[:, j].mean() <= hi + slack, (key, t, e[:, j].mean(), lo, hi, sigma)
