#!/usr/bin/env python
"""Turn the scratch files a GPU pass left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/make_profiles.py r1     # writes profiles/r1_*.{csv,txt,json}
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launch_share(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# per-kernel share of the ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n")
        f.write("kernel,launches,avg_us,total_ms,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k},{len(v)},{sum(v)/len(v)/1e3:.2f},{sum(v)/1e6:.3f},{100*sum(v)/tot:.1f}\n")


def ncu_text(rep, dst):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(dst, "w").write(out)
    return out


def traffic_json(rep):
    """dram bytes per launch of the hot kernels, for bench.py's roofline.traffic."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    acc = collections.defaultdict(list)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        key = "agents" if "k_agents" in name else ("trail" if "k_trail_rows" in name else None)
        if key is None:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(r[hdr.index(m)].replace(",", ""))
            u = units[hdr.index(m)].lower()
            tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        acc[key].append(tot)
    return {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v)} for k, v in acc.items()}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(P, exist_ok=True)
    if os.path.exists(os.path.join(G, "launches.csv")):
        shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches.csv"))
        launch_share(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launch_share.csv"))
    traffic = {}
    for name in ("prof_agents", "prof_trail", "prof_insitu"):
        rep = os.path.join(G, name + ".ncu-rep")
        if os.path.exists(rep):
            ncu_text(rep, os.path.join(P, f"{tag}_{name}_summary.txt"))
            for k, v in traffic_json(rep).items():
                v["source"] = f"{tag}_{name}"
                traffic[k] = v
    if traffic:
        json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
    for f in ("kernel_sweep.jsonl", "bench.log"):
        if os.path.exists(os.path.join(G, f)):
            shutil.copy(os.path.join(G, f), os.path.join(P, f"{tag}_{f}"))
    print(os.listdir(P))


if __name__ == "__main__" and "--readme" not in sys.argv:
    main()


# ----------------------------------------------------------------------------------------------
# README.md generator: python tools/make_profiles.py r1 --readme
# ----------------------------------------------------------------------------------------------
def _json_line(path):
    try:
        for l in open(path):
            if l.startswith("{"):
                return json.loads(l)
    except FileNotFoundError:
        pass
    return None


def _ncu_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"name": r[hdr.index("Kernel Name")]}
        for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                  "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
                  "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
                  "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum"):
            if k in hdr:
                v = float(r[hdr.index(k)].replace(",", ""))
                u = units[hdr.index(k)].lower()
                if u in ("kbyte", "mbyte", "gbyte"):
                    v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
                if u == "ms":
                    v *= 1e3
                if u == "ns":
                    v /= 1e3
                d[k] = v
        res.append(d)
    return res


def readme(tag):
    L = []
    A = L.append
    A(f"# profiles/ -- round-1 measurements ({tag})\n")
    A("All numbers from `gpurun` on one B200 (or N for the multi-GPU table) of this pool, image as in the task; "
      "timing by CUDA events on the engine's stream, never under a profiler. ncu captures use `--clock-control none`. "
      "Regenerate with `bash tools/gpu_profile.sh` (GPU box) then `python tools/make_profiles.py r1 --readme` (here).\n")
    peak = 6553.9
    b = _json_line(os.path.join(P, f"{tag}_bench.log"))
    if b:
        A("## Headline (bench.py, BASELINE configs[1]: 16,777,216 agents, 4096x4096, Default preset)\n")
        A("| quantity | value |\n|---|---|")
        A(f"| agent-steps/s, state resident in HBM (`value`) | **{b['value']:.3e}** ({b['ms_per_step']*1e3:.1f} us/step) |")
        A(f"| agent-steps/s, host-driven frame loop (`e2e`: 56 B uniform in, 32 B statistic out, one sync per step) | **{b['e2e']['value']:.3e}** |")
        k = b["kernels"]
        A(f"| k_agents per launch | {k['agents']['ms']*1e3:.1f} us -> {k['agents']['gbs']:.0f} GB/s algorithmic = {k['agents']['gbs']/peak:.2f} of measured HBM peak (issue-bound, see ncu) |")
        A(f"| k_trail_rows (full step) per launch | {k['trail']['ms']*1e3:.1f} us -> {k['trail']['gbs']:.0f} GB/s algorithmic (8 B/cell) |")
        A(f"| cell sort, amortised | {k['sort_ms_per_step']*1e3:.1f} us/step |")
        A(f"| diffusion-only on the same 4096^2 map (L2-resident: 64 MiB) | {b['diffusion']['gbs']:.0f} GB/s |")
        c = b.get("cpu_baseline")
        if c:
            A(f"| CPU restatement of compute.wgsl (oracle, OpenMP, {c['cores']} threads of the GPU box) | {c['value']:.3e} agent-steps/s ({c['ms_per_step']:.0f} ms/step) -> GPU/CPU = {b['value']/c['value']:.0f}x (`value`), {b['e2e']['value']/c['value']:.0f}x (`e2e`) |")
        A(f"| clocks during the timed region | {b['clocks']} |")
        A(f"| kernels launched in the timed region (`gpu_launches`) | {b['gpu_launches']} |\n")
    r = _json_line(os.path.join(G, "bench_reference.log"))
    if r:
        A(f"`bench.py --impl reference` (same box): {r['value']:.3e} agent-steps/s, sample: {r['cpu_baseline']['sample']}.\n")
    for name, title in (("bench_config1.log", "configs[0]-sized run (1,000,000 agents, 1920x1080, Default): launch-bound"),
                        ("bench_config3.log", "configs[2] (100,000,000 agents, 8192x8192, sensor distance 225 = Snake)")):
        d = _json_line(os.path.join(G, name))
        if d:
            k = d["kernels"]
            A(f"* {title}: **{d['value']:.3e}** agent-steps/s ({d['ms_per_step']*1e3:.1f} us/step; agents {k['agents']['ms']*1e3:.1f} us, "
              f"trail {k['trail']['ms']*1e3:.1f} us, sort {k['sort_ms_per_step']*1e3:.1f} us/step), e2e {d['e2e']['value']:.3e}.")
    A("")
    # multi-GPU
    rows = []
    for n in (1, 2, 4, 8):
        for x in ("", "_p2p", "_nccl"):
            d = _json_line(os.path.join(G, f"bench_n{n}{x}.log"))
            if d and x == "":
                rows.append((n, "-" if n == 1 else "p2p, overlapped", d))
    if rows:
        A("## Multi-GPU strips (weak scaling: 16.7 M agents and 4096 map rows per GPU)\n")
        A("| GPUs | exchange | agent-steps/s | us/step | agents us | trail us | exchange us/step | efficiency vs N=1 |\n|---|---|---|---|---|---|---|---|")
        base = next((d for n, x, d in rows if n == 1), None)
        for n, x, d in rows:
            k = d["kernels"]
            eff = d["value"] / (n * base["value"]) if base else float("nan")
            A(f"| {n} | {x} | {d['value']:.3e} | {d['ms_per_step']*1e3:.1f} | {k['agents']['ms']*1e3:.1f} | {k['trail']['ms']*1e3:.1f} | "
              f"{k.get('exchange_ms_per_step', 0)*1e3:.1f} | {eff:.2f} |")
            shutil.copy(os.path.join(G, f"bench_n{n}.log"), os.path.join(P, f"{tag}_bench_n{n}.log"))
        A("\n(`trail us` is the interior pass on strips; `exchange us/step` is what the overlap leaves exposed: the join and the "
          "refresh of the sampler's ghost rows. Session-1 numbers with the serial exchange and the first agent kernel: "
          "2 GPUs p2p 1.03e11 / nccl 8.9e10, 4 GPUs p2p 2.05e11 / nccl 1.87e11.)\n")
        c4 = _json_line(os.path.join(G, "bench_config4_n8.log"))
        if c4:
            k, df = c4["kernels"], c4["diffusion"]
            shutil.copy(os.path.join(G, "bench_config4_n8.log"), os.path.join(P, f"{tag}_bench_config4_n8.log"))
            A(f"**BASELINE configs[3]** (1,000,000,000 agents, 32768x32768, Default, 8 strips of 4096 rows on 8 B200): "
              f"**{c4['value']:.3e}** agent-steps/s ({c4['ms_per_step']*1e3:.0f} us/step; agents {k['agents']['ms']*1e3:.0f} us, interior trail "
              f"{k['trail']['ms']*1e3:.0f} us, sort {k['sort_ms_per_step']*1e3:.0f} us/step, exposed exchange {k['exchange_ms_per_step']*1e3:.0f} us/step), "
              f"e2e {c4['e2e']['value']:.3e}; target of the north star: >= 1e11. Diffusion-only on the same strips: "
              f"{df['gbs']/1e3:.1f} TB/s aggregate = {df['frac_of_peak']:.2f} of 8 x the measured copy peak ({df['ms_per_pass']*1e3:.0f} us/pass).\n")
    # sweeps
    sw = os.path.join(P, f"{tag}_kernel_sweep.jsonl")
    if os.path.exists(sw):
        rows = [json.loads(l) for l in open(sw)]
        A("## Diffusion-only sweep (`sm_diffuse_only`, BASELINE config 5, radius-1 box = the reference's parity mode)\n")
        A("8 algorithmic bytes per cell-pass. `rpc` = rows per chunk (0 = engine default; 16 when this sweep ran, 8 since). Peak = 6553.9 GB/s measured copy (MEASURED_PEAKS.json); nominal 8 TB/s.\n")
        A("| map | rpc | us/pass | GB/s | of measured peak | of 8 TB/s |\n|---|---|---|---|---|---|")
        for r in rows:
            if r["sweep"] == "diffusion":
                A(f"| {r['size']}^2 | {r['rows_per_chunk']} | {r['ms_per_pass']*1e3:.1f} | {r['gbs']:.0f} | {r['frac_of_measured_peak']:.3f} | {r['frac_of_8TBs']:.3f} |")
        A("\n4096^2 (64 MiB in + 64 MiB out) is L2-sized on B200: flagged, not a DRAM number.\n")
        A("## Cell-sort tile shape / interval (config 2, Default preset, steady state)\n")
        A("| tile | sort every | us/step | k_agents us | sort us/step |\n|---|---|---|---|---|")
        for r in rows:
            if r["sweep"] == "agents":
                if r.get("tile"):
                    A(f"| {r['tile'][0]}x{r['tile'][1]} | {r['sort_interval']} | {r['ms_per_step']*1e3:.1f} | {r['agents_ms']*1e3:.1f} | {r['sort_ms_per_step']*1e3:.1f} |")
                else:
                    A(f"| never sorted | - | {r['ms_per_step']*1e3:.1f} | | |")
        if any(r["sweep"] == "gauss" for r in rows):
            A("\n## EXTENSION: separable Gaussian, radius 2 / 4 / 8 (no reference semantics; BASELINE config 5 names radii 1-8)\n")
            A("`k_gauss_fused` (one pass, tiles + halos in shared memory) vs the two-pass form it replaced; 8 algorithmic bytes per cell-pass, sigma = R/2.\n")
            A("| map | radius | kernel | us/pass | GB/s | of measured peak |\n|---|---|---|---|---|---|")
            last = {}
            for r in rows:
                if r["sweep"] == "gauss":
                    last[(r["size"], r["radius"], r["kernel"])] = r          # the file is appended to: keep the latest run
            for (size, rad, kern), r in sorted(last.items()):
                A(f"| {size}^2 | {rad} | {kern} | {r['ms_per_pass']*1e3:.1f} | {r['gbs']:.0f} | {r['frac_of_measured_peak']:.2f} |")
            A("\n`fused_scalar` is the default. `fused_packed` halves the tap instructions (FFMA2) and is still slower: the kernel waits on its two "
              "barriers and on load latency at 4 CTAs/SM, not on FMA issue.")
        A("\n## All presets at config-2 size (16.7 M agents, 4096^2): uniform-random start vs steady state (>= 500 steps in)\n")
        A("| preset | initial agent-steps/s | steady agent-steps/s | steady us/step | trail mean | occupied cells |\n|---|---|---|---|---|---|")
        for r in rows:
            if r["sweep"] == "presets":
                A(f"| {r['preset']} | {r['initial_agent_steps_per_s']:.3e} | {r['steady_agent_steps_per_s']:.3e} | {r['steady_ms_per_step']*1e3:.1f} | {r['trail_mean']:.3f} | {r['occupied_frac']:.2f} |")
        A("")
    # launch share
    ls = os.path.join(P, f"{tag}_launch_share.csv")
    if os.path.exists(ls):
        A("## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), share of the stepped region\n")
        A("```\n" + open(ls).read().strip() + "\n```\n")
        A("(k_trail_stats and k_trail_rows<0,..> belong to bench.py's e2e / diffusion legs, not to the step.) "
          "The share agrees with the CUDA-event split above: agents ~3/4, trail ~1/5, sort ~5 %.\n")
    # ncu key counters
    A("## ncu --set full, key counters (full text: `*_summary.txt`)\n")
    A("| capture | kernel | us | DRAM rd+wr MB | DRAM % | SM % | issue active % | L1TEX % | L2 % | L1 hit % | L2 hit % | warps active % | regs | warp-instr |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for name in ("prof_agents", "prof_trail", "prof_insitu", "prof_diffusion"):
        rep = os.path.join(G, name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        seen = set()
        for d in _ncu_rows(rep):
            key = d["name"].split("(")[0]
            if key in seen:
                continue
            seen.add(key)
            g = lambda k: d.get(k, float("nan"))
            A(f"| {name} | `{key[:46]}` | {g('gpu__time_duration.sum'):.1f} | {(g('dram__bytes_read.sum')+g('dram__bytes_write.sum'))/1e6:.0f} | "
              f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g('l1tex__t_sector_hit_rate.pct'):.0f} | {g('lts__t_sector_hit_rate.pct'):.0f} | "
              f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {g('launch__registers_per_thread'):.0f} | {g('smsp__inst_executed.sum')/1e6:.1f} M |")
    A("\n`prof_agents` / `prof_trail` / `prof_diffusion`: default cache control (cold L2 per replay); `prof_insitu`: `--cache-control none`, "
      "kernels in their place in the step loop.\n")
    extra = os.path.join(P, "NOTES.md")
    if os.path.exists(extra):
        A(open(extra).read())
    open(os.path.join(P, "README.md"), "w").write("\n".join(L) + "\n")
    print("wrote profiles/README.md")


if __name__ == "__main__" and "--readme" in sys.argv:
    readme(sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "r1")
