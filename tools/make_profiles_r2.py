#!/usr/bin/env python
"""Round 2: turn the scratch files the GPU calls left in gpurun_out/ into the tracked evidence under profiles/.

    python tools/make_profiles_r2.py

* profiles/r2_prof_*_summary.txt     the ncu counters DESIGN.md quotes (tools/ncu_summary.py) + the L1TEX / L2 request counters
                                     that explain the agent kernel at sensor distance 225
* profiles/r2_sass_mix_k_agents.txt  static (cuobjdump) and dynamic (ncu source page) instruction mix of k_agents / k_trail_rows
* profiles/r2_micro_*.log            the micro-benchmarks that decided the layouts (tools/microbench)
* profiles/r2_probe_*.jsonl          the A/B runs (tools/probe.py), one JSON line per run
* profiles/r2_bench_*.log, r2_parity_*.log, r2_statistics_config1.log
* profiles/roofline_traffic.json     DRAM bytes per launch of the dominant kernels (bench.py's roofline.traffic)
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

EXTRA = [
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__t_sectors_pipe_tex_mem_texture.sum", "l1tex__t_requests_pipe_tex_mem_texture.sum",
    "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
    "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def summary(rep, dst, note=""):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    hdr, units, rows = raw(rep)
    extra = []
    for r in rows:
        for k in EXTRA:
            if k in hdr:
                extra.append(f"  {k:88s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    with open(dst, "w") as f:
        if note:
            f.write(note.rstrip() + "\n")
        f.write(txt)
        f.write("--- L1TEX / L2 request counters\n" + "\n".join(extra) + "\n")


def dram_bytes(rep):
    hdr, units, rows = raw(rep)
    tot = []
    for r in rows:
        t = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(r[hdr.index(m)].replace(",", ""))
            t += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[units[hdr.index(m)].lower()]
        tot.append(t)
    return sum(tot) / len(tot), len(tot)


def dynamic_mix(rep, per):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    c, cs = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= iE:
            continue
        try:
            e = int(r[iE])
        except ValueError:
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip()).split()[0].split(".")[0] if r[iS].strip() else "?"
        c[op] += e
        cs[op] += int(r[iN] or 0)
    tot, ts = sum(c.values()), max(sum(cs.values()), 1)
    lines = [f"kernel: {rows[0][1][:110]}", f"warp instructions executed: {tot}  ({tot / per:.1f} per {('agent-warp' if per < 1e6 else 'unit')})",
             "opcode      executed/unit   share   stall-sample share"]
    for op, e in c.most_common(32):
        lines.append(f"{op:10s} {e / per:12.2f} {100 * e / tot:7.1f}% {100 * cs[op] / ts:10.1f}%")
    return "\n".join(lines)


def static_mix(sym_regex):
    lib = os.path.join(ROOT, "slime_mold_b200", "libslime_b200.so")
    names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    out, cur, c = [], None, None
    for line in names.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if cur and c:
                out.append((cur, c))
            cur = m.group(1) if re.search(sym_regex, m.group(1)) else None
            c = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
            if m:
                c[m.group(1)] += 1
    if cur and c:
        out.append((cur, c))
    return out


def main():
    os.makedirs(P, exist_ok=True)
    reps = {
        "r2_prof_agents_sd225": "k_agents at BASELINE configs[2] (100 M agents, 8192^2, sensor distance 225 / angle 1.34); tools/r2/gpu_01.sh",
        "r2_prof_agents_sd225_dual": "the same launch with the (+4,+2)-shifted second sampler copy (measured and dropped); tools/r2/gpu_04.sh",
        "r2_prof_agents_c2": "k_agents at BASELINE configs[1] (16.7 M agents, 4096^2, Default preset); tools/r2/gpu_11.sh",
        "r2_prof_agents_c2_pipe": "the same with the software-pipelined loop (measured and dropped); tools/r2/gpu_12.sh",
        "r2_prof_agents_p2p": "k_agents<XM_P2P> (strip instantiation) in one process, self-peer (sm_tuning.debug_single_rank_strip); tools/r2/gpu_11.sh",
        "r2_prof_trail_8192": "k_trail_rows full step at 8192^2, row-by-row surface writes (before the whole-sector form); tools/r2/gpu_01.sh",
        "r2_prof_trail_4096_pairs": "k_trail_rows full step at 4096^2, whole-sector surface writes; tools/r2/gpu_06.sh",
        "r2_prof_agents_sd225_tiled": "k_agents at BASELINE configs[2] with the u8 deposit flags in 8x8 tiles (final tree); tools/r2/gpu_34.sh",
        "r2_prof_trail_8192_tiled": "k_trail_rows full step at 8192^2, tiled flags: one sector per tile and batch, bulk zeroing (final tree); tools/r2/gpu_34.sh",
        "r2_prof_agents_c2_tiled": "k_agents at BASELINE configs[1] with tiled flags (final tree); tools/r2/gpu_34.sh",
    }
    traffic = {}
    for name, note in reps.items():
        rep = os.path.join(G, name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        summary(rep, os.path.join(P, name + "_summary.txt"), "# " + note)
        b, n = dram_bytes(rep)
        traffic[name] = (b, n)
    tj = os.path.join(P, "roofline_traffic.json")
    t = json.load(open(tj)) if os.path.exists(tj) else {}
    for key, name in (("agents_sd225", "r2_prof_agents_sd225"), ("agents", "r2_prof_agents_c2"), ("trail", "r2_prof_trail_4096_pairs"),
                      ("trail_8192_rowwise", "r2_prof_trail_8192"),
                      # final tree (tiled flag field): these replace the entries above where both exist
                      ("agents_sd225", "r2_prof_agents_sd225_tiled"), ("agents", "r2_prof_agents_c2_tiled"), ("trail_8192", "r2_prof_trail_8192_tiled")):
        if name in traffic:
            t[key] = {"dram_bytes_per_launch": traffic[name][0], "launches": traffic[name][1], "source": name}
    json.dump(t, open(tj, "w"), indent=1)

    with open(os.path.join(P, "r2_sass_mix_k_agents.txt"), "w") as f:
        f.write("# Instruction mix of the hot kernels.  Dynamic = ncu source page (warp instructions executed per opcode), static = cuobjdump -sass\n")
        for name, per, what in (("r2_prof_agents_c2", 16777216 / 32, "k_agents, config 2 (per agent-warp = 32 agents)"),
                                ("r2_prof_agents_sd225", 100000000 / 32, "k_agents, configs[2] sd 225 (per agent-warp)"),
                                ("r2_prof_trail_4096_pairs", 4096 * 4096 / 128, "k_trail_rows full step, 4096^2 (per warp-row = 128 cells)")):
            rep = os.path.join(G, name + ".ncu-rep")
            if os.path.exists(rep):
                f.write(f"\n== dynamic: {what}\n" + dynamic_mix(rep, per) + "\n")
        for sym, c in static_mix(r"k_agentsILi0EiNS_8FetchTexELb1E|k_trail_rowsILi2ELb1ELi4ELb0E"):
            dem = subprocess.run(["c++filt", sym], capture_output=True, text=True).stdout.strip()
            f.write(f"\n== static: {dem[:120]}\n  total {sum(c.values())} SASS instructions; " +
                    ", ".join(f"{k} {v}" for k, v in c.most_common(24)) + "\n")
            f.write("  packed FP32 (FFMA2/FADD2/FMUL2): %d, TLD4: %d, SUST: %d, SHFL: %d\n" %
                    (c["FFMA2"] + c["FADD2"] + c["FMUL2"], c["TLD4"], c["SUST"], c["SHFL"]))

    for fn in sorted(os.listdir(G)):
        if re.match(r"r2_(micro_.*\.log|probe_.*\.jsonl|bench_.*\.log|parity_.*\.log|statistics_.*\.log|gauss_wring_sweep\.jsonl|launches_(config1|headline)\.csv)$", fn):
            shutil.copy(os.path.join(G, fn), os.path.join(P, fn))
    print("profiles/ updated:", len([f for f in os.listdir(P) if f.startswith("r2_")]), "round-2 files")


if __name__ == "__main__":
    main()
