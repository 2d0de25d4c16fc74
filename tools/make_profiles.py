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


if __name__ == "__main__":
    main()
