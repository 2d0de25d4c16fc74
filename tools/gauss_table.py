#!/usr/bin/env python
"""Markdown table of a Gaussian-kernel sweep (tools/bench_kernels.py gauss_rows -> *.jsonl): best chunk per
(map, radius, kernel), fraction of the measured HBM peak.   python tools/gauss_table.py profiles/r1_gauss_rows_sweep.jsonl"""
import json
import sys


def main(path):
    rows = [json.loads(l) for l in open(path)]
    best = {}
    for r in rows:
        if r["sweep"] != "gauss_rows":
            continue
        k = (r["size"], r["radius"], r["kernel"])
        if k not in best or r["ms_per_pass"] < best[k]["ms_per_pass"]:
            best[k] = r
    order = ["rows", "rows_packed", "rows_pk0", "rows_pk1", "rows_pk2", "stream", "stream_pk0", "stream_pk1", "wring", "tile"]
    kernels = sorted({k[2] for k in best}, key=lambda n: order.index(n) if n in order else 99)
    print("| map | radius | " + " | ".join(f"{k}: us/pass, GB/s, of measured peak" for k in kernels) + " |")
    print("|---|---|" + "---|" * len(kernels))
    for size in sorted({k[0] for k in best}):
        for rad in sorted({k[1] for k in best if k[0] == size}):
            cells = []
            for kern in kernels:
                r = best.get((size, rad, kern))
                cells.append("-" if r is None else f"{r['ms_per_pass'] * 1e3:.1f}, {r['gbs']:.0f}, **{r['frac_of_measured_peak']:.2f}**" + (f" (chunk {r['chunk']})" if r.get("chunk") else ""))
            print(f"| {size}^2 | {rad} | " + " | ".join(cells) + " |")
    full = [r for r in rows if r["sweep"] == "gauss_full_step"]
    if full:
        print("\nFull step in Gaussian mode at config-2 size (16.7 M agents, 4096^2; u8 deposit flags + sampler copy kept by the pass):\n")
        print("| radius | kernel | agent-steps/s | us/step | k_agents us | Gaussian pass us |\n|---|---|---|---|---|---|")
        for r in full:
            print(f"| {r['radius']} | {r['kernel']} | {r['agent_steps_per_s']:.3e} | {r['ms_per_step'] * 1e3:.1f} | {r['agents_ms'] * 1e3:.1f} | {r['trail_ms'] * 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
