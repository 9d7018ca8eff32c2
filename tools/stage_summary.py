#!/usr/bin/env python3
"""Per-stage summary of one wave from the ncu per-launch metric CSV of tools/gpu_profile.sh (SURVEY / north_star: HBM and L2 GB/s,
bytes per ray or path, lane utilisation, issue utilisation, occupancy PER STAGE).  Durations are ncu's (serialised launches); rates are
bytes / that duration; `of HBM` / `of L2` use MEASURED_PEAKS.json and profiles/l2_peak.json.

    python tools/stage_summary.py gpurun_out/r02c_wave_metrics.csv gpurun_out/r02c_wave_counts.json > profiles/r02_stage_counters.md
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    counts = json.load(open(sys.argv[2]))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        per.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    l2 = json.load(open(os.path.join(ROOT, "profiles", "l2_peak.json")))["l2_read_gbs"]
    # PathTracer stages and the VolumePathTracer's (volumeTraverseKernel<0 / 1 / 2>: merged extend / surface shadow / scatter shadow)
    stage_of = lambda n: ("extend" if ("volumeTraverseKernel<0" in n or n.startswith("void traverseKernel<0")) else
                          "shadow" if ("volumeTraverseKernel<" in n or n.startswith("void traverseKernel<1")) else
                          "logic" if ("logicKernel" in n or "volumeLogicKernel" in n or "containerKernel" in n) else
                          "material" if ("materialKernel" in n or "volumeMaterialKernel" in n) else "generate" if "generateKernel" in n else
                          "resolve" if "accumulateKernel" in n else None)
    units = {"extend": sum(counts["extend_rays"]), "shadow": sum(counts["shadow_rays"]), "logic": sum(counts["extend_rays"]),
             "material": sum(counts["extend_rays"][1:]), "generate": counts["extend_rays"][0], "resolve": counts["extend_rays"][0]}
    unit_name = {"extend": "ray", "shadow": "ray", "logic": "path vertex", "material": "path vertex", "generate": "path", "resolve": "sample"}
    agg = collections.OrderedDict()
    for m in per.values():
        s = stage_of(m["name"])
        if not s:
            continue
        a = agg.setdefault(s, collections.defaultdict(float))
        t = m["gpu__time_duration.sum"]
        a["ns"] += t; a["launches"] += 1
        a["dram"] += m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        a["l2"] += m["lts__t_bytes.sum"]; a["l1"] += m.get("l1tex__t_bytes.sum", 0.0)
        a["inst"] += m["smsp__inst_executed.sum"]
        for k, metric in (("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("lanes", "smsp__thread_inst_executed_per_inst_executed.ratio"),
                          ("l1hit", "l1tex__t_sector_hit_rate.pct"), ("l2hit", "lts__t_sector_hit_rate.pct"), ("warps", "sm__warps_active.avg.pct_of_peak_sustained_active")):
            a[k] += m[metric] * t
    total = sum(a["ns"] for a in agg.values())
    print("| stage | launches | ms (ncu) | share | units | DRAM B / unit | DRAM GB/s (of HBM %.0f) | L2 B / unit | L2 GB/s (of L2 %.0f) | warp instr / unit | lanes / instr | issue active | warps active | L1 hit | L2 hit |" % (hbm, l2))
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for s, a in agg.items():
        n = max(units[s], 1)
        sec = a["ns"] * 1e-9
        print("| %s | %d | %.2f | %.1f %% | %.1f M %ss | %.0f | %.0f (%.2f) | %.0f | %.0f (%.2f) | %.1f | %.1f | %.0f %% | %.0f %% | %.0f %% | %.0f %% |" % (
            s, a["launches"], a["ns"] / 1e6, 100 * a["ns"] / total, n / 1e6, unit_name[s], a["dram"] / n, a["dram"] / sec / 1e9, a["dram"] / sec / 1e9 / hbm,
            a["l2"] / n, a["l2"] / sec / 1e9, a["l2"] / sec / 1e9 / l2, a["inst"] / n, a["lanes"] / a["ns"], a["issue"] / a["ns"], a["warps"] / a["ns"],
            a["l1hit"] / a["ns"], a["l2hit"] / a["ns"]))
    print("\nwave: %s, %d spp; %.2f ms of kernels under ncu (launches serialised, no overlap)" % (counts["workload"], counts["spp"], total / 1e6))


if __name__ == "__main__":
    main()
