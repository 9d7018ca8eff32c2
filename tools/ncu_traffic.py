#!/usr/bin/env python3
"""Merges an ncu per-launch metric CSV of tools/profile_wave.py with the wave's per-bounce ray counts into profiles/traffic.json:
DRAM bytes, L2 bytes, issue utilisation and lane utilisation of the extend and shadow kernels, per ray and per launch.
bench.py reads it for `roofline.traffic` / `l2_frac` / `issue_active` and refuses it when the source hash is not this build's.

    python tools/ncu_traffic.py gpurun_out/wave_metrics.csv gpurun_out/wave_counts.json profiles/traffic.json [copy-of-csv-in-profiles]
"""
import csv
import json
import sys


def main():
    csv_path, counts_path, out_path = sys.argv[1:4]
    counts = json.load(open(counts_path))
    rows = []
    header = None
    for r in csv.reader(open(csv_path, errors="replace")):
        if "Kernel Name" in r:
            header = r
            continue
        if header and len(r) == len(header):
            rows.append(dict(zip(header, r)))
    launches = {}
    for r in rows:
        if "traverseKernel" not in r["Kernel Name"]:
            continue
        launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    ordered = [launches[k] for k in sorted(launches)]
    extend = [l for l in ordered if "traverseKernel<0" in l["name"] or "traverseKernel<false" in l["name"]]
    shadow = [l for l in ordered if l not in extend]
    out = {"workload": counts["workload"], "spp": counts["spp"], "source_hash": counts["source_hash"]}
    for tag, ls, rays in (("extend", extend, counts["extend_rays"]), ("shadow", shadow, [c for c in counts["shadow_rays"] if c])):
        rays = [c for c in rays if c][:len(ls)]
        n = float(sum(rays))
        if not ls or not n:
            continue
        tot = lambda m: sum(l.get(m, 0.0) for l in ls)
        ns = tot("gpu__time_duration.sum")
        out[tag + "_launches"] = len(ls)
        out[tag + "_rays"] = int(n)
        out[tag + "_dram_bytes_per_ray"] = (tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")) / n
        out[tag + "_l2_bytes_per_ray"] = tot("lts__t_bytes.sum") / n
        out[tag + "_ns_under_ncu"] = ns
        # time-weighted means of the ratio metrics
        for key, metric in (("issue_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                            ("lanes_per_instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
                            ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
                            ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active")):
            if any(metric in l for l in ls):
                v = sum(l.get(metric, 0.0) * l.get("gpu__time_duration.sum", 0.0) for l in ls) / max(ns, 1.0)
                out[tag + "_" + key] = v / 100.0 if key == "issue_active" else v
        out[tag + "_per_launch"] = [{"rays": int(c), "us": l.get("gpu__time_duration.sum", 0.0) / 1e3,
                                     "dram_bytes": l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0),
                                     "l2_bytes": l.get("lts__t_bytes.sum", 0.0),
                                     "lanes": l.get("smsp__thread_inst_executed_per_inst_executed.ratio")} for c, l in zip(rays, ls)]
    out["source"] = "%s (ncu per-launch metrics over the %d traversal launches of one %d-spp wave of %s, build %s)" % (
        sys.argv[4] if len(sys.argv) > 4 else csv_path, len(ordered), counts["spp"], counts["workload"], counts["source_hash"])
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if not k.endswith("per_launch")}, indent=1))


if __name__ == "__main__":
    main()
