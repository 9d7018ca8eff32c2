#!/usr/bin/env python3
"""One wave of a bench workload and nothing else: the command ncu wraps for per-launch metrics (tools/ncu_traffic.py merges its
CSV with the per-bounce ray counts this script writes).  A number printed by a run under ncu is never a bench value.

    ncu --metrics <list> --clock-control none -k regex:traverseKernel --csv --log-file gpurun_out/wave_metrics.csv \
        python tools/profile_wave.py dragon 64 gpurun_out/wave_counts.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "dragon"
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else bench.SPP_PER_STEP
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "wave_counts.json")
    import torch
    from pathed_b200 import load_scene
    w = bench.WORKLOADS[workload]
    ctx = load_scene(w["scene"], w["width"], w["height"], integrator=1 if w.get("integrator") == "VolumePathTracer" else 0)
    # one launch sequence (under ncu the launches run one after the other anyway): per-bounce ray counts pair with the launches, and the
    # launches are the ones bench.py's stage-timed pass measures
    ctx.set_option("lanes", int(sys.argv[4]) if len(sys.argv) > 4 else 1)
    accum = torch.zeros(w["width"] * w["height"] * 3, dtype=torch.float32, device="cuda")
    ctx.render_device(0x5EED, 0, spp, 0, w["last_bounce"], accum.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    extend, shadow = ctx.wave_counts(w["last_bounce"] + 2)
    json.dump({"workload": workload, "spp": spp, "extend_rays": extend, "shadow_rays": shadow, "source_hash": bench.source_hash()}, open(out, "w"))
    print("wave done:", sum(extend), "extend rays,", sum(shadow), "shadow rays")


if __name__ == "__main__":
    main()
