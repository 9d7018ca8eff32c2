#!/usr/bin/env python3
"""In-process sweep over run-time knobs on the GPU box: every configuration gets a fresh context (environment hooks such as
PTC_PLOC_TOP, PTC_TRAVERSE_PER_SM are read at ptc_create / ptc_commit; `lanes`, `paths_per_wave` are ptc_set_option names), renders a
few 64-spp steps of each workload and reports ms per step, Msamples/s, the BVH build time and whether the image equals the image of the
first configuration bit for bit (closest hits do not depend on the tree, sample order does not depend on the lanes).

    python tools/sweep_env.py dragon,teapot 4 "" "lanes=2" "lanes=2,PTC_TRAVERSE_PER_SM=6" "PTC_PLOC_TOP=64"
        ->  gpurun_out/sweep_env.txt
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OPTIONS = ("lanes", "paths_per_wave", "overlap_shadow")


def main():
    workloads = sys.argv[1].split(",")
    steps = int(sys.argv[2])
    configs = sys.argv[3:] or [""]
    import torch
    import bench
    from pathed_b200 import load_scene
    out = open(os.path.join(ROOT, "gpurun_out", "sweep_env.txt"), "a")
    spp, seed = 64, 0x5EED
    hooks = set()
    for c in configs:
        hooks |= {kv.split("=")[0] for kv in c.split(",") if kv and kv.split("=")[0] not in OPTIONS}
    for workload in workloads:
        w = bench.WORKLOADS[workload]
        width, height, last = w["width"], w["height"], w["last_bounce"]
        integrator = 1 if w.get("integrator") == "VolumePathTracer" else 0
        reference = None
        for c in configs:
            pairs = dict(kv.split("=") for kv in c.split(",") if kv)
            for h in hooks:
                os.environ.pop(h, None)
            for k, v in pairs.items():
                if k not in OPTIONS:
                    os.environ[k] = v
            ctx = load_scene(w["scene"], width, height, integrator=integrator)
            for k, v in pairs.items():
                if k in OPTIONS:
                    ctx.set_option(k, int(v))
            build_ms = ctx.stats().bvh_build_ms
            buf = torch.zeros(height * width * 3, dtype=torch.float32, device="cuda")
            stream = torch.cuda.current_stream().cuda_stream
            ctx.render_device(seed, 0, spp, 0, last, buf.data_ptr(), stream)
            torch.cuda.synchronize()
            image = torch.nan_to_num(buf.clone())
            same = None
            if reference is None:
                reference = image
            else:
                same = bool(torch.equal(image, reference))
                if not same:
                    same = "no: %d values differ, max |d| %.3g" % (int((image != reference).sum()), float((image - reference).abs().max()))
            for i in range(2):
                ctx.render_device(seed, (i + 1) * spp, spp, 0, last, buf.data_ptr(), stream)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(steps):
                ctx.render_device(seed, (i + 3) * spp, spp, 0, last, buf.data_ptr(), stream)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / steps
            line = "%-14s %-44s %8.2f ms/step %8.1f Msamples/s  bvh build %6.2f ms  same image: %s" % (
                workload, c or "(default)", ms, width * height * spp / ms * 1e-3, build_ms, same)
            print(line, flush=True)
            out.write(line + "\n"); out.flush()
            ctx.close()
            del ctx, buf
            torch.cuda.empty_cache()


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("sweep took %.1f s" % (time.time() - t0))
