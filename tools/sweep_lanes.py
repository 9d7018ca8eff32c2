#!/usr/bin/env python3
"""Sweep of the interleaved-lane launch scheme (ptc_set_option("lanes"), pathed_cuda.cu: launchWave) on the GPU box, in ONE process:
every configuration gets a fresh context (the grid hooks PTC_TRAVERSE_PER_SM / PTC_LOGIC_PER_SM / PTC_SHADE_PER_SM are read by
ptc_create), renders a few 64-spp steps of the workload, and its image is compared bit for bit with the one-lane image.

    python tools/sweep_lanes.py [workload] [steps] [quick|full]   ->  gpurun_out/sweep_lanes_<workload>.txt
"""
import itertools
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "dragon"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    mode = sys.argv[3] if len(sys.argv) > 3 else "full"
    import torch
    import bench
    from pathed_b200 import load_scene
    w = bench.WORKLOADS[workload]
    width, height, last = w["width"], w["height"], w["last_bounce"]
    integrator = 1 if w.get("integrator") == "VolumePathTracer" else 0
    spp, seed = 64, 0x5EED
    out = open(os.path.join(ROOT, "gpurun_out", "sweep_lanes_%s.txt" % workload), "w")
    if mode == "full":
        configs = [(l, t, s, g) for l, t, s, g in itertools.product((1, 2, 3, 4), (8, 6, 5, 4), (8, 4), (3, 2, 1))]
    else:
        configs = [(1, 8, 8, 3)] + [tuple(int(x) for x in c.split(",")) for c in mode.split(";")]
    reference = None
    for lanes, trav, shade, logic in configs:
        os.environ["PTC_TRAVERSE_PER_SM"] = str(trav)
        os.environ["PTC_SHADE_PER_SM"] = str(shade)
        os.environ["PTC_LOGIC_PER_SM"] = str(logic)
        ctx = load_scene(w["scene"], width, height, integrator=integrator)
        ctx.set_option("lanes", lanes)
        buf = torch.zeros(height * width * 3, dtype=torch.float32, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        ctx.render_device(seed, 0, spp, 0, last, buf.data_ptr(), stream)
        torch.cuda.synchronize()
        image = buf.clone()
        same = None
        if lanes == 1 and reference is None:
            reference = image
        elif reference is not None:
            same = bool(torch.equal(torch.nan_to_num(image), torch.nan_to_num(reference)))
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
        line = "lanes %d traverse/SM %d shade/SM %d logic/SM %d : %7.2f ms/step %7.1f Msamples/s  image == one-lane image: %s" % (
            lanes, trav, shade, logic, ms, width * height * spp / ms * 1e-3, same)
        print(line, flush=True)
        out.write(line + "\n"); out.flush()
        ctx.close()
        del ctx, buf
        torch.cuda.empty_cache()


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("sweep took %.1f s" % (time.time() - t0))
