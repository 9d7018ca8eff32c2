#!/usr/bin/env python3
"""Summarises `ncu -i <rep> --page source --csv --print-source sass` output: per kernel, executed instructions, average active lanes
and stall samples in blocks of N SASS instructions (to find where the issue slots of an issue-bound kernel go).

    ncu -i gpurun_out/x.ncu-rep --page source --csv --print-source sass > /tmp/sass.csv; python tools/sass_profile.py /tmp/sass.csv [block]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    block = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kernels = []
    for r in rows:
        if r and r[0] == "Kernel Name":
            kernels.append([r[1], None, []])
        elif r and r[0] == "Address":
            kernels[-1][1] = r
        elif kernels and kernels[-1][1] and len(r) == len(kernels[-1][1]):
            kernels[-1][2].append(r)
    seen = set()
    for name, hdr, data in kernels:
        ix = {h: i for i, h in enumerate(hdr)}
        key = (name, len(data), sum(float(r[ix["Instructions Executed"]] or 0) for r in data))
        if key in seen:  # ncu prints every launch twice (SASS view of the source page and of the PTX page)
            continue
        seen.add(key)
        f = lambda r, k: float(r[ix[k]] or 0)
        ti = sum(f(r, "Instructions Executed") for r in data)
        tt = sum(f(r, "Thread Instructions Executed") for r in data)
        ts = sum(f(r, "# Samples") for r in data)
        print("==", name[:90], "SASS", len(data), "inst %.3g" % ti, "lanes %.1f" % (tt / max(ti, 1)), "samples", ts)
        for k in range(0, len(data), block):
            blk = data[k:k + block]
            i = sum(f(r, "Instructions Executed") for r in blk)
            t = sum(f(r, "Thread Instructions Executed") for r in blk)
            s = sum(f(r, "# Samples") for r in blk)
            ops = {}
            for r in blk:
                src = r[ix["Source"]].split()
                op = src[1] if src and src[0].startswith("@") else (src[0] if src else "")
                op = op.split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            top = sorted(ops.items(), key=lambda x: -x[1])[:6]
            stalls = {h: sum(f(r, h) for r in blk) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
            st = sorted(stalls.items(), key=lambda x: -x[1])[:3]
            print("%5d inst %5.1f%% lanes %4.1f samples %5.1f%%  %s  %s" % (k, 100 * i / max(ti, 1), t / max(i, 1), 100 * s / max(ts, 1),
                  " ".join("%s:%d" % o for o in top), " ".join("%s:%.0f%%" % (a[6:], 100 * b / max(s, 1)) for a, b in st)))


if __name__ == "__main__":
    main()
