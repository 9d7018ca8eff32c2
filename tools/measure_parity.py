#!/usr/bin/env python3
"""Measures, on a GPU box, how far the CUDA path is from the reference on the SURVEY 8(d)-sized BSDF batches (2^16 tuples per
material configuration) and writes the per-configuration table the GPU tests gate on (tests/golden/bsdf_error_table.json).

    python tools/measure_parity.py [out.json]

Reference = the compiled reference (oracle/_ref/libpathed_ref_probe.so) when present, else the pinned CPU oracle.
For every configuration and every output (f, pdf, sample wi / pdf / throughput): the fraction of tuples within 1e-5 relative,
the 99.99th percentile and the maximum of the relative error -- for the CUDA path and, beside it, for the CPU oracle.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import reference_live as rl  # noqa: E402
from golden_inputs import BSDF_CONFIGS, bsdf_inputs, material_desc  # noqa: E402
from oracle_binding import oracle_context  # noqa: E402
from parity import GOLDEN, frac_within, make_isects  # noqa: E402

N = 1 << 16


def errors(got, want):
    out = {}
    for key in ("f", "pdf", "sample_wi", "sample_pdf", "sample_throughput"):
        ok, e = frac_within(got[key], want[key])
        out[key] = {"within_1e-5": ok, "p9999": float(np.quantile(e, 0.9999)), "max": float(e.max())}
        if e.max() > 1e-5:  # the worst tuples, for diagnosis
            worst = np.argsort(e)[-3:][::-1]
            out[key]["worst"] = [{"index": int(i), "err": float(e[i]), "got": np.asarray(got[key][i]).tolist(), "want": np.asarray(want[key][i]).tolist()} for i in worst]
    return out


def answers(api, name):
    wo, ng, ns, uv, wi, xi = bsdf_inputs(name, N)
    mat = api.add_material(material_desc(BSDF_CONFIGS[name], api))
    if api.prefix == "ptc_":
        api.add_triangle_mesh([[0, 0, 0], [1, 0, 0], [0, 1, 0]], None, None, [[0, 1, 2]], mat)
        api.set_camera((0, 0, 5), (0, 0, 0), (0, 1, 0), 0.5, 8, 8)
        api.commit()
    isects = make_isects(wo, ng, ns, uv, mat)
    f, pdf = api.bsdf_eval(mat, isects, wi)
    swi, spdf, sthr = api.bsdf_sample(mat, isects, xi)
    return dict(f=f, pdf=pdf, sample_wi=swi, sample_pdf=spdf, sample_throughput=sthr)


def main():
    from pathed_b200 import create_context
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(GOLDEN, "bsdf_error_table.json")
    table = {"n_tuples": N, "tolerance": 1e-5, "configs": {}}
    table["reference"] = "compiled reference (oracle/_ref)" if rl.have_probe() else "CPU oracle"
    for name in sorted(BSDF_CONFIGS):
        orc = answers(oracle_context(), name)
        want = rl.reference_bsdf(name, N, os.path.join(GOLDEN, "texture_test.png")) if rl.have_probe() else orc
        table["configs"][name] = {"oracle": errors(orc, want)}
        try:
            table["configs"][name]["cuda"] = errors(answers(create_context(0), name), want)
        except Exception as e:  # no GPU here: the oracle's column alone (what the CPU suite pins)
            print("no CUDA column:", e)
        c = table["configs"][name].get("cuda", table["configs"][name]["oracle"])
        print("%-22s f %.5f (max %.1e)  pdf %.5f (max %.1e)  wi %.5f (max %.1e)  spdf %.5f (max %.1e)  thr %.5f (max %.1e)" % (
            name, c["f"]["within_1e-5"], c["f"]["max"], c["pdf"]["within_1e-5"], c["pdf"]["max"], c["sample_wi"]["within_1e-5"],
            c["sample_wi"]["max"], c["sample_pdf"]["within_1e-5"], c["sample_pdf"]["max"], c["sample_throughput"]["within_1e-5"],
            c["sample_throughput"]["max"]))
    with open(out_path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
