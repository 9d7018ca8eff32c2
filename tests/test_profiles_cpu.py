"""The measurement artefacts bench.py reads are of the committed build (no GPU needed)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_committed_ncu_capture_is_of_this_build():
    """roofline.traffic / l2_frac / issue_active come from profiles/traffic.json, which carries the hash of the CUDA sources its ncu capture
    was taken from; bench.py reports null for a capture of another build, so the committed pair has to match"""
    facts, source = bench.profile_facts("dragon")
    assert facts is not None, source
    assert facts["source_hash"] == bench.source_hash()
    # per-ray DRAM traffic far below the algorithmic bytes (the BVH is L2-resident), L2 traffic above them
    assert 20 < facts["extend_dram_bytes_per_ray"] < 200 and 800 < facts["extend_l2_bytes_per_ray"] < 3000
    assert 0.5 < facts["extend_issue_active"] < 1.0 and 8 < facts["extend_lanes_per_instruction"] <= 32


def test_measured_ceilings_are_recorded():
    l2 = json.load(open(os.path.join(ROOT, "profiles", "l2_peak.json")))
    assert l2["l2_read_gbs"] > l2["hbm_read_gbs"] > 1000
    peak, note = bench.measured_peak()
    assert peak > 1000 and note


def test_both_arms_describe_the_same_workload():
    """config is a pure function of the command line: the reference arm and ours print the same object (the driver compares them)"""
    import argparse
    args = argparse.Namespace(workload="dragon", spp_per_step=bench.SPP_PER_STEP, spp_per_step_given=False, scaling="weak", paths_per_wave=0)
    one = bench.workload_config(args, 1)
    assert one == bench.workload_config(args, 1) and one["workload"].startswith("scenes/dragon.json 1024x1024 PathTracer")
    assert bench.workload_config(args, 8)["spp_per_step"] == 8 * one["spp_per_step"]
