"""world_size-2 gloo test of the multi-GPU host logic (pathed_b200/distributed.py): the spp split over ranks followed by one
framebuffer reduce reproduces the single-rank image.  The per-rank renderer here is the CPU oracle (test infrastructure);
on the GPU box bench.py runs the same functions with ptc_render_device and NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

from pathed_b200.distributed import sample_block, split_samples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_blocks_tile_the_sample_axis():
    for world in (1, 2, 4, 8):
        seen = []
        for step in range(3):
            for rank in range(world):
                first, n = sample_block(step, rank, world, 16)
                seen += list(range(first, first + n))
        assert seen == list(range(3 * world * 16))
    for first, count, world in [(0, 8, 2), (5, 7, 4), (0, 1, 8), (3, 0, 2), (16, 16, 3)]:
        blocks = split_samples(first, count, world)
        assert len(blocks) == world
        flat = [s for a, n in blocks for s in range(a, a + n)]
        assert flat == list(range(first, first + count))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    from oracle_binding import oracle_scene
    from pathed_b200.distributed import init_from_env, reduce_framebuffer, sample_block
    r, _, w = init_from_env("gloo")
    assert (r, w) == (rank, world)
    scene = oracle_scene("scenes/cornell.json", 24, 24)
    fb = torch.zeros(24, 24, 3, dtype=torch.float32)
    for step in range(2):
        first, n = sample_block(step, r, w, 2)
        local = scene.render(77, first, n, 0, 4)
        fb += torch.from_numpy(local)
    # the per-step pattern of bench.py: the cumulative framebuffer keeps accumulating, a staging copy is reduced every step
    from pathed_b200.distributed import reduce_cumulative
    cumulative = torch.zeros(24, 24, 3, dtype=torch.float32)
    staging = torch.zeros_like(cumulative)
    for step in range(2):
        first, n = sample_block(step, r, w, 2)
        cumulative += torch.from_numpy(scene.render(77, first, n, 0, 4))
        reduce_cumulative(cumulative, staging, dst=0)
        if r == 0:
            np.save(os.path.join(out_dir, "per_step_%d.npy" % step), staging.numpy())
    reduce_framebuffer(fb, dst=0)
    if r == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), fb.numpy())
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_spp_split_and_reduce_match_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from oracle_binding import oracle_scene
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    reduced = np.load(str(tmp_path / "reduced.npy"))
    single = oracle_scene("scenes/cornell.json", 24, 24).render(77, 0, 8, 0, 4)
    # same samples, different fp32 summation order
    assert np.allclose(reduced, single, rtol=1e-5, atol=1e-6)
    assert reduced.sum() > 0
    # reduced every step without double counting (the N-rank image after k steps = the 1-rank image of the same samples)
    orc = oracle_scene("scenes/cornell.json", 24, 24)
    assert np.allclose(np.load(str(tmp_path / "per_step_0.npy")), orc.render(77, 0, 4, 0, 4), rtol=1e-5, atol=1e-6)
    assert np.allclose(np.load(str(tmp_path / "per_step_1.npy")), single, rtol=1e-5, atol=1e-6)
