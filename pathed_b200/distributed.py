"""Multi-GPU plumbing: one process per GPU, samples-per-pixel split across ranks, one reduce of the fp32 framebuffer.

The path shards over samples: rank r of R renders its own block of global sample indices of every pixel,
Philox-keyed by (pixel, sample, bounce), so the union over ranks is the 1-GPU result up to fp32 summation order.  The only
communication is one sum-reduce of the 3*W*H framebuffer to rank 0 per step (NCCL over NVLink on GPUs; the same code runs
over gloo on CPU tensors in the tests).  The reference has no counterpart: its only parallelism is OpenMP over image rows
(/root/reference/src/sample_integrator.cpp:99).
"""
import os


def sample_block(step, rank, world, spp_per_rank):
    """First global sample index and count for `rank` in `step`: steps are contiguous runs of world*spp_per_rank samples,
    ranks take contiguous blocks inside a step."""
    return (step * world + rank) * spp_per_rank, spp_per_rank


def split_samples(first, count, world):
    """Partition [first, first+count) into `world` contiguous blocks (sizes differ by at most one; empty blocks allowed),
    as CudaPathTracer::run does for the GPUs of one process."""
    per = (count + world - 1) // world
    out = []
    for rank in range(world):
        a = min(first + rank * per, first + count)
        b = min(a + per, first + count)
        out.append((a, b - a))
    return out


def init_from_env(backend):
    """torchrun / torch.distributed.run environment (RANK, LOCAL_RANK, WORLD_SIZE, MASTER_*) -> (rank, local_rank, world)."""
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend)
    return rank, local_rank, world


def reduce_framebuffer(framebuffer, dst=0):
    """Sum the per-rank framebuffers into rank `dst` (in place there; other ranks' buffers are left unspecified).
    A buffer that went through this call must not be accumulated into again: rank `dst` now holds every rank's samples
    (use reduce_cumulative for a framebuffer that keeps accumulating across steps)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(framebuffer, dst=dst, op=dist.ReduceOp.SUM)
    return framebuffer


def reduce_cumulative(local, staging, dst=0):
    """Per-step reduce of framebuffers that keep accumulating (radianceLookup +=, src/sample_integrator.cpp:61-63): every rank's
    cumulative sum `local` is left untouched; its copy in `staging` is sum-reduced, so after the call `staging` on rank `dst` is the
    image of ALL samples rendered so far by all ranks (a sum of cumulative sums = the cumulative sum of the per-step sums)."""
    staging.copy_(local)
    return reduce_framebuffer(staging, dst)
