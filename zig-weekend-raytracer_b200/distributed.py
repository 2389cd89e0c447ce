"""Multi-GPU plumbing: one process per GPU, rows of the image interleaved over the ranks, one gather at the end.

The render path shards naturally (pixels and samples are independent; the reference already renders disjoint
(row, 32-column) jobs without synchronisation, render.zig:55-73).  Rank r renders image rows r, r + N, r + 2N, ... into a
dense local buffer (wrt_params.row_shard_index / row_shard_count); the only exchange is the final framebuffer assembly on
rank 0 — `torch.distributed.gather` over NCCL / NVLink on the GPU box, gloo in the CPU tests.  Sobol jitter and the
Philox stream are keyed by the global pixel index, so the assembled frame is bit-identical to a single-GPU render.
"""
from __future__ import annotations


def local_rows(height: int, rank: int, world: int) -> int:
    """Number of image rows rank `rank` renders (rows rank, rank + world, ...)."""
    if rank >= height:
        return 0
    return (height - rank + world - 1) // world


def padded_rows(height: int, world: int) -> int:
    """Row count of the equal-sized gather buffers (the largest shard)."""
    return (height + world - 1) // world


def gather_frame(dist, local, height: int, rank: int, world: int, out=None, gather_list=None):
    """Gather the ranks' dense row shards to rank 0 and interleave them into the full frame.

    `local`: tensor (padded_rows, width, lanes) whose first local_rows(...) rows are valid, on the backend's device.
    Returns the (height, width, lanes) frame on rank 0 (written into `out` when given), None elsewhere.
    """
    import torch

    if world == 1:
        frame = local[:height]
        if out is not None:
            out.copy_(frame)
            return out
        return frame
    if rank == 0:
        if gather_list is None:
            gather_list = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, gather_list, dst=0)
        if out is None:
            out = torch.empty((height,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        for r in range(world):
            n_r = local_rows(height, r, world)
            if n_r:
                out[r::world] = gather_list[r][:n_r]
        return out
    dist.gather(local, None, dst=0)
    return None
