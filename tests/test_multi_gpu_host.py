"""The N > 1 path on the CPU: world_size 2 and 3 over gloo.  Each rank produces its row shard (with the oracle standing in
for the device renderer, through the same wrt_params sharding fields) and the product's gather/interleave code
(zig-weekend-raytracer_b200/distributed.py) assembles the frame on rank 0, which must equal the unsharded render bit for
bit — the partition-independence the GPU path relies on."""
from __future__ import annotations

import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, height: int, width: int, out_path: str):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    import torch
    import torch.distributed as dist
    import wro_py as wro
    distributed = importlib.import_module("zig-weekend-raytracer_b200.distributed")

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = wro.OracleScene("emissive")
    cam = sc.camera(width, height)
    p = sc.params(width, height, 4, 8, seed=21, row_shard_index=rank, row_shard_count=world)
    part, st = sc.render(cam, p, wro.RNG_COUNTER, threads=2)
    assert part.shape[0] == distributed.local_rows(height, rank, world)
    local = torch.zeros((distributed.padded_rows(height, world), width, 4), dtype=torch.float64)
    local[: part.shape[0]] = torch.from_numpy(part)
    rays = torch.tensor([float(st.rays)], dtype=torch.float64)
    dist.all_reduce(rays, op=dist.ReduceOp.SUM)  # whole-job ray count, as bench.py aggregates it
    frame = distributed.gather_frame(dist, local, height, rank, world)
    if rank == 0:
        np.savez(out_path, frame=frame.numpy(), rays=rays.numpy())
    else:
        assert frame is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height,width", [(2, 20, 33), (3, 20, 33), (2, 7, 40)])
def test_gloo_row_shards_assemble_to_the_unsharded_frame(tmp_path, wro, world, height, width):
    import torch.multiprocessing as mp

    out = tmp_path / "frame.npz"
    mp.spawn(_worker, args=(world, _free_port(), height, width, str(out)), nprocs=world, join=True)
    got = np.load(out)
    sc = wro.OracleScene("emissive")
    full, st = sc.render(sc.camera(width, height), sc.params(width, height, 4, 8, seed=21), wro.RNG_COUNTER)
    np.testing.assert_array_equal(got["frame"].view(np.uint64), full.view(np.uint64))
    assert int(got["rays"][0]) == int(st.rays)
    sc.close()


def test_shard_row_arithmetic(wrt):
    distributed = importlib.import_module("zig-weekend-raytracer_b200.distributed")
    for height in (1, 7, 20, 1080):
        for world in (1, 2, 3, 4, 8):
            rows = [distributed.local_rows(height, r, world) for r in range(world)]
            assert sum(rows) == height
            assert max(rows) == distributed.padded_rows(height, world)
            # identical to the C ABI's own definition (wrt_params.row_shard_*)
            for r in range(world):
                p = wrt.Params(width=4, height=height, row_shard_index=r, row_shard_count=world)
                assert wrt.Context.local_rows(p) == rows[r]
