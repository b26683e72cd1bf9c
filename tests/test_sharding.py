"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard the rays / the image rows,
compute their share with the oracle standing in for the device, and one collective reassembles
the result (rodent_b200/sharding.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from rodent_b200 import sharding  # noqa: E402


def test_ray_ranges_tile_the_job():
    for n in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            edges = [sharding.ray_range(r, world, n) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.ray_range(2, 2, 10)


def test_row_bands_partition_the_image():
    for height, parts, band in ((720, 8, 8), (67, 3, 8), (5, 2, 1), (1080, 4, 16)):
        rows = [sharding.rows_of(p, parts, height, band) for p in range(parts)]
        assert sorted(np.concatenate(rows).tolist()) == list(range(height))
        assert all(((r // band) % parts == p).all() for p, r in enumerate(rows))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import oracle
    from rodent_b200 import formats, render as R, testdata
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # traversal: each rank traces its contiguous range of a ragged ray set
        nodes, tris = formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)
        rays = formats.load_rays(testdata.rays("random"), 0.0, 1.0)[:10_001]
        b, e = sharding.ray_range(rank, world, len(rays))
        local = oracle.traverse(nodes, tris, np.ascontiguousarray(rays[b:e]), threads=2)
        hits = sharding.gather_hits(local, rank, world, len(rays))
        # the tensor form bench.py uses on the GPUs (equal shares): every rank sends (n, 4) int32 records
        even = torch.from_numpy(np.ascontiguousarray(local[:5000]).view(np.int32).reshape(-1, 4).copy())
        parts = [torch.empty_like(even) for _ in range(world)] if rank == 0 else None
        sharding.gather_hits_device(even, parts)
        if rank == 0:
            assert torch.equal(parts[0], even) and parts[1].shape == even.shape
            b1, _ = sharding.ray_range(1, world, len(rays))
            want1 = oracle.traverse(nodes, tris, np.ascontiguousarray(rays[b1:b1 + 5000]), threads=2)
            assert parts[1].numpy().tobytes() == want1.tobytes()
        # rendering: each rank keeps only the rows it owns, one reduce puts the film together
        scene = R.Scene.load_obj(ROOT / "tests" / "golden" / "cornell_box.obj")
        W, H = 48, 40
        cam = R.camera((0, 1, 2.7), (0, 0, -1), (0, 1, 0), 60.0, W, H)
        full, _ = oracle.render(scene.view, cam, W, H, 2, 4, 0, threads=2)
        mine = np.zeros_like(full)
        rows = sharding.rows_of(rank, world, H, band=8)
        mine[rows] = full[rows]
        film = sharding.reduce_film(torch.from_numpy(mine))
        if rank == 0:
            np.savez(Path(out_dir) / "rank0.npz", hits=hits.view(np.int32), film=film.numpy(), full=full,
                     want=oracle.traverse(nodes, tris, rays, threads=2).view(np.int32))
    finally:
        dist.destroy_process_group()


def test_two_ranks_reassemble_hits_and_film(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / "rank0.npz")
    assert z["hits"].tobytes() == z["want"].tobytes(), "gathered hit records differ from the single-process run"
    assert np.array_equal(z["film"], z["full"]), "reduced film differs from the single-process film"


def test_numa_binding_is_best_effort_without_a_gpu():
    info = sharding.pin_to_gpu_numa_node(0)
    assert set(info) >= {"numa_node", "cpus", "bound"} and (info["bound"] or "error" in info or info["numa_node"] in (None, -1))


def test_bench_shards_are_generated_by_the_ray_gen_rules():
    """bench.py --gpus N: rank r > 0 traces shard r of the job -- the reference's generator with the camera turned by
    3 r degrees / the random seed 42 + r; shard 0 is the pair of reference files."""
    sys.path.insert(0, str(ROOT))
    import bench
    a, b = bench.load_rays(0), bench.load_rays(1)
    for name in ("primary", "random"):
        assert len(a[name]) == len(b[name]) == 1 << 20
        assert a[name]["tmax"][0] == b[name]["tmax"][0]
        assert not np.array_equal(a[name]["dir"], b[name]["dir"])
    assert np.array_equal(a["primary"]["org"], b["primary"]["org"])                 # same eye, turned camera
    lo, hi = a["random"]["org"].min(0), a["random"]["org"].max(0)
    assert ((b["random"]["org"] >= lo - 1) & (b["random"]["org"] <= hi + 1)).all()   # same bounding box
    assert bench.load_rays(1)["random"].tobytes() == b["random"].tobytes()
