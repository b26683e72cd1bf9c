"""The sharded paths behind the C ABI, driven from ONE process over several devices (needs >= 2 GPUs on the box):
the host-pointer traversal entry points cut a call into contiguous ray ranges (no collective), the multi-device
renderer deals the row bands out and sums the films with ncclReduce called from C++."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from rodent_b200 import formats

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLDEN = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def devices():
    from rodent_b200 import lib
    n = lib.load().rodent_b200_device_count()
    if n < 2:
        pytest.skip(f"{n} CUDA device(s): the multi-device tests need two")
    return list(range(min(n, 4)))


def test_host_entry_points_over_several_devices(devices, sponza, ray_sets, oracle_hits):
    from rodent_b200 import lib, traversal
    L = lib.load()
    nodes, tris = sponza
    devs = (ctypes.c_int32 * len(devices))(*devices)
    L.rodent_b200_set_devices(devs, len(devices))
    try:
        for name in ("primary", "random"):
            got = traversal.intersect_host(nodes, tris, ray_sets[name])
            assert got.tobytes() == oracle_hits[name].tobytes(), name
        ragged = np.ascontiguousarray(ray_sets["random"][:100_003])
        assert traversal.intersect_host(nodes, tris, ragged).tobytes() == oracle_hits["random"][:100_003].tobytes()
        occl = traversal.intersect_host(nodes, tris, ragged, any_hit=True)
        assert ((occl["tri_id"] >= 0) == (oracle_hits["random"][:100_003]["tri_id"] >= 0)).all()
    finally:
        L.rodent_b200_set_device(0)


def test_multi_device_renderer_matches_one_device(devices):
    from rodent_b200 import render as R
    scene = R.Scene.load_obj(GOLDEN / "cornell_box.obj")
    W, H, spp, depth = 320, 200, 4, 6
    cam = R.camera((0, 1, 2.7), (0, 0, -1), (0, 1, 0), 60.0, W, H)
    one = R.Renderer(scene, 0, W, H, spp, depth)
    many = R.Renderer(scene, devices, W, H, spp, depth)
    for it in range(3):                      # the film accumulates over iterations on both
        one.render(cam, it)
        many.render(cam, it)
    a, b = one.film().copy(), many.film().copy()
    sa, sb = one.stats(), many.stats()
    one.free(); many.free()
    assert np.abs(a - b).max() <= 1e-5 * max(a.max(), 1.0), np.abs(a - b).max()
    assert sa["samples"] == sb["samples"] == W * H * spp
    assert sa["primary_rays"] == sb["primary_rays"] and sa["shadow_rays"] == sb["shadow_rays"]
    many2 = R.Renderer(scene, devices, W, H, spp, depth)
    many2.render(cam, 0)
    many2.clear()
    many2.render(cam, 0)
    one2 = R.Renderer(scene, 0, W, H, spp, depth)
    one2.render(cam, 0)
    assert np.abs(one2.film() - many2.film()).max() <= 1e-5 * max(a.max(), 1.0)
    one2.free(); many2.free()


def test_tools_take_gpus(devices, tmp_path):
    from rodent_b200 import testdata
    n = str(len(devices))
    r = subprocess.run([str(ROOT / "tools" / "bin" / "bench_traversal"), "-bvh", str(testdata.sponza_bvh8()), "-ray", str(testdata.rays("random")),
                        "--tmax", "1", "-s", "--bvh-width", "8", "--gpus", n, "--warmup", "2", "--bench", "3", "-o", str(tmp_path / "r.fbuf")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "959359 intersection(s)" in r.stdout
    png = tmp_path / "c.png"
    r = subprocess.run([str(ROOT / "tools" / "bin" / "rodent"), "--scene", str(GOLDEN / "cornell_box.obj"), "--eye", "0", "1", "2.7", "--dir", "0", "0", "-1",
                        "--up", "0", "1", "0", "--width", "256", "--height", "256", "--spp", "4", "--max-path-len", "8", "--gpus", n, "--bench", "4",
                        "-o", str(png)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "min/med/max Msamples/s" in r.stdout and png.stat().st_size > 1000


def _peer_worker(rank, world, port, out_dir):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import oracle
    from rodent_b200 import formats, lib, sharding, testdata, traversal
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        nodes, tris = formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)
        rays = formats.load_rays(testdata.rays("random"), 0.0, 1.0)[:200_000]
        n = len(rays)
        bvh = traversal.Bvh8(rank, nodes, tris)
        b, e = sharding.ray_range(rank, world, n)
        d_rays = traversal.DeviceArray.from_host(rank, np.ascontiguousarray(rays[b:e]))
        pb = sharding.PeerBuffer(rank, rank, n * 16)
        assert pb.ptr, "CUDA IPC / peer access not available between the two devices"

        class View:
            def __init__(self, ptr, count):
                self.ptr, self.count = ptr, count

        traversal.intersect(bvh, d_rays, View(pb.ptr + b * 16, e - b))       # records land in rank 0's memory
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            got = np.empty(n, formats.HIT1)
            lib.load().rodent_b200_copy_to_host(0, got.ctypes.data, pb.ptr, n * 16)
            want = oracle.traverse(nodes, tris, rays, threads=4)
            np.savez(Path(out_dir) / "peer.npz", got=got.view(np.int32), want=want.view(np.int32))
        dist.barrier()
        pb.close()
    finally:
        dist.destroy_process_group()


def test_peer_buffer_between_processes(devices, tmp_path):
    """One process per GPU: rank 1's traversal kernel writes its hit records straight into a buffer in rank 0's HBM
    (sharding.PeerBuffer: CUDA IPC mapping over NVLink), rank 0 finds the whole job's records there -- bit-identical to
    the oracle's."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_peer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / "peer.npz")
    assert z["got"].tobytes() == z["want"].tobytes()
