import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

REFERENCE = Path("/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_unavailable():
    """Reason why the gpu-marked tests cannot run here, or None.  On a box that has an NVIDIA device node nothing is
    skipped: there a missing or broken library must fail the tests, not hide them."""
    import os
    if os.path.exists("/dev/nvidia0"):
        return None
    from rodent_b200 import lib
    if not lib.LIB_PATH.exists():
        return f"{lib.LIB_PATH.name} is not built (python -m rodent_b200.build)"
    try:
        if lib.load().rodent_b200_device_count() <= 0:
            return "no CUDA device"
    except Exception as exc:                       # a library that does not load is a failure on a GPU box, a skip elsewhere
        return f"librodent_b200.so does not load: {exc}"
    return None


def pytest_collection_modifyitems(config, items):
    """`-m gpu` on a box without a device (or without the built library) skips instead of failing."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    why = _gpu_unavailable()
    if why:
        skip = pytest.mark.skip(reason=why)
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def sponza():
    """(nodes, tris) of the Sponza BVH8 block."""
    from rodent_b200 import formats, testdata
    return formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)


@pytest.fixture(scope="session")
def ray_sets():
    """name -> Ray1 array (1 Mi rays each) with the reference's tmin/tmax."""
    from rodent_b200 import formats, testdata
    return {name: formats.load_rays(testdata.rays(name), tmin, tmax) for name, (tmin, tmax) in testdata.RAY_SETS.items()}


@pytest.fixture(scope="session")
def oracle_hits(sponza, ray_sets):
    """Oracle closest-hit records for both full ray sets (a few seconds on 8 threads)."""
    from oracle import oracle
    nodes, tris = sponza
    return {name: oracle.traverse(nodes, tris, rays) for name, rays in ray_sets.items()}
