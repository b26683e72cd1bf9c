import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

REFERENCE = Path("/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
