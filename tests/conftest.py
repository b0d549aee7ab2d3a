import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def orc():
    mod = graft.load_oracle()
    mod.build()
    return mod


@pytest.fixture(scope="session")
def workloads(pkg):
    from kcbs_b200 import workloads as w
    return w


@pytest.fixture(scope="session")
def gpu_ctx_factory(pkg):
    """Creates (and caches) one C-ABI context per scene on cuda:0.  Fails loudly without a GPU."""
    cache = {}

    def make(scene_or_world):
        key = scene_or_world if isinstance(scene_or_world, str) else id(scene_or_world)
        if key not in cache:
            world = pkg.scenes.world(scene_or_world) if isinstance(scene_or_world, str) else scene_or_world
            cache[key] = pkg.Context(world, device=0)
        return cache[key]

    yield make
    for c in cache.values():
        c.close()
