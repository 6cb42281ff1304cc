import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """The checker and the product library are build artefacts (git-ignored)."""
    from oracle import pyoracle as po
    from or_cdchomp_b200 import capi
    if not po.available("port") or not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


_ensure_built()


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle as po
    return po


@pytest.fixture(scope="session")
def flavour(oracle):
    return oracle.best_flavour()


@pytest.fixture(scope="session")
def wam7():
    from or_cdchomp_b200 import models
    return models.wam7_robot()


@pytest.fixture(scope="session")
def table(oracle, flavour):
    """Config-1 scene: table slab SDF built by the oracle.  Returns a dict."""
    from or_cdchomp_b200 import capi, models
    kin_pose, prims, apos, aext = models.table_scene()
    sizes, lengths, gpose = models.field_geometry(apos, aext, 0.02, 0.2)
    gprims = models.prims_to_grid_frame(prims, gpose)
    pa = capi.make_prims(gprims)
    obs, sdf = oracle.computedistancefield(pa, len(gprims), sizes, lengths, 0.02, flavour=flavour)
    pose_world = models.pose_compose(kin_pose, gpose)
    return dict(sizes=sizes, lengths=lengths, gprims=gprims, obs=obs, sdf=sdf, pose_world=pose_world,
                desc=capi.SdfDesc(sdf, lengths, pose_world))


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from or_cdchomp_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def golden_path(name):
    return os.path.join(ROOT, "tests", "golden", name)
